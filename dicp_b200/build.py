"""Build libdicp_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m dicp_b200.build            # or: python dicp_b200/build.py

The library has no torch / Python dependency (plain C ABI, static cudart); it is loaded
with ctypes by dicp_b200/_lib.py.  The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdicp_b200.so")
SOURCES = ["nn_kernels.cu", "gn_kernels.cu", "bwd_kernels.cu", "kabsch_kernels.cu", "grid_nn_kernels.cu", "gumbel_kernels.cu"]
HEADERS = ["common.cuh", "dicp_math.cuh", "tile_stream.cuh", os.path.join("..", "..", "include", "dicp_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]   # IEEE sqrt/div are required for parity


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stdout.write(out)
        if p.returncode != 0:
            failed = True
            sys.stdout.write("nvcc failed for %s\n" % src)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    t0 = time.time()
    lib = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built %s in %.1fs" % (lib, time.time() - t0))
