"""Build tuning variants of libdicp_b200.so from the LIBRARY sources with -D knobs (no second copy of any kernel).

    python tools/build_variants.py gn_kernels.cu  u2m4:-DGN_U=2,-DGN_MINB=4  u4s6:-DGN_STAGES=6 ...

Every variant recompiles only the named source file with the extra flags and links it with the other objects
of dicp_b200/lib/ into variants/<name>/libdicp_b200.so (git-ignored; it travels to the GPU box).  Select one with
DICP_B200_LIB=variants/<name>/libdicp_b200.so (dicp_b200/_lib.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dicp_b200 import build as B  # noqa: E402


def main():
    src = sys.argv[1]
    B.build()
    procs = []
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition(":")
        out = os.path.join(ROOT, "variants", name)
        os.makedirs(out, exist_ok=True)
        obj = os.path.join(out, src.replace(".cu", ".o"))
        cmd = [B.NVCC] + B.FLAGS + [f for f in flags.split(",") if f] + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        procs.append((name, out, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, out, obj, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            sys.stdout.write(log)
            raise SystemExit("variant %s failed" % name)
        objs = [obj if s == src else os.path.join(B.LIBDIR, s.replace(".cu", ".o")) for s in B.SOURCES]
        subprocess.check_call([B.NVCC, "-shared", "-o", os.path.join(out, "libdicp_b200.so")] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"])
        print("built variants/%s" % name)


if __name__ == "__main__":
    main()
