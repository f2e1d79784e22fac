"""CPU: invariants of the persistent tile schedule of the HBM-bound kernels (dicp_b200/csrc/tile_stream.cuh),
checked by compiling the very same header for the host (tests/host_emul/plan_check.cpp, test infrastructure)."""
import ctypes
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "plan_check.cpp")
LIB = os.path.join(HERE, "host_emul", "libdicp_plan.so")
HDR = os.path.join(HERE, "..", "dicp_b200", "csrc", "tile_stream.cuh")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-x", "c++", SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


SHAPES = [(1, 1), (1, 65), (3, 65), (1000, 65), (64, 4096), (256, 16384), (255, 16384), (257, 16380), (2048, 16384), (1, 130000),
          (2, 130000), (7, 5000), (300, 513), (65535, 7), (1, 512), (1, 513), (5, 100000), (1480, 2048), (31, 16384)]


@pytest.mark.parametrize("N,n", SHAPES)
@pytest.mark.parametrize("tile,cps,k", [(512, 3, 4), (512, 3, 1), (256, 4, 8), (512, 2, 3), (1024, 1, 2)])
def test_schedule_covers_every_tile_once_with_unique_rows(lib, N, n, tile, cps, k):
    out = (ctypes.c_int * 4)()
    for sms in (148, 132, 1):
        rc = lib.dicp_plan_check(N, n, tile, cps, k, sms, out)
        assert rc == 0, "plan check failed with code %d for N=%d n=%d tile=%d cps=%d k=%d sms=%d" % (rc, N, n, tile, cps, k, sms)
        ctas, rows, max_t, min_t = out
        tiles = N * ((n + tile - 1) // tile)
        assert ctas <= sms * cps and ctas <= tiles
        assert max_t - min_t <= max(1, k) or tiles < ctas * k          # the final wave evens out what is left
        assert max_t <= -(-tiles // ctas) + k                           # nobody works more than a chunk beyond the average


def test_benchmark_shape_is_balanced(lib):
    out = (ctypes.c_int * 4)()
    assert lib.dicp_plan_check(256, 16384, 512, 3, 4, 148, out) == 0
    ctas, rows, max_t, min_t = out
    assert ctas == 444 and max_t == 19 and min_t >= 16 and rows <= 12
