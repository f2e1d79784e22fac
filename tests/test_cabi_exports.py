"""CPU: libdicp_b200.so loads and exports every symbol include/dicp_b200.h declares
(no compute calls: there is no GPU here), and argument validation works without a device."""
import ctypes
import os
import re

import pytest

from dicp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dicp_b200.h")).read()
    return sorted(set(re.findall(r"DICP_API\s+[\w\s\*]+?\b(dicp_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    if not os.path.exists(_lib.LIB_PATH):
        from dicp_b200 import build
        build.build()
    names = declared_symbols()
    assert len(names) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "python binding and header disagree"


def test_version_and_arg_errors():
    lib = _lib.load_library()
    assert lib.dicp_version() == 201
    assert lib.dicp_icp_workspace_bytes(256, 16384, 16384, 4) > 0
    assert lib.dicp_nn_workspace_bytes(1, 1, 1, 4) > 0
    # null pointers / bad sizes are rejected before any CUDA call
    rc = lib.dicp_nn_argmin_f32(None, None, 1, 1, 1, 3, None, None, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.dicp_last_error()
    p = _lib.DicpParams()
    rc = lib.dicp_icp_forward_f32(None, None, None, None, None, None, 1, 1, 1, 6, ctypes.byref(p), None, None, None, None, None, None, None,
                                  None, None, None, None, None, 0, None)
    assert rc == -1


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dICP.ICP import ICP
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ICP().icp(torch.zeros(4, 3), torch.zeros(4, 6), torch.eye(4))
    from dICP.nn import nn
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nn(differentiable=False).find_nn(torch.zeros(4, 3), torch.zeros(4, 3))


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/dicp_b200.h is the drop-in boundary: it must compile as C99 (no C++ / torch types) and a
    C program must link against the library through it (dicp_version needs no device)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "use_header.c"
    src.write_text('#include "dicp_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %zu\\n", dicp_version(), dicp_icp_workspace_bytes(2, 100, 100, 4)); return 0; }\n')
    exe = tmp_path / "use_header"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-ldicp_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) >= 100 and int(out[1]) > 0
