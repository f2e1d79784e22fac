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


def test_header_constants_match_the_python_binding():
    """The flag / status-slot / layout constants of include/dicp_b200.h against dicp_b200/_lib.py, and the size and field
    offsets of struct dicp_params against the ctypes mirror (compiled with gcc from the header itself)."""
    import re
    import subprocess
    import tempfile
    from dicp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dicp_b200.h")).read()

    def define(name):
        return int(re.search(r"#define\s+%s\s+(-?\d+)" % name, hdr).group(1))

    assert define("DICP_FLAG_NO_SYNC") == _lib.DICP_FLAG_NO_SYNC
    assert define("DICP_STATE_STRIDE") == _lib.STATE_STRIDE
    assert define("DICP_STATE_ITERS_RUN") == _lib.STATE_ITERS_RUN and define("DICP_STATE_SINGULAR") == _lib.STATE_SINGULAR
    assert define("DICP_ERR_HOOK") == _lib.DICP_ERR_HOOK
    assert define("DICP_VERSION") == _lib.load_library().dicp_version()
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "layout.c")
        fields = [f[0] for f in _lib.DicpParams._fields_]
        with open(src, "w") as f:
            f.write('#include <stddef.h>\n#include <stdio.h>\n#include "dicp_b200.h"\nint main(void) {\n  printf("%zu", sizeof(dicp_params));\n')
            for name in fields:
                f.write('  printf(" %%zu", offsetof(dicp_params, %s));\n' % name)
            f.write("  return 0;\n}\n")
        exe = os.path.join(tmp, "layout")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        nums = [int(x) for x in subprocess.check_output([exe]).split()]
    assert nums[0] == ctypes.sizeof(_lib.DicpParams)
    assert nums[1:] == [getattr(_lib.DicpParams, name).offset for name in fields]


def test_sync_free_flag_is_only_set_for_constant_iteration_loops():
    from dicp_b200 import _functions, _lib
    mk = lambda const_iter, sf: _functions.make_params("pt2pl", 3, True, "cauchy", 0.5, 2.0, 5.0, 1e-12, 10, const_iter, 0.0, 1, "grid", sf)  # noqa: E731
    assert mk(True, True).flags == _lib.DICP_FLAG_NO_SYNC
    assert mk(False, True).flags == 0 and mk(True, False).flags == 0
    from dICP.ICP import ICP
    icp = ICP()
    assert icp.sync_free is False and icp.check_status() is None       # the reference's error timing by default
