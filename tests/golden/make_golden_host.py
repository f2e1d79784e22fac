"""Golden vectors for the HOST side of the path: `ICP.batch_size_handling` (dICP/ICP.py:305-511) of the
UNMODIFIED reference on list / tensor / ragged / empty / weighted inputs, for both ICP types.

    python tests/golden/make_golden_host.py        # build container only (needs /root/reference)

Writes tests/golden/host_batch_handling.npz: for every case the inputs and the four outputs
(source, target, T_init, weight) the reference returns, or the exception type it raises."""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
import dICP  # noqa: E402
assert os.path.abspath(dICP.__file__).startswith(REF), dICP.__file__
from dICP.ICP import ICP  # noqa: E402

sys.path.insert(0, HERE)
from host_cases import host_cases  # noqa: E402  (the inputs, shared with tests/test_host_logic.py)

out = {}
for name, icp_type, make in host_cases():
    source, target, T_init, weight = make()
    icp = ICP(icp_type=icp_type, max_iterations=3)
    try:
        res = icp.batch_size_handling(source, target, T_init, weight)
    except Exception as exc:                       # noqa: BLE001 - the exception type is part of the contract
        out[name + "/raises"] = np.array(type(exc).__name__)
        print(name, "raises", type(exc).__name__)
        continue
    for key, val in zip(("source", "target", "T_init", "weight"), res):
        if val is not None:
            out[name + "/" + key] = val.detach().cpu().numpy()
            out[name + "/" + key + "_dtype"] = np.array(str(val.dtype))
    print(name, [tuple(v.shape) if v is not None else None for v in res])
path = os.path.join(HERE, "host_batch_handling.npz")
np.savez_compressed(path, **out)
print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))
