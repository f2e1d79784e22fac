"""CPU: the host side of the path — `ICP.batch_size_handling` (dICP/ICP.py:305-511: list / tensor / ragged /
empty / weighted inputs -> padded batch) and the constructor / YAML surface — against golden outputs of the
unmodified reference (tests/golden/make_golden_host.py).  No compute: nothing here needs a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from host_cases import host_cases  # noqa: E402

from dicp_b200.ICP import ICP  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "host_batch_handling.npz"))
CASES = host_cases()


@pytest.mark.parametrize("name,icp_type,make", CASES, ids=[c[0] for c in CASES])
def test_batch_size_handling_matches_reference(name, icp_type, make):
    source, target, T_init, weight = make()
    icp = ICP(icp_type=icp_type, max_iterations=3)
    if name + "/raises" in GOLD:
        with pytest.raises(Exception) as err:
            icp.batch_size_handling(source, target, T_init, weight)
        assert type(err.value).__name__ == str(GOLD[name + "/raises"])
        return
    res = icp.batch_size_handling(source, target, T_init, weight)
    assert len(res) == 4
    for key, val in zip(("source", "target", "T_init", "weight"), res):
        if name + "/" + key not in GOLD:
            assert val is None, key
            continue
        want = GOLD[name + "/" + key]
        assert val is not None, key
        assert tuple(val.shape) == want.shape, (key, tuple(val.shape), want.shape)
        assert str(val.dtype) == str(GOLD[name + "/" + key + "_dtype"]), key
        got = val.detach().cpu().numpy()
        assert np.array_equal(got, want, equal_nan=True), (key, np.abs(got - want).max())


def test_constructor_surface_and_yaml_defaults():
    """dICP/ICP.py:15-44: constructor arguments and the attributes callers mutate (tests/test_ICP_inputs.py:263)."""
    icp = ICP()
    assert icp.icp_type == "pt2pl" and icp.max_iterations == 100 and icp.tolerance == 1e-12 and icp.diff is True
    for attr in ("const_iter", "verbose", "target_pad_val", "source_zeroes_are_pad", "match_ratio_thresh", "config"):
        assert hasattr(icp, attr), attr
    assert icp.nn_search == "brute" and icp.check_every == 1          # B200-side additions keep reference behaviour by default
    icp2 = ICP(icp_type="pt2pt", max_iterations=7, tolerance=1e-6, differentiable=False)
    assert (icp2.icp_type, icp2.max_iterations, icp2.tolerance, icp2.diff) == ("pt2pt", 7, 1e-6, False)
    assert (icp.const_iter, icp.verbose, icp.target_pad_val, icp.source_zeroes_are_pad, icp.match_ratio_thresh) == (False, False, 1000, False, 0.0)
    assert ICP(icp_type="plane2plane").icp_type == "plane2plane"      # like the reference, the constructor does not validate the type


def test_input_validation_happens_before_any_device_work():
    """dICP/ICP.py:79, 96, 103: dim, dtype and column asserts fire on the host (there is no GPU in this suite)."""
    icp = ICP(icp_type="pt2pl", max_iterations=2)
    s, t, T0 = torch.rand(1, 5, 3), torch.rand(1, 6, 6), torch.eye(4)[None]
    with pytest.raises(AssertionError):
        icp.icp(s, t, T0, dim=4)
    with pytest.raises(AssertionError):
        icp.icp(s, t.double(), T0)
    with pytest.raises(AssertionError):
        icp.icp(s, t[:, :, :3], T0)                                   # pt2pl needs normals


def test_loss_class_matches_reference_known_answers():
    """dicp_b200.loss.loss (the public mirror of dICP/loss.py:3-58, plain tensor math on the caller's device) on the
    golden errors: Huber / Cauchy / trim, soft and hard, 3-D and 2-D inputs.  The reference's hard trim on a 2-D
    input returns an accidental (n,n) broadcast (loss.py:46-58); every row of it equals the (n,) answer."""
    from dicp_b200.loss import loss
    g = np.load(os.path.join(HERE, "golden", "loss_known_answers.npz"))
    for name, metric in (("huber", 1.3), ("cauchy", 0.7), ("trim", 1.9)):
        for diff in (True, False):
            L = loss(name=name, metric=metric, differentiable=diff, tanh_steepness=5.0)
            for key in ("e3", "e1", "e2d"):
                got = L.get_weight(torch.from_numpy(g[key])).numpy()
                want = g["%s_%d_%s" % (name, diff, key)]
                if want.ndim == 2 and got.ndim == 1 and want.shape == (got.shape[0], got.shape[0]):
                    assert np.allclose(want, np.broadcast_to(got, want.shape), rtol=1e-12, atol=1e-15), (name, diff, key)
                else:
                    assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=1e-15), (name, diff, key)
    with pytest.raises(ValueError):
        loss(name="tukey").get_weight(torch.zeros(3, 3))                # loss.py:19
