"""Multi-GPU tests on real hardware (skipped below two devices): the split-target search of one huge pair over NVLink
peers and the sharded learning-mode step of bench.py, launched with torchrun as the driver does."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, port=29541, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    out = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_split_targets_over_two_gpus_is_bit_identical_to_one_gpu(transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rec = _torchrun(2, "tools/run_c4_split.py", transport, port=29541 if transport == "peer" else 29542)
    assert rec["transport"] == transport and rec["n_gpus"] == 2
    assert rec["bit_identical_T"] and rec["bit_identical_grad_source"]
    assert rec["split_targets_ms"] < rec["single_gpu_brute_ms"]


def test_bench_two_gpus_learning_mode_line():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rec = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", port=29543)
    assert rec["n_gpus"] == 2 and rec["pairs_total"] == 512 and rec["scaling"] == "weak"
    assert rec["learning_mode_allreduce"]["bytes"] > 1_000_000 and rec["learning_mode_allreduce"]["ms"] > 0.0
    assert rec["c4_split"]["bit_identical_to_single_gpu"] and rec["cpu_baseline"] is None
