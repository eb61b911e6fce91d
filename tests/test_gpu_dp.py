"""NCCL data parallelism on real GPUs: N ranks (sharded global batch, global loss normaliser, one summed all-reduce of
the flat gradient buffer) == one GPU on the whole batch.  Needs >= 2 GPUs (skipped on a 1-GPU box; evidence of a
2-GPU run is kept under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision,tol", [("bf16x3", 2e-4), ("f16", 2e-2)])
def test_two_ranks_equal_one_gpu(precision, tol):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SAMK_PRECISION=precision)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == 2
    assert abs(out["loss_dp"] - out["loss_single"]) <= 1e-5 * abs(out["loss_single"]) + tol * 1e-2 * abs(out["loss_single"])
    # strict mode: only the summation order differs; product mode: the bf16 / scaled-half gradient operands are
    # rounded per shard instead of per batch
    assert out["grad_rel_err"] < tol, out
