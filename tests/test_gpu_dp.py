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


def test_peer_exchange_equals_nccl_allreduce_on_two_gpus():
    """samk_exchange_sum (NVSwitch multicast and NVLink peer loads / stores, bf16 and fp32 on the wire) == ncclAllReduce:
    eager, back to back with changing data, on sub-ranges, and replayed from a CUDA graph beside a GEMM
    (tools/xchg_check.py asserts; a rank that misses a barrier raises instead of hanging)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, XCHG_N=str(8 * 1024 * 1024))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tools", "xchg_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == 2 and len(out["cases"]) == 4
    for c in out["cases"]:
        if c["multicast"] == "unavailable":
            continue
        assert c["max_rel_err"] <= c["tol"] and c["graph_rel_err"] <= c["tol"], c


def test_two_ranks_with_the_overlapped_peer_exchange_equal_one_gpu():
    """the bucketed exchange under the backward pass (dp.FlatGradBuffer.enable_overlap(transport="peer")) gives the
    gradients of the whole batch on one GPU (bf16 on the wire: one extra rounding per gradient)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SAMK_PRECISION="bf16x3", SAMK_DP_TRANSPORT="peer")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29651", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["transport"] == "peer" and out["buckets"] >= 2, out
    assert out["grad_rel_err"] < 6e-3, out                      # bf16 wire: 2^-9 per addend
