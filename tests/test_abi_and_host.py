"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the host logic
(config, state_dict contract, head-bit tables, sector table, data-parallel plumbing) is right."""
import os
import re

import numpy as np
import pytest
import torch

from sam_textvqa_b200 import _lib, synth
from sam_textvqa_b200.config import BertConfig, c3_config
from tests._util import sam4c_state_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from sam_textvqa_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "samk.h")).read()
    declared = set(re.findall(r"\b(samk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.samk_version() >= 100


def test_compiled_sector_table_matches_this_hosts_numpy(built):
    from sam_textvqa_b200 import spatial_utils
    assert np.array_equal(spatial_utils.default_sector_table(), spatial_utils.derive_sector_table_from_numpy())


def test_sector_steps_are_the_only_steps_near_the_boundaries(built):
    """The kernel assumes ceil(angle/(pi/4)) is a 2-step function of sin/cos per quadrant; check
    NumPy's arcsin/arccos exhaustively +-300 ulps around each step position."""
    import math
    from sam_textvqa_b200 import spatial_utils
    tab = spatial_utils.default_sector_table()
    PI = math.pi
    fs = [lambda s: np.arcsin(s), lambda s: PI + np.arcsin(s), lambda s: np.arcsin(s) + 2 * PI,
          lambda s: (np.arcsin(s) + 2 * PI) - PI, lambda c: np.arccos(c), lambda c: np.arccos(c) + PI,
          lambda c: 2 * PI - np.arccos(c), lambda c: (2 * PI - np.arccos(c)) - PI]
    for k in range(8):
        base, step, t1, t2 = tab[k]
        for t in (t1, t2):
            x = np.array([t], dtype=np.float64)
            lo, hi = x.copy(), x.copy()
            xs = [x]
            for _ in range(300):
                lo = np.nextafter(lo, -np.inf); hi = np.nextafter(hi, np.inf)
                xs += [lo, hi]
            xs = np.concatenate(xs)
            dom = (xs >= (0.0 if k < 2 else -1.0)) & (xs <= (1.0 if k < 2 else -5e-324))
            xs = xs[dom]
            got = np.ceil(fs[k](xs) / (PI / 4)) + 3
            want = base + step * ((xs >= t1).astype(int) + (xs >= t2).astype(int))
            assert np.array_equal(got, want), k


def test_state_dict_contract():
    from sam_textvqa_b200.registry import registry
    from sam_textvqa_b200.sa_m4c import SAM4C
    registry.answer_vocab = ["w"] * 5000
    mmt, tb = c3_config()
    model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb))
    sd = model.state_dict()
    want = dict(sam4c_state_shapes(mmt, tb, 5000))
    assert len(sd) == 179 and set(sd) == set(want)
    assert all(tuple(sd[k].shape) == tuple(want[k]) for k in want)
    assert sum(v.numel() for v in sd.values()) == 96633224
    groups = model.get_optimizer_parameters(1e-4)
    assert len(groups) == 2 and "lr" not in groups[0] and groups[1]["lr"] == 1e-4
    assert sum(len(g["params"]) for g in groups) == len(list(model.parameters()))


def test_config_semantics_and_errors():
    cfg = BertConfig.from_dict({"hidden_size": 768, "foo": 3, "layer_type_list": ["s"]})
    assert cfg.foo == 3 and cfg.num_attention_heads == 12 and cfg.layer_norm_eps == 1e-12
    from sam_textvqa_b200.sa_m4c import SpatialBertSelfAttention
    mmt, _ = c3_config(attention_mask_quadrants=[3])
    with pytest.raises(ValueError):
        SpatialBertSelfAttention(BertConfig.from_dict(mmt))
    mmt, _ = c3_config(attention_mask_quadrants=[1, 2, 9])
    assert SpatialBertSelfAttention(BertConfig.from_dict(mmt)).quadrant_bits == 0b100000011


def test_cpu_inputs_fail_loudly():
    from sam_textvqa_b200 import ops
    with pytest.raises(_lib.SamkError):
        ops.linear(torch.zeros(8, 8), torch.zeros(8, 8), torch.zeros(8))


def test_head_bit_lut_matches_oracle_chain():
    from oracle import graph_oracle as G
    rs = np.random.RandomState(0)
    types = rs.randint(0, 13, (40, 40)).astype(np.int8)
    for c in (1, 3, 5, 7, 9):
        lut = synth.head_bits_for_context(c)
        assert np.array_equal(lut[types.astype(np.int64)], G.head_bits_closed_form(types, c))
        heads = synth.expand_types_to_heads(torch.from_numpy(types)[None], c)[0].numpy()
        assert np.array_equal(heads, ((lut[types.astype(np.int64)][..., None] >> np.arange(12)) & 1).astype(np.int8))


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from sam_textvqa_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    buf = dp.FlatGradBuffer(model.parameters())
    g = torch.Generator().manual_seed(1)
    batch = {"x": torch.randn(8, 16, generator=g), "y": torch.randn(8, 4, generator=g), "d": {"m": torch.ones(8, 2)}}
    shard = dp.shard_batch(batch, rank, world)
    assert shard["x"].shape[0] == 4 and shard["d"]["m"].shape[0] == 4
    buf.zero()
    ((model(shard["x"]) - shard["y"]) ** 2).mean().backward()
    buf.all_reduce(average=True)
    got = buf.flat.clone()
    model.zero_grad(set_to_none=True)
    buf.zero()                                    # re-attaches the views
    ((model(batch["x"]) - batch["y"]) ** 2).mean().backward()
    full = buf.flat.clone()
    # overlapped, bucketed exchange: step 1 learns how often each parameter is reported (weight of layer 2 twice,
    # like the classifier weight shared with PrevPredEmbeddings), later steps reduce ready runs early
    from sam_textvqa_b200 import ops
    buf.enable_overlap(average=True, bucket_bytes=256)
    params = list(model.parameters())
    errs = []
    for _ in range(3):
        buf.zero()
        buf.begin_step()
        ((model(shard["x"]) - shard["y"]) ** 2).mean().backward()
        for p in reversed(params):                 # what the backward ops of `ops` report as they finish
            ops.grad_ready_hook([p])
        ops.grad_ready_hook([params[2]])
        buf.finish_step()
        errs.append(float((buf.flat - full).abs().max()))
    ops.grad_ready_hook = None
    q.put((rank, max(float((got - full).abs().max()), max(errs)), dp.global_loss_scale(torch.tensor(float(rank + 1)))))
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err < 1e-6 for _, err, _ in res)
    assert abs(res[0][2] - 1.0 / 3.0) < 1e-6 and abs(res[1][2] - 2.0 / 3.0) < 1e-6


def test_relation_bits_contract_on_host():
    """Reference-format masks are preferred; without them the padded boxes are required (on-device batch
    preparation) -- the failure mode is a KeyError naming the missing inputs, not a silent default."""
    from sam_textvqa_b200 import sa_m4c
    with pytest.raises(KeyError):
        sa_m4c._relation_bits({"question_mask": torch.ones(1, 20)}, "3", torch.device("cpu"))
    with pytest.raises(KeyError):
        sa_m4c._relation_bits({"spatial_adj_matrices": {"1": torch.zeros(1, 4, 4, 12, dtype=torch.int8)}}, "3",
                              torch.device("cpu"))


def test_graph_step_module_imports_without_gpu():
    import sam_textvqa_b200.graph_step as gs
    assert hasattr(gs.GraphedTrainStep, "run") and hasattr(gs.GraphedTrainStep, "load")


def test_warmup_schedule_matches_reference_formula_and_flat_adam_needs_gpu():
    """sam/task_utils.py:46-55 restated in optim.warmup_lr_lambda; FlatAdam has no CPU path."""
    from bisect import bisect
    from sam_textvqa_b200 import optim
    from sam_textvqa_b200.dp import FlatGradBuffer
    warm, factor, miles, decay = 1000, 0.2, [14000, 19000], 0.1
    fn = optim.warmup_lr_lambda(warm, factor, miles, decay)
    for it in (0, 1, 500, 1000, 1001, 13999, 14000, 14001, 19000, 24000):
        if it <= warm:
            a = float(it) / float(warm)
            ref = factor * (1.0 - a) + a
        else:
            ref = pow(decay, bisect(miles, it))
        assert fn(it) == ref, it
    w = torch.nn.Parameter(torch.zeros(4, 4))
    b = torch.nn.Parameter(torch.zeros(4))
    groups = [{"params": [b]}, {"params": [w], "lr": 1e-5}]          # first group without lr, as get_optimizer_parameters
    grads = optim.flat_grad_buffer_for(groups)
    assert [p is q for p, q in zip(grads.params, [b, w])] == [True, True]
    with pytest.raises(ValueError):          # no default lr for the group that has none
        optim.FlatAdam(groups, grads)
    with pytest.raises(RuntimeError):
        optim.FlatAdam(groups, grads, lr=1e-4)
    with pytest.raises(ValueError):          # layout not grouped
        optim.FlatAdam(groups, FlatGradBuffer([w, b]), lr=1e-4)


def test_peer_exchange_buckets_are_whole_wire_vectors_and_cover_every_gradient_once():
    """Bucket logic of the overlapped exchange with the peer transport (dp.FlatGradBuffer._flush / _launch): every range
    handed to samk_exchange_sum starts and ends on a wire-vector boundary (8 bf16 elements), parameters whose sizes
    are not multiples of 8 wait for their neighbours, and after finish_step every element went out exactly once."""
    from sam_textvqa_b200 import dp, ops

    class Stub(object):
        vec = 8

        def __init__(self):
            self.ranges = []

        def exchange_sum(self, storage, lo, hi):
            self.ranges.append((lo, hi))

    sizes = [24, 5, 3, 64, 7, 9, 40, 13]                       # offsets 0 24 29 32 96 103 112 152 -> 165 elements
    params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    buf = dp.FlatGradBuffer(params)
    assert buf._storage.numel() % 64 == 0 and buf.flat.numel() == sum(sizes)
    stub = Stub()
    saved = ops.grad_ready_hook
    try:
        buf.enable_overlap(bucket_bytes=8 * 4, transport=stub)
        for step in range(2):                                  # step 0 learns the report counts
            buf.begin_step()
            for p in reversed(params):                         # backward order
                ops.grad_ready_hook([p])
            if step == 0:
                buf._ov["expected"] = [buf._ov["seen"].get(i, 0) for i in range(len(params))]
            else:
                before_final = list(stub.ranges)
                buf._flush(final=True)
    finally:
        ops.grad_ready_hook = saved
    total = buf.flat.numel()
    padded = (total + 7) // 8 * 8
    assert before_final, "nothing went out before the final flush"
    covered = [0] * padded
    for lo, hi in stub.ranges:
        assert lo % 8 == 0 and hi % 8 == 0 and lo < hi <= padded, (lo, hi)
        for i in range(lo, hi):
            covered[i] += 1
    assert covered == [1] * padded, stub.ranges


def test_flat_grad_buffer_views_padding_and_reattachment_on_host():
    """dp.FlatGradBuffer: every .grad is a view of one flat buffer whose storage is padded to whole wire vectors; zero()
    clears it and re-attaches views that a caller replaced (optimizer.zero_grad(set_to_none=True), train.py:144)."""
    from sam_textvqa_b200 import dp
    params = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 2))]
    buf = dp.FlatGradBuffer(params)
    assert buf.flat.numel() == 30 and buf._storage.numel() == 64 and buf.flat.data_ptr() == buf._storage.data_ptr()
    off = 0
    for p in params:
        assert p.grad.data_ptr() == buf.flat.data_ptr() + 4 * off and p.grad.shape == p.shape
        off += p.numel()
    sum((p * p).sum() for p in params).backward()
    assert torch.allclose(buf.flat[:21].view(7, 3), 2 * params[0].detach())
    params[1].grad = None
    buf.zero()
    assert float(buf._storage.abs().sum()) == 0.0
    assert params[1].grad is not None and params[1].grad.data_ptr() == buf.flat.data_ptr() + 4 * 21


def test_bench_numa_binding_is_a_no_op_without_topology_information():
    """bench.bind_to_gpu_numa_node must never raise (no GPU here, single-node VMs on the GPU pool): it returns None."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("samk_bench", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    before = os.sched_getaffinity(0)
    assert mod.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
