"""GPU parity of the individual kernels (through the C ABI) against torch fp32 / the SIMT cross-check."""
import ctypes

import numpy as np
import pytest
import torch

from sam_textvqa_b200 import synth
from tests._util import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from sam_textvqa_b200 import ops as _ops
    return _ops


_FMT = {"bf16": (torch.bfloat16, 1), "f16": (torch.float16, 2)}


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("fmt", [("f16", "f16"), ("bf16", "bf16")])
@pytest.mark.parametrize("shape", [(128, 256, 64), (200, 136, 72), (1000, 768, 768), (768, 3072, 1184)])
def test_tcgen05_gemm_all_operand_majors_and_formats(ops, a_mn, b_mn, fmt, shape):
    """Every operand major in both 16-bit formats (forward products are half x half, dgrad / wgrad bf16 x bf16 or
    scaled-half x half)."""
    from sam_textvqa_b200._lib import GemmEpilogue, check, lib, ptr, stream_ptr
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    (ta, ca), (tb, cb) = _FMT[fmt[0]], _FMT[fmt[1]]
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV).to(ta)
    B = (torch.randn(N, K, generator=g) * 0.5).to(DEV).to(tb)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    outs = []
    for impl in (0, 1):
        out = torch.zeros(M, N, device=DEV)
        ep = GemmEpilogue()
        ep.out, ep.ldo, ep.out_dtype, ep.alpha = out.data_ptr(), N, 0, 1.0
        check(lib().samk_gemm_16(ptr(As), ca, a_mn, As.stride(0), ptr(Bs), cb, b_mn, Bs.stride(0), M, N, K,
                                 ctypes.byref(ep), 1, impl, stream_ptr()))
        outs.append(out)
    ref = A.float() @ B.float().t()
    assert rel_err(outs[0], ref) < 1e-5          # fp32 accumulation of exact 16-bit products
    assert rel_err(outs[0], outs[1]) < 1e-5      # tensor-core kernel == SIMT cross-check


def test_gemm_rejects_mixed_operand_formats(ops):
    """tcgen05.mma kind::f16 with a_format != b_format faults on sm_100a (measured): the C ABI refuses it."""
    from sam_textvqa_b200._lib import GemmEpilogue, lib, ptr, stream_ptr
    A = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    B = torch.zeros(128, 64, device=DEV, dtype=torch.float16)
    out = torch.zeros(128, 128, device=DEV)
    ep = GemmEpilogue()
    ep.out, ep.ldo, ep.out_dtype, ep.alpha = out.data_ptr(), 128, 0, 1.0
    rc = lib().samk_gemm_16(ptr(A), 1, 0, 64, ptr(B), 2, 0, 64, 128, 128, 64, ctypes.byref(ep), 1, 0, stream_ptr())
    assert rc == -3 and b"one 16-bit format" in lib().samk_last_error()


def test_scaled_half_gradient_operand_reproduces_the_weight_gradient(ops):
    """wgrad where the saved activation exists in half only: dY is re-expressed as half(dY * S), S an exact power of
    two found on the device, and the GEMM multiplies by 1/S (alpha_dev) -- tiny gradients keep their precision."""
    M, N, K = 2048, 768, 3072
    g = torch.Generator().manual_seed(3)
    for mag in (1.0, 3e-6, 4e4):
        dy = (mag * torch.randn(M, N, generator=g)).to(DEV).bfloat16()
        dy[5, 7] = 37.0 * mag                                   # an outlier sets the scale
        x = torch.randn(M, K, generator=g).to(DEV).half()
        dyh, inv = ops.scaled_f16(dy)
        S = 1.0 / inv.item()
        assert 2048.0 <= dy.float().abs().max().item() * S < 4096.0 and S == 2.0 ** round(np.log2(S))
        gw = torch.zeros(N, K, device=DEV)
        ops.gemm(dyh, True, ops.operand(x, "b", True), True, N, K, M, gw, accumulate=True, alpha_dev=inv)
        ref = dy.float().t() @ x.float()
        assert rel_err(gw, ref) < 1e-3, mag      # half has 3 more bits than the bf16 source: only subnormal tails round


def test_gemm_16bit_outputs_in_both_formats(ops):
    """Forward epilogues write half, backward epilogues bf16 (specialised masks), anything else the dynamic epilogue."""
    M, N, K = 512, 768, 768
    x = torch.randn(M, K, device=DEV)
    w = 0.05 * torch.randn(N, K, device=DEV)
    b = torch.randn(N, device=DEV)
    xo, wo = ops.operand(x, "a", False), ops.operand(w, "b", False)
    assert xo.t.dtype == torch.float16 and wo.t.dtype == torch.float16
    base = xo.t.float() @ wo.t.float().t()
    for dt, tol in ((torch.float16, 6e-4), (torch.bfloat16, 5e-3)):
        out = torch.empty(M, N, dtype=dt, device=DEV)
        ops.gemm(xo, False, wo, False, M, N, K, out, bias=b)                  # M_QKV (half) / dynamic (bf16)
        assert rel_err(out.float(), base + b) < tol
        pre = torch.empty(M, N, dtype=dt, device=DEV)
        ops.gemm(xo, False, wo, False, M, N, K, out, bias=b, act=3, pre=pre)  # gelu pair
        hh = (base + b).clone().requires_grad_(True)
        gg = torch.nn.functional.gelu(hh)
        gg.sum().backward()
        assert rel_err(out.float(), gg.detach()) < tol and rel_err(pre.float(), hh.grad) < tol
        dy = ops.operand(torch.randn(M, N, device=DEV), "a", False, fmt="bf16")
        wb = ops.operand(w, "b", True, fmt="bf16")
        assert dy.t.dtype == torch.bfloat16 and wb.t.dtype == torch.bfloat16
        dx = torch.empty(M, K, dtype=torch.bfloat16, device=DEV)
        ops.gemm(dy, False, wb, True, M, K, N, dx, act=4, aux=pre)            # M_MULAUX with half / bf16 aux
        assert rel_err(dx.float(), (dy.t.float() @ wb.t.float()) * pre.float()) < 5e-3


def test_gemm_fused_epilogues_and_split_k(ops):
    M, N, K = 1024, 768, 1024
    x = torch.randn(M, K, device=DEV)
    w = 0.05 * torch.randn(N, K, device=DEV)
    b = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    xo, wo = ops.operand(x, "a", False), ops.operand(w, "b", False)
    xr, wr = xo.t.float()[:, :K], wo.t.float()[:, :K]
    base = xr @ wr.t() + b
    pre = torch.empty(M, N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(xo, False, wo, False, M, N, K, out, bias=b, act=1, pre=pre)
    assert rel_err(pre, base) < 1e-5 and rel_err(out, torch.nn.functional.gelu(base)) < 1e-5
    ops.gemm(xo, False, wo, False, M, N, K, out, bias=b, residual=res)
    assert rel_err(out, base + res) < 1e-5
    h = torch.randn(M, N, device=DEV)
    ops.gemm(xo, False, wo, False, M, N, K, out, act=2, aux=h)
    hh = h.clone().requires_grad_(True)
    torch.nn.functional.gelu(hh).sum().backward()
    assert rel_err(out, (xr @ wr.t()) * hh.grad) < 1e-5
    acc = torch.ones(M, N, device=DEV)
    ops.gemm(xo, False, wo, False, M, N, K, acc, accumulate=True, split_k=4)
    assert rel_err(acc, xr @ wr.t() + 1.0) < 1e-5
    # dropout epilogue: keep-rate and scaling, identical mask for identical (seed, offset)
    o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(xo, False, wo, False, M, N, K, o1, bias=b, drop_p=0.1, drop=(123, 7))
    ops.gemm(xo, False, wo, False, M, N, K, o2, bias=b, drop_p=0.1, drop=(123, 7))
    assert torch.equal(o1, o2)
    kept = o1 != 0
    assert abs(kept.float().mean().item() - 0.9) < 5e-3
    assert rel_err(o1[kept], (base / 0.9)[kept]) < 1e-5


def test_bf16x3_split_reaches_fp32_accuracy(ops):
    ops.set_precision("bf16x3")
    try:
        x = torch.randn(512, 768, device=DEV, requires_grad=True)
        w = (0.05 * torch.randn(768, 768, device=DEV)).requires_grad_(True)
        b = torch.randn(768, device=DEV, requires_grad=True)
        y = ops.linear(x, w, b)
        ref = (x.double() @ w.double().t() + b.double())
        assert rel_err(y, ref) < 2e-5
        gy = torch.randn_like(y)
        gx, gw, gb = torch.autograd.grad((y * gy).sum(), (x, w, b))
        rx, rw, rb = torch.autograd.grad((ref * gy.double()).sum(), (x, w, b))
        assert rel_err(gx, rx) < 2e-5 and rel_err(gw, rw) < 2e-5 and rel_err(gb, rb) < 1e-5
    finally:
        ops.set_precision("f16")


def test_layernorm_forward_backward(ops):
    x = torch.randn(1111, 768, device=DEV, requires_grad=True)
    g = (1 + 0.1 * torch.randn(768, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(768, device=DEV)).requires_grad_(True)
    y = ops.layer_norm(x, g, b, 1e-12)
    ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-12)
    assert rel_err(y, ref) < 1e-5
    w = torch.randn_like(y)
    got = torch.autograd.grad((y * w).sum(), (x, g, b))
    want = torch.autograd.grad((ref * w).sum(), (x, g, b))
    for a, r in zip(got, want):
        assert rel_err(a, r) < 1e-5


def _attn_reference(qkv, valid, adj, T, A, D, quadrants):
    """Explicit boolean-mask softmax in torch (fp64)."""
    B, L = valid.shape
    H = 12
    q3 = qkv.view(B, L, 3, H, 64).double()
    qq, kk, vv = (q3[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ext = valid.double()[:, None, :].repeat(1, L, 1)
    if D:
        ext[:, -D:, -D:] = torch.tril(torch.ones(D, D, device=qkv.device, dtype=torch.double))
    allow = (ext[:, None] > 0).expand(B, H, L, L).clone()
    if adj is not None:
        m = torch.ones(B, L, L, H, device=qkv.device)
        m[:, T:T + A, T:T + A] = adj.float()
        if 1 in quadrants:
            m[:, :T, :T] = 0
        if 2 in quadrants:
            m[:, :T, T:T + A] = 0
        allow &= m.permute(0, 3, 1, 2) > 0
    s = (qq @ kk.transpose(-1, -2)) / 8.0
    p = torch.softmax(s.masked_fill(~allow, float("-inf")), -1)
    p = torch.where(allow.any(-1, keepdim=True), p, torch.zeros_like(p))
    return (p @ vv).permute(0, 2, 1, 3).reshape(B, L, H * 64)


def test_attention_matches_reference_module_golden(ops):
    """ctx of the unmodified reference SpatialBertSelfAttention (tests/golden/attn_unit.npz)."""
    from sam_textvqa_b200.sa_m4c import pack_relation_bits
    g = load_golden("attn_unit.npz")
    T, A, D = int(g["T"]), int(g["A"]), int(g["D"])
    hidden = torch.from_numpy(g["hidden"]).to(DEV)
    B, L, d = hidden.shape
    sd = synth.seeded_state([("%s.%s" % (n, s), (768, 768) if s == "weight" else (768,))
                             for n in ("query", "key", "value") for s in ("weight", "bias")], 3)
    W = torch.cat([sd["query.weight"], sd["key.weight"], sd["value.weight"]]).to(DEV).double()
    bias = torch.cat([sd["query.bias"], sd["key.bias"], sd["value.bias"]]).to(DEV).double()
    qkv = (hidden.view(B * L, d).double() @ W.t() + bias).float().contiguous()
    bits = pack_relation_bits(torch.from_numpy(g["adj"]), torch.device(DEV))
    valid = torch.from_numpy(g["valid"]).to(DEV).to(torch.uint8).contiguous()
    dims = (B, L, 12, T, A, D)
    ctx, _ = ops.attention_fwd(qkv, valid, bits, dims, True, 0b11, 0.0, (0, 0))
    want = torch.from_numpy(g["ctx"]).to(DEV)
    assert rel_err(ctx.view(B, L, d), want) < 2e-5
    assert ctx.view(B, L, d)[:, :T].abs().max().item() == 0.0        # dead text rows are exactly zero
    ctx16, _ = ops.attention_fwd(qkv.half(), valid, bits, dims, True, 0b11, 0.0, (0, 0))      # tensor-core kernel
    assert ctx16.dtype == torch.float16
    assert rel_err(ctx16.float().view(B, L, d), want) < 2e-3
    assert ctx16.view(B, L, d)[:, :T].abs().max().item() == 0.0


@pytest.mark.parametrize("spatial", [True, False])
@pytest.mark.parametrize("L_cfg", [(20, 30, 12), (20, 150, 12), (5, 0, 0), (20, 300, 12)])
def test_attention_forward_backward_vs_torch(ops, spatial, L_cfg):
    T, A, D = L_cfg
    if spatial and A == 0:
        pytest.skip("spatial layers need entities")
    B, L = 2, T + A + D
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(B * L, 3 * 768, generator=g).to(DEV)
    valid = (torch.rand(B, L, generator=g) < 0.8).to(torch.uint8).to(DEV)
    valid[:, 0] = 1
    if D:
        valid[:, -D:] = 0
    rs = np.random.RandomState(L)
    types = torch.from_numpy(rs.randint(0, 13, (B, max(A, 1), max(A, 1))).astype(np.int8))
    if A:
        types[0, 3] = 0
    adj = synth.expand_types_to_heads(types, 3).to(DEV) if A else None
    from sam_textvqa_b200.sa_m4c import pack_relation_bits
    bits = pack_relation_bits(adj, torch.device(DEV)) if (spatial and A) else None
    dims = (B, L, 12, T, A, D)
    ctx, lse = ops.attention_fwd(qkv, valid, bits, dims, spatial, 0b11 if spatial else 0, 0.0, (0, 0))
    q = qkv.clone().requires_grad_(True)
    ref = _attn_reference(q, valid, adj if spatial else None, T, A, D, (1, 2) if spatial else ())
    assert rel_err(ctx.view(B, L, 768), ref) < 2e-5
    w = torch.randn(B, L, 768, generator=g).to(DEV)
    (gref,) = torch.autograd.grad((ref * w.double()).sum(), q)
    dqkv = ops.attention_bwd(w.view(B * L, 768).contiguous(), qkv, ctx, lse, valid, bits, dims, spatial,
                             0b11 if spatial else 0, 0.0, (0, 0))
    assert rel_err(dqkv, gref) < 5e-5


def test_attention_dropout_is_consistent_between_forward_and_backward(ops):
    """With dropout, d(ctx . w)/d(qkv) from the kernel must match finite differences of the kernel's own
    forward under the same (seed, offset) mask."""
    T, A, D = 8, 16, 4
    B, L = 1, T + A + D
    g = torch.Generator().manual_seed(5)
    qkv = (0.5 * torch.randn(B * L, 3 * 768, generator=g)).to(DEV)
    valid = torch.ones(B, L, dtype=torch.uint8, device=DEV)
    valid[:, -D:] = 0
    dims = (B, L, 12, T, A, D)
    drop = (99, 3)
    w = torch.randn(B * L, 768, generator=g).to(DEV)
    ctx, lse = ops.attention_fwd(qkv, valid, None, dims, False, 0, 0.3, drop)
    dqkv = ops.attention_bwd(w, qkv, ctx, lse, valid, None, dims, False, 0, 0.3, drop)
    direction = torch.randn(B * L, 3 * 768, generator=g).to(DEV)
    eps = 1e-2
    f = lambda t: (ops.attention_fwd(t, valid, None, dims, False, 0, 0.3, drop)[0].double() * w.double()).sum()
    fd = (f(qkv + eps * direction) - f(qkv - eps * direction)) / (2 * eps)
    an = (dqkv.double() * direction.double()).sum()
    assert abs(fd.item() - an.item()) < 2e-2 * max(1.0, abs(an.item()))
    assert (ctx == 0).float().mean().item() < 0.05       # rows still populated


def test_embeddings_prevpred_pointer_and_loss_vs_torch(ops):
    B, T, D, R, V, d = 3, 20, 12, 50, 200, 768
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: (0.1 * torch.randn(*s, generator=g)).to(DEV).requires_grad_(True)
    # TextBert embeddings
    ids = torch.randint(0, 1000, (B, T), generator=g).to(DEV)
    ids[0, -3:] = 0
    word, pos, typ, ga, be = mk(1000, d), mk(512, d), mk(2, d), mk(d), mk(d)
    y = ops.BertEmbedFn.apply(ids, word, pos, typ, ga, be, 1e-12, 0.0)
    ref = torch.nn.functional.layer_norm(word[ids] + pos[:T][None] + typ[0], (d,), ga, be, 1e-12)
    assert rel_err(y, ref) < 1e-5
    w = torch.randn_like(ref)
    got = torch.autograd.grad((y * w).sum(), (word, pos, typ, ga, be))
    # padding_idx=0 rows get no gradient in nn.Embedding
    ref2 = torch.nn.functional.layer_norm(torch.nn.functional.embedding(ids, word, padding_idx=0) + pos[:T][None] + typ[0], (d,), ga, be, 1e-12)
    want = torch.autograd.grad((ref2 * w).sum(), (word, pos, typ, ga, be))
    for a, r in zip(got, want):
        assert rel_err(a, r) < 2e-5
    # PrevPredEmbeddings
    prev = torch.randint(0, V + R, (B, D), generator=g).to(DEV)
    cls_w, ocr_in, ppos, ptyp = mk(V, d), mk(B, R, d), mk(100, d), mk(5, d)
    lns = [mk(d) for _ in range(6)]
    out = ops.PrevPredFn.apply(prev, cls_w, ocr_in, ppos, ptyp, *lns, 1e-12, 0.0)
    ln = lambda x, gg, bb: torch.nn.functional.layer_norm(x, (d,), gg, bb, 1e-12)
    table = torch.cat([ln(cls_w, lns[0], lns[1])[None].expand(B, -1, -1), ln(ocr_in, lns[2], lns[3])], 1)
    raw = torch.gather(table, 1, prev[..., None].expand(-1, -1, d))
    refp = raw + ln(ppos[:D][None] + ptyp[(prev >= V).long()], lns[4], lns[5])
    assert rel_err(out, refp) < 1e-5
    w = torch.randn_like(refp)
    allp = [cls_w, ocr_in, ppos, ptyp] + lns
    got = torch.autograd.grad((out * w).sum(), allp)
    want = torch.autograd.grad((refp * w).sum(), allp)
    for a, r in zip(got, want):
        assert rel_err(a, r) < 2e-5
    # masked BCE loss (sam/task_utils.py:19-30)
    scores = (3 * torch.randn(B, D, V + R, generator=g)).to(DEV).requires_grad_(True)
    targets = (torch.rand(B, D, V + R, generator=g) < 0.01).float().to(DEV)
    mask = (torch.rand(B, D, generator=g) < 0.7).float().to(DEV)
    loss = ops.bce_with_mask_loss(scores, targets, mask)
    refl = (torch.nn.functional.binary_cross_entropy_with_logits(scores, targets, reduction="none") * mask[..., None]).sum() / mask.sum().clamp(min=1)
    assert abs(loss.item() - refl.item()) < 1e-4 * abs(refl.item())
    (gs,) = torch.autograd.grad(loss * 2.0, scores)
    (rs_,) = torch.autograd.grad(refl * 2.0, scores)
    assert rel_err(gs, rs_) < 1e-5
    zero_mask_loss = ops.bce_with_mask_loss(scores, targets, torch.zeros_like(mask))
    assert zero_mask_loss.item() == 0.0


@pytest.mark.parametrize("case", [(2, 20, 150, 12, True, 0.0), (2, 20, 150, 12, False, 0.0), (3, 20, 0, 0, False, 0.0),
                                  (2, 20, 86, 12, True, 0.1), (2, 20, 200, 12, True, 0.0), (1, 20, 442, 12, False, 0.1),
                                  (1, 20, 1004, 12, True, 0.0),
                                  # more (sample, head) items than SMs: every persistent CTA walks several items
                                  (40, 20, 150, 12, True, 0.1), (64, 20, 0, 0, False, 0.1), (30, 20, 86, 12, False, 0.0),
                                  (28, 20, 200, 12, True, 0.1), (2, 20, 150, 12, True, 0.1), (40, 20, 150, 12, True, 0.0),
                                  (40, 20, 150, 12, False, 0.0), (13, 20, 150, 12, False, 0.0)])
def test_tcgen05_attention_matches_exact_fp32_kernel(ops, case):
    """Tensor-core attention (half q|k|v / P / ctx, bf16 gradients; fwd + bwd, masks, dead rows, dropout from the
    precomputed keep bits, multi-tile online softmax up to L=1036) against the exact-fp32 SIMT kernel on the same
    half-representable inputs and the same Philox stream (drawn inline there)."""
    from sam_textvqa_b200.sa_m4c import pack_relation_bits
    B, T, A, D, spatial, p = case
    L = T + A + D
    g = torch.Generator().manual_seed(L)
    qkv16 = (0.7 * torch.randn(B * L, 2304, generator=g)).to(DEV).half()
    qkv32 = qkv16.float()
    valid = (torch.rand(B, L, generator=g) < 0.85).to(torch.uint8).to(DEV)
    valid[:, 0] = 1
    if D:
        valid[:, -D:] = 0
    bits = None
    if spatial:
        types = torch.from_numpy(np.random.RandomState(L).randint(0, 13, (B, A, A)).astype(np.int8))
        types[0, 3] = 0
        bits = pack_relation_bits(synth.expand_types_to_heads(types, 3), torch.device(DEV))
    dims, quad, drop = (B, L, 12, T, A, D), (0b11 if spatial else 0), (77, 5)
    ctx_ref, lse_ref = ops.attention_fwd(qkv32, valid, bits, dims, spatial, quad, p, drop)
    ctx, lse = ops.attention_fwd(qkv16, valid, bits, dims, spatial, quad, p, drop)
    assert rel_err(ctx.float(), ctx_ref) < 1.5e-3                  # half rounding of P and of the output
    assert torch.equal(torch.isinf(lse), torch.isinf(lse_ref))     # dead rows agree
    fin = torch.isfinite(lse_ref)
    assert (lse[fin] - lse_ref[fin]).abs().max().item() < 1e-4
    if spatial:
        assert ctx.view(B, L, 768)[:, :T].abs().max().item() == 0.0
    w16 = torch.randn(B * L, 768, generator=g).to(DEV).bfloat16()
    dq_ref = ops.attention_bwd(w16.float(), qkv32, ctx_ref, lse_ref, valid, bits, dims, spatial, quad, p, drop)
    dq = ops.attention_bwd(w16, qkv16, ctx, lse, valid, bits, dims, spatial, quad, p, drop)
    assert rel_err(dq.float(), dq_ref) < 1e-2


def test_fused_clip_adam_matches_torch_adam_and_clip_grad_norm(ops):
    """optim.FlatAdam (samk_sumsq + samk_adam_step) against the reference's optimizer side (train.py:139-143):
    nn.utils.clip_grad_norm_ + torch.optim.Adam with two learning-rate groups, four updates."""
    from sam_textvqa_b200 import optim
    g = torch.Generator().manual_seed(11)
    shapes = [(300, 768), (768,), (5, 7, 3), (1,), (2304, 768)]
    ours = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    split = 2
    groups = [{"params": ours[:split], "lr": 1e-3}, {"params": ours[split:], "lr": 1e-4}]
    grads = optim.flat_grad_buffer_for(groups)
    opt = optim.FlatAdam(groups, grads, max_grad_norm=0.25)
    ref_opt = torch.optim.Adam([{"params": ref[:split], "lr": 1e-3}, {"params": ref[split:], "lr": 1e-4}], lr=1e-3)
    for step in range(4):
        scale = 10.0 if step % 2 == 0 else 1e-3          # clipped and unclipped updates
        for p, q in zip(ours, ref):
            gr = (scale * torch.randn(*p.shape, generator=g)).to(DEV)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref, 0.25)
        assert abs(opt.grad_norm().item() - norm_ref.item()) <= 1e-5 * norm_ref.item()
        opt.step()
        ref_opt.step()
        for p, q in zip(ours, ref):
            assert (p.detach() - q.detach()).abs().max().item() <= 2e-6 * max(1.0, q.detach().abs().max().item()), step
    assert ours[0].data.data_ptr() == opt.flat_params.data_ptr()       # parameters are views of the flat buffer
    assert ours[0]._version >= 4                                        # weight-operand caches see the updates
    assert opt.step_count == 4


def test_flat_adam_state_dict_round_trip_and_detached_grads(ops):
    """train.py:144 / :179-181: `model.zero_grad()` (set_to_none) between steps and an optimizer state_dict that
    torch.optim.Adam can load."""
    from sam_textvqa_b200 import optim
    g = torch.Generator().manual_seed(2)
    ps = [torch.nn.Parameter(torch.randn(64, 32, generator=g).to(DEV)), torch.nn.Parameter(torch.randn(32, generator=g).to(DEV))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    groups = [{"params": ps}]
    grads = optim.flat_grad_buffer_for(groups)
    opt = optim.FlatAdam(groups, grads, lr=1e-2)
    ropt = torch.optim.Adam(ref, lr=1e-2)
    for it in range(3):
        for p in ps:
            p.grad = None                                  # what model.zero_grad() does in torch 2.x
        loss = sum((p * p).sum() for p in ps)
        loss.backward()                                    # fresh .grad tensors, not views of the flat buffer
        sum((p * p).sum() for p in ref).backward()
        opt.step()
        ropt.step()
        ropt.zero_grad()
        for p, q in zip(ps, ref):
            assert (p.detach() - q.detach()).abs().max().item() < 1e-6, it
    sd = opt.state_dict()
    ropt2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1e-2)
    ropt2.load_state_dict({"state": sd["state"], "param_groups": sd["param_groups"]})      # Adam accepts it
    opt.load_state_dict(sd)
    assert opt.step_count == 3


@pytest.mark.parametrize("B,D,R,dq", [(5, 12, 50, 768), (3, 1, 7, 64), (2, 20, 33, 1024)])
def test_pointer_scores_forward_and_backward_match_torch(ops, B, D, R, dq):
    """OcrPtrNet scoring (sa_m4c.py:878-897): scores[b,t,V+r] = q.k / sqrt(dq) + (1 - mask) * -1e4 written into the
    pointer columns of the shared [B*D, V+R] buffer, and the gradients of q and k (block-per-sample kernels)."""
    from sam_textvqa_b200._lib import check, lib, ptr, stream_ptr
    V = 24
    g = torch.Generator().manual_seed(B * 131 + D)
    q = torch.randn(B, D, dq, generator=g).to(DEV)
    k = torch.randn(B, R, dq, generator=g).to(DEV)
    mask = (torch.rand(B, R, generator=g) > 0.3).long().to(DEV)
    out = torch.full((B * D, V + R), 7.0, device=DEV)
    check(lib().samk_ptr_scores_fwd(ptr(q), ptr(k), ptr(mask), ptr(out), V + R, V, B, D, R, dq, stream_ptr()), "ptr fwd")
    ref = torch.einsum("btc,brc->btr", q.double(), k.double()) / dq ** 0.5 + (1.0 - mask.double())[:, None, :] * -10000.0
    got = out.view(B, D, V + R)
    assert torch.all(got[..., :V] == 7.0)                               # the vocabulary columns are not touched
    assert (got[..., V:].double() - ref).abs().max().item() < 2e-3 * max(1.0, ref[ref > -5000].abs().max().item())
    ds = torch.randn(B * D, V + R, generator=g).to(DEV)
    dq_, dk_ = torch.empty_like(q), torch.empty_like(k)
    check(lib().samk_ptr_scores_bwd(ptr(ds), V + R, V, ptr(q), ptr(k), ptr(dq_), ptr(dk_), B, D, R, dq, stream_ptr()), "ptr bwd")
    dsr = ds.view(B, D, V + R)[..., V:].double() / dq ** 0.5
    assert rel_err(dq_.cpu(), torch.einsum("btr,brc->btc", dsr, k.double()).float().cpu()) < 1e-5
    assert rel_err(dk_.cpu(), torch.einsum("btr,btc->brc", dsr, q.double()).float().cpu()) < 1e-5


@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_output_rows_only_8_byte_aligned(ops, accumulate):
    """fp32 output with an even pitch that is not a multiple of 4 (the classifier columns of the [rows, 5050] score buffer,
    the OCR projection's [768, 3002] weight gradient): the vector epilogue with 8-byte stores / red.v2 must equal torch."""
    g = torch.Generator().manual_seed(11)
    M, N, K, pitch = 512, 264, 192, 270
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV).half()
    B = (torch.randn(N, K, generator=g) * 0.5).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    buf = torch.full((M, pitch), 3.0, device=DEV)
    O = lambda t: ops.Operand(t, t.stride(0), 1)
    ref = A.float() @ B.float().t()
    if accumulate:
        At, Bt = A.t().contiguous(), B.t().contiguous()              # stored [K, M] / [K, N]: both MN-major (the wgrad form)
        ops.gemm(O(At), True, O(Bt), True, M, N, K, buf[:, :N], accumulate=True)
        want = ref + 3.0
    else:
        ops.gemm(O(A), False, O(B), False, M, N, K, buf[:, :N], bias=bias)
        want = ref + bias
    assert torch.all(buf[:, N:] == 3.0)
    assert rel_err(buf[:, :N].cpu(), want.cpu()) < 2e-3
