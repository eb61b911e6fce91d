"""GPU probe: tensor-core attention (bf16) vs the exact SIMT kernel on bf16-rounded inputs."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_textvqa_b200 import ops, synth
from sam_textvqa_b200.sa_m4c import pack_relation_bits
from tests._util import rel_err
dev = torch.device("cuda:0")

def case(B, T, A, D, spatial, p, seed=0):
    L = T + A + D
    g = torch.Generator().manual_seed(seed + L)
    qkv16 = (0.7 * torch.randn(B * L, 2304, generator=g)).to(dev).bfloat16()
    qkv32 = qkv16.float()
    valid = (torch.rand(B, L, generator=g) < 0.85).to(torch.uint8).to(dev)
    valid[:, 0] = 1
    if D: valid[:, -D:] = 0
    bits = None
    if spatial:
        rs = np.random.RandomState(seed + L)
        types = torch.from_numpy(rs.randint(0, 13, (B, A, A)).astype(np.int8))
        types[0, 3] = 0
        bits = pack_relation_bits(synth.expand_types_to_heads(types, 3), dev)
    dims = (B, L, 12, T, A, D)
    quad = 0b11 if spatial else 0
    drop = (77, 5)
    ctx_ref, lse_ref = ops.attention_fwd(qkv32, valid, bits, dims, spatial, quad, p, drop)
    ctx, lse = ops.attention_fwd(qkv16, valid, bits, dims, spatial, quad, p, drop)
    torch.cuda.synchronize()
    fin = torch.isfinite(lse_ref)
    e_ctx = rel_err(ctx.float(), ctx_ref)
    e_lse = (lse[fin] - lse_ref[fin]).abs().max().item() if fin.any() else 0.0
    dead_ok = bool((torch.isinf(lse) == torch.isinf(lse_ref)).all())
    w16 = torch.randn(B * L, 768, generator=g).to(dev).bfloat16()
    dq_ref = ops.attention_bwd(w16.float(), qkv32, ctx_ref, lse_ref, valid, bits, dims, spatial, quad, p, drop)
    dq = ops.attention_bwd(w16, qkv16, ctx, lse, valid, bits, dims, spatial, quad, p, drop)
    torch.cuda.synchronize()
    r = dq_ref.view(-1, 3, 768); o = dq.float().view(-1, 3, 768)
    print("B=%d T=%d A=%d D=%d spatial=%d p=%.1f | ctx %.2e lse %.2e dead_ok %s | dq %.2e dk %.2e dv %.2e" % (
        B, T, A, D, spatial, p, e_ctx, e_lse, dead_ok, rel_err(o[:, 0], r[:, 0]), rel_err(o[:, 1], r[:, 1]), rel_err(o[:, 2], r[:, 2])), flush=True)

cases = [(2, 20, 150, 12, True, 0.0), (2, 20, 150, 12, False, 0.0), (3, 20, 0, 0, False, 0.0), (2, 20, 86, 12, True, 0.0),
         (2, 20, 150, 12, True, 0.1), (2, 20, 200, 12, True, 0.0), (1, 20, 442, 12, True, 0.0), (1, 20, 442, 12, False, 0.1),
         (1, 20, 1004, 12, True, 0.0)]
for c in cases:
    try:
        case(*c)
    except Exception:
        traceback.print_exc(); sys.stdout.flush()

# timing at bench shape
B, T, A, D = 128, 20, 150, 12
L = T + A + D
qkv = torch.randn(B * L, 2304, device=dev).bfloat16()
valid = torch.ones(B, L, dtype=torch.uint8, device=dev); valid[:, -D:] = 0
types = torch.from_numpy(np.random.RandomState(0).randint(0, 13, (B, A, A)).astype(np.int8))
bits = pack_relation_bits(synth.expand_types_to_heads(types, 3), dev)
dims = (B, L, 12, T, A, D)
allow = ops.build_attn_mask(valid, bits, dims, True, 0b11)
w = torch.randn(B * L, 768, device=dev).bfloat16()
for p in (0.0, 0.1):
    for _ in range(3):
        ctx, lse = ops.attention_fwd(qkv, valid, bits, dims, True, 0b11, p, (1, 1), allow)
        dq = ops.attention_bwd(w, qkv, ctx, lse, valid, bits, dims, True, 0b11, p, (1, 1), allow)
    s, e, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ctx, lse = ops.attention_fwd(qkv, valid, bits, dims, True, 0b11, p, (1, 1), allow)
    e.record()
    for _ in range(10):
        dq = ops.attention_bwd(w, qkv, ctx, lse, valid, bits, dims, True, 0b11, p, (1, 1), allow)
    e2.record(); torch.cuda.synchronize()
    print("p=%.1f  fwd %.1f us  bwd %.1f us (B=128 L=182 H=12)" % (p, s.elapsed_time(e) * 100, e.elapsed_time(e2) * 100), flush=True)
