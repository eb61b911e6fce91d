#!/usr/bin/env python
"""Developer tool: per-item clock stamps of CTA 0 of the forward attention kernel (needs a build with
SAMK_NVCC_EXTRA=-DSAMK_TIMELINE).  Prints, per item, cycles relative to the first stamp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_textvqa_b200 import ops, synth, _lib
from sam_textvqa_b200.sa_m4c import pack_relation_bits
dev = torch.device("cuda:0")
B, T, O, R, D = 128, 20, 100, 50, 12
A, L, H, d = O + R, T + O + R + D, 12, 768
qkv = torch.randn(B * L, 3 * d, device=dev).half()
valid = torch.ones(B, L, dtype=torch.uint8, device=dev); valid[:, -D:] = 0
dims = (B, L, H, T, A, D)
allow = ops.build_attn_mask(valid, None, dims, False, 0)
p = float(os.environ.get("P", "0.1"))
bwd = os.environ.get("BWD", "0") == "1"
w = torch.randn(B * L, d, device=dev).bfloat16()
keep = ops.build_attn_keep(dims, p, (1, 1), dev) if p > 0 else None
for _ in range(3):
    ctx, lse = ops.attention_fwd(qkv, valid, None, dims, False, 0, p, (1, 1), allow, keep=keep)
    if bwd:
        ops.attention_bwd(w, qkv, ctx, lse, valid, None, dims, False, 0, p, (1, 1), allow, keep=keep)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
lib = _lib.lib()
lib.samk_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.samk_debug_timeline(buf, 4096) == 0
tl = np.array(buf[:]).reshape(64, 64)
names = {0: "mma:top", 1: "mma:kv_full", 2: "mma:S issued", 3: "mma:p_ready", 8: "sm:top", 9: "sm:s_full", 10: "sm:pass1 done",
         11: "sm:max xchg", 12: "sm:P arrived", 13: "sm:l xchg", 14: "sm:o_done", 15: "sm:epilogue end"}
if bwd:
    names = {20: "ew:top", 21: "ew:s_full", 22: "ew:computed", 23: "ew:stage_free", 24: "ew:arrived", 25: "ew:drained",
             26: "mma:top", 27: "mma:blk_done", 28: "mma:issued"}
    t0 = tl[20, 0]
    for it in range(20):
        print("sub %d: " % it + "  ".join("%s=%d" % (names[s], tl[s, it] - t0) for s in sorted(names)))
    print("per-warp (warp 2..17 = quarter w%4, slice (w-2)//4): cycles after warp 2's top of the same sub")
    for it in range(8):
        print("sub %d s_full: " % it + " ".join("%5d" % (tl[30 + w, it] - tl[20, it]) for w in range(16)))
        print("sub %d arrive: " % it + " ".join("%5d" % (tl[46 + w, it] - tl[20, it]) for w in range(16)))
else:
    t0 = tl[8, 0]
    for it in range(10):
        print("item %d: " % it + "  ".join("%s=%d" % (names[s], tl[s, it] - t0) for s in sorted(names)))
