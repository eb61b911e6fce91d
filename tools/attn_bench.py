#!/usr/bin/env python
"""Attention-kernel benchmark: fwd / bwd time, achieved HBM GB/s and dense-equivalent TFLOP/s
(SURVEY.md section 8d formulas) at the shipped geometry or over the joint-token sweep of BASELINE config 5.

    python tools/attn_bench.py [--sweep] [--B 128] [--p 0.1] [--reps 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_textvqa_b200 import ops, synth
from sam_textvqa_b200.sa_m4c import pack_relation_bits

ap = argparse.ArgumentParser()
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--p", type=float, default=0.1)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--context", type=int, default=3)
ap.add_argument("--only", type=int, default=0, help="one point of the sweep: number of objects O (B=64)")
args = ap.parse_args()
dev = torch.device("cuda:0")
HBM = 6514.2
if os.path.exists("MEASURED_PEAKS.json"):
    HBM = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", HBM)

def run(B, T, O, R, D, spatial=True):
    A, L, H, d = O + R, T + O + R + D, 12, 768
    qkv = torch.randn(B * L, 3 * d, device=dev).half()
    valid = torch.ones(B, L, dtype=torch.uint8, device=dev); valid[:, -D:] = 0
    rs = np.random.RandomState(0)
    bits = None
    if spatial:
        types = torch.from_numpy(rs.choice(13, size=(B, A, A), p=[0.62] + [0.38 / 12] * 12).astype(np.int8))
        bits = pack_relation_bits(synth.expand_types_to_heads(types, args.context), dev)
    dims = (B, L, H, T, A, D)
    allow = ops.build_attn_mask(valid, bits, dims, spatial, 0b11 if spatial else 0)
    w = torch.randn(B * L, d, device=dev).bfloat16()
    keep = ops.build_attn_keep(dims, args.p, (1, 1), dev) if args.p > 0 else None
    f = lambda: ops.attention_fwd(qkv, valid, bits, dims, spatial, 0b11 if spatial else 0, args.p, (1, 1), allow, keep=keep)
    ctx, lse = f()
    g = lambda: ops.attention_bwd(w, qkv, ctx, lse, valid, bits, dims, spatial, 0b11 if spatial else 0, args.p, (1, 1), allow, keep=keep)
    for _ in range(3): f(); g()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(args.reps): f()
    ev[1].record()
    for _ in range(args.reps): g()
    ev[2].record(); torch.cuda.synchronize()
    tf, tb = ev[0].elapsed_time(ev[1]) / args.reps * 1e3, ev[1].elapsed_time(ev[2]) / args.reps * 1e3
    mask_b = B * ((H if spatial else 1) + (H if keep is not None else 0)) * L * ((L + 31) // 32) * 4
    bytes_f = B * (4 * L * d * 2 + H * L * 4) + mask_b
    bytes_b = B * (8 * L * d * 2 + 2 * H * L * 4) + mask_b
    fl_f, fl_b = B * 4 * L * L * d, B * 10 * L * L * d
    print("B=%d L=%d (T=%d O=%d R=%d D=%d) spatial=%d p=%.1f | fwd %7.1f us %6.0f GB/s (%4.1f%% of %.0f) %6.1f TFLOP/s | bwd %7.1f us %6.0f GB/s (%4.1f%%) %6.1f TFLOP/s"
          % (B, L, T, O, R, D, spatial, args.p, tf, bytes_f / tf / 1e3, 100 * bytes_f / tf / 1e3 / HBM, HBM, fl_f / tf / 1e6,
             tb, bytes_b / tb / 1e3, 100 * bytes_b / tb / 1e3 / HBM, fl_b / tb / 1e6), flush=True)

if args.only:
    run(64, 20, args.only, 50, 12)
elif args.sweep:
    for O in (36, 186, 442, 954):          # joint tokens 106 / 256 / 512 / 1024 (BASELINE config 5), B=64
        run(64, 20, O, 50, 12)
else:
    run(args.B, 20, 100, 50, 12)
    run(args.B, 20, 100, 50, 12, spatial=False)
