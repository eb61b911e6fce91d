#!/usr/bin/env python
"""Times the GEMM shapes of one SA-M4C encoder layer (B=128, L=182) through the C ABI, in isolation."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_textvqa_b200 import ops
dev = torch.device("cuda:0")
M = 23296
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(os.environ.get("REPS", "20"))

def bf(*s): return (0.1 * torch.randn(*s, device=dev)).bfloat16()      # gradients
def hf(*s): return (0.1 * torch.randn(*s, device=dev)).half()          # forward activations / weights
def f32(*s): return 0.1 * torch.randn(*s, device=dev)

def run(name, fn, flops):
    if only and only != name: return
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print("%-14s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flops / ms / 1e9), flush=True)

x768, x3072, W_qkv, W_o, W_1, W_2 = hf(M, 768), hf(M, 3072), hf(2304, 768), hf(768, 768), hf(3072, 768), hf(768, 3072)
dy768, dy3072, dy2304 = bf(M, 768), bf(M, 3072), bf(M, 2304)
b768, b2304, b3072 = f32(768), f32(2304), f32(3072)
res = f32(M, 768)
o768f, o768b, o2304b, o3072b, pre3072 = torch.empty(M, 768, device=dev), bf(M, 768), hf(M, 2304), hf(M, 3072), hf(M, 3072)
o3072g = bf(M, 3072)
dgelu3072 = bf(M, 3072)            # gelu'(h): bf16 in the product mode (ops.BertLayerFn)
W_qkvb, W_ob, W_1b, W_2b = W_qkv.bfloat16(), W_o.bfloat16(), W_1.bfloat16(), W_2.bfloat16()   # dgrad reads bf16 weight copies
x768b = bf(M, 768)                 # bf16 activation copy the weight-gradient products read
dy768h = hf(M, 768)                # scaled-half gradient (ops.scaled_f16) for the products against half activations
O = lambda t: ops.Operand(t, t.stride(0), 1)
g = ops.gemm
run("plain_3072", lambda: g(O(x768), False, O(W_1), False, M, 3072, 768, o3072b), 2 * M * 3072 * 768)
o768h = hf(M, 768)
run("plain_768k3072", lambda: g(O(x3072), False, O(W_2), False, M, 768, 3072, o768h), 2 * M * 3072 * 768)
run("qkv_fwd", lambda: g(O(x768), False, O(W_qkv), False, M, 2304, 768, o2304b, bias=b2304), 2 * M * 2304 * 768)
run("o_fwd", lambda: g(O(x768), False, O(W_o), False, M, 768, 768, o768f, bias=b768, drop_p=0.1, drop=(1, 2), residual=res), 2 * M * 768 * 768)
run("o_f32", lambda: g(O(x768), False, O(W_o), False, M, 768, 768, o768f), 2 * M * 768 * 768)
run("o_f32_bias", lambda: g(O(x768), False, O(W_o), False, M, 768, 768, o768f, bias=b768), 2 * M * 768 * 768)
run("o_f32_res", lambda: g(O(x768), False, O(W_o), False, M, 768, 768, o768f, residual=res), 2 * M * 768 * 768)
run("o_f32_drop", lambda: g(O(x768), False, O(W_o), False, M, 768, 768, o768f, drop_p=0.1, drop=(1, 2)), 2 * M * 768 * 768)
run("ffn1_fwd", lambda: g(O(x768), False, O(W_1), False, M, 3072, 768, o3072b, bias=b3072, act=3, pre=dgelu3072), 2 * M * 3072 * 768)
run("ffn1_gelu", lambda: g(O(x768), False, O(W_1), False, M, 3072, 768, o3072b, bias=b3072, act=1), 2 * M * 3072 * 768)
run("ffn1_bias", lambda: g(O(x768), False, O(W_1), False, M, 3072, 768, o3072b, bias=b3072), 2 * M * 3072 * 768)
run("ffn2_fwd", lambda: g(O(x3072), False, O(W_2), False, M, 768, 3072, o768f, bias=b768, drop_p=0.1, drop=(1, 2), residual=res), 2 * M * 3072 * 768)
run("ffn2_dgrad", lambda: g(O(dy768), False, O(W_2b), True, M, 3072, 768, o3072g, act=4, aux=dgelu3072), 2 * M * 3072 * 768)
run("ffn1_dgrad", lambda: g(O(dy3072), False, O(W_1b), True, M, 768, 3072, o768f, residual=res), 2 * M * 3072 * 768)
run("o_dgrad", lambda: g(O(dy768), False, O(W_ob), True, M, 768, 768, o768b), 2 * M * 768 * 768)
run("qkv_dgrad", lambda: g(O(dy2304), False, O(W_qkvb), True, M, 768, 2304, o768f, residual=res), 2 * M * 2304 * 768)
gw = torch.zeros(768, 3072, device=dev)
run("ffn2_wgrad", lambda: g(O(dy768h), True, O(x3072), True, 768, 3072, M, gw, accumulate=True), 2 * M * 3072 * 768)
gw1 = torch.zeros(3072, 768, device=dev)
run("ffn1_wgrad", lambda: g(O(dy3072), True, O(x768b), True, 3072, 768, M, gw1, accumulate=True), 2 * M * 3072 * 768)
gw2 = torch.zeros(768, 768, device=dev)
run("sq_wgrad", lambda: g(O(dy768h), True, O(x768), True, 768, 768, M, gw2, accumulate=True), 2 * M * 768 * 768)
# classifier head: 1536 decoder rows x 5000 answers, 3-term split (K = 3 x 768), fp32 into the [rows, 5050] score buffer
Mc, Vc, Rc = 1536, 5000, 50
xs3, wc3 = bf(Mc, 2304), bf(Vc, 2304)
scores = torch.empty(Mc, Vc + Rc, device=dev)
bV = f32(Vc)
run("cls_fwd_5050", lambda: g(O(xs3), False, O(wc3), False, Mc, Vc, 2304, scores[:, :Vc], bias=bV), 2 * Mc * Vc * 2304)
scores8 = torch.empty(Mc, Vc + Rc + 6, device=dev)
run("cls_fwd_5056", lambda: g(O(xs3), False, O(wc3), False, Mc, Vc, 2304, scores8[:, :Vc], bias=bV), 2 * Mc * Vc * 2304)
