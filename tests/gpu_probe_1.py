"""First GPU probe: graph kernel vs oracle, tcgen05 GEMM (all operand majors) vs torch."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_textvqa_b200 import _lib
from sam_textvqa_b200._lib import lib, check, ptr, stream_ptr, GemmEpilogue
from oracle import graph_oracle as G
from sam_textvqa_b200 import synth

dev = torch.device("cuda:0")
L = lib()
print("sm_count", L.samk_sm_count(), torch.cuda.get_device_name(0), flush=True)

# ---------------- graph ----------------
g = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_kat.npz"))
names = sorted({k.split("/")[0] for k in g.files})
bad = 0
for name in names:
    b = g[name + "/boxes"]
    N = b.shape[0]
    for dt in (torch.float64, torch.float32):
        boxes = torch.from_numpy(b).to(dev, dt).contiguous().view(1, N, 4)
        types = torch.empty(1, N, N, dtype=torch.int8, device=dev)
        shared = torch.empty(8, 1, N, N, dtype=torch.int8, device=dev)
        bits = torch.empty(1, N, N, dtype=torch.int16, device=dev)
        fn = L.samk_build_graph_f64 if dt == torch.float64 else L.samk_build_graph_f32
        check(fn(ptr(boxes), ptr(types), ptr(shared), ptr(bits), 1, N, 0.5, 3, None, stream_ptr()), "graph")
        torch.cuda.synchronize()
        if dt == torch.float32 and not np.array_equal(b.astype(np.float32).astype(np.float64), b):
            continue
        ok = np.array_equal(types[0].cpu().numpy(), g[name + "/m1"])
        for i, k in enumerate(("31", "32", "51", "52", "71", "72", "91", "92")):
            ok = ok and np.array_equal(shared[i, 0].cpu().numpy(), g[name + "/m" + k])
        hb = bits[0].cpu().numpy().astype(np.uint16)
        heads = ((hb[..., None] >> np.arange(12)) & 1).astype(np.int8)
        ok = ok and np.array_equal(heads, g[name + "/heads3"])
        if not ok:
            bad += 1
            print("GRAPH MISMATCH", name, dt)
print("graph golden sets:", len(names), "mismatches:", bad, flush=True)
# big random batch vs numpy oracle
rs = np.random.RandomState(0)
B, N = 16, 150
bx = synth.make_boxes(rs, B, N)[..., :4]
bx[:, 140:] = 0
boxes = torch.from_numpy(bx).to(dev)
types = torch.empty(B, N, N, dtype=torch.int8, device=dev)
check(L.samk_build_graph_f32(ptr(boxes), ptr(types), None, None, B, N, 0.5, 1, None, stream_ptr()))
torch.cuda.synchronize()
ref = np.stack([G.build_graph(bx[i].astype(np.float64))["1"] for i in range(B)])
print("graph random batch equal:", np.array_equal(ref, types.cpu().numpy()), flush=True)

# ---------------- gemm ----------------
def run_gemm(M, N, K, a_mn, b_mn, impl, split=1, bias=False, act=0, out_bf16=False, residual=False, atomic=False):
    gen = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=gen) * 0.5).to(dev).bfloat16()
    Bm = (torch.randn(N, K, generator=gen) * 0.5).to(dev).bfloat16()
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = Bm.t().contiguous() if b_mn else Bm.contiguous()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ep = GemmEpilogue()
    ep.out = out.data_ptr(); ep.ldo = N; ep.out_dtype = 1 if out_bf16 else 0; ep.atomic_add = 1 if atomic else 0
    ep.alpha = 1.0
    bias_t = torch.randn(N, device=dev) if bias else None
    ep.bias = bias_t.data_ptr() if bias else None
    ep.act = act
    res_t = torch.randn(M, N, device=dev) if residual else None
    ep.residual = res_t.data_ptr() if residual else None
    ep.ldres = N
    check(L.samk_gemm_bf16(ptr(As), a_mn, As.stride(0), ptr(Bs), b_mn, Bs.stride(0), M, N, K,
                           ctypes.byref(ep), split, impl, stream_ptr()), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t()
    if bias: ref = ref + bias_t
    if act == 1: ref = torch.nn.functional.gelu(ref)
    if residual: ref = ref + res_t
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    return err

for impl in (1, 0):
    for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        for (M, N, K) in ((128, 256, 64), (256, 512, 256), (200, 136, 72), (1000, 768, 768)):
            if impl == 1 and M * N * K > 3e8: continue
            try:
                e = run_gemm(M, N, K, a_mn, b_mn, impl)
                print("gemm impl=%d a_mn=%d b_mn=%d %dx%dx%d rel_err=%.3e %s" % (impl, a_mn, b_mn, M, N, K, e, "OK" if e < 2e-2 else "BAD"), flush=True)
            except Exception as ex:
                print("gemm impl=%d a_mn=%d b_mn=%d %dx%dx%d EXC %s" % (impl, a_mn, b_mn, M, N, K, ex), flush=True)
for kw in (dict(bias=True), dict(bias=True, act=1, out_bf16=True), dict(bias=True, residual=True), dict(split=4, atomic=True)):
    e = run_gemm(1024, 768, 1024, 0, 0, 0, **kw)
    print("gemm epi", kw, "rel_err=%.3e" % e, flush=True)

# timing of the big FFN shape
M, N, K = 23296, 3072, 768
A = torch.randn(M, K, device=dev).bfloat16(); Bm = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
ep = GemmEpilogue(); ep.out = out.data_ptr(); ep.ldo = N; ep.out_dtype = 1; ep.alpha = 1.0
for _ in range(3):
    check(L.samk_gemm_bf16(ptr(A), 0, K, ptr(Bm), 0, K, M, N, K, ctypes.byref(ep), 1, 0, stream_ptr()))
s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    check(L.samk_gemm_bf16(ptr(A), 0, K, ptr(Bm), 0, K, M, N, K, ctypes.byref(ep), 1, 0, stream_ptr()))
e_.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e_) / 20
print("FFN1 gemm %dx%dx%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2 * M * N * K / ms / 1e9), flush=True)
s.record()
for _ in range(20):
    torch.matmul(A, Bm.t())
e_.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e_) / 20
print("cuBLAS same shape: %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * N * K / ms / 1e9), flush=True)
