"""Operators of the SA-M4C hot path: thin Python over the C ABI of libsamk.so.

Every function here launches hand-written CUDA kernels on the caller's current device and stream
(`torch.cuda.current_stream()`); PyTorch only owns the memory and the autograd tape.  There is no
fallback: CPU tensors are rejected and a missing library raises.

Precision modes (``set_precision`` / env ``SAMK_PRECISION``):
  "f16"     the product mode.  Forward activations and weight operands that feed a contraction are IEEE half
            (11 significant bits), gradients bfloat16 (fp32 exponent range); fp32 accumulation in TMEM, fp32 residual
            stream / LayerNorm / softmax statistics.  tcgen05 needs both operands of a product in one format (mixed
            f16 x bf16 faults on sm_100a), so in the backward pass dgrad multiplies bf16 dY with a bf16 weight copy,
            wgrad multiplies bf16 dY with the bf16 activation copy LayerNorm wrote next to the f16 one -- or, where the
            saved activation is the big one (GELU output, attention context), an exactly scaled f16 copy of the small
            dY with the f16 activation -- and attention backward runs in f16 with one exact power-of-two scale per
            (sample, head).  No global loss scale.  The small projections whose rounding dominates the logit error
            (OCR pointer network, obj / ocr input encoders, classifier: 2 % of the FLOPs) run as the 3-term split
            below.  Measured on the c3 golden: logits within 1e-3 of the fp32 reference, argmax identical
            (error budget per component in DESIGN.md).
  "bf16x3"  the strict mode.  Activations stay fp32; every contraction runs as a 3-term bf16 split
            (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, one tcgen05 GEMM over a 3x longer K) and attention
            runs in exact fp32: ~2e-5 on the logits.
(The round-1 "bf16" mode -- bf16 forward operands, 4e-3 on the logits -- is gone: half operands cost the same.)
"""
import ctypes
import math
import os
import threading

import torch

from . import _lib
from ._lib import DT_BF16, DT_F16, DT_F32, GemmEpilogue, check, lib, ptr, stream_ptr

_PRECISION = os.environ.get("SAMK_PRECISION", "f16")
if _PRECISION not in ("f16", "bf16x3"):
    raise ValueError("SAMK_PRECISION must be 'f16' or 'bf16x3'")
_GEMM_IMPL = int(os.environ.get("SAMK_GEMM_IMPL", "0"))
_ATTN_IMPL = int(os.environ.get("SAMK_ATTN_IMPL", "0"))
launch_count = 0  # kernels launched through this module (bench.py reports it)
class TimingEvent(object):
    """CUDA event that may be recorded inside a stream capture (samk_timing_event_*): same record() / elapsed_time()
    surface as torch.cuda.Event, readable after a replay of the captured graph."""

    def __init__(self):
        h = ctypes.c_void_p()
        check(lib().samk_timing_event_create(ctypes.byref(h)), "timing_event_create")
        self.h = h

    def record(self):
        check(lib().samk_timing_event_record(self.h, stream_ptr()), "timing_event_record")

    def elapsed_time(self, end):
        ms = ctypes.c_float()
        check(lib().samk_timing_event_elapsed_ms(self.h, end.h, ctypes.byref(ms)), "timing_event_elapsed")
        return ms.value

    def __del__(self):
        try:
            lib().samk_timing_event_destroy(self.h)
        except Exception:
            pass


profile_in_graph = False     # bench.py: the profile lists below take TimingEvent pairs (recorded inside the captured step)


def _prof_events():
    if profile_in_graph:
        return TimingEvent(), TimingEvent()
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


gemm_profile = None  # bench.py: list collecting (start_event, end_event, algorithmic_flops, (M, N, K, a_mn, b_mn)) per GEMM launch
grad_ready_hook = None  # dp.FlatGradBuffer.enable_overlap: called with the parameters a backward op has just finished
launch_log = None    # bench.py --profile-only: list collecting ("gemm", M, N, K, a_mn, b_mn) / ("attn_fwd" | "attn_bwd", L) in launch order
attn_profile = None  # bench.py: list collecting (kind, L, start_event, end_event, algorithmic_bytes, dense_flops) per launch


def set_precision(mode):
    global _PRECISION
    if mode not in ("f16", "bf16x3"):
        raise ValueError("precision must be 'f16' or 'bf16x3'")
    _PRECISION = mode


def get_precision():
    return _PRECISION


def act_dtype():
    """storage dtype of forward activations that feed a contraction"""
    return torch.float16 if _PRECISION == "f16" else torch.float32


def grad_dtype():
    """storage dtype of gradients that feed a contraction"""
    return torch.bfloat16 if _PRECISION == "f16" else torch.float32


_DT = {torch.bfloat16: DT_BF16, torch.float32: DT_F32, torch.float16: DT_F16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("unsupported dtype %s" % t.dtype)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SamkError("samk ops need CUDA tensors (no CPU path); got a tensor on %s" % t.device)


def _count(n=1):
    global launch_count
    launch_count += n


# ---- gradient targets ---------------------------------------------------------------------------------
direct_grad_accumulation = True   # write weight gradients straight into existing fp32 .grad buffers


def _gbuf(p):
    """(buffer, direct): where the gradient of parameter p is accumulated.  When p.grad already exists
    (e.g. views of dp.FlatGradBuffer) the kernels add into it and autograd gets None for that input;
    otherwise a zeroed buffer is returned to autograd."""
    g = p.grad if p.is_leaf else None
    if (direct_grad_accumulation and g is not None and g.dtype == torch.float32 and g.is_contiguous()
            and g.shape == p.shape and g.device == p.device):
        return g, True
    return zeros(p.shape, torch.float32, p.device), False


def _ret(buf_direct):
    return None if buf_direct[1] else buf_direct[0]


# parameters that also receive a gradient contribution from autograd AFTER the op that owns them has finished (the
# out-projection weight and the context biases of a `use_bias` spatial layer: the folded bias W_o b + b_o is formed by a
# torch op outside BertLayerFn); they are never reported early, the final flush of the exchange sends them
late_grad_param_ids = set()


def _grads_done(*params):
    if grad_ready_hook is not None:
        grad_ready_hook([p for p in params if p is not None and p.is_leaf and id(p) not in late_grad_param_ids])


def _ln_partials(dev, cols):
    """scratch of the LayerNorm backward's two-stage reduction: one buffer per (width, STREAM) -- the input encoders and
    TextBert run their backward passes on different streams at the same time (sa_m4c.SAM4C.forward)"""
    ws = _state(dev).ln_ws
    key = (cols, torch.cuda.current_stream().cuda_stream)
    if key not in ws:
        ws[key] = torch.empty(int(lib().samk_layernorm_bwd_partials(cols)), dtype=torch.float32, device=dev)
    return ws[key]


def _ln_bwd(dy, x, gamma, eps, dx, dxd, p, drop, dg, db, dbias, rows, cols, amax=None, side_final=False):
    """side_final: the reduction of the block partial sums into dgamma / dbeta / dbias runs on the side branch (beside
    the GEMM the caller issues next; joined by the caller's join_side()) from a scratch buffer of its own."""
    if side_final and _SIDE_BRANCH and rows > 0:
        part = torch.empty(int(lib().samk_layernorm_bwd_partials(cols)), dtype=torch.float32, device=dy.device)
        check(lib().samk_layernorm_bwd_main(ptr(dy), ptr(x), ptr(gamma), eps, ptr(dx), ptr(dxd),
                                            _dt(dxd) if dxd is not None else 0, p, drop[0], drop[1], ptr(dg), ptr(db),
                                            ptr(dbias), ptr(part), rows, cols, ptr(amax), stream_ptr()), "layernorm_bwd_main")
        with side_branch():
            check(lib().samk_layernorm_bwd_finalize(ptr(part), rows, cols, ptr(dg), ptr(db), ptr(dbias), stream_ptr()),
                  "layernorm_bwd_finalize")
            _state(dy.device).side_keep.append(part)       # alive until the branch is joined
        _count(2)
        return
    check(lib().samk_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), eps, ptr(dx), ptr(dxd),
                                   _dt(dxd) if dxd is not None else 0, p, drop[0], drop[1], ptr(dg), ptr(db),
                                   ptr(dbias), ptr(_ln_partials(dy.device, cols)), rows, cols, ptr(amax), stream_ptr()),
          "layernorm_bwd")
    _count(2)


# ---- dropout stream ------------------------------------------------------------------------------
_BASE_SEED = 0x5A17C0DE


def _rank_salt():
    """Data-parallel replicas must not draw the same dropout masks: the stream seed mixes in the process rank
    (torch.distributed, else the RANK variable torchrun sets) and the device index (nn.DataParallel threads)."""
    rank = 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank = dist.get_rank()
        else:
            rank = int(os.environ.get("RANK", "0"))
    except Exception:
        rank = 0
    return rank


class _Rng(object):
    def __init__(self, dev=0):
        self.dev = int(dev)
        self.reseed(_BASE_SEED)

    def reseed(self, seed):
        z = (int(seed) + 0x9E3779B97F4A7C15 * (1 + _rank_salt()) + 0xD1B54A32D192ED03 * self.dev) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        self.seed = z ^ (z >> 31)
        self.offset = 0

    def next(self):
        self.offset += 1
        return self.seed, self.offset


def _next_drop(dev=None):
    return _state(dev).rng.next()


def manual_seed(seed):
    """Re-seed the dropout stream of the current device (rank and device index are mixed in)."""
    _state().rng.reseed(seed)


# ---- operand preparation ---------------------------------------------------------------------------
def _ceil8(n):
    return (n + 7) // 8 * 8


def cast_16(x2d, dtype):
    """fp32 [rows, cols] (row stride arbitrary, unit column stride) -> half / bf16 [rows, ceil8(cols)] buffer."""
    rows, cols = x2d.shape
    assert x2d.stride(1) == 1 and cols % 4 == 0
    y = torch.empty(rows, _ceil8(cols), dtype=dtype, device=x2d.device)
    check(lib().samk_cast_16(ptr(x2d), x2d.stride(0), ptr(y), y.stride(0), _DT[dtype], rows, cols, stream_ptr()), "cast_16")
    _count()
    return y


def cast_flat(src, dst):
    """contiguous fp32 <-> 16-bit copy of a flat buffer (gradient wire format of dp.GradExchange)"""
    _cuda(src, dst)
    n = src.numel()
    assert dst.numel() == n and src.is_contiguous() and dst.is_contiguous()
    check(lib().samk_cast_flat(ptr(src), _dt(src), ptr(dst), _dt(dst), n, stream_ptr()), "cast_flat")
    _count()


def _split3(x2d, order, along_rows):
    rows, cols = x2d.shape
    assert x2d.stride(1) == 1 and cols % 4 == 0
    if along_rows:
        y = torch.empty(3 * rows, _ceil8(cols), dtype=torch.bfloat16, device=x2d.device)
    else:
        y = torch.empty(rows, _ceil8(3 * cols), dtype=torch.bfloat16, device=x2d.device)
    check(lib().samk_split3_bf16(ptr(x2d), x2d.stride(0), ptr(y), y.stride(0), rows, cols, order, along_rows,
                                 stream_ptr()), "split3")
    _count()
    return y


class Operand(object):
    """A GEMM operand ready for the tensor-core kernel: 16-bit storage, its format, leading dimension, K multiplier."""
    __slots__ = ("t", "ld", "kmul", "dt")

    def __init__(self, t, ld, kmul):
        self.t, self.ld, self.kmul, self.dt = t, ld, kmul, _DT[t.dtype]


class _State(object):
    """Per-device mutable state of the operators: dropout stream, operand caches, side stream, LayerNorm workspace.
    `nn.DataParallel` (the reference's multi-GPU path, train.py:111-112) runs replicas concurrently in Python threads
    on DIFFERENT devices, so nothing here is shared between them; the forward (caller's thread) and the backward
    (autograd's device thread) of one replica share their device's state, which is what the operand caches need."""

    def __init__(self, dev):
        self.rng = _Rng(dev)
        self.act_cache = {}
        self.weight_cache = {}
        self.ln_ws = {}
        self.side_stream = None
        self.side_open = False
        self.side_keep = []            # tensors a side-branch kernel still reads (released by join_side)


_states = {}
_states_lock = threading.Lock()


def _state(dev=None):
    if dev is None:
        dev = torch.cuda.current_device()
    elif isinstance(dev, torch.device):
        dev = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _states.get(dev)
    if st is None:
        with _states_lock:
            st = _states.get(dev)
            if st is None:
                st = _states[dev] = _State(dev)
    return st


# 16-bit operand copies of fp32 activations made during the current forward pass, keyed on the fp32 tensor's
# storage.  The entry keeps the fp32 tensor alive (so the address cannot be recycled under the key) and is
# dropped at the next forward (`begin_forward`).  Serves the backward pass (wgrad re-reads the layer input) and the
# next layer (LayerNorm writes the half copy of its output in the same pass, no separate cast kernel).
def begin_forward():
    _state().act_cache.clear()


def _act_key(x2d):
    return (x2d.data_ptr(), tuple(x2d.shape), x2d.stride(0), x2d._version)


def remember_act(x2d, *copies):
    """register 16-bit copies (any formats) of the fp32 activation x2d made by the kernel that produced it"""
    if _PRECISION == "f16":
        cache = _state(x2d.device).act_cache
        for y in copies:
            if y is not None:
                cache[(_act_key(x2d), y.dtype)] = (x2d, y)


def cached_copy(x2d, dtype):
    hit = _state(x2d.device).act_cache.get((_act_key(x2d), dtype))
    return None if hit is None else hit[1]


_FMT_DTYPE = {"f16": torch.float16, "bf16": torch.bfloat16}


def operand(x2d, role, mn_major, fmt="f16", split=None):
    """x2d: logical [rows(MN), K] if not mn_major else stored [K, MN].  role 'a' or 'b'.
    fmt: 16-bit format an fp32 tensor is cast to in the product mode ("f16": forward values, "bf16": gradients and
    the activation copies weight-gradient products read); a 16-bit tensor is used as it is.
    split: force (True) / forbid (False) the 3-term bf16 split; None = by precision mode."""
    _cuda(x2d)
    if split is None:
        split = _PRECISION != "f16"
    if x2d.dtype != torch.float32:
        if split:
            raise _lib.SamkError("16-bit activation reached a split (fp32-accurate) contraction")
        assert x2d.stride(1) == 1 and x2d.stride(0) % 8 == 0 and x2d.data_ptr() % 16 == 0
        return Operand(x2d, x2d.stride(0), 1)
    if split:
        y = _split3(x2d, 0 if role == "a" else 1, 1 if mn_major else 0)
        return Operand(y, y.stride(0), 3)
    dtype = _FMT_DTYPE[fmt]
    y = cached_copy(x2d, dtype)
    if y is None:
        y = cast_16(x2d, dtype)
        if x2d.shape[0] >= 1024 and x2d.shape[1] % 8 == 0:       # activations worth remembering (wgrad re-reads them)
            remember_act(x2d, y)
    return Operand(y, y.stride(0), 1)


def scaled_f16(x2d, amax=None, out=None):
    """(Operand over half(x * S), device pointer to 1/S) for a gradient tensor x2d (bf16 or fp32, contiguous): S is the
    power of two, found on the device, that puts max|x| into [2^11, 2^12).  The consumer passes the pointer as
    `alpha_dev` of its GEMM.  Used where a gradient meets a big saved f16 activation in a weight-gradient product.
    amax: device float already holding max|x| (written by the kernel that produced x), saves the reduction pass."""
    _cuda(x2d)
    assert x2d.is_contiguous() and x2d.numel() % 4 == 0
    if out is not None:          # (half [shape of x2d], 4 floats) allocated by the caller, e.g. before it forks a side branch
        y, sc = out
    else:
        y = torch.empty(x2d.shape, dtype=torch.float16, device=x2d.device)
        sc = torch.empty(4, dtype=torch.float32, device=x2d.device)
    check(lib().samk_cast_scaled_f16(ptr(x2d), _dt(x2d), x2d.numel(), ptr(amax), ptr(y), ptr(sc), stream_ptr()),
          "cast_scaled_f16")
    _count(1 if amax is not None else 2)
    return Operand(y, y.stride(0), 1), sc[1:2]


def weight_operand(params, mn_major, split=None, kdim=None, fmt="f16"):
    """Cached operand for one weight (or the row-concatenation of several, e.g. fused q|k|v); kdim: use the first
    kdim input columns only (the OCR projection drops its 50 always-zero columns, sa_m4c.py:240-242).

    params: list of [n_i, K] fp32 parameters.  The cache entry holds the parameters themselves (their storage cannot
    be freed and re-allocated under the key) and is stamped with their version counters, so an optimizer step (in-place
    update) invalidates it; updates that bypass the version counter (`p.data.copy_`, raw-pointer kernels such as
    optim.FlatAdam) must call `clear_weight_cache()` / `bump_versions`.  Non-leaf tensors (e.g. `nn.DataParallel`
    replicas, re-broadcast every step into recycled addresses with version 0) are never cached."""
    if split is None:
        split = _PRECISION != "f16"
    cacheable = all(p.is_leaf for p in params)
    layout = bool(mn_major) and split            # plain 16-bit copies serve both majors
    slot = (tuple(id(p) for p in params), tuple(tuple(p.shape) + tuple(p.stride()) for p in params), layout, split, kdim,
            None if split else fmt)
    stamp = tuple((p._version, p.data_ptr()) for p in params)
    cache = _state(params[0].device).weight_cache
    hit = cache.get(slot) if cacheable else None
    if hit is not None and hit[0] == stamp:
        return hit[1]
    with torch.no_grad():
        w = params[0] if len(params) == 1 else torch.cat(list(params), dim=0)
        if kdim is not None and kdim != w.shape[1]:
            w = w[:, :kdim]
        w = w.detach()
        if not split and cacheable and w.dtype == torch.float32 and w.stride(1) == 1 and w.shape[1] % 4 == 0:
            # product mode: the forward reads the half copy, dgrad the bf16 copy -- both from ONE pass over the weight
            rows, cols = w.shape
            yh = torch.empty(rows, _ceil8(cols), dtype=torch.float16, device=w.device)
            yb = torch.empty(rows, _ceil8(cols), dtype=torch.bfloat16, device=w.device)
            check(lib().samk_cast_dual(ptr(w), w.stride(0), ptr(yh), ptr(yb), yh.stride(0), rows, cols, stream_ptr()), "cast_dual")
            _count()
            ops_ = {"f16": Operand(yh, yh.stride(0), 1), "bf16": Operand(yb, yb.stride(0), 1)}
            for f_, o_ in ops_.items():
                cache[slot[:-1] + (f_,)] = (stamp, o_, tuple(params))
            return ops_[fmt]
        op = operand(w, "b", mn_major, fmt=fmt, split=split)
    if cacheable:
        cache[slot] = (stamp, op, tuple(params))      # the parameters stay alive with the entry: ids cannot be reused
    return op


def clear_weight_cache():
    for st in list(_states.values()):
        st.weight_cache.clear()


# ---- GEMM -------------------------------------------------------------------------------------------
_SMS = {}


def sm_count():
    dev = torch.cuda.current_device()
    if dev not in _SMS:
        _SMS[dev] = max(1, int(lib().samk_sm_count()))
    return _SMS[dev]


def _pick_split_k(M, N, K, sms=None):
    if sms is None:
        sms = sm_count()
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    if tiles >= sms or K <= 1024:
        return 1
    kb = (K + 63) // 64
    return int(max(1, min(kb // 8, (2 * sms) // tiles)))


def gemm(a, a_mn, b, b_mn, M, N, K, out, bias=None, act=0, pre=None, aux=None, drop_p=0.0, drop=None,
         residual=None, alpha=1.0, accumulate=False, split_k=None, out_parts=None, alpha_dev=None):
    """out[M,N] = epilogue(alpha * A.B^T).  a, b: Operand.  out: fp32 / half / bf16 2-D view (unit column stride)."""
    assert a.kmul == b.kmul and a.dt == b.dt, "tcgen05 products need both operands in one format"
    ep = GemmEpilogue()
    ep.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    ep.out = out.data_ptr()
    ep.ldo = out.stride(0)
    ep.out_dtype = _dt(out)
    ep.alpha = alpha
    ep.bias = bias.data_ptr() if bias is not None else None
    if pre is not None:
        ep.pre, ep.ldpre, ep.pre_dtype = pre.data_ptr(), pre.stride(0), _dt(pre)
    ep.act = act
    if aux is not None:
        ep.aux, ep.ldaux, ep.aux_dtype = aux.data_ptr(), aux.stride(0), _dt(aux)
    if drop_p > 0.0:
        ep.drop_p, ep.drop_seed, ep.drop_offset = drop_p, drop[0], drop[1]
    if residual is not None:
        ep.residual, ep.ldres = residual.data_ptr(), residual.stride(0)
    Kk = K * a.kmul
    if split_k is None:
        split_k = _pick_split_k(M, N, Kk) if accumulate else 1
    ep.atomic_add = 1 if (accumulate or split_k > 1) else 0
    if out_parts is not None:       # (rows per part, out1, out2): M = 3 * rows, destinations out / out1 / out2
        ep.part_rows, ep.out_part1, ep.out_part2 = out_parts[0], out_parts[1].data_ptr(), out_parts[2].data_ptr()
    if launch_log is not None:
        launch_log.append(("gemm", int(M), int(N), int(K), bool(a_mn), bool(b_mn)))
    if gemm_profile is not None:
        ev0, ev1 = _prof_events()
        ev0.record()
    check(lib().samk_gemm_16(ptr(a.t), a.dt, 1 if a_mn else 0, a.ld, ptr(b.t), b.dt, 1 if b_mn else 0, b.ld, M, N, Kk,
                             ctypes.byref(ep), split_k, _GEMM_IMPL, stream_ptr()), "gemm")
    if gemm_profile is not None:
        ev1.record()
        gemm_profile.append((ev0, ev1, 2.0 * M * N * K, (int(M), int(N), int(K), bool(a_mn), bool(b_mn))))
    _count()
    return out


def colsum_into(x2d, out):
    check(lib().samk_colsum(ptr(x2d), _dt(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], ptr(out), stream_ptr()),
          "colsum")
    _count()


# ---- side branch -------------------------------------------------------------------------------------
# Bias-gradient column sums are L2/HBM-bound readers whose result nothing inside the backward pass consumes.  They are
# issued on a second stream right behind the kernel that produced their input and joined before the gradients are
# declared final, so they run beside the tensor-bound dgrad / wgrad GEMMs (a parallel branch of the captured graph;
# colsum_kernel uses no shared memory so that its blocks fit on an SM whose shared memory a GEMM CTA owns).
_SIDE_BRANCH = os.environ.get("SAMK_SIDE_BRANCH", "1") != "0"
# also on the branch (SAMK_SIDE_OVERLAP=0 keeps them in line): LayerNorm-backward parameter-gradient reductions, the scaled
# half copies of dY for the weight-gradient products, the attention-backward preparation kernel
# (bit 0: after LayerNorm 2, beside the FFN2 dgrad; bit 1: after LayerNorm 1, beside the out-projection dgrad; bit 2: the
#  attention preparation, beside the out-projection weight-gradient product)
_SIDE_OVERLAP = int(os.environ.get("SAMK_SIDE_OVERLAP", "0")) if _SIDE_BRANCH else 0


class side_branch(object):

    def __enter__(self):
        self._cm = None
        if not _SIDE_BRANCH:
            return self
        st = _state()
        if st.side_stream is None:
            st.side_stream = torch.cuda.Stream(device=torch.cuda.current_device())
        side = st.side_stream
        side.wait_stream(torch.cuda.current_stream())
        self._cm = torch.cuda.stream(side)
        self._cm.__enter__()
        st.side_open = True
        return self

    def __exit__(self, *exc):
        if self._cm is not None:
            self._cm.__exit__(*exc)
        return False


def branch_stream(i=0):
    """extra compute streams of this device (sa_m4c.SAM4C.forward runs TextBert and the OCR encoder on them beside
    the region encoder)"""
    st = _state()
    if getattr(st, "branches", None) is None:
        st.branches = {}
    if i not in st.branches:
        st.branches[i] = torch.cuda.Stream(device=torch.cuda.current_device())
    return st.branches[i]


def join_side():
    st = _state()
    if st.side_open:
        st.side_open = False
        torch.cuda.current_stream().wait_stream(st.side_stream)
    del st.side_keep[:]


# ---- simple differentiable ops -----------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b with fp32 output (input projections, pointer-net projections).
    strict: the forward product runs as the 3-term split in every precision mode (x must then be fp32)."""

    @staticmethod
    def forward(ctx, x2d, weight, bias, kdim, strict):
        _cuda(x2d, weight)
        M = x2d.shape[0]
        N = weight.shape[0]
        K = kdim
        split = True if strict else None
        xs = x2d[:, :K] if x2d.shape[1] != K else x2d
        if (_PRECISION == "f16" and not strict and xs.dtype == torch.float32 and weight.requires_grad and M >= 1024
                and K % 4 == 0 and xs.stride(1) == 1 and cached_copy(xs, torch.float16) is None):
            # the half copy this product reads and the bf16 copy its weight-gradient product will read, from ONE pass
            # over the fp32 input (region / OCR features: 180 MB per step)
            yh = torch.empty(M, _ceil8(K), dtype=torch.float16, device=xs.device)
            yb = torch.empty(M, _ceil8(K), dtype=torch.bfloat16, device=xs.device)
            check(lib().samk_cast_dual(ptr(xs), xs.stride(0), ptr(yh), ptr(yb), yh.stride(0), M, K, stream_ptr()), "cast_dual")
            _count()
            remember_act(xs, yh, yb)
        x_op = operand(xs, "a", False, split=split)
        w_op = weight_operand([weight], False, split=split, kdim=K)
        y = torch.empty(M, N, dtype=torch.float32, device=x2d.device)
        gemm(x_op, False, w_op, False, M, N, K, y, bias=bias)
        ctx.save_for_backward(x2d, weight)
        ctx.bias_ref = bias
        ctx.dims = (M, N, K)
        ctx.x_needs_grad = x2d.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, weight = ctx.saved_tensors
        M, N, K = ctx.dims
        dy = dy.contiguous()
        ctx_bias = ctx.bias_ref
        dW, db = _gbuf(weight), _gbuf(ctx_bias)
        with side_branch():
            colsum_into(dy, db[0])
        if _PRECISION != "f16":
            dy_act = dy
        else:
            dy_act = cached_copy(dy, torch.bfloat16) if dy.shape[1] % 8 == 0 else None      # (written by the LayerNorm backward)
            if dy_act is None:
                dy_act = cast_16(dy, torch.bfloat16)[:, :N]
        dy_k = operand(dy_act, "a", False, fmt="bf16")          # [M, N] K-major for dgrad
        dy_mn = operand(dy_act, "a", True, fmt="bf16")          # stored [tokens, N]: MN-major for wgrad
        xs = x2d[:, :K] if x2d.shape[1] != K else x2d
        if xs.dtype == torch.float16:
            # a half input made by its producer together with a bf16 copy (sa_m4c._feature_buffers): gradients are bf16
            xb = cached_copy(xs, torch.bfloat16)
            if xb is None:
                raise _lib.SamkError("a half input of ops.linear needs a registered bf16 copy for its weight gradient")
            xs = xb
        x_mn = operand(xs, "b", True, fmt="bf16")
        dx = None
        if ctx.x_needs_grad:
            dx = torch.zeros_like(x2d, dtype=torch.float32) if x2d.shape[1] != K else torch.empty(
                M, K, dtype=torch.float32, device=dy.device)
            w_op = weight_operand([weight], True, kdim=K, fmt="bf16")
            gemm(dy_k, False, w_op, True, M, K, N, dx[:, :K])
        gemm(dy_mn, True, x_mn, True, N, K, M, dW[0][:, :K], accumulate=True)
        join_side()
        _grads_done(weight, ctx_bias)
        return dx, _ret(dW), _ret(db), None, None


def linear(x2d, weight, bias, kdim=None, strict=False):
    return LinearFn.apply(x2d, weight, bias, weight.shape[1] if kdim is None else kdim, bool(strict))


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, gamma, beta, eps):
        _cuda(x2d, gamma, beta)
        x2d = x2d.contiguous()
        y = torch.empty_like(x2d)
        check(lib().samk_layernorm_fwd(ptr(x2d), ptr(gamma), ptr(beta), eps, ptr(y), None, 0, None, 0, x2d.shape[0],
                                       x2d.shape[1], stream_ptr()), "layernorm_fwd")
        _count()
        ctx.save_for_backward(x2d, gamma)
        ctx.beta_ref = beta
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x2d)
        dg, db = _gbuf(gamma), _gbuf(ctx.beta_ref)
        dxd = None
        if _PRECISION == "f16" and x2d.shape[0] >= 1024 and x2d.shape[1] % 8 == 0:
            # the producer of x is a projection whose backward reads dx in bf16 (LinearFn.backward): written by this
            # kernel in the same pass instead of a separate cast
            dxd = torch.empty(x2d.shape, dtype=torch.bfloat16, device=x2d.device)
        _ln_bwd(dy, x2d, gamma, ctx.eps, dx, dxd, 0.0, (0, 0), dg[0], db[0], None, x2d.shape[0], x2d.shape[1])
        if dxd is not None:
            remember_act(dx, dxd)
        _grads_done(gamma, ctx.beta_ref)
        return dx, _ret(dg), _ret(db), None


def layer_norm(x2d, gamma, beta, eps):
    return LayerNormFn.apply(x2d, gamma, beta, eps)


class DropoutAddFn(torch.autograd.Function):
    """dropout(a + b)"""

    @staticmethod
    def forward(ctx, a, b, p):
        _cuda(a, b)
        a = a.contiguous()
        b = b.contiguous()
        out = torch.empty_like(a)
        ctx.p = p
        ctx.drop = _next_drop() if p > 0 else (0, 0)
        rows, cols = a.numel() // a.shape[-1], a.shape[-1]
        check(lib().samk_dropout_add(ptr(a), ptr(b), ptr(out), None, 0, rows, cols, p, ctx.drop[0], ctx.drop[1],
                                     stream_ptr()), "dropout_add")
        _count()
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        if ctx.p <= 0:
            return dout, dout, None
        d = torch.empty_like(dout)
        rows, cols = dout.numel() // dout.shape[-1], dout.shape[-1]
        check(lib().samk_dropout_add(ptr(dout), None, ptr(d), None, 0, rows, cols, ctx.p, ctx.drop[0], ctx.drop[1],
                                     stream_ptr()), "dropout_bwd")
        _count()
        return d, d, None


def dropout_add(a, b, p):
    return DropoutAddFn.apply(a, b, float(p))


def l2norm_into(x3d, out2d, col_off, normalize):
    """out2d[:, col_off:col_off+d] = F.normalize(x) (or a plain copy/cast); no gradient (inputs are data)."""
    _cuda(x3d, out2d)
    x2d = x3d.reshape(-1, x3d.shape[-1])
    assert x2d.stride(1) == 1
    dst = out2d[:, col_off:col_off + x2d.shape[1]]
    check(lib().samk_l2norm(ptr(x2d), x2d.stride(0), ptr(dst), out2d.stride(0), _dt(out2d), x2d.shape[0],
                            x2d.shape[1], 1 if normalize else 0, stream_ptr()), "l2norm")
    _count()


def l2norm_into2(x3d, out_a, out_b, col_off, normalize):
    """two 16-bit destinations from one pass over x (see samk_l2norm2); out_b may be None"""
    if out_b is None:
        return l2norm_into(x3d, out_a, col_off, normalize)
    _cuda(x3d, out_a, out_b)
    x2d = x3d.reshape(-1, x3d.shape[-1])
    assert x2d.stride(1) == 1
    d = x2d.shape[1]
    check(lib().samk_l2norm2(ptr(x2d), x2d.stride(0), ptr(out_a[:, col_off:col_off + d]), out_a.stride(0), _dt(out_a),
                             ptr(out_b[:, col_off:col_off + d]), out_b.stride(0), _dt(out_b), x2d.shape[0], d,
                             1 if normalize else 0, stream_ptr()), "l2norm2")
    _count()


# ---- embeddings ---------------------------------------------------------------------------------------
class BertEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, word, pos, type_, gamma, beta, eps, p):
        _cuda(ids, word)
        if ids.dtype != torch.long:
            raise TypeError("question_indices must be int64 (got %s)" % ids.dtype)
        ids = ids.contiguous()
        B, T = ids.shape
        d = word.shape[1]
        out = torch.empty(B * T, d, dtype=torch.float32, device=word.device)
        ctx.drop = _next_drop() if p > 0 else (0, 0)
        ctx.p, ctx.eps = p, eps
        check(lib().samk_bert_embed_fwd(ptr(ids), ptr(word), ptr(pos), ptr(type_), ptr(gamma), ptr(beta), eps,
                                        ptr(out), None, 0, B * T, T, d, p, ctx.drop[0], ctx.drop[1], stream_ptr()),
              "bert_embed_fwd")
        _count()
        ctx.save_for_backward(ids, word, pos, type_, gamma)
        ctx.beta_ref = beta
        return out.view(B, T, d)

    @staticmethod
    def backward(ctx, dout):
        ids, word, pos, type_, gamma = ctx.saved_tensors
        B, T = ids.shape
        d = word.shape[1]
        dout = dout.contiguous()
        gs = [_gbuf(t) for t in (word, pos, type_, gamma, ctx.beta_ref)]
        check(lib().samk_bert_embed_bwd(ptr(dout), ptr(ids), ptr(word), ptr(pos), ptr(type_), ptr(gamma), ctx.eps,
                                        ptr(gs[0][0]), ptr(gs[1][0]), ptr(gs[2][0]), ptr(gs[3][0]), ptr(gs[4][0]),
                                        B * T, T, d, ctx.p, ctx.drop[0], ctx.drop[1], stream_ptr()), "bert_embed_bwd")
        _count()
        return (None,) + tuple(_ret(g) for g in gs) + (None, None)


class PrevPredFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prev, cls_w, ocr_in, pos, type_, ag, ab, og, ob, eg, eb, eps, p):
        _cuda(prev, cls_w, ocr_in)
        prev = prev.contiguous()
        ocr_in = ocr_in.contiguous()
        B, D = prev.shape
        V, d = cls_w.shape
        R = ocr_in.shape[1]
        out = torch.empty(B, D, d, dtype=torch.float32, device=cls_w.device)
        ctx.drop = _next_drop() if p > 0 else (0, 0)
        ctx.p, ctx.eps, ctx.dims = p, eps, (B, D, V, R, d)
        ln6 = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in (ag, ab, og, ob, eg, eb)])
        check(lib().samk_prevpred_fwd(ptr(prev), ptr(cls_w), ptr(ocr_in), ptr(pos), ptr(type_), ln6, eps, ptr(out),
                                      B, D, V, R, d, p, ctx.drop[0], ctx.drop[1], stream_ptr()), "prevpred_fwd")
        _count()
        ctx.save_for_backward(prev, cls_w, ocr_in, pos, type_, ag, ab, og, ob, eg, eb)
        return out

    @staticmethod
    def backward(ctx, dout):
        prev, cls_w, ocr_in, pos, type_, ag, ab, og, ob, eg, eb = ctx.saved_tensors
        B, D, V, R, d = ctx.dims
        dout = dout.contiguous()
        gs = [_gbuf(t) if t.is_leaf else (zeros(t.shape, torch.float32, t.device), False)
              for t in (cls_w, ocr_in, pos, type_, ag, ab, og, ob, eg, eb)]
        grads = [g[0] for g in gs]
        ln6 = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in (ag, ab, og, ob, eg, eb)])
        g10 = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in grads])
        check(lib().samk_prevpred_bwd(ptr(dout), ptr(prev), ptr(cls_w), ptr(ocr_in), ptr(pos), ptr(type_), ln6,
                                      ctx.eps, g10, B, D, V, R, d, ctx.p, ctx.drop[0], ctx.drop[1], stream_ptr()),
              "prevpred_bwd")
        _count()
        _grads_done(*[t for t in (cls_w, ocr_in, pos, type_, ag, ab, og, ob, eg, eb) if t.is_leaf])
        return (None,) + tuple(_ret(g) for g in gs) + (None, None)


# ---- attention ------------------------------------------------------------------------------------------
def _attn_params(qkv, ctx_t, lse, valid, rel, dims, spatial, quad_mask, p, drop, dctx=None, dqkv=None, delta=None,
                 allow=None, dq_accum=None, keep=None):
    B, L, H, T, A, D = dims
    ap = _lib.AttnParams()
    if qkv is not None:
        ap.qkv, ap.ctx, ap.lse = qkv.data_ptr(), ctx_t.data_ptr(), lse.data_ptr()
        ap.dtype = _dt(qkv)
        ap.grad_dtype = DT_BF16 if qkv.dtype == torch.float16 else DT_F32
    if dctx is not None:
        ap.dctx, ap.dqkv, ap.delta = dctx.data_ptr(), dqkv.data_ptr(), delta.data_ptr()
        ap.grad_dtype = _dt(dctx)
    ap.keep_bits = keep.data_ptr() if keep is not None else None
    ap.B, ap.H, ap.head_dim = B, H, 64
    ap.T, ap.A, ap.D = T, A, D
    ap.key_valid = valid.data_ptr()
    ap.rel_bits = rel.data_ptr() if (spatial and rel is not None) else None
    ap.quadrant_mask = quad_mask if spatial else 0
    ap.spatial = 1 if spatial else 0
    ap.scale = 1.0 / math.sqrt(64.0)
    ap.drop_p = p
    ap.drop_seed, ap.drop_offset = drop
    ap.allow_bits = allow.data_ptr() if allow is not None else None
    ap.dq_accum = dq_accum.data_ptr() if dq_accum is not None else None
    return ap


def uses_tensor_core_attention(dtype):
    return dtype == torch.float16 and _ATTN_IMPL == 0


def build_attn_keep(dims, p, drop, device):
    """Dropout keep bits [B, H, L, ceil(L/32)] of one layer's attention probabilities (stream `drop`), shared by the
    tensor-core forward and backward kernels."""
    B, L, H, T, A, D = dims
    keep = torch.empty(B * H * L * ((L + 31) // 32), dtype=torch.int32, device=device)
    fill_attn_keep(keep, dims, p, drop)
    return keep


def fill_attn_keep(keep, dims, p, drop):
    B, L, H, T, A, D = dims
    ap = _lib.AttnParams()
    ap.B, ap.H, ap.head_dim, ap.T, ap.A, ap.D = B, H, 64, T, A, D
    ap.drop_p = p
    ap.drop_seed, ap.drop_offset = drop
    check(lib().samk_attn_build_keep(ctypes.byref(ap), ptr(keep), stream_ptr()), "attn_build_keep")
    _count()


def build_attn_mask(valid, rel, dims, spatial, quad_mask):
    """Packed allow-bits [B, H|1, L, ceil(L/32)] for the tensor-core attention (built once per step and
    mask kind, shared by every layer and by the backward pass)."""
    B, L, H, T, A, D = dims
    words = lib().samk_attn_mask_words(B, H, T, A, D, 1 if spatial else 0)
    allow = torch.empty(max(int(words), 1), dtype=torch.int32, device=valid.device)
    ap = _attn_params(None, None, None, valid, rel, dims, spatial, quad_mask, 0.0, (0, 0))
    check(lib().samk_attn_build_mask(ctypes.byref(ap), ptr(allow), stream_ptr()), "attn_build_mask")
    _count()
    return allow


def attention_fwd(qkv, valid, rel, dims, spatial, quad_mask, p, drop, allow=None, q_begin=0, keep=None):
    B, L, H, T, A, D = dims
    if uses_tensor_core_attention(qkv.dtype):
        if allow is None:
            allow = build_attn_mask(valid, rel, dims, spatial, quad_mask)
        if p > 0 and keep is None:
            keep = build_attn_keep(dims, p, drop, qkv.device)
    ctx_t = torch.empty(B * L, H * 64, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device)
    ap = _attn_params(qkv, ctx_t, lse, valid, rel, dims, spatial, quad_mask, p, drop, allow=allow, keep=keep)
    ap.q_begin = int(q_begin)
    if attn_profile is not None:
        ev0, ev1 = _prof_events()
        ev0.record()
    if launch_log is not None:
        launch_log.append(("attn_fwd", int(L)))
    check(lib().samk_attn_fwd(ctypes.byref(ap), _ATTN_IMPL, stream_ptr()), "attn_fwd")
    if attn_profile is not None:
        ev1.record()
        # Q+K+V+O 16-bit, allow bits (+ dropout keep bits), LSE; dense-equivalent 4 L^2 d FLOP
        mask_b = B * ((H if spatial else 1) + (H if keep is not None else 0)) * L * ((L + 31) // 32) * 4
        attn_profile.append(("fwd", L, ev0, ev1, B * (4 * L * H * 64 * 2 + H * L * 4) + mask_b, 4.0 * B * L * L * H * 64))
    _count()
    return ctx_t, lse


def attention_bwd_workspaces(qkv, lse, dims):
    """(do16, inv_scale, delta) of the tensor-core backward, or None when the path has no separate preparation"""
    B, L, H, T, A, D = dims
    if not uses_tensor_core_attention(qkv.dtype):
        return None
    return (torch.empty(B * L, H * 64, dtype=torch.float16, device=qkv.device),
            torch.empty(B * H, dtype=torch.float32, device=qkv.device), torch.empty_like(lse))


def attention_bwd_prepare(dctx, qkv, ctx_t, lse, dims, ws):
    """First kernel of the tensor-core backward on its own (dO re-expressed in half with one exact scale per (sample,
    head), delta = rowsum(dO . O)): needs only dctx and ctx, so a layer's backward issues it on the side branch beside
    the weight-gradient product that sits between the out-projection dgrad and the attention backward.  Returns the
    ws: attention_bwd_workspaces(...) (allocated on the caller's stream); hand it to attention_bwd(prepared=ws)."""
    B, L, H, T, A, D = dims
    do16, inv_scale, delta = ws
    ap = _lib.AttnParams()
    ap.qkv, ap.ctx, ap.lse = qkv.data_ptr(), ctx_t.data_ptr(), lse.data_ptr()
    ap.dtype, ap.grad_dtype = _dt(qkv), _dt(dctx)
    ap.dctx, ap.dqkv, ap.delta = dctx.data_ptr(), dctx.data_ptr(), delta.data_ptr()     # (dqkv is not touched in this phase)
    ap.B, ap.H, ap.head_dim, ap.T, ap.A, ap.D = B, H, 64, T, A, D
    ap.scale = 1.0 / math.sqrt(64.0)
    ap.do_f16, ap.do_inv_scale = do16.data_ptr(), inv_scale.data_ptr()
    ap.bwd_phase = 2
    check(lib().samk_attn_bwd(ctypes.byref(ap), _ATTN_IMPL, stream_ptr()), "attn_bwd(prepare)")
    return ws


def attention_bwd(dctx, qkv, ctx_t, lse, valid, rel, dims, spatial, quad_mask, p, drop, allow=None, keep=None, prepared=None):
    """dctx [B*L, H*64] (bf16 in the product mode) -> dq|dk|dv [B*L, 3*H*64] in the same dtype."""
    B, L, H, T, A, D = dims
    dq_accum = do16 = inv_scale = None
    tc = uses_tensor_core_attention(qkv.dtype)
    delta = None
    if tc:
        if allow is None:
            allow = build_attn_mask(valid, rel, dims, spatial, quad_mask)
        if p > 0 and keep is None:
            keep = build_attn_keep(dims, p, drop, qkv.device)
        if L > 256:      # long sequences: key-tile CTAs reduce dQ through an fp32 buffer
            dq_accum = torch.empty(B * L, H * 64, dtype=torch.float32, device=qkv.device)
        if prepared is not None:
            do16, inv_scale, delta = prepared
        else:
            do16 = torch.empty(B * L, H * 64, dtype=torch.float16, device=qkv.device)     # dO in half, one scale per (b, h)
            inv_scale = torch.empty(B * H, dtype=torch.float32, device=qkv.device)
    dqkv = torch.empty(qkv.shape, dtype=dctx.dtype, device=qkv.device)
    if delta is None:
        delta = torch.empty_like(lse)
    ap = _attn_params(qkv, ctx_t, lse, valid, rel, dims, spatial, quad_mask, p, drop, dctx, dqkv, delta,
                      allow=allow, dq_accum=dq_accum, keep=keep)
    if tc:
        ap.do_f16, ap.do_inv_scale = do16.data_ptr(), inv_scale.data_ptr()
        ap.bwd_phase = 1 if prepared is not None else 0
    if attn_profile is not None:
        ev0, ev1 = _prof_events()
        ev0.record()
    if launch_log is not None:
        launch_log.append(("attn_bwd", int(L)))
    check(lib().samk_attn_bwd(ctypes.byref(ap), _ATTN_IMPL, stream_ptr()), "attn_bwd")
    if attn_profile is not None:
        ev1.record()
        # Q,K,V,O,dO read + dQ,dK,dV written (16-bit), allow (+ keep) bits, LSE + delta; 10 L^2 d FLOP
        mask_b = B * ((H if spatial else 1) + (H if keep is not None else 0)) * L * ((L + 31) // 32) * 4
        attn_profile.append(("bwd", L, ev0, ev1, B * (8 * L * H * 64 * 2 + 2 * H * L * 4) + mask_b, 10.0 * B * L * L * H * 64))
    _count(4 if dq_accum is not None else 2)   # prep (+ memset) + main kernel (+ dq conversion)
    return dqkv


def zeros(shape, dtype, device):
    """zero-filled tensor through a stream-ordered memset (a memset node in a captured step, not an aten fill kernel)"""
    t = torch.empty(shape, dtype=dtype, device=device)
    check(lib().samk_memset0(ptr(t), t.numel() * t.element_size(), stream_ptr()), "memset0")
    return t


def zero_(t):
    assert t.is_contiguous()
    check(lib().samk_memset0(ptr(t), t.numel() * t.element_size(), stream_ptr()), "memset0")
    return t


def _row_segments(joint, segs, offs, to_joint):
    """segs: contiguous fp32 [B, rows_k, d] tensors; offs: first row of each inside joint [B, L, d]"""
    B, L, d = joint.shape
    n = len(segs)
    ptrs = (ctypes.c_void_p * 4)(*([t.data_ptr() for t in segs] + [None] * (4 - n)))
    rows = (ctypes.c_int * 4)(*([t.shape[1] for t in segs] + [0] * (4 - n)))
    offa = (ctypes.c_int * 4)(*(list(offs) + [0] * (4 - n)))
    check(lib().samk_row_segments_f32(ptr(joint), ptrs, rows, offa, n, B, L, d, 1 if to_joint else 0, stream_ptr()), "row_segments")
    _count()


class JoinFn(torch.autograd.Function):
    """[txt ; obj ; ocr ; dec] along the token axis (MMT.forward, sa_m4c.py:790) in one launch; the backward hands every
    segment its contiguous gradient in one launch (instead of four strided slice copies)."""

    @staticmethod
    def forward(ctx, *segs):
        _cuda(*segs)
        segs = [t.float().contiguous() for t in segs]
        B, d = segs[0].shape[0], segs[0].shape[2]
        offs, L = [], 0
        for t in segs:
            offs.append(L)
            L += t.shape[1]
        out = torch.empty(B, L, d, dtype=torch.float32, device=segs[0].device)
        _row_segments(out, segs, offs, True)
        ctx.meta = ([t.shape[1] for t in segs], offs)
        return out

    @staticmethod
    def backward(ctx, dout):
        rows, offs = ctx.meta
        dout = dout.contiguous()
        B, L, d = dout.shape
        outs = [torch.empty(B, r, d, dtype=torch.float32, device=dout.device) for r in rows]
        _row_segments(dout, outs, offs, False)
        return tuple(outs)


def join_segments(*segs):
    return JoinFn.apply(*segs)


def key_valid_bytes(q_mask, obj_mask, ocr_mask, D):
    """uint8 [B, T+O+R+D] key-valid map of MMT.forward (sa_m4c.py:793-795) from the int64 masks of the batch"""
    def m(t):
        return None if t is None else t.long().contiguous()
    q_mask, obj_mask, ocr_mask = m(q_mask), m(obj_mask), m(ocr_mask)
    _cuda(q_mask, obj_mask, ocr_mask)
    B, T = q_mask.shape
    O = obj_mask.shape[1] if obj_mask is not None else 0
    R = ocr_mask.shape[1] if ocr_mask is not None else 0
    out = torch.empty(B, T + O + R + D, dtype=torch.uint8, device=q_mask.device)
    check(lib().samk_key_valid(ptr(q_mask), ptr(obj_mask), ptr(ocr_mask), ptr(out), B, T, O, R, D, stream_ptr()), "key_valid")
    _count()
    return out


def _bias3(qb, kb, vb):
    """q|k|v biases as one [3d] vector for the fused projection (concatenated by one of our kernels, not aten::cat)."""
    d = qb.shape[0]
    out = torch.empty(3 * d, dtype=torch.float32, device=qb.device)
    check(lib().samk_concat3_f32(ptr(qb), ptr(kb), ptr(vb), ptr(out), d, stream_ptr()), "concat3")
    _count()
    return out


# ---- one post-LN BERT block (plain or spatial) -------------------------------------------------------------
class BertLayerFn(torch.autograd.Function):
    """BertLayer / SpatialBertLayer (sa_m4c.py:660-684 and the third-party BertLayer):
       a = LN(dropout(W_o . attn(x) + b_o) + x);  out = LN(dropout(W_2 . gelu(W_1 a + b_1) + b_2) + a)

    params order: q.w q.b k.w k.b v.w v.b  o.w o.b ln1.g ln1.b  i.w i.b  o2.w o2.b ln2.g ln2.b
    """

    @staticmethod
    def forward(ctx, x, valid, rel, cfg, *P):
        (dims, spatial, quad_mask, p_attn, p_hid, eps, mask_cache) = cfg
        B, L, H, T, A, D = dims
        qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, o2w, o2b, g2, b2 = P
        _cuda(x, valid, qw)
        dev = x.device
        adt = act_dtype()
        M, d, F = B * L, x.shape[-1], iw.shape[0]
        x2 = x.contiguous().view(M, d)
        drops = [_next_drop() if pp > 0 else (0, 0) for pp in (p_attn, p_hid, p_hid)]
        want_bf = adt == torch.float16 and any(ctx.needs_input_grad)    # bf16 copies for the weight-gradient products
        x_bf = cached_copy(x2, torch.bfloat16) if want_bf else None     # (written by the LayerNorm that produced x)

        x_op = operand(x2, "a", False)
        wqkv = weight_operand([qw, kw, vw], False)
        bqkv = _bias3(qb, kb, vb)
        qkv = torch.empty(M, 3 * d, dtype=adt, device=dev)
        keep = None
        if uses_tensor_core_attention(adt) and p_attn > 0:
            # the dropout keep bits of this layer's attention are drawn beside the tensor-bound q|k|v projection
            keep = torch.empty(B * H * L * ((L + 31) // 32), dtype=torch.int32, device=dev)
            with side_branch():
                fill_attn_keep(keep, dims, p_attn, drops[0])
        gemm(x_op, False, wqkv, False, M, 3 * d, d, qkv, bias=bqkv)
        if keep is not None:
            join_side()
        allow = None
        if uses_tensor_core_attention(adt):
            mkey = (bool(spatial), quad_mask if spatial else 0, rel.data_ptr() if (spatial and rel is not None) else 0, dims)
            allow = mask_cache.get(mkey) if mask_cache is not None else None
            if allow is None:
                allow = build_attn_mask(valid, rel, dims, spatial, quad_mask)
                if mask_cache is not None:
                    mask_cache[mkey] = allow
        ctx_t, lse = attention_fwd(qkv, valid, rel, dims, spatial, quad_mask, p_attn, drops[0], allow, keep=keep)

        y1 = torch.empty(M, d, dtype=torch.float32, device=dev)
        gemm(operand(ctx_t, "a", False), False, weight_operand([ow], False), False, M, d, d, y1, bias=ob,
             drop_p=p_hid, drop=drops[1], residual=x2)
        a = torch.empty(M, d, dtype=torch.float32, device=dev)
        a_act = torch.empty(M, d, dtype=adt, device=dev) if adt != torch.float32 else None
        a_bf = torch.empty(M, d, dtype=torch.bfloat16, device=dev) if want_bf else None
        check(lib().samk_layernorm_fwd(ptr(y1), ptr(g1), ptr(b1), eps, ptr(a), ptr(a_act), _DT[adt], ptr(a_bf), DT_BF16, M, d,
                                       stream_ptr()), "ln1")
        _count()
        a_in = a_act if a_act is not None else a
        h = torch.empty(M, F, dtype=grad_dtype(), device=dev)     # gelu'(pre-activation): read by the FFN2 dgrad epilogue only
        g = torch.empty(M, F, dtype=adt, device=dev)
        gemm(operand(a_in, "a", False), False, weight_operand([iw], False), False, M, F, d, g, bias=ib, act=3, pre=h)
        y2 = torch.empty(M, d, dtype=torch.float32, device=dev)
        gemm(operand(g, "a", False), False, weight_operand([o2w], False), False, M, d, F, y2, bias=o2b, drop_p=p_hid,
             drop=drops[2], residual=a)
        out = torch.empty(M, d, dtype=torch.float32, device=dev)
        out_act = torch.empty(M, d, dtype=adt, device=dev) if adt != torch.float32 else None
        out_bf = torch.empty(M, d, dtype=torch.bfloat16, device=dev) if want_bf else None
        check(lib().samk_layernorm_fwd(ptr(y2), ptr(g2), ptr(b2), eps, ptr(out), ptr(out_act), _DT[adt], ptr(out_bf), DT_BF16,
                                       M, d, stream_ptr()), "ln2")
        _count()
        if out_act is not None:
            remember_act(out, out_act, out_bf)   # the next layer's q|k|v projection (half) and its wgrad (bf16) read these

        ctx.cfg, ctx.drops = cfg[:6], drops
        ctx.save_for_backward(x2, valid, rel, qkv, ctx_t, lse, y1, a_in, h, g, y2, allow, keep, x_bf, a_bf, *P)
        return out.view(B, L, d)

    @staticmethod
    def backward(ctx, dout):
        (dims, spatial, quad_mask, p_attn, p_hid, eps) = ctx.cfg
        B, L, H, T, A, D = dims
        x2, valid, rel, qkv, ctx_t, lse, y1, a_in, h, g, y2, allow, keep, x_bf, a_bf = ctx.saved_tensors[:15]
        P = ctx.saved_tensors[15:]
        qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, o2w, o2b, g2, b2 = P
        G = [_gbuf(p) for p in P]            # same order as P
        (Gqw, Gqb, Gkw, Gkb, Gvw, Gvb, Gow, Gob, Gg1, Gb1, Giw, Gib, Go2w, Go2b, Gg2, Gb2) = [x[0] for x in G]
        dev = dout.device
        adt, gdt = act_dtype(), grad_dtype()
        M, d, F = B * L, x2.shape[1], iw.shape[0]
        dout = dout.contiguous().view(M, d)

        # Product mode: dY tensors are bf16.  dgrad = bf16 dY x bf16 weight copy; wgrad = bf16 dY x bf16 activation copy
        # (layer input, LN1 output: written by the LayerNorms) or, for the big saved activations that exist in half only
        # (GELU output, attention context), half(dY * S) x half activation with alpha = 1/S from the device.
        half = adt == torch.float16
        wfmt = "bf16"
        # ---- LN2 backward (+ dropout mask of the FFN output, + b_2 gradient)
        dy2 = torch.empty(M, d, dtype=torch.float32, device=dev)      # grad wrt (dropout(dense)+a)
        dY2 = torch.empty(M, d, dtype=gdt, device=dev)                 # grad wrt dense output
        amax2 = torch.empty(2, dtype=torch.float32, device=dev) if half else None     # max|dY2|, max|dY1|
        # (side branch, beside the dgrad GEMM below: the reduction of the LayerNorm parameter gradients and the scaled
        #  half copy of dY2 that the weight-gradient product reads)
        _ln_bwd(dout, y2, g2, eps, dy2, dY2, p_hid, ctx.drops[2], Gg2, Gb2, Go2b, M, d, amax=amax2, side_final=bool(_SIDE_OVERLAP & 1))
        if half and (_SIDE_OVERLAP & 1):
            out_h = (torch.empty(M, d, dtype=torch.float16, device=dev), torch.empty(4, dtype=torch.float32, device=dev))
            with side_branch():
                dY2h, inv2 = scaled_f16(dY2, amax2, out=out_h)
        # ---- FFN2: dgrad (fused with the stored GELU') and wgrad
        dh = torch.empty(M, F, dtype=gdt, device=dev)
        gemm(operand(dY2, "a", False, fmt=wfmt), False, weight_operand([o2w], True, fmt=wfmt), True, M, F, d, dh, act=4, aux=h)
        if _SIDE_OVERLAP & 1:
            join_side()
        with side_branch():                       # b_1 gradient beside the GEMMs that follow
            colsum_into(dh, Gib)
        if half:
            if not (_SIDE_OVERLAP & 1):
                dY2h, inv2 = scaled_f16(dY2, amax2)
            gemm(dY2h, True, operand(g, "b", True), True, d, F, M, Go2w, accumulate=True, alpha_dev=inv2)
        else:
            gemm(operand(dY2, "a", True), True, operand(g, "b", True), True, d, F, M, Go2w, accumulate=True)
        # ---- FFN1
        da = torch.empty(M, d, dtype=torch.float32, device=dev)
        gemm(operand(dh, "a", False, fmt=wfmt), False, weight_operand([iw], True, fmt=wfmt), True, M, d, F, da, residual=dy2)
        a_w = a_bf if half else a_in
        gemm(operand(dh, "a", True, fmt=wfmt), True, operand(a_w, "b", True), True, F, d, M, Giw, accumulate=True)
        # ---- LN1 backward (+ dropout mask of the attention output dense, + b_o gradient)
        dy1 = torch.empty(M, d, dtype=torch.float32, device=dev)
        dY1 = torch.empty(M, d, dtype=gdt, device=dev)
        _ln_bwd(da, y1, g1, eps, dy1, dY1, p_hid, ctx.drops[1], Gg1, Gb1, Gob, M, d, amax=amax2[1:] if half else None,
                side_final=bool(_SIDE_OVERLAP & 2))
        if half and (_SIDE_OVERLAP & 2):
            out_h = (torch.empty(M, d, dtype=torch.float16, device=dev), torch.empty(4, dtype=torch.float32, device=dev))
            with side_branch():
                dY1h, inv1 = scaled_f16(dY1, amax2[1:], out=out_h)
        # ---- attention output dense
        dctx = torch.empty(M, d, dtype=gdt, device=dev)
        gemm(operand(dY1, "a", False, fmt=wfmt), False, weight_operand([ow], True, fmt=wfmt), True, M, d, d, dctx)
        prepared = None
        if _SIDE_OVERLAP & 2:
            join_side()
        if _SIDE_OVERLAP & 4:
            prepared = attention_bwd_workspaces(qkv, lse, dims)
            if prepared is not None:
                with side_branch():               # attention-backward preparation beside the weight-gradient product
                    attention_bwd_prepare(dctx, qkv, ctx_t, lse, dims, prepared)
        if half:
            if not (_SIDE_OVERLAP & 2):
                dY1h, inv1 = scaled_f16(dY1, amax2[1:])
            gemm(dY1h, True, operand(ctx_t, "b", True), True, d, d, M, Gow, accumulate=True, alpha_dev=inv1)
        else:
            gemm(operand(dY1, "a", True), True, operand(ctx_t, "b", True), True, d, d, M, Gow, accumulate=True)
        if _SIDE_OVERLAP & 4:
            join_side()
        # ---- attention core
        dqkv = attention_bwd(dctx, qkv, ctx_t, lse, valid, rel, dims, spatial, quad_mask, p_attn, ctx.drops[0], allow,
                             keep=keep, prepared=prepared)
        # ---- fused q|k|v projection: one dgrad, three wgrads (separate parameter gradients)
        with side_branch():                       # q, k, v bias gradients beside the dgrad / wgrad below
            check(lib().samk_colsum3(ptr(dqkv), _dt(dqkv), dqkv.stride(0), M, d, ptr(Gqb), ptr(Gkb), ptr(Gvb),
                                     stream_ptr()), "colsum3")
            _count()
        dx = torch.empty(M, d, dtype=torch.float32, device=dev)
        gemm(operand(dqkv, "a", False, fmt=wfmt), False, weight_operand([qw, kw, vw], True, fmt=wfmt), True, M, d, 3 * d, dx,
             residual=dy1)
        x_mn = operand(x_bf if x_bf is not None else x2, "b", True, fmt=wfmt)
        if Gqw.stride(0) == Gkw.stride(0) == Gvw.stride(0) and d % 32 == 0:
            # the three weight gradients in one launch: M = 3d output rows routed to three destinations
            gemm(operand(dqkv, "a", True, fmt=wfmt), True, x_mn, True, 3 * d, d, M, Gqw, accumulate=True, out_parts=(d, Gkw, Gvw))
        else:
            for part, Gw in enumerate((Gqw, Gkw, Gvw)):
                sl = dqkv[:, part * d:(part + 1) * d]
                gemm(operand(sl, "a", True, fmt=wfmt), True, x_mn, True, d, d, M, Gw, accumulate=True)
        join_side()
        _grads_done(*P)
        return (dx.view(B, L, d), None, None, None) + tuple(_ret(x) for x in G)


def bert_layer_infer(x2, valid, rel, dims, spatial, quad_mask, eps, P, allow, cache=None):
    """Inference-only (no autograd, no dropout) forward of one post-LN block, same kernels as BertLayerFn.

    cache is None : x2 = all rows [B*L, d]; returns (out [B*L, d], {"qkv": fused q|k|v of every row}).
    cache given   : x2 = the D decoder rows of every sample [B*D, d]; only those rows are recomputed: their
                    q|k|v replace the decoder rows of cache["qkv"] (encoder rows never depend on the decoder,
                    sa_m4c.py:793-795), attention runs for the query tiles that hold decoder rows, and the
                    dense layers see B*D rows instead of B*L.  Returns (out [B*D, d], cache)."""
    B, L, H, T, A, D = dims
    qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, o2w, o2b, g2, b2 = P
    dev, adt = x2.device, act_dtype()
    d, F = x2.shape[1], iw.shape[0]
    rows = x2.shape[0]
    wqkv = weight_operand([qw, kw, vw], False)
    bqkv = _bias3(qb, kb, vb)
    qkv_rows = torch.empty(rows, 3 * d, dtype=adt, device=dev)
    gemm(operand(x2, "a", False), False, wqkv, False, rows, 3 * d, d, qkv_rows, bias=bqkv)
    if cache is None:
        qkv = qkv_rows
        ctx_t, _ = attention_fwd(qkv, valid, rel, dims, spatial, quad_mask, 0.0, (0, 0), allow)
        ctx_rows = ctx_t
        cache = {"qkv": qkv}
    else:
        qkv = cache["qkv"]
        qkv.view(B, L, 3 * d)[:, L - D:, :] = qkv_rows.view(B, D, 3 * d)
        ctx_t, _ = attention_fwd(qkv, valid, rel, dims, spatial, quad_mask, 0.0, (0, 0), allow, q_begin=L - D)
        ctx_rows = ctx_t.view(B, L, d)[:, L - D:, :].reshape(rows, d)
    y1 = torch.empty(rows, d, dtype=torch.float32, device=dev)
    gemm(operand(ctx_rows, "a", False), False, weight_operand([ow], False), False, rows, d, d, y1, bias=ob, residual=x2)
    a = torch.empty(rows, d, dtype=torch.float32, device=dev)
    a_act = torch.empty(rows, d, dtype=adt, device=dev) if adt != torch.float32 else None
    check(lib().samk_layernorm_fwd(ptr(y1), ptr(g1), ptr(b1), eps, ptr(a), ptr(a_act), _DT[adt], None, 0, rows, d,
                                   stream_ptr()), "ln1")
    _count()
    g = torch.empty(rows, F, dtype=adt, device=dev)
    gemm(operand(a_act if a_act is not None else a, "a", False), False, weight_operand([iw], False), False, rows, F, d,
         g, bias=ib, act=1)
    y2 = torch.empty(rows, d, dtype=torch.float32, device=dev)
    gemm(operand(g, "a", False), False, weight_operand([o2w], False), False, rows, d, F, y2, bias=o2b, residual=a)
    out = torch.empty(rows, d, dtype=torch.float32, device=dev)
    check(lib().samk_layernorm_fwd(ptr(y2), ptr(g2), ptr(b2), eps, ptr(out), None, 0, None, 0, rows, d, stream_ptr()), "ln2")
    _count()
    return out, cache


# ---- output heads + loss -------------------------------------------------------------------------------------
class OutputFn(torch.autograd.Function):
    """scores = cat[classifier(dec), OcrPtrNet(dec, ocr, mask)] written into one [B,D,V+R] buffer
    (sa_m4c.py:270-278, 878-897).  Takes the whole MMT output [B,L,d] and the row ranges of its OCR / decoder
    segments (sa_m4c.py:852-862): the rows are gathered, and their gradients scattered into the gradient of the whole
    sequence, by one launch each (no strided slice copies, no gradient adds on the autograd side)."""

    @staticmethod
    def forward(ctx, seq, ocr_off, R, D, ocr_mask, cw, cb, qw, qb, kw, kb):
        _cuda(seq, cw)
        seq = seq.float().contiguous()
        B, L, d = seq.shape
        V, dq = cw.shape[0], qw.shape[0]
        dev = seq.device
        ocr = torch.empty(B, R, d, dtype=torch.float32, device=dev)
        dec = torch.empty(B, D, d, dtype=torch.float32, device=dev)
        _row_segments(seq, [ocr, dec], [ocr_off, L - D], False)
        dec2 = dec.view(B * D, d)
        ocr2 = ocr.view(B * R, d)
        ocr_mask = ocr_mask.long().contiguous()      # any mask dtype (the reference multiplies it as a float, sa_m4c.py:879)
        scores = torch.empty(B * D, V + R, dtype=torch.float32, device=dev)
        # the vocabulary and pointer projections run as 3-term splits in every precision mode: they are 0.3 % of the
        # FLOPs and their operand rounding is the largest single term of the logit error (DESIGN.md error budget)
        dec_op = operand(dec2, "a", False, split=True)
        gemm(dec_op, False, weight_operand([cw], False, split=True), False, B * D, V, d, scores[:, :V], bias=cb)
        q = torch.empty(B * D, dq, dtype=torch.float32, device=dev)
        k = torch.empty(B * R, dq, dtype=torch.float32, device=dev)
        gemm(dec_op, False, weight_operand([qw], False, split=True), False, B * D, dq, d, q, bias=qb)
        gemm(operand(ocr2, "a", False, split=True), False, weight_operand([kw], False, split=True), False, B * R, dq, d, k, bias=kb)
        check(lib().samk_ptr_scores_fwd(ptr(q), ptr(k), ptr(ocr_mask), ptr(scores), V + R, V, B, D, R, dq,
                                        stream_ptr()), "ptr_scores_fwd")
        _count()
        ctx.dims = (B, D, R, V, d, dq)
        ctx.seq_meta = (L, ocr_off)
        ctx.bias_refs = (cb, qb, kb)
        ctx.save_for_backward(dec2, ocr2, q, k, cw, qw, kw)
        return scores.view(B, D, V + R)

    @staticmethod
    def backward(ctx, ds):
        dec2, ocr2, q, k, cw, qw, kw = ctx.saved_tensors
        cb, qb, kb = ctx.bias_refs
        B, D, R, V, d, dq = ctx.dims
        dev = ds.device
        ds = ds.contiguous().view(B * D, V + R)
        G = [_gbuf(p) for p in (cw, cb, qw, qb, kw, kb)]
        Gcw, Gcb, Gqw, Gqb, Gkw, Gkb = [x[0] for x in G]
        dq_ = torch.empty_like(q)
        dk_ = torch.empty_like(k)
        check(lib().samk_ptr_scores_bwd(ptr(ds), V + R, V, ptr(q), ptr(k), ptr(dq_), ptr(dk_), B, D, R, dq,
                                        stream_ptr()), "ptr_scores_bwd")
        _count()
        dsv = ds[:, :V]
        if _PRECISION == "f16":
            dsv = cast_16(dsv, torch.bfloat16)[:, :V]
            dq_a, dk_a = cast_16(dq_, torch.bfloat16), cast_16(dk_, torch.bfloat16)
        else:
            dq_a, dk_a = dq_, dk_
        dec_mn, ocr_mn = operand(dec2, "b", True, fmt="bf16"), operand(ocr2, "b", True, fmt="bf16")
        # classifier
        ddec = torch.empty(B * D, d, dtype=torch.float32, device=dev)
        gemm(operand(dsv, "a", False, fmt="bf16"), False, weight_operand([cw], True, fmt="bf16"), True, B * D, d, V, ddec)
        gemm(operand(dsv, "a", True, fmt="bf16"), True, dec_mn, True, V, d, B * D, Gcw, accumulate=True)
        colsum_into(dsv if dsv.dtype != torch.float32 else ds[:, :V], Gcb)      # (the bf16 copy: vector loads, half the bytes)
        # pointer query / key projections
        ddec2 = torch.empty(B * D, d, dtype=torch.float32, device=dev)
        gemm(operand(dq_a, "a", False, fmt="bf16"), False, weight_operand([qw], True, fmt="bf16"), True, B * D, d, dq, ddec2, residual=ddec)
        gemm(operand(dq_a, "a", True, fmt="bf16"), True, dec_mn, True, dq, d, B * D, Gqw, accumulate=True)
        colsum_into(dq_, Gqb)
        docr = torch.empty(B * R, d, dtype=torch.float32, device=dev)
        gemm(operand(dk_a, "a", False, fmt="bf16"), False, weight_operand([kw], True, fmt="bf16"), True, B * R, d, dq, docr)
        gemm(operand(dk_a, "a", True, fmt="bf16"), True, ocr_mn, True, dq, d, B * R, Gkw, accumulate=True)
        colsum_into(dk_, Gkb)
        _grads_done(cw, cb, qw, qb, kw, kb)
        L, ocr_off = ctx.seq_meta
        dseq = zeros((B, L, d), torch.float32, dev)
        _row_segments(dseq, [docr.view(B, R, d), ddec2.view(B, D, d)], [ocr_off, L - D], True)
        return (dseq, None, None, None, None) + tuple(_ret(x) for x in G)


class BceLossFn(torch.autograd.Function):
    """M4CDecodingBCEWithMaskLoss (sam/task_utils.py:19-30): forward value and gradient in one pass."""

    @staticmethod
    def forward(ctx, scores, targets, loss_mask):
        _cuda(scores, targets, loss_mask)
        scores = scores.float().contiguous()
        targets = targets.float().contiguous()
        loss_mask = loss_mask.contiguous().float()
        rows, ncls = scores.numel() // scores.shape[-1], scores.shape[-1]
        loss = torch.empty(2, dtype=torch.float32, device=scores.device)
        ds = torch.empty_like(scores)
        check(lib().samk_bce_loss(ptr(scores), ptr(targets), ptr(loss_mask), ptr(ds), ptr(loss), ptr(loss[1:]), rows,
                                  ncls, stream_ptr()), "bce_loss")
        _count(2)
        ctx.save_for_backward(ds)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        (ds,) = ctx.saved_tensors
        g = gout.reshape(1).float().contiguous()
        check(lib().samk_scale_inplace(ptr(ds), ds.numel(), ptr(g), stream_ptr()), "scale")
        _count()
        return ds, None, None


def argmax_rows(scores, targets=None):
    """Greedy tokens on the device: scores [..., ncls] fp32 -> int64 [...] (first maximum, like torch.argmax on distinct
    values); with `targets` also returns hit[...] = targets[..., argmax] (a token-level accuracy count that never moves
    the [B, D, V+R] logits to the host)."""
    _cuda(scores, targets)
    scores = scores.float().contiguous()
    ncls = scores.shape[-1]
    rows = scores.numel() // ncls
    idx = torch.empty(scores.shape[:-1], dtype=torch.long, device=scores.device)
    hit = None
    if targets is not None:
        targets = targets.float().contiguous()
        hit = torch.empty(scores.shape[:-1], dtype=torch.float32, device=scores.device)
    check(lib().samk_argmax_rows(ptr(scores), ncls, rows, ncls, ptr(targets), ncls, ptr(idx), ptr(hit), stream_ptr()), "argmax_rows")
    _count()
    return idx if targets is None else (idx, hit)


def beam_step(scores_t, row_stride, beam_scores, completed, eos, first_step, B, K):
    """One beam-search step (samk_beam_step); scores_t = scores[:, t, :] view of the contiguous [B K, D, ncls] logits."""
    _cuda(scores_t, beam_scores)
    ncls = scores_t.shape[-1]
    dev = scores_t.device
    prev_pos = torch.empty(B * K, dtype=torch.long, device=dev)
    new_pos = torch.empty(B * K, dtype=torch.long, device=dev)
    new_scores = torch.empty(B * K, dtype=torch.float32, device=dev)
    check(lib().samk_beam_step(ptr(scores_t), row_stride, ncls, ptr(beam_scores), ptr(completed), int(eos), 1 if first_step else 0,
                               B, K, ptr(prev_pos), ptr(new_pos), ptr(new_scores), stream_ptr()), "beam_step")
    _count()
    return prev_pos, new_pos, new_scores


def bce_with_mask_loss(scores, targets, loss_mask):
    return BceLossFn.apply(scores, targets, loss_mask)
