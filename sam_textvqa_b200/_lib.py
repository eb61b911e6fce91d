"""ctypes binding of libsamk.so (the C ABI declared in include/samk.h).

The product path has no fallback: if the shared object is missing or a call fails, this
raises.  `lib()` loads lazily so that CPU-only tooling (config, synthetic data, host logic
tests) can import the package on a machine where the library has not been built yet.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAMK_LIB") or os.path.join(HERE, "libsamk.so")     # SAMK_LIB: an instrumented developer build

c_void_p, c_int, c_ll, c_float, c_double, c_ull = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                                   ctypes.c_float, ctypes.c_double, ctypes.c_ulonglong)

DT_F32, DT_BF16, DT_F16 = 0, 1, 2


class GemmEpilogue(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p), ("ldo", c_ll), ("out_dtype", c_int), ("atomic_add", c_int), ("alpha", c_float),
        ("bias", c_void_p), ("pre", c_void_p), ("ldpre", c_ll), ("pre_dtype", c_int), ("act", c_int),
        ("aux", c_void_p), ("ldaux", c_ll), ("aux_dtype", c_int), ("drop_p", c_float),
        ("drop_seed", c_ull), ("drop_offset", c_ull), ("residual", c_void_p), ("ldres", c_ll),
        ("part_rows", c_int), ("out_part1", c_void_p), ("out_part2", c_void_p), ("alpha_dev", c_void_p),
    ]


class AttnParams(ctypes.Structure):
    _fields_ = [
        ("qkv", c_void_p), ("ctx", c_void_p), ("lse", c_void_p), ("dctx", c_void_p), ("dqkv", c_void_p),
        ("delta", c_void_p), ("dtype", c_int), ("grad_dtype", c_int), ("B", c_int), ("H", c_int), ("head_dim", c_int), ("T", c_int),
        ("A", c_int), ("D", c_int), ("key_valid", c_void_p), ("rel_bits", c_void_p), ("quadrant_mask", ctypes.c_uint),
        ("spatial", c_int), ("scale", c_float), ("drop_p", c_float), ("drop_seed", c_ull), ("drop_offset", c_ull),
        ("allow_bits", c_void_p), ("dq_accum", c_void_p), ("q_begin", c_int), ("bwd_phase", c_int),
        ("keep_bits", c_void_p), ("do_f16", c_void_p), ("do_inv_scale", c_void_p),
    ]


class PeerWire(ctypes.Structure):
    _fields_ = [("rank", c_int), ("world", c_int), ("wire_dtype", c_int), ("wire_peers", ctypes.POINTER(c_void_p)),
                ("wire_mc", c_void_p), ("flag_peers", ctypes.POINTER(c_void_p)), ("epoch", c_void_p), ("error", c_void_p),
                ("timeout_clocks", c_ll)]


# name -> (restype, argtypes); must list every symbol include/samk.h declares
SIGNATURES = {
    "samk_version": (c_int, []),
    "samk_last_error": (ctypes.c_char_p, []),
    "samk_sm_count": (c_int, []),
    "samk_reserve_sms": (c_int, [c_int]),
    "samk_set_dropout_salt": (c_int, [c_ull, c_void_p]),
    "samk_advance_dropout_salt": (c_int, [c_void_p]),
    "samk_build_graph_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p, c_void_p]),
    "samk_build_graph_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p, c_void_p]),
    "samk_graph_default_sectors": (ctypes.POINTER(c_double), []),
    "samk_pack_adj": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "samk_unpack_bits": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "samk_types_to_bits": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "samk_gemm_bf16": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int, c_int,
                               ctypes.POINTER(GemmEpilogue), c_int, c_int, c_void_p]),
    "samk_gemm_16": (c_int, [c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int,
                             ctypes.POINTER(GemmEpilogue), c_int, c_int, c_void_p]),
    "samk_cast_bf16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "samk_cast_16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "samk_cast_dual": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "samk_cast_flat": (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_void_p]),
    "samk_exchange_sum": (c_int, [ctypes.POINTER(PeerWire), c_void_p, c_ll, c_ll, c_void_p]),
    "samk_timing_event_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "samk_timing_event_record": (c_int, [c_void_p, c_void_p]),
    "samk_timing_event_elapsed_ms": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_float)]),
    "samk_timing_event_destroy": (c_int, [c_void_p]),
    "samk_cast_scaled_f16": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "samk_split3_bf16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_l2norm": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_l2norm2": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                   c_int, c_void_p]),
    "samk_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_float, c_ull,
                                   c_ull, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "samk_layernorm_bwd_partials": (c_ll, [c_int]),
    "samk_layernorm_bwd_main": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_float, c_ull,
                                        c_ull, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "samk_layernorm_bwd_finalize": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "samk_dropout_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_ull, c_ull, c_void_p]),
    "samk_colsum": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "samk_colsum3": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "samk_bert_embed_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_int, c_int, c_int, c_int, c_float, c_ull, c_ull, c_void_p]),
    "samk_bert_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_ull, c_ull,
                                    c_void_p]),
    "samk_prevpred_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                  c_int, c_int, c_int, c_int, c_int, c_float, c_ull, c_ull, c_void_p]),
    "samk_prevpred_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                  c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_ull, c_ull, c_void_p]),
    "samk_ptr_scores_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_ptr_scores_bwd": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_void_p]),
    "samk_bce_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "samk_scale_inplace": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "samk_row_segments_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_key_valid": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "samk_memset0": (c_int, [c_void_p, c_ll, c_void_p]),
    "samk_argmax_rows": (c_int, [c_void_p, c_ll, c_ll, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "samk_beam_step": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "samk_concat3_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "samk_sumsq": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "samk_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_double, c_double, c_double, c_double, c_int,
                               c_void_p, c_double, c_void_p]),
    "samk_attn_fwd": (c_int, [ctypes.POINTER(AttnParams), c_int, c_void_p]),
    "samk_attn_bwd": (c_int, [ctypes.POINTER(AttnParams), c_int, c_void_p]),
    "samk_attn_build_keep": (c_int, [ctypes.POINTER(AttnParams), c_void_p, c_void_p]),
    "samk_attn_mask_words": (c_ll, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "samk_attn_build_mask": (c_int, [ctypes.POINTER(AttnParams), c_void_p, c_void_p]),
}

_LIB = None


class SamkError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SamkError(
                "libsamk.so not found at %s -- build it with `python -m sam_textvqa_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


_DEBUG_CAPTURE = os.environ.get("SAMK_DEBUG_CAPTURE", "0") != "0"
_cudart = None


def _capture_status():
    """cudaStreamCaptureStatus of torch's current stream (0 none, 1 active, 2 invalidated); debugging aid."""
    global _cudart
    import torch
    if _cudart is None:
        _cudart = ctypes.CDLL("libcudart.so.12")
    st = ctypes.c_int(0)
    _cudart.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return st.value


def check(rc, what=""):
    if rc != 0:
        msg = lib().samk_last_error()
        raise SamkError("%s failed (%d): %s" % (what or "samk call", rc, (msg or b"").decode()))
    if _DEBUG_CAPTURE and _capture_status() == 2:
        raise SamkError("stream capture was invalidated at or before %r" % what)


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
