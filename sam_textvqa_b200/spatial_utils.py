"""Spatial-relation graph on the GPU: mirror of /root/reference/sam/spatial_utils.py.

`build_graph_using_normalized_boxes` and `torch_broadcast_adj_matrix` keep the reference's names,
arguments and return types (a dict of nine int8 numpy matrices; an int8 [N,N,12] tensor), so
`SpatialProcessor` (sam/datasets/processors.py:460-464) and the dataset's context expansion
(sam/datasets/textvqa_dataset.py:373-409) can call them unchanged.  `build_graph_batch` is the
batched on-device form the B200 path itself uses: boxes [B,N,4] -> types / packed head bits without
leaving HBM.  All results are bit-identical to the NumPy reference (tests/test_gpu_graph.py).
"""
import ctypes
import struct

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

SHARED_KEYS = ("1", "31", "32", "51", "52", "71", "72", "91", "92")
_sector_table = None   # None -> the table compiled into libsamk.so


def set_sector_table(table):
    """Override the angle-sector step table (8 rows of base, step, t1, t2); None restores the default."""
    global _sector_table
    if table is None:
        _sector_table = None
    else:
        arr = np.ascontiguousarray(np.asarray(table, dtype=np.float64).reshape(8, 4))
        _sector_table = arr


def default_sector_table():
    p = lib().samk_graph_default_sectors()
    return np.array([p[i] for i in range(32)], dtype=np.float64).reshape(8, 4)


def derive_sector_table_from_numpy():
    """Re-derive the sector table from THIS machine's np.arcsin / np.arccos by bisection over the
    doubles (the reference's angles come from NumPy, sam/spatial_utils.py:172-189, and NumPy's
    SIMD libm differs from glibc/CUDA in the last ulp exactly at octant boundaries)."""
    import math
    PI = math.pi

    def ce(label):
        q = np.ceil(label / (PI / 4))
        return 4 if math.isnan(q) else int(q) + 3

    asin = lambda v: np.arcsin(np.array([v]))[0]
    acos = lambda v: np.arccos(np.array([v]))[0]
    tiny = 5e-324
    funcs = [
        (lambda s: ce(asin(s)), 0.0, 1.0), (lambda s: ce(PI + asin(s)), 0.0, 1.0),
        (lambda s: ce(asin(s) + 2 * PI), -1.0, -tiny), (lambda s: ce((asin(s) + 2 * PI) - PI), -1.0, -tiny),
        (lambda c: ce(acos(c)), -1.0, -tiny), (lambda c: ce(acos(c) + PI), -1.0, -tiny),
        (lambda c: ce(2 * PI - acos(c)), -1.0, -tiny), (lambda c: ce((2 * PI - acos(c)) - PI), -1.0, -tiny),
    ]

    def key(x):
        i = struct.unpack("<q", struct.pack("<d", x))[0]
        return i if i >= 0 else -(i & 0x7FFFFFFFFFFFFFFF)

    def unkey(k):
        i = k if k >= 0 else ((-k) | (1 << 63))
        return struct.unpack("<d", struct.pack("<Q", i & 0xFFFFFFFFFFFFFFFF))[0]

    rows = []
    for f, lo, hi in funcs:
        steps = []

        def rec(a, b, fa, fb):
            if fa == fb:
                return
            if b - a == 1:
                steps.append((unkey(b), fb - fa))
                return
            m = (a + b) // 2
            fm = f(unkey(m))
            rec(a, m, fa, fm)
            rec(m, b, fm, fb)

        rec(key(lo), key(hi), f(lo), f(hi))
        if len(steps) != 2 or steps[0][1] != steps[1][1] or abs(steps[0][1]) != 1:
            raise RuntimeError("unexpected sector step structure: %r" % (steps,))
        rows.append([float(f(lo)), float(steps[0][1]), steps[0][0], steps[1][0]])
    return np.array(rows, dtype=np.float64)


def build_graph_batch(boxes, distance_threshold=0.5, context=None, want_shared=False):
    """boxes: [B,N,4] float32/float64 tensor or ndarray -> (types int8 [B,N,N], shared int8 [8,B,N,N] | None,
    bits int16 [B,N,N] | None) on the GPU."""
    if not torch.is_tensor(boxes):
        boxes = torch.from_numpy(np.ascontiguousarray(boxes))
    if boxes.dtype not in (torch.float32, torch.float64):
        boxes = boxes.double()
    if not boxes.is_cuda:
        boxes = boxes.cuda(non_blocking=True)
    boxes = boxes.contiguous()
    B, N, four = boxes.shape
    assert four == 4
    dev = boxes.device
    types = torch.empty(B, N, N, dtype=torch.int8, device=dev)
    shared = torch.empty(8, B, N, N, dtype=torch.int8, device=dev) if want_shared else None
    bits = torch.empty(B, N, N, dtype=torch.int16, device=dev) if context is not None else None
    fn = lib().samk_build_graph_f64 if boxes.dtype == torch.float64 else lib().samk_build_graph_f32
    sect = _sector_table.ctypes.data_as(ctypes.c_void_p) if _sector_table is not None else None
    check(fn(ptr(boxes), ptr(types), ptr(shared), ptr(bits), B, N, float(distance_threshold),
             int(context) if context is not None else 1, sect, stream_ptr()), "build_graph")
    ops._count()
    return types, shared, bits


def build_graph_using_normalized_boxes(bbox, label_num=11, distance_threshold=0.5, build_gauss_bias=False):
    """Same contract as sam/spatial_utils.py:92-218: bbox [N,4] -> {"1","31",...,"92": int8 [N,N]}."""
    bbox = np.asarray(bbox, dtype=np.float64)
    types, shared, _ = build_graph_batch(bbox[None], distance_threshold, None, want_shared=True)
    out = {"1": types[0].cpu().numpy()}
    sh = shared[:, 0].cpu().numpy()
    for i, k in enumerate(SHARED_KEYS[1:]):
        out[k] = sh[i]
    return out


def torch_broadcast_adj_matrix(adj_matrix):
    """int8 [N,N] relation types -> int8 [N,N,12] one-hot heads (sam/spatial_utils.py:33-52)."""
    src = adj_matrix
    t = src.to(device="cuda", dtype=torch.int8).contiguous()
    n = t.numel()
    bits = torch.empty(t.shape, dtype=torch.int16, device=t.device)
    out = torch.empty(tuple(t.shape) + (12,), dtype=torch.int8, device=t.device)
    check(lib().samk_types_to_bits(ptr(t), ptr(bits), n, 1, stream_ptr()), "types_to_bits")
    check(lib().samk_unpack_bits(ptr(bits), ptr(out), n, 12, stream_ptr()), "unpack_bits")
    ops._count(2)
    return out.to(src.device).to(src.dtype)


def expand_context(types, context):
    """types int8 [...,N,N] (cuda) -> reference-layout head masks int8 [...,N,N,12] for context c."""
    t = types.to(device="cuda", dtype=torch.int8).contiguous()
    n = t.numel()
    bits = torch.empty(t.shape, dtype=torch.int16, device=t.device)
    out = torch.empty(tuple(t.shape) + (12,), dtype=torch.int8, device=t.device)
    check(lib().samk_types_to_bits(ptr(t), ptr(bits), n, int(context), stream_ptr()), "types_to_bits")
    check(lib().samk_unpack_bits(ptr(bits), ptr(out), n, 12, stream_ptr()), "unpack_bits")
    ops._count(2)
    return out
