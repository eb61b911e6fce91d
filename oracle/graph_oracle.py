"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference spatial-graph builder.

Follows /root/reference/sam/spatial_utils.py:
  * pair classification            build_graph_using_normalized_boxes  :92-218
  * IoU                            bb_intersection_over_union           :7-30
  * shared ("31".."92") matrices   _build_replace_dict                  :55-89, :205-213
  * type -> 12-head one-hot        torch_broadcast_adj_matrix           :33-52
  * context expansion c=3/5/7/9    sam/datasets/textvqa_dataset.py      :373-409

The reference walks the i<j pairs with a Python double loop over float64 NumPy scalars; this
restatement evaluates the same IEEE-754 float64 operations in the same order on whole [N,N]
arrays (every add / multiply / divide / sqrt below is one correctly-rounded numpy op, so the
values are bit-identical to the scalar loop; `np.arcsin` / `np.arccos` are NumPy's own, as in
the reference :174-189).  Pinned against the unmodified reference by tests/test_graph_oracle.py
(random, grid-aligned adversarial and Appendix-A known-answer boxes) and tests/golden/graph_*.npz.

Documented deviation: the reference raises AssertionError when a non-contained pair has
union area 0 (:22).  Here that pair gets IoU = NaN, which compares false against 0.5 and falls
through to the directional branch (coincident centres -> NaN angle -> type 4, exactly the
reference's own NaN rule :192-203).  The CUDA kernel does the same.
"""
import math

import numpy as np

SHARED_KEYS = ("1", "31", "32", "51", "52", "71", "72", "91", "92")
_SHIFT = {"31": 1, "32": -1, "51": 2, "52": -2, "71": 3, "72": -3, "91": 4, "92": -4}


def _shift_sector(t, k):
    """Directional type t in 4..11 moved k sectors round the compass (:68-87)."""
    return (t - 4 + k) % 8 + 4


def _ceil_sector(label):
    """int(ceil(label / (pi/4))) + 3, NaN -> 4 (:192-203)."""
    q = np.ceil(label / (math.pi / 4))
    out = np.where(np.isnan(q), 1.0, q) + 3.0
    return out.astype(np.int64)


def build_graph(bbox, distance_threshold=0.5):
    """bbox float64 [N,4] (x1,y1,x2,y2) -> dict of nine int8 [N,N] matrices."""
    bbox = np.asarray(bbox, dtype=np.float64)
    n = bbox.shape[0]
    x1, y1, x2, y2 = (bbox[:, k] for k in range(4))
    # Python sum(): ((((0 + a) + b) + c) + d), :134
    pad = ((((0.0 + x1) + y1) + x2) + y2) == 0
    cx = 0.5 * (x1 + x2)
    cy = 0.5 * (y1 + y2)
    I = np.arange(n)[:, None]
    J = np.arange(n)[None, :]
    upper = (I < J) & ~pad[:, None] & ~pad[None, :]

    def a(v):  # value of box i broadcast along rows
        return v[:, None]

    def b(v):  # value of box j
        return v[None, :]

    i_covers_j = (a(x1) < b(x1)) & (a(x2) > b(x2)) & (a(y1) < b(y1)) & (a(y2) > b(y2))
    j_covers_i = (b(x1) < a(x1)) & (b(x2) > a(x2)) & (b(y1) < a(y1)) & (b(y2) > a(y2))

    with np.errstate(all="ignore"):
        ix = np.maximum(0, np.minimum(a(x2), b(x2)) - np.maximum(a(x1), b(x1)))
        iy = np.maximum(0, np.minimum(a(y2), b(y2)) - np.maximum(a(y1), b(y1)))
        inter = ix * iy
        area_a = (a(x2) - a(x1)) * (a(y2) - a(y1))
        area_b = (b(x2) - b(x1)) * (b(y2) - b(y1))
        iou = inter / ((area_a + area_b) - inter)
        overlap = iou >= 0.5

        yd = a(cy) - b(cy)
        xd = a(cx) - b(cx)
        diag = np.sqrt(yd * yd + xd * xd)
        near = diag < distance_threshold * math.sqrt(1.0 ** 2 + 1.0 ** 2)
        s = yd / diag
        c = xd / diag
        q1 = (s >= 0) & (c >= 0)
        q4 = (s < 0) & (c >= 0)
        q2 = (s >= 0) & (c < 0)
        asin_s = np.arcsin(s)
        acos_c = np.arccos(c)
        lab_i = np.where(q1, asin_s,
                 np.where(q4, asin_s + 2 * math.pi,
                  np.where(q2, acos_c, 2 * math.pi - acos_c)))
        lab_j = np.where(q1 | q2, lab_i + math.pi, lab_i - math.pi)
        # the reference writes label_j = math.pi + label_i in Q1 and label_i + math.pi in Q2:
        # IEEE addition commutes, so one expression covers both (:175,183).
        t_ij = _ceil_sector(lab_i)
        t_ji = _ceil_sector(lab_j)

    cls1 = upper & i_covers_j
    cls2 = upper & ~i_covers_j & j_covers_i
    rest = upper & ~i_covers_j & ~j_covers_i
    cls3 = rest & overlap
    direc = rest & ~overlap & near

    m = np.zeros((n, n), dtype=np.int64)
    m[np.arange(n)[~pad], np.arange(n)[~pad]] = 12
    mt = m.T  # view: writes through mt[i,j] land in m[j,i]
    m[cls1] = 1
    mt[cls1] = 2
    m[cls2] = 2
    mt[cls2] = 1
    m[cls3] = 3
    mt[cls3] = 3
    m[direc] = t_ij[direc]
    mt[direc] = t_ji[direc]

    out = {"1": m.astype(np.int8)}
    dir_full = direc | direc.T
    for key, k in _SHIFT.items():
        sh = np.zeros((n, n), dtype=np.int64)
        # only directional types 4..11 have an entry in the replace dict (.get(t, 0), :205-213);
        # a directional pair whose angle is exactly 0 gets type 3 and therefore 0 here.
        ok = dir_full & (m >= 4) & (m <= 11)
        sh[ok] = _shift_sector(m[ok], k)
        out[key] = sh.astype(np.int8)
    return out


def onehot_heads(types):
    """int8 [...,N,N] types 0..12 -> int8 [...,N,N,12]; type t>0 sets head t-1 (:33-52)."""
    types = np.asarray(types)
    heads = np.arange(1, 13, dtype=types.dtype)
    return (types[..., None] == heads).astype(np.int8)


def expand_context(shared, context):
    """Head masks for context c in {1,3,5,7,9}: the max-chain of textvqa_dataset.py:378-409."""
    m = onehot_heads(shared["1"])
    for c in (3, 5, 7, 9):
        if c > context:
            break
        m = np.maximum(m, onehot_heads(shared["%d1" % c]))
        m = np.maximum(m, onehot_heads(shared["%d2" % c]))
    return m


def head_bits_closed_form(types, context):
    """uint16 bit h set iff head h may attend for this type.

    Closed form of expand_context(): types 1,2,3,12 -> head t-1 only; directional types
    4..11 -> heads 3..10 within cyclic distance (context-1)/2 of t-4.
    """
    types = np.asarray(types).astype(np.int64)
    r = (context - 1) // 2
    bits = np.zeros(types.shape, dtype=np.uint16)
    for t in range(1, 13):
        if 4 <= t <= 11:
            v = 0
            for d in range(-r, r + 1):
                v |= 1 << (3 + (t - 4 + d) % 8)
        else:
            v = 1 << (t - 1)
        bits[types == t] = v
    return bits
