"""Seeded synthetic SA-M4C batches.

Shapes, dtypes and padding follow what the reference data pipeline hands to `SAM4C.forward`
(/root/reference/sam/datasets/textvqa_dataset.py:285-305, 338-348; processors.py:684-691):
trailing rows of every per-sample array are zero padding and the `*_mask` tensors mark the
valid prefix.  Boxes are built in float32 like `_image_features_reader.py:155-169`.

The relation matrices are the REAL output of a graph builder on these boxes; the builder is a
parameter (`graph_fn(boxes float64 [B,N,4]) -> int8 types [B,N,N]`) so that the product path
passes the CUDA builder and CPU-only tooling passes the numpy oracle.
"""
import numpy as np
import torch


def head_bits_for_context(context):
    """uint16 lookup table [13]: bit h set iff head h is active for relation type t at context c.

    Closed form of torch_broadcast_adj_matrix + the max-chain of
    /root/reference/sam/datasets/textvqa_dataset.py:373-409.
    """
    r = (int(context) - 1) // 2
    lut = np.zeros(13, dtype=np.uint16)
    for t in range(1, 13):
        if 4 <= t <= 11:
            v = 0
            for d in range(-r, r + 1):
                v |= 1 << (3 + (t - 4 + d) % 8)
        else:
            v = 1 << (t - 1)
        lut[t] = v
    return lut


def expand_types_to_heads(types, context):
    """int8 [B,N,N] relation types -> int8 [B,N,N,12] reference-layout head masks."""
    lut = torch.from_numpy(head_bits_for_context(context).astype(np.int64))
    bits = lut[types.long()]
    heads = torch.arange(12)
    return ((bits.unsqueeze(-1) >> heads) & 1).to(torch.int8)


def make_boxes(rs, B, N):
    x1 = rs.uniform(0, 0.8, (B, N)).astype(np.float32)
    y1 = rs.uniform(0, 0.8, (B, N)).astype(np.float32)
    w = rs.uniform(0.01, 0.31, (B, N)).astype(np.float32)
    h = rs.uniform(0.01, 0.31, (B, N)).astype(np.float32)
    x2 = np.minimum(x1 + w, np.float32(1.0))
    y2 = np.minimum(y1 + h, np.float32(1.0))
    area = (x2 - x1) * (y2 - y1)
    return np.stack([x1, y1, x2, y2, area], axis=-1).astype(np.float32)


def make_batch(B, T=20, O=100, R=50, D=12, V=5000, seed=0, contexts=(1, 3), graph_fn=None,
               targets=True):
    """CPU batch dict with every key `SAM4C.forward` and the loss read."""
    rs = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    n_obj = rs.randint(max(O // 2, 1), O + 1, B)
    n_ocr = rs.randint(1, R + 1, B)
    n_q = rs.randint(3, T + 1, B)
    obj_mask = (np.arange(O)[None] < n_obj[:, None])
    ocr_mask = (np.arange(R)[None] < n_ocr[:, None])
    q_mask = (np.arange(T)[None] < n_q[:, None])

    obj_box = make_boxes(rs, B, O) * obj_mask[..., None]
    ocr_box = make_boxes(rs, B, R) * ocr_mask[..., None]

    def feats(n, d, mask, kind):
        if kind == "relu":
            x = torch.randn(B, n, d, generator=g).abs_()
        elif kind == "normal":
            x = torch.randn(B, n, d, generator=g)
        else:
            x = (torch.rand(B, n, d, generator=g) < 0.05).float()
        return x * torch.from_numpy(mask.astype(np.float32))[..., None]

    batch = {
        "pad_obj_features": feats(O, 2048, obj_mask, "relu"),
        "pad_obj_bboxes": torch.from_numpy(obj_box.astype(np.float32)),
        "pad_obj_mask": torch.from_numpy(obj_mask.astype(np.int64)),
        "pad_ocr_features": feats(R, 2048, ocr_mask, "relu"),
        "pad_ocr_bboxes": torch.from_numpy(ocr_box.astype(np.float32)),
        "pad_ocr_mask": torch.from_numpy(ocr_mask.astype(np.int64)),
        "ocr_fasttext": feats(R, 300, ocr_mask, "normal"),
        "ocr_phoc": feats(R, 604, ocr_mask, "bern"),
        "question_indices": torch.from_numpy((rs.randint(1000, 30000, (B, T)) * q_mask).astype(np.int64)),
        "question_mask": torch.from_numpy(q_mask.astype(np.int64)),
    }
    prev = rs.randint(0, V + R, (B, D)).astype(np.int64)
    prev[:, 0] = 1
    batch["train_prev_inds"] = torch.from_numpy(prev)
    batch["train_loss_mask"] = torch.ones(B, D)
    if targets:
        batch["targets"] = (torch.rand(B, D, V + R, generator=g) < 1e-3).float()

    boxes = np.concatenate([obj_box[..., :4], ocr_box[..., :4]], axis=1).astype(np.float64)
    batch["boxes"] = torch.from_numpy(boxes)
    if graph_fn is not None:
        types = graph_fn(boxes)
        if not torch.is_tensor(types):
            types = torch.from_numpy(np.asarray(types))
        types = types.cpu()
        batch["spatial_types"] = types
        adj = {str(c): expand_types_to_heads(types, c) for c in contexts}
        adj["full_spatial"] = (types != 0).int()
        batch["spatial_adj_matrices"] = adj
    return batch


def seeded_state(named_shapes, seed=0, classifier_std=0.03, ptr_std=0.2):
    """Deterministic weights keyed by parameter NAME (independent of module construction order).

    Used by the golden fixtures and every parity test so that the reference module, the oracle
    and the CUDA module can be given identical weights without shipping a checkpoint.
    Linear / embedding weights ~ N(0, 0.02) (classifier and pointer-net wider so greedy decoding leaves BOS and copies varied OCR tokens),
    biases ~ N(0, 0.05), LayerNorm weights ~ 1 + N(0, 0.05).
    """
    import zlib
    out = {}
    for name, shape in named_shapes:
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        x = torch.randn(tuple(shape), generator=g)
        is_ln = ("LayerNorm" in name) or ("layer_norm" in name)
        if name.endswith(".bias"):
            x = 0.05 * x
        elif is_ln:
            x = 1.0 + 0.05 * x
        elif name.startswith("classifier."):
            x = classifier_std * x
        elif name.startswith("ocr_ptr_net."):
            x = ptr_std * x
        else:
            x = 0.02 * x
        out[name] = x
    return out
