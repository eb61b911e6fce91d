"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the SA-M4C forward / loss.

A functional (state_dict in, tensors out) restatement of the reference algorithm, op for op,
so that (i) the CUDA path can be checked on the GPU box, where /root/reference does not exist,
and (ii) `bench.py` has a CPU baseline ("kind": "port") that performs the same work as the
reference, including its dense [B,L,L,H] mask algebra, its `torch.unique` sanity check and
its dropout draws in train mode.

Each function cites the reference lines it restates (paths relative to /root/reference):
  obj / ocr input encodings        sam/sa_m4c.py:204-257
  TextBert                         sam/sa_m4c.py:382-396  (+ pytorch-transformers 1.x BertEmbeddings/BertLayer)
  PrevPredEmbeddings, gather       sam/sa_m4c.py:919-948, 970-982
  MMT mask + encoder loop          sam/sa_m4c.py:782-863, 730-770
  spatial self-attention           sam/sa_m4c.py:453-610
  pointer network, classifier      sam/sa_m4c.py:878-897, 270-278
  greedy decoding loop             sam/sa_m4c.py:285-302
  masked BCE loss                  sam/task_utils.py:19-30
The third-party BERT blocks (pytorch-transformers==1.0.0, requirements.txt:1, not vendored)
are restated from their published semantics: post-LN block, erf-GELU, TF-style LayerNorm with
eps inside the square root, additive -10000 masks.

Pinned by tests/test_sam4c_oracle.py against the UNMODIFIED reference module (build container)
and against tests/golden/sam4c_cfg1.npz (everywhere).
"""
import math

import torch
import torch.nn.functional as F

MATRIX_KEY = {"none": "1", "share3": "3", "share5": "5", "share7": "7", "share9": "9"}


def _ln(x, P, name, eps=1e-12):
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return P[name + ".weight"] * ((x - u) / torch.sqrt(s + eps)) + P[name + ".bias"]


def _lin(x, P, name):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


def _drop(x, p, train):
    return F.dropout(x, p, training=True) if (train and p > 0) else x


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _heads(x, nh):
    b, l, d = x.shape
    return x.view(b, l, nh, d // nh).permute(0, 2, 1, 3)


def _attend(P, pre, h, add_mask, nh, p_attn, train, spatial_add=None):
    """Self-attention core.  spatial_add None -> plain BertSelfAttention;
    else the SpatialBertSelfAttention steps (4)-(6) of sa_m4c.py:566-588."""
    q = _heads(_lin(h, P, pre + "query"), nh)
    k = _heads(_lin(h, P, pre + "key"), nh)
    v = _heads(_lin(h, P, pre + "value"), nh)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.size(-1))
    if spatial_add is None:
        probs = torch.softmax(scores + add_mask, dim=-1)
    else:
        combined = torch.min(add_mask, spatial_add)
        assert len(torch.unique(combined)) == 2                      # :569
        alive = ((combined.max(dim=-1)[0] + 10000.0) / 10000.0).unsqueeze(-1)
        probs = torch.softmax(scores + combined, dim=-1) * alive
    probs = _drop(probs, p_attn, train)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(ctx.size(0), ctx.size(1), -1)
    if spatial_add is not None and (pre + "biases.weight") in P:     # use_bias, sa_m4c.py:439-443, 600-603
        ctx = ctx + P[pre + "biases.weight"][0]
    return ctx


def _bert_layer(P, pre, h, add_mask, cfg, train, spatial_add=None, nh=12):
    eps = cfg["layer_norm_eps"]
    ph = cfg["hidden_dropout_prob"]
    ctx = _attend(P, pre + "attention.self.", h, add_mask, nh,
                  cfg["attention_probs_dropout_prob"], train, spatial_add)
    a = _ln(_drop(_lin(ctx, P, pre + "attention.output.dense"), ph, train) + h,
            P, pre + "attention.output.LayerNorm", eps)
    g = _gelu(_lin(a, P, pre + "intermediate.dense"))
    return _ln(_drop(_lin(g, P, pre + "output.dense"), ph, train) + a,
               P, pre + "output.LayerNorm", eps)


def _spatial_additive_mask(adj, like, T, L, quadrants):
    """sa_m4c.py:475-552: ones[B,L,L,H]; entity block <- adj; zero quadrants; (1-m)*-1e4."""
    B, A, _, H = adj.shape
    m = like.new_ones((B, L, L, H))
    m[:, T:T + A, T:T + A, :] = adj
    for quad in quadrants:
        rows = {1: slice(0, T), 2: slice(0, T), 4: slice(T, T + A),
                7: slice(T + A, None), 8: slice(T + A, None), 9: slice(T + A, None)}
        cols = {1: slice(0, T), 2: slice(T, T + A), 4: slice(0, T),
                7: slice(0, T), 8: slice(T, T + A), 9: slice(T + A, None)}
        if quad not in rows:
            raise ValueError(quad)
        m[:, rows[quad], cols[quad], :] = 0
    return ((1.0 - m) * -10000.0).permute(0, 3, 1, 2)


def encode_inputs(P, batch, mmt, train):
    """obj_mmt_in, ocr_mmt_in (sa_m4c.py:204-257)."""
    def nrm(x):
        return F.normalize(x, dim=-1) if mmt["normalize"] else x
    obj = (_ln(_lin(nrm(batch["pad_obj_features"]), P, "linear_obj_feat_to_mmt_in"), P, "obj_feat_layer_norm")
           + _ln(_lin(batch["pad_obj_bboxes"][:, :, :-1], P, "linear_obj_bbox_to_mmt_in"), P, "obj_bbox_layer_norm"))
    obj = _drop(obj, mmt["obj_drop"], train)
    fc6 = batch["pad_ocr_features"]
    order = fc6.new_zeros((fc6.size(0), fc6.size(1), 50))   # :242 generalised from (B,50,50) to (B,R,50)
    if mmt["use_phoc_fasttext"]:
        feat = torch.cat([nrm(batch["ocr_fasttext"]), nrm(batch["ocr_phoc"]), nrm(fc6), order], dim=-1)
    else:
        feat = torch.cat([nrm(fc6), order], dim=-1)
    ocr = (_ln(_lin(feat, P, "linear_ocr_feat_to_mmt_in"), P, "ocr_feat_layer_norm")
           + _ln(_lin(batch["pad_ocr_bboxes"][:, :, :-1], P, "linear_ocr_bbox_to_mmt_in"), P, "ocr_bbox_layer_norm"))
    return obj, _drop(ocr, mmt["ocr_drop"], train)


def text_bert(P, batch, tb, train):
    """sa_m4c.py:382-396."""
    ids = batch["question_indices"]
    pos = torch.arange(ids.size(1), device=ids.device)
    e = (P["text_bert.embeddings.word_embeddings.weight"][ids]
         + P["text_bert.embeddings.position_embeddings.weight"][pos][None]
         + P["text_bert.embeddings.token_type_embeddings.weight"][0][None, None])
    h = _drop(_ln(e, P, "text_bert.embeddings.LayerNorm", tb["layer_norm_eps"]), tb["hidden_dropout_prob"], train)
    add = ((1.0 - batch["question_mask"][:, None, None, :]) * -10000.0).to(h.dtype)
    for i in range(tb["num_hidden_layers"]):
        h = _bert_layer(P, "text_bert.encoder.layer.%d." % i, h, add, tb, train,
                        nh=tb["num_attention_heads"])
    return h


def prev_pred_embeddings(P, ocr_mmt_in, prev_inds, mmt, train):
    """sa_m4c.py:919-948 with the [B,V+R,768] concat of :932-934 kept (same gather result)."""
    eps = mmt["layer_norm_eps"]
    pre = "mmt.prev_pred_embeddings."
    ans = _ln(P["classifier.weight"], P, pre + "ans_layer_norm", eps)
    ocr = _ln(ocr_mmt_in, P, pre + "ocr_layer_norm", eps)
    B, V = prev_inds.size(0), ans.size(0)
    table = torch.cat([ans.unsqueeze(0).expand(B, -1, -1), ocr], dim=1)
    raw = torch.gather(table, 1, prev_inds.unsqueeze(-1).expand(-1, -1, table.size(-1)))
    D = prev_inds.size(1)
    emb = (P[pre + "position_embeddings.weight"][:D][None]
           + P[pre + "token_type_embeddings.weight"][(prev_inds >= V).long()])
    return raw + _drop(_ln(emb, P, pre + "emb_layer_norm", eps), mmt["hidden_dropout_prob"], train)


def mmt_forward(P, batch, txt, obj, ocr, prev_inds, mmt, train):
    """sa_m4c.py:782-863 + encoder loop :730-770.  Returns the [B,L,768] joint output."""
    dec = prev_pred_embeddings(P, ocr, prev_inds, mmt, train)
    x = torch.cat([txt, obj, ocr, dec], dim=1)
    T, D = txt.size(1), dec.size(1)
    L = x.size(1)
    valid = torch.cat([batch["question_mask"], batch["pad_obj_mask"], batch["pad_ocr_mask"],
                       torch.zeros_like(prev_inds)], dim=1)
    ext = valid[:, None, None, :].repeat(1, 1, L, 1).to(x.dtype)
    ext[:, :, -D:, -D:] = torch.tril(torch.ones(D, D, dtype=x.dtype, device=x.device))
    add = (1.0 - ext) * -10000.0
    n_i, s_i = 0, 0
    for kind, mix in zip(mmt["layer_type_list"], mmt.get("mix_list") or ["none"] * len(mmt["layer_type_list"])):
        if kind == "n":
            x = _bert_layer(P, "mmt.encoder.normal_layers.%d." % n_i, x, add, mmt, train,
                            nh=mmt["num_attention_heads"])
            n_i += 1
        elif kind == "s":
            adj = batch["spatial_adj_matrices"][MATRIX_KEY[mix]].to(x.device)
            sp = _spatial_additive_mask(adj, add, T, L, mmt["attention_mask_quadrants"])
            x = _bert_layer(P, "mmt.encoder.spatial_layers.%d." % s_i, x, add, mmt, train,
                            spatial_add=sp, nh=mmt["num_spatial_relations"])
            s_i += 1
        else:
            raise ValueError(kind)
    return x


def output_scores(P, seq, batch, T, O, R, D):
    """classifier + OcrPtrNet (sa_m4c.py:270-278, 878-897)."""
    dec = seq[:, -D:]
    ocr = seq[:, T + O:T + O + R]
    fixed = _lin(dec, P, "classifier")
    q = _lin(dec, P, "ocr_ptr_net.query")
    k = _lin(ocr, P, "ocr_ptr_net.key")
    ptr = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.size(-1))
    ptr = ptr + ((1.0 - batch["pad_ocr_mask"].to(ptr.dtype)) * -10000.0).unsqueeze(1)
    return torch.cat([fixed, ptr], dim=-1)


def forward(P, batch, mmt, tb, train=False, bos_idx=1, teacher_forced=None):
    """SAM4C.forward (sa_m4c.py:179-202, 280-302).  Returns (scores, prev_inds_used, seq_output)."""
    mmt = _with_defaults(mmt)
    tb = _with_defaults(tb)
    obj, ocr = encode_inputs(P, batch, mmt, train)
    T, O, R = batch["question_mask"].size(1), obj.size(1), ocr.size(1)
    prev = batch["train_prev_inds"]
    D = prev.size(1)
    if teacher_forced is None:
        teacher_forced = train
    if teacher_forced:
        txt = text_bert(P, batch, tb, train)
        seq = mmt_forward(P, batch, txt, obj, ocr, prev, mmt, train)
        return output_scores(P, seq, batch, T, O, R, D), prev, seq
    prev = torch.zeros_like(prev)
    prev[:, 0] = bos_idx
    for _ in range(D):
        txt = text_bert(P, batch, tb, train)
        seq = mmt_forward(P, batch, txt, obj, ocr, prev, mmt, train)
        scores = output_scores(P, seq, batch, T, O, R, D)
        prev[:, 1:] = scores.argmax(dim=-1)[:, :-1]
    return scores, prev, seq


def bce_with_mask_loss(scores, targets, loss_mask):
    """sam/task_utils.py:19-30."""
    losses = F.binary_cross_entropy_with_logits(scores, targets, reduction="none")
    losses = losses * loss_mask.unsqueeze(-1)
    count = torch.max(torch.sum(loss_mask), torch.ones(1, device=losses.device))
    return torch.sum(losses) / count


_BERT_DEFAULTS = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072,
                      hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                      layer_norm_eps=1e-12, num_hidden_layers=12)


def _with_defaults(cfg):
    out = dict(_BERT_DEFAULTS)
    out.update(dict(cfg))
    return out
