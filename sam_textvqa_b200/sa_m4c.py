"""Drop-in `SAM4C` for /root/reference/sam/sa_m4c.py, running on hand-written sm_100a kernels.

Same constructor (`SAM4C(mmt_config, text_bert_config)`), same `forward(batch_dict,
use_beam_search=False) -> {"textvqa_scores": [B,D,V+R]}`, same config keys, same `state_dict`
names and shapes (179 tensors for the shipped yml), same side keys left in
`batch_dict`, same `get_optimizer_parameters` / `finetune_modules`, so the reference `train.py` /
`evaluator.py` run unchanged with `sys.modules["sam.sa_m4c"]` pointing here.

The nn.Module tree below only HOLDS parameters under the reference's names; the arithmetic is in
`ops.py` -> libsamk.so:
  * masks are never materialised: `[B,L]` key-valid bytes + packed 12-bit relation words
    `[B,A,A]` replace the reference's fp32 `[B,1,L,L]` and `[B,L,L,12]` tensors (sa_m4c.py:475-552,
    834-844), and the `torch.unique` debug sort (:569) does not exist;
  * q|k|v are one fused 768->2304 projection; bias/GELU/dropout/residual live in GEMM epilogues;
  * PrevPredEmbeddings gathers and normalises only the D rows it needs instead of
    concatenating a [B,V+R,768] table (:932-934).
There is no CPU path: CPU inputs are moved to the module's device like `forward_model` does
(sam/task_utils.py:113-115), a missing libsamk.so raises.
"""
import logging
import os
from collections import Counter

import torch
from torch import nn

from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .config import BertConfig  # noqa: F401  (train.py imports BertConfig from sam.sa_m4c)

logger = logging.getLogger(__name__)

try:  # inside the reference tree: share its global registry (tools/registry.py:1-3)
    from tools.registry import registry  # type: ignore
except Exception:  # standalone
    from .registry import registry

_LEGAL_QUADRANTS = (1, 2, 4, 7, 8, 9)   # sa_m4c.py:505-549 raises ValueError for anything else
_MATRIX_KEY = {"none": "1", "share3": "3", "share5": "5", "share7": "7", "share9": "9"}


class BertLayerNorm(nn.Module):
    """Parameter holder + TF-style LayerNorm (sa_m4c.py:1016-1028)."""

    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        shape = x.shape
        return ops.layer_norm(x.reshape(-1, shape[-1]), self.weight, self.bias, self.variance_epsilon).view(shape)


def _init_bert_weights(module, std):
    """BertPreTrainedModel._init_weights (pytorch-transformers): N(0,std) weights, zero bias, LN=(1,0)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, BertLayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()


class _SelfAttentionParams(nn.Module):
    def __init__(self, config, heads):
        super().__init__()
        if config.hidden_size % heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (config.hidden_size, heads))
        self.num_attention_heads = heads
        self.attention_head_size = config.hidden_size // heads
        if self.attention_head_size != 64:
            raise NotImplementedError("samk attention kernels are built for head size 64")
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(config.hidden_size, config.hidden_size)
        self.value = nn.Linear(config.hidden_size, config.hidden_size)
        self.attention_probs_dropout_prob = config.attention_probs_dropout_prob


class SpatialBertSelfAttention(_SelfAttentionParams):
    """Parameters of sa_m4c.py:399-451.  One head per spatial relation (`num_spatial_relations`)."""

    def __init__(self, config, use_implicit=False):
        assert hasattr(config, "num_spatial_relations")
        if use_implicit:
            raise ValueError("implicit layers are not constructible in the reference either (sa_m4c.py:751-752)")
        super().__init__(config, config.num_spatial_relations)
        self.num_spatial_relations = config.num_spatial_relations
        self.max_seq_len = config.max_seq_length
        self.mask_quadrants = list(config.attention_mask_quadrants)
        self.max_decoding_steps = config.num_decoding_steps
        if getattr(config, "no_drop", False):
            self.attention_probs_dropout_prob = 0.0
        self.use_bias = bool(getattr(config, "use_bias", False))
        if self.use_bias:   # sa_m4c.py:439-443, 600-603: one learned d-vector added to every context row
            self.biases = nn.Embedding(1, config.hidden_size)
        qm = 0
        for q in self.mask_quadrants:
            if q not in _LEGAL_QUADRANTS:
                raise ValueError(q)
            qm |= 1 << (q - 1)
        self.quadrant_bits = qm


class _SelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _Attention(nn.Module):
    def __init__(self, config, spatial):
        super().__init__()
        self.self = (SpatialBertSelfAttention(config) if spatial
                     else _SelfAttentionParams(config, config.num_attention_heads))
        self.output = _SelfOutput(config)


class _Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError("only erf-GELU is fused into the FFN epilogue")
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class _Output(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLayer(nn.Module):
    """One post-LN block.  spatial=False: the third-party BertLayer ('n' layers, TextBert);
    spatial=True: SpatialBertLayer (sa_m4c.py:660-684)."""

    def __init__(self, config, spatial=False):
        super().__init__()
        self.spatial = spatial
        self.attention = _Attention(config, spatial)
        self.intermediate = _Intermediate(config)
        self.output = _Output(config)
        self.hidden_dropout_prob = config.hidden_dropout_prob
        self.layer_norm_eps = config.layer_norm_eps

    def _params(self):
        a, s = self.attention, self.attention.self
        out_bias = a.output.dense.bias
        if getattr(s, "use_bias", False):
            # W_o (ctx + b) + b_o = W_o ctx + (W_o b + b_o): the context bias of sa_m4c.py:600-603 (added to every row,
            # dead ones included) folds into the out-projection bias; autograd carries its gradient to b and W_o
            out_bias = out_bias + torch.nn.functional.linear(s.biases.weight, a.output.dense.weight)[0]
            ops.late_grad_param_ids.update((id(a.output.dense.weight), id(s.biases.weight), id(a.output.dense.bias)))
        return (s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
                a.output.dense.weight, out_bias, a.output.LayerNorm.weight, a.output.LayerNorm.bias,
                self.intermediate.dense.weight, self.intermediate.dense.bias,
                self.output.dense.weight, self.output.dense.bias, self.output.LayerNorm.weight,
                self.output.LayerNorm.bias)

    def forward(self, hidden, key_valid, rel_bits, seg, mask_cache=None):
        """hidden [B,L,d] fp32; key_valid uint8 [B,L]; rel_bits uint16 [B,A,A] or None; seg = (T,A,D);
        mask_cache: dict shared by the layers of one forward pass (packed allow-bits per mask kind)."""
        B, L, _ = hidden.shape
        s = self.attention.self
        T, A, D = seg
        dims = (B, L, s.num_attention_heads, T, A, D)
        train = self.training
        cfg = (dims, self.spatial, getattr(s, "quadrant_bits", 0),
               float(s.attention_probs_dropout_prob) if train else 0.0,
               float(self.hidden_dropout_prob) if train else 0.0, float(self.layer_norm_eps), mask_cache)
        return ops.BertLayerFn.apply(hidden, key_valid, rel_bits if self.spatial else None, cfg, *self._params())


    def infer(self, x2, key_valid, rel_bits, seg5, mask_cache, cache=None):
        """No-autograd forward on 2-D rows (see ops.bert_layer_infer); seg5 = (B, L, T, A, D)."""
        s = self.attention.self
        B, L, T, A, D = seg5
        dims = (B, L, s.num_attention_heads, T, A, D)
        spatial, quad = self.spatial, getattr(s, "quadrant_bits", 0)
        rel = rel_bits if spatial else None
        allow = None
        if ops.uses_tensor_core_attention(ops.act_dtype()):
            mkey = (bool(spatial), quad if spatial else 0, rel.data_ptr() if rel is not None else 0, dims)
            allow = mask_cache.get(mkey)
            if allow is None:
                allow = mask_cache[mkey] = ops.build_attn_mask(key_valid, rel, dims, spatial, quad)
        return ops.bert_layer_infer(x2, key_valid, rel, dims, spatial, quad, float(self.layer_norm_eps),
                                    self._params(), allow, cache)


class SpatialBertLayer(BertLayer):
    def __init__(self, config, use_implicit=False):
        if use_implicit:
            raise ValueError("implicit layers are not supported")
        super().__init__(config, spatial=True)


class _BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.hidden_dropout_prob = config.hidden_dropout_prob

    def forward(self, input_ids):
        p = float(self.hidden_dropout_prob) if self.training else 0.0
        return ops.BertEmbedFn.apply(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                                     self.token_type_embeddings.weight, self.LayerNorm.weight, self.LayerNorm.bias,
                                     float(self.LayerNorm.variance_epsilon), p)


class _BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class TextBert(nn.Module):
    """sa_m4c.py:374-396: BertEmbeddings + num_hidden_layers BertLayers over the question tokens."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = _BertEmbeddings(config)
        self.encoder = _BertEncoder(config)
        _init_bert_weights(self, config.initializer_range)

    def forward(self, batch_dict):
        ids = batch_dict["question_indices"]
        x = self.embeddings(ids)
        valid = ops.key_valid_bytes(batch_dict["question_mask"], None, None, 0)
        T = ids.shape[1]
        cache = {}
        for layer in self.encoder.layer:
            x = layer(x, valid, None, (T, 0, 0), cache)
        return x


class PrevPredEmbeddings(nn.Module):
    """sa_m4c.py:900-948."""

    def __init__(self, config):
        super().__init__()
        MAX_DEC_LENGTH, MAX_TYPE_NUM = 100, 5
        d, eps = config.hidden_size, config.layer_norm_eps
        self.position_embeddings = nn.Embedding(MAX_DEC_LENGTH, d)
        self.token_type_embeddings = nn.Embedding(MAX_TYPE_NUM, d)
        self.ans_layer_norm = BertLayerNorm(d, eps=eps)
        self.ocr_layer_norm = BertLayerNorm(d, eps=eps)
        self.emb_layer_norm = BertLayerNorm(d, eps=eps)
        self.hidden_dropout_prob = config.hidden_dropout_prob
        self.eps = eps

    def forward(self, ans_emb, ocr_emb, prev_inds):
        assert prev_inds.dim() == 2 and prev_inds.dtype == torch.long
        assert ans_emb.dim() == 2
        p = float(self.hidden_dropout_prob) if self.training else 0.0
        return ops.PrevPredFn.apply(prev_inds, ans_emb, ocr_emb, self.position_embeddings.weight,
                                    self.token_type_embeddings.weight, self.ans_layer_norm.weight,
                                    self.ans_layer_norm.bias, self.ocr_layer_norm.weight, self.ocr_layer_norm.bias,
                                    self.emb_layer_norm.weight, self.emb_layer_norm.bias, float(self.eps), p)


class BertSpatialEncoder(nn.Module):
    """sa_m4c.py:687-770: the layer_type_list / mix_list schedule over 'n' and 's' layers."""

    def __init__(self, config):
        super().__init__()
        self.layer_type_list = list(config.layer_type_list)
        counter = Counter(self.layer_type_list)
        self.num_spatial_layers, self.num_normal_layers = counter["s"], counter["n"]
        self.num_implicit_layers = counter["i"]
        if getattr(config, "mix_list", None) is None:
            self.mix_list = ["none"] * len(self.layer_type_list)
        else:
            self.mix_list = list(config.mix_list)
        assert len(self.mix_list) == len(self.layer_type_list)
        self.matrix_type_map = dict(_MATRIX_KEY)
        self.normal_layers = nn.ModuleList([BertLayer(config) for _ in range(self.num_normal_layers)])
        self.spatial_layers = nn.ModuleList([SpatialBertLayer(config) for _ in range(self.num_spatial_layers)])
        if self.num_implicit_layers:
            raise ValueError("layer type 'i' raises in the reference forward (sa_m4c.py:751-752)")
        self.implicit_layers = nn.ModuleList([])

    def forward(self, hidden, key_valid, rel_lookup, seg):
        normal_iter, spatial_iter = iter(self.normal_layers), iter(self.spatial_layers)
        cache = {}
        for layer_type, mix_type in zip(self.layer_type_list, self.mix_list):
            if layer_type == "n":
                hidden = next(normal_iter)(hidden, key_valid, None, seg, cache)
            elif layer_type == "s":
                hidden = next(spatial_iter)(hidden, key_valid, rel_lookup(self.matrix_type_map[mix_type]), seg, cache)
            else:
                raise ValueError
        return hidden


def _encoder_infer(enc, x2, key_valid, rel_lookup, dims, mask_cache, caches=None):
    """BertSpatialEncoder schedule without autograd.  caches None: all rows, returns (out, [per-layer cache]);
    else x2 = decoder rows only and the per-layer caches are updated in place."""
    normal_iter, spatial_iter = iter(enc.normal_layers), iter(enc.spatial_layers)
    new = []
    for li, (layer_type, mix_type) in enumerate(zip(enc.layer_type_list, enc.mix_list)):
        if layer_type == "n":
            layer, rel = next(normal_iter), None
        elif layer_type == "s":
            layer, rel = next(spatial_iter), rel_lookup(enc.matrix_type_map[mix_type])
        else:
            raise ValueError
        x2, c = layer.infer(x2, key_valid, rel, dims, mask_cache, None if caches is None else caches[li])
        new.append(c)
    return x2, new


class MMT(nn.Module):
    """sa_m4c.py:773-863."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.prev_pred_embeddings = PrevPredEmbeddings(config)
        self.encoder = BertSpatialEncoder(config)
        _init_bert_weights(self, config.initializer_range)

    def forward(self, batch_dict, fixed_ans_emb):
        dec_emb = self.prev_pred_embeddings(fixed_ans_emb, batch_dict["ocr_mmt_in"], batch_dict["train_prev_inds"])
        txt, obj, ocr = batch_dict["text_bert_emb"], batch_dict["obj_mmt_in"], batch_dict["ocr_mmt_in"]
        x = ops.join_segments(txt, obj, ocr, dec_emb)
        T, O, R, D = txt.size(1), obj.size(1), ocr.size(1), dec_emb.size(1)
        dev = x.device
        key_valid = ops.key_valid_bytes(batch_dict["question_mask"], batch_dict["pad_obj_mask"], batch_dict["pad_ocr_mask"], D)
        rel_cache = batch_dict.setdefault("_samk_rel_bits", {})

        def rel_lookup(key):
            if key not in rel_cache:
                rel_cache[key] = _relation_bits(batch_dict, key, dev)
            bits = rel_cache[key]
            if bits.shape[1] != O + R:
                raise ValueError("spatial_adj_matrices[%r] is %d x %d, expected %d entities"
                                 % (key, bits.shape[1], bits.shape[2], O + R))
            return bits

        seq = self.encoder(x, key_valid, rel_lookup, (T, O + R, D))
        return {
            "mmt_seq_output": seq,
            "mmt_txt_output": seq[:, :T],
            "mmt_ocr_output": seq[:, T + O:T + O + R],
            "mmt_dec_output": seq[:, -D:],
        }


def _relation_bits(batch_dict, key, dev):
    """Packed head bits uint16 [B,A,A] for relation-matrix key "1"/"3"/"5"/"7"/"9".

    Reference contract: `spatial_adj_matrices[key]` = int8 [B,A,A,12] prepared by the dataset (hours of CPU graph
    building + a pickle cache, sam/datasets/textvqa_dataset.py:228-280, 373-409).  On-device batch preparation
    (SURVEY 8f): when the dict (or the key) is absent the graph is built here, on the GPU, from the padded boxes the
    batch already carries -- [pad_obj_bboxes ; pad_ocr_bboxes][..., :4], exactly the array process_spatials feeds
    to the graph builder -- bit-identical to the reference builder + context expansion (tests/test_gpu_graph.py)."""
    adj = batch_dict.get("spatial_adj_matrices")
    if isinstance(adj, dict) and key in adj:
        return pack_relation_bits(adj[key], dev)
    if "pad_obj_bboxes" not in batch_dict or "pad_ocr_bboxes" not in batch_dict:
        raise KeyError("spatial_adj_matrices[%r] missing and no pad_obj_bboxes / pad_ocr_bboxes to build it from" % key)
    from . import spatial_utils
    boxes = torch.cat([batch_dict["pad_obj_bboxes"][..., :4], batch_dict["pad_ocr_bboxes"][..., :4]], dim=1)
    boxes = boxes.to(device=dev, dtype=torch.float32)
    _, _, bits = spatial_utils.build_graph_batch(boxes, 0.5, context=int(key))
    return bits


def pack_relation_bits(adj, device):
    """int8 [B,A,A,12] head masks (any device; the reference leaves them on the CPU when n_gpu == 1,
    SURVEY.md section 3.1) -> uint16 [B,A,A] on `device`, bit h = head h may attend."""
    if adj.dtype == torch.int16 and adj.dim() == 3:   # already packed (on-device batch preparation)
        return adj.to(device)
    adj = adj.to(device=device, dtype=torch.int8, non_blocking=True).contiguous()
    B, A, A2, H = adj.shape
    bits = torch.empty(B, A, A2, dtype=torch.int16, device=device)
    check(lib().samk_pack_adj(ptr(adj), ptr(bits), B * A * A2, H, stream_ptr()), "pack_adj")
    ops._count()
    return bits


class OcrPtrNet(nn.Module):
    """Parameter holder of sa_m4c.py:866-897; scoring is fused into ops.OutputFn."""

    def __init__(self, hidden_size, query_key_size=None):
        super().__init__()
        if query_key_size is None:
            query_key_size = hidden_size
        self.hidden_size, self.query_key_size = hidden_size, query_key_size
        self.query = nn.Linear(hidden_size, query_key_size)
        self.key = nn.Linear(hidden_size, query_key_size)


class GeLU(nn.Module):
    def forward(self, x):
        return x * 0.5 * (1.0 + torch.erf(x / 1.4142135623730951))


class SimpleClassifier(nn.Module):
    """sa_m4c.py:1031-1042 (aux heads only; off in every shipped config, plain torch)."""

    def __init__(self, in_dim, hid_dim, out_dim, dropout=0):
        super().__init__()
        self.logit_fc = nn.Sequential(nn.Linear(in_dim, hid_dim), GeLU(), nn.LayerNorm(hid_dim, eps=1e-12),
                                      nn.Linear(hid_dim, out_dim))

    def forward(self, hidden_states):
        return self.logit_fc(hidden_states)


_STRICT_INPUT_ENC = os.environ.get("SAMK_STRICT_INPUT_ENC", "0") == "1"
# 1: TextBert on its own stream (measured 11.49 vs 11.60 ms/step); a third stream for the OCR encoder measured no
# further gain (11.74 vs 11.76) and is not built
_BRANCH_STREAMS = os.environ.get("SAMK_BRANCH_STREAMS", "1") == "1"


class SAM4C(nn.Module):
    """SAM4C has two transformers, MMT and TextBert (sa_m4c.py:20-371)."""

    def __init__(self, mmt_config, text_bert_config):
        super().__init__()
        self.mmt_config = mmt_config
        self.text_bert_config = text_bert_config
        self.frcn_encoder_type = "default"
        self.normalize = self.mmt_config.normalize
        self.aux_spatial_fusion = getattr(self.mmt_config, "aux_spatial_fusion", "mul")
        self.use_aux_heads = getattr(self.mmt_config, "use_aux_heads", False)
        self.spatial_type = getattr(self.mmt_config, "spatial_type", "top")
        self.build()

    # ---- construction (same order and names as the reference) ----------------------------------
    def set_beam_size(self, beam_size):
        self.beam_size = beam_size

    def build(self):
        self.finetune_modules = []
        self._build_txt_encoding()
        self._build_obj_encoding()
        self._build_ocr_encoding()
        self._build_mmt()
        self._build_output()
        if self.use_aux_heads:
            self._build_aux_heads()

    def _build_txt_encoding(self):
        TEXT_BERT_HIDDEN_SIZE = 768
        self.text_bert = TextBert(self.text_bert_config)
        if self.text_bert_config.text_bert_init_from_bert_base:
            path = os.environ.get("SAMK_BERT_BASE_STATE")
            if not path:
                raise RuntimeError(
                    "text_bert_init_from_bert_base=true needs the bert-base-uncased weights; the reference "
                    "downloads them (sa_m4c.py:74-77). Offline: export them once with torch.save(state_dict) "
                    "and set SAMK_BERT_BASE_STATE=<file>, or set text_bert_init_from_bert_base: false.")
            sd = torch.load(path, map_location="cpu")
            # what pytorch_transformers' from_pretrained does to a bert-base-uncased checkpoint (sa_m4c.py:74-77): strip
            # the "bert." prefix, rename the TF-style LayerNorm gamma / beta, keep the first num_hidden_layers layers
            ren = {}
            for k, v in sd.items():
                k = k[len("bert."):] if k.startswith("bert.") else k
                if k.endswith(".gamma"):
                    k = k[:-len("gamma")] + "weight"
                elif k.endswith(".beta"):
                    k = k[:-len("beta")] + "bias"
                ren[k] = v
            res = self.text_bert.load_state_dict(ren, strict=False)
            if res.missing_keys:           # every TextBert parameter must come from the checkpoint
                raise RuntimeError("SAMK_BERT_BASE_STATE=%s lacks TextBert parameters: %s" % (path, res.missing_keys[:8]))
            extra = [k for k in res.unexpected_keys if not (k.startswith("encoder.layer.") or k.startswith("pooler.")
                                                           or k.startswith("cls."))]
            if extra:
                logger.warning("bert-base state: %d unexpected keys ignored, e.g. %s", len(extra), extra[:4])
            self.finetune_modules.append({"module": self.text_bert,
                                          "lr_scale": self.text_bert_config.lr_scale_text_bert})
        if self.mmt_config.hidden_size != TEXT_BERT_HIDDEN_SIZE:
            self.text_bert_out_linear = nn.Linear(TEXT_BERT_HIDDEN_SIZE, self.mmt_config.hidden_size)
        else:
            self.text_bert_out_linear = nn.Identity()

    def _build_obj_encoding(self):
        assert self.frcn_encoder_type == "default"
        self.obj_faster_rcnn_fc7 = nn.Identity()       # ImageEncoder("default") is an identity (textvqa_encoders.py:17-33)
        d = self.mmt_config.hidden_size
        self.linear_obj_feat_to_mmt_in = nn.Linear(self.mmt_config.obj_feature_size, d)
        self.linear_obj_bbox_to_mmt_in = nn.Linear(4, d)
        self.obj_feat_layer_norm = BertLayerNorm(d)
        self.obj_bbox_layer_norm = BertLayerNorm(d)
        self.obj_drop_prob = float(self.mmt_config.obj_drop)

    def _build_ocr_encoding(self):
        assert self.frcn_encoder_type == "default"
        self.ocr_faster_rcnn_fc7 = nn.Identity()
        d = self.mmt_config.hidden_size
        self.linear_ocr_feat_to_mmt_in = nn.Linear(self.mmt_config.ocr_feature_size, d)
        self.linear_ocr_bbox_to_mmt_in = nn.Linear(4, d)
        self.ocr_feat_layer_norm = BertLayerNorm(d)
        self.ocr_bbox_layer_norm = BertLayerNorm(d)
        self.ocr_drop_prob = float(self.mmt_config.ocr_drop)

    def _build_mmt(self):
        self.mmt = MMT(self.mmt_config)
        self.finetune_modules.append({"module": self.mmt, "lr_scale": self.mmt_config.lr_scale_mmt})

    def _build_output(self):
        self.ocr_ptr_net = OcrPtrNet(hidden_size=self.mmt_config.hidden_size,
                                     query_key_size=self.mmt_config.ptr_query_size)
        num_outputs = len(registry.answer_vocab)
        self.classifier = nn.Linear(self.mmt_config.hidden_size, num_outputs)

    def _build_aux_heads(self):
        self.origin_transform = SimpleClassifier(self.mmt_config.hidden_size, 128, 32)
        self.dest_transform = SimpleClassifier(self.mmt_config.hidden_size, 128, 32)
        self.spatial_classifier = nn.Linear(32, 12)

    # ---- forward --------------------------------------------------------------------------------
    def _device(self):
        return self.classifier.weight.device

    def _to_device(self, batch_dict):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("SAM4C (samk) runs on CUDA only: call model.cuda() first (there is no CPU path)")
        for k, v in list(batch_dict.items()):
            if torch.is_tensor(v) and v.device != dev and k != "spatial_adj_matrices":
                batch_dict[k] = v.to(dev, non_blocking=True)

    def forward(self, batch_dict, use_beam_search=False):
        if use_beam_search and (self.training or torch.is_grad_enabled()):
            raise RuntimeError("beam search is an inference path: call model.eval() and run under torch.no_grad()")
        self._to_device(batch_dict)
        ops.begin_forward()
        batch_dict.pop("_samk_rel_bits", None)
        # The three input encoders are independent until the MMT joins their rows.  In training TextBert (three small
        # layers, 2560 rows: its GEMMs fill 40-90 of the 148 SMs) runs on a second stream beside the region / OCR
        # encoders; autograd runs every backward node on its forward stream, so the backward pass gets the same
        # concurrency, and in a captured step the streams become parallel branches of the graph.  Python order (and
        # with it the order of the dropout streams) is unchanged.
        fork = None
        if _BRANCH_STREAMS and self.training and not use_beam_search and torch.is_grad_enabled():
            fork = torch.cuda.Event()
            fork.record()
        self._forward_obj_encoding(batch_dict)
        self._forward_ocr_encoding(batch_dict)
        if fork is not None:
            main, branch = torch.cuda.current_stream(), ops.branch_stream()
            branch.wait_event(fork)
            with torch.cuda.stream(branch):
                self._forward_text_bert(batch_dict)
            main.wait_stream(branch)
        if use_beam_search:
            self._forward_beam_search(batch_dict)
        else:
            self._forward_mmt_and_output(batch_dict, have_text=fork is not None)
        if self.use_aux_heads:
            self._forward_aux(batch_dict)
        return {"textvqa_scores": batch_dict["scores"]}

    def _encode(self, feat_buf, bbox, lin_feat, lin_bbox, ln_feat, ln_bbox, kdim, drop_p):
        B, N = bbox.shape[0], bbox.shape[1]
        d = self.mmt_config.hidden_size
        # SAMK_STRICT_INPUT_ENC=1: the 2048 / 2952-long feature projections as 3-term splits in every precision mode
        # (their half rounding is ~1.5e-4 of the 1e-3 logit budget; the split costs ~2 % of the step, so it is off)
        f = ops.layer_norm(ops.linear(feat_buf, lin_feat.weight, lin_feat.bias, kdim, strict=_STRICT_INPUT_ENC),
                           ln_feat.weight, ln_feat.bias, ln_feat.variance_epsilon)
        bb = bbox.reshape(B * N, bbox.shape[-1])[:, :4]                    # remove bbox-area (sa_m4c.py:214,252)
        g = ops.layer_norm(ops.linear(bb, lin_bbox.weight, lin_bbox.bias, 4), ln_bbox.weight, ln_bbox.bias,
                           ln_bbox.variance_epsilon)
        out = ops.dropout_add(f, g, drop_p if self.training else 0.0)
        return out.view(B, N, d)

    def _feature_buffers(self, rows, cols, device):
        """(destination(s) of the L2-normalised features, tensor handed to ops.linear).  Product mode: the projection reads
        a half copy and its weight-gradient product a bf16 copy, both written by the normalisation kernel itself -- the
        normalised fp32 features are never materialised (region + OCR features: 0.36 GB of traffic per step)."""
        if ops.get_precision() == "f16" and not _STRICT_INPUT_ENC and cols % 8 == 0:
            h = torch.empty(rows, cols, dtype=torch.float16, device=device)
            b = None
            if self.training and torch.is_grad_enabled():
                b = torch.empty(rows, cols, dtype=torch.bfloat16, device=device)
                ops.remember_act(h, b)
            return h, b, h
        buf = torch.empty(rows, cols, dtype=torch.float32, device=device)
        return buf, None, buf

    def _forward_obj_encoding(self, batch_dict):
        feats = batch_dict["pad_obj_features"].float()
        B, O, dfeat = feats.shape
        dst_a, dst_b, buf = self._feature_buffers(B * O, dfeat, feats.device)
        ops.l2norm_into2(feats, dst_a, dst_b, 0, self.normalize)
        batch_dict["obj_mmt_in"] = self._encode(
            buf, batch_dict["pad_obj_bboxes"].float(), self.linear_obj_feat_to_mmt_in, self.linear_obj_bbox_to_mmt_in,
            self.obj_feat_layer_norm, self.obj_bbox_layer_norm, dfeat, self.obj_drop_prob)

    def _forward_ocr_encoding(self, batch_dict):
        fc6 = batch_dict["pad_ocr_features"].float()
        B, R, _ = fc6.shape
        parts = [fc6]
        if self.mmt_config.use_phoc_fasttext:
            ft, ph = batch_dict["ocr_fasttext"].float(), batch_dict["ocr_phoc"].float()
            assert ft.size(-1) == 300 and ph.size(-1) == 604
            parts = [ft, ph, fc6]
        # the trailing 50 "order vector" columns are always zero (sa_m4c.py:240-242): they are not
        # materialised and the matching weight columns never enter the contraction.
        kdim = sum(p.size(-1) for p in parts)
        if kdim + 50 != self.linear_ocr_feat_to_mmt_in.weight.shape[1]:
            raise RuntimeError("ocr_feature_size %d does not match the concatenated OCR features (%d + 50)"
                               % (self.linear_ocr_feat_to_mmt_in.weight.shape[1], kdim))
        dst_a, dst_b, buf = self._feature_buffers(B * R, kdim, fc6.device)
        off = 0
        for p in parts:
            ops.l2norm_into2(p, dst_a, dst_b, off, self.normalize)
            off += p.size(-1)
        batch_dict["ocr_mmt_in"] = self._encode(
            buf, batch_dict["pad_ocr_bboxes"].float(), self.linear_ocr_feat_to_mmt_in, self.linear_ocr_bbox_to_mmt_in,
            self.ocr_feat_layer_norm, self.ocr_bbox_layer_norm, kdim, self.ocr_drop_prob)

    def _forward_text_bert(self, batch_dict):
        text_bert_out = self.text_bert(batch_dict)
        if isinstance(self.text_bert_out_linear, nn.Identity):
            batch_dict["text_bert_emb"] = text_bert_out
        else:
            B, T, d = text_bert_out.shape
            lin = self.text_bert_out_linear
            batch_dict["text_bert_emb"] = ops.linear(text_bert_out.reshape(B * T, d), lin.weight, lin.bias).view(B, T, -1)

    def _forward_mmt(self, batch_dict, have_text=False):
        if not have_text:
            self._forward_text_bert(batch_dict)
        batch_dict.update(self.mmt(batch_dict, fixed_ans_emb=self.classifier.weight))

    def _forward_output(self, batch_dict):
        p = self.ocr_ptr_net
        seq = batch_dict["mmt_seq_output"]
        R, D = batch_dict["mmt_ocr_output"].size(1), batch_dict["mmt_dec_output"].size(1)
        ocr_off = seq.size(1) - D - R                          # [txt ; obj ; ocr ; dec] (sa_m4c.py:852-862)
        batch_dict["scores"] = ops.OutputFn.apply(
            seq, ocr_off, R, D, batch_dict["pad_ocr_mask"],
            self.classifier.weight, self.classifier.bias, p.query.weight, p.query.bias, p.key.weight, p.key.bias)

    def _forward_mmt_and_output(self, batch_dict, have_text=False):
        if self.training:
            self._forward_mmt(batch_dict, have_text)
            self._forward_output(batch_dict)
            return
        if os.environ.get("SAMK_GREEDY", "cached") != "reference" and not torch.is_grad_enabled():
            return self._greedy_decode_cached(batch_dict)
        # greedy decoding exactly as the reference runs it (sa_m4c.py:285-302): D full passes
        dec_step_num = batch_dict["train_prev_inds"].size(1)
        batch_dict["train_prev_inds"] = torch.zeros_like(batch_dict["train_prev_inds"])
        batch_dict["train_prev_inds"][:, 0] = registry.BOS_IDX
        for _ in range(dec_step_num):
            self._forward_mmt(batch_dict)
            self._forward_output(batch_dict)
            argmax_inds = ops.argmax_rows(batch_dict["scores"].detach())
            batch_dict["train_prev_inds"][:, 1:] = argmax_inds[:, :-1]

    def _greedy_decode_cached(self, batch_dict):
        """Greedy decoding with the encoder computed once.

        The reference re-runs TextBert and all 182 rows of every layer for each of the D steps
        (sa_m4c.py:294-296).  Rows of the question / object / OCR segments never see decoder keys
        (dec_mask = 0, :793-795), so they are identical in every step: step 0 runs all rows and keeps each
        layer's fused q|k|v; steps 1..D-1 recompute only the D decoder rows of every layer against the cached
        keys/values.  Same kernels, same arithmetic per row -> same tokens and logits as the D-pass loop."""
        mmt = self.mmt
        prev = torch.zeros_like(batch_dict["train_prev_inds"])
        prev[:, 0] = registry.BOS_IDX
        batch_dict["train_prev_inds"] = prev
        B, D = prev.shape
        text_bert_out = self.text_bert(batch_dict)
        if not isinstance(self.text_bert_out_linear, nn.Identity):
            lin = self.text_bert_out_linear
            text_bert_out = ops.linear(text_bert_out.reshape(-1, text_bert_out.shape[-1]), lin.weight, lin.bias).view(B, -1, lin.weight.shape[0])
        batch_dict["text_bert_emb"] = txt = text_bert_out
        obj, ocr = batch_dict["obj_mmt_in"], batch_dict["ocr_mmt_in"]
        T, O, R = txt.size(1), obj.size(1), ocr.size(1)
        L, d = T + O + R + D, txt.size(2)
        dev = txt.device
        key_valid = ops.key_valid_bytes(batch_dict["question_mask"], batch_dict["pad_obj_mask"], batch_dict["pad_ocr_mask"], D)
        rel_cache = batch_dict.setdefault("_samk_rel_bits", {})

        def rel_lookup(key):
            if key not in rel_cache:
                rel_cache[key] = _relation_bits(batch_dict, key, dev)
            return rel_cache[key]

        dims = (B, L, T, O + R, D)
        mask_cache, caches, seq = {}, None, None
        for step in range(D):
            dec_emb = mmt.prev_pred_embeddings(self.classifier.weight, ocr, prev)
            if step == 0:
                x = ops.join_segments(txt, obj, ocr, dec_emb).view(B * L, d)
                out, caches = _encoder_infer(mmt.encoder, x, key_valid, rel_lookup, dims, mask_cache)
                seq = out.view(B, L, d)
            else:
                out, caches = _encoder_infer(mmt.encoder, dec_emb.reshape(B * D, d).contiguous(), key_valid, rel_lookup,
                                             dims, mask_cache, caches)
                seq[:, L - D:, :] = out.view(B, D, d)
            batch_dict.update({"mmt_seq_output": seq, "mmt_txt_output": seq[:, :T],
                               "mmt_ocr_output": seq[:, T + O:T + O + R], "mmt_dec_output": seq[:, L - D:]})
            self._forward_output(batch_dict)
            prev[:, 1:] = ops.argmax_rows(batch_dict["scores"])[:, :-1]

    def _forward_beam_search(self, batch_dict):
        """Beam-search decoding (sa_m4c.py:304-314 + sam/beam_search.py), on the cached decoder: the rows of the question /
        object / OCR segments do not depend on the decoded prefix, so they are encoded once (step 0), not once per
        step; every beam's decoder rows are recomputed from its current prefix each step, and since the encoder rows
        of a sample's beams are identical, re-ordering the beams only permutes the prefixes (the caches stay put).  Selection per step is samk_beam_step (log sigmoid + accumulated beam score, completed
        beams continue with EOS only, step 0 looks at one beam per sample).
        Outputs as the reference's evaluator reads them (evaluator.py:137-160): `complete_seqs` [B*K, D] (BOS-first
        token sequences), `topkscores` [B*K, 1], `train_prev_inds` = the sequences, `scores` of the last pass.
        Two defects of the reference's disabled decoder are not reproduced: `indices / vocab_size` is an integer
        division here (beam_search.py:109 yields fractional beam indices on torch >= 1.5) and a beam's score is the
        accumulated log-probability once (beam_search.py:124 adds the parent's score a second time)."""
        mmt = self.mmt
        K = int(getattr(self, "beam_size", 1) or 1)
        eos = int(getattr(registry, "EOS_IDX", 2))
        B, D = batch_dict["train_prev_inds"].shape

        def rep(v):
            return v.repeat_interleave(K, dim=0) if torch.is_tensor(v) else v
        prev = batch_dict["train_prev_inds"].new_zeros((B * K, D))
        prev[:, 0] = registry.BOS_IDX
        text_bert_out = self.text_bert(batch_dict)
        if not isinstance(self.text_bert_out_linear, nn.Identity):
            lin = self.text_bert_out_linear
            text_bert_out = ops.linear(text_bert_out.reshape(-1, text_bert_out.shape[-1]), lin.weight, lin.bias).view(B, -1, lin.weight.shape[0])
        batch_dict["text_bert_emb"] = text_bert_out
        txt, obj, ocr = rep(text_bert_out), rep(batch_dict["obj_mmt_in"]), rep(batch_dict["ocr_mmt_in"])
        T, O, R = txt.size(1), obj.size(1), ocr.size(1)
        L, d = T + O + R + D, txt.size(2)
        dev = txt.device
        key_valid = ops.key_valid_bytes(rep(batch_dict["question_mask"]), rep(batch_dict["pad_obj_mask"]),
                                        rep(batch_dict["pad_ocr_mask"]), D)
        ocr_mask = rep(batch_dict["pad_ocr_mask"])
        rel_cache = {}

        def rel_lookup(key):
            if key not in rel_cache:
                rel_cache[key] = _relation_bits(batch_dict, key, dev).repeat_interleave(K, dim=0).contiguous()
            return rel_cache[key]

        dims = (B * K, L, T, O + R, D)
        beam_scores = torch.zeros(B * K, dtype=torch.float32, device=dev)
        completed = None
        mask_cache, caches, seq = {}, None, None
        p = self.ocr_ptr_net
        for t in range(D):
            dec_emb = mmt.prev_pred_embeddings(self.classifier.weight, ocr, prev)
            if t == 0:
                x = ops.join_segments(txt, obj, ocr, dec_emb).view(B * K * L, d)
                out, caches = _encoder_infer(mmt.encoder, x, key_valid, rel_lookup, dims, mask_cache)
                seq = out.view(B * K, L, d)
            else:
                out, caches = _encoder_infer(mmt.encoder, dec_emb.reshape(B * K * D, d).contiguous(), key_valid, rel_lookup,
                                             dims, mask_cache, caches)
                seq[:, L - D:, :] = out.view(B * K, D, d)
            scores = ops.OutputFn.apply(seq, T + O, R, D, ocr_mask, self.classifier.weight, self.classifier.bias,
                                        p.query.weight, p.query.bias, p.key.weight, p.key.bias)
            ncls = scores.shape[-1]
            prev_pos, new_pos, beam_scores = ops.beam_step(scores[:, t, :], D * ncls, beam_scores, completed, eos, t == 0, B, K)
            prev = prev[prev_pos]
            if t + 1 < D:
                prev[:, t + 1] = new_pos
                completed = (prev[:, t + 1] == eos).to(torch.uint8).contiguous()
                if bool(completed.all()):
                    break
        batch_dict.update({"mmt_seq_output": seq, "mmt_txt_output": seq[:, :T], "mmt_ocr_output": seq[:, T + O:T + O + R],
                           "mmt_dec_output": seq[:, L - D:], "scores": scores, "train_prev_inds": prev,
                           "complete_seqs": prev, "topkscores": beam_scores.view(-1, 1)})

    def _forward_aux(self, batch_dict):
        T = batch_dict["question_mask"].size(-1)
        A = batch_dict["pad_obj_mask"].size(-1) + batch_dict["pad_ocr_mask"].size(-1)
        ent = batch_dict["mmt_seq_output"][:, T:T + A, :]
        o = self.origin_transform(ent).unsqueeze(-2)
        dst = self.dest_transform(ent).unsqueeze(-3)
        if self.aux_spatial_fusion == "mul":
            fused = o * dst
        elif self.aux_spatial_fusion == "add":
            fused = o + dst
        else:
            raise ValueError
        batch_dict["spatial_head_out"] = self.spatial_classifier(fused)

    def get_optimizer_parameters(self, base_lr):
        """sa_m4c.py:349-371: [rest @ base lr] + one group per finetune module."""
        groups, seen = [], set()
        for m in self.finetune_modules:
            groups.append({"params": list(m["module"].parameters()), "lr": base_lr * m["lr_scale"]})
            seen.update(list(m["module"].parameters()))
        groups.insert(0, {"params": [p for p in self.parameters() if p not in seen]})
        return groups
