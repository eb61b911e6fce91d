"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from sam_textvqa_b200 import synth
from sam_textvqa_b200.config import c3_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name))


def cfg1(V=500):
    mmt, tb = c3_config(layer_type_list=["s"], mix_list=["share3"], hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    return mmt, tb


def golden_batch(g):
    batch = {}
    for k in g.files:
        if k.startswith("batch/") and k not in ("batch/adj1", "batch/adj3"):
            batch[k[6:]] = torch.from_numpy(g[k])
    batch["spatial_adj_matrices"] = {"1": torch.from_numpy(g["batch/adj1"]),
                                     "3": torch.from_numpy(g["batch/adj3"])}
    return batch


def sam4c_state_shapes(mmt, tb, V, T=20):
    """(name, shape) list of the SAM4C state_dict contract."""
    d, f = mmt["hidden_size"], 3072
    out = []

    def lin(name, o, i):
        out.extend([(name + ".weight", (o, i)), (name + ".bias", (o,))])

    def ln(name):
        out.extend([(name + ".weight", (d,)), (name + ".bias", (d,))])

    def layer(pre, spatial=False):
        for n in ("query", "key", "value"):
            lin(pre + "attention.self." + n, d, d)
        if spatial and mmt.get("use_bias"):          # sa_m4c.py:439-443
            out.append((pre + "attention.self.biases.weight", (1, d)))
        lin(pre + "attention.output.dense", d, d)
        ln(pre + "attention.output.LayerNorm")
        lin(pre + "intermediate.dense", f, d)
        lin(pre + "output.dense", d, f)
        ln(pre + "output.LayerNorm")

    out.append(("text_bert.embeddings.word_embeddings.weight", (tb["vocab_size"], d)))
    out.append(("text_bert.embeddings.position_embeddings.weight", (512, d)))
    out.append(("text_bert.embeddings.token_type_embeddings.weight", (2, d)))
    ln("text_bert.embeddings.LayerNorm")
    for i in range(tb["num_hidden_layers"]):
        layer("text_bert.encoder.layer.%d." % i)
    lin("linear_obj_feat_to_mmt_in", d, mmt["obj_feature_size"])
    lin("linear_obj_bbox_to_mmt_in", d, 4)
    ln("obj_feat_layer_norm")
    ln("obj_bbox_layer_norm")
    lin("linear_ocr_feat_to_mmt_in", d, mmt["ocr_feature_size"])
    lin("linear_ocr_bbox_to_mmt_in", d, 4)
    ln("ocr_feat_layer_norm")
    ln("ocr_bbox_layer_norm")
    out.append(("mmt.prev_pred_embeddings.position_embeddings.weight", (100, d)))
    out.append(("mmt.prev_pred_embeddings.token_type_embeddings.weight", (5, d)))
    for n in ("ans_layer_norm", "ocr_layer_norm", "emb_layer_norm"):
        ln("mmt.prev_pred_embeddings." + n)
    for i in range(mmt["layer_type_list"].count("n")):
        layer("mmt.encoder.normal_layers.%d." % i)
    for i in range(mmt["layer_type_list"].count("s")):
        layer("mmt.encoder.spatial_layers.%d." % i, spatial=True)
    lin("ocr_ptr_net.query", mmt["ptr_query_size"], d)
    lin("ocr_ptr_net.key", mmt["ptr_query_size"], d)
    lin("classifier", V, d)
    return out


def rel_err(a, b, mask=None):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    if mask is not None:
        a, b = a[mask], b[mask]
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def usebias_case():
    """Config, seeded weights and batch of tests/golden/sam4c_usebias.npz (oracle/make_golden.py:make_usebias_golden);
    the relation graph comes from the oracle builder and is checked against the types the reference produced."""
    from oracle import graph_oracle
    from sam_textvqa_b200 import synth
    from sam_textvqa_b200.config import c3_config
    g = load_golden("sam4c_usebias.npz")
    mmt, tb = c3_config(layer_type_list=["s", "s"], mix_list=["share3", "none"], use_bias=True,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    state = synth.seeded_state(sam4c_state_shapes(mmt, tb, 500), 4)
    graph_fn = lambda boxes: np.stack([graph_oracle.build_graph(b)["1"] for b in np.asarray(boxes)])
    batch = synth.make_batch(2, T=20, O=36, R=50, D=12, V=500, seed=3, contexts=(1, 3), graph_fn=graph_fn)
    assert np.array_equal(batch["spatial_types"].numpy(), g["types"])
    return g, mmt, tb, state, batch
