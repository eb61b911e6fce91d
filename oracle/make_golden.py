"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The fixtures are small (cfg1 geometry, B=2..4) and are what pins parity on the GPU box, where the
reference tree does not exist.

  graph_kat.npz     Appendix-A known-answer boxes + random + grid-aligned adversarial sets
                    -> the nine int8 matrices of build_graph_using_normalized_boxes and the
                    [N,N,12] head masks for c = 1,3,5 built exactly like
                    sam/datasets/textvqa_dataset.py:373-409.
  sam4c_cfg1.npz    BASELINE config 0: 1 spatial layer (share3), T=20,O=36,R=50,D=12, B=4,
                    V=500 (vocab size is a runtime value, sa_m4c.py:169; small keeps the fixture
                    under 10 MB), seed-0 weights: teacher-forced logits, loss, selected
                    gradients, greedy-decoded tokens, per-stage activations.
  attn_unit.npz     SpatialBertSelfAttention.forward on random hidden states, B=2.
  sam4c_c3.npz      the shipped layer schedule (n,n,s,s,s,s; share3), 100 objects (L = 182), B=3, V=500, seed-1
                    weights, batch = synth.make_batch(3, seed=5): logits, loss, gradients.  Same model and batch as
                    tests/test_gpu_model.py::test_full_c3_stack_vs_oracle_on_fresh_batch_with_cpu_resident_masks, so the
                    oracle output that GPU test compares against is itself pinned to the unmodified reference.
  sam4c_usebias.npz two spatial layers with `use_bias: true` (sa_m4c.py:439-443, 600-603; off in the shipped
                    configs), B=2, cfg1 geometry: logits, loss, gradients of the context biases and of the
                    out-projection they fold into.  The batch is regenerated in the tests from
                    synth.make_batch(seed=3) with the oracle graph builder; its relation types are stored
                    here as a cross-check.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle.ref_loader import load_reference  # noqa: E402
from sam_textvqa_b200 import synth  # noqa: E402
from sam_textvqa_b200.config import c3_config  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

KAT_BOXES = {
    "identical": [[.1, .1, .5, .5], [.1, .1, .5, .5]],
    "same_centre_cross": [[.1, .4, .9, .6], [.4, .1, .6, .9]],
    "strict_containment": [[.1, .1, .9, .9], [.2, .2, .3, .3]],
    "touching_edge": [[.1, .1, .9, .9], [.1, .2, .3, .3]],
    "dy0_right": [[.5, .1, .6, .2], [.1, .1, .2, .2]],
    "dy0_left": [[.1, .1, .2, .2], [.5, .1, .6, .2]],
    "dx0_below": [[.1, .5, .2, .6], [.1, .1, .2, .2]],
    "diag": [[.3, .3, .4, .4], [.1, .1, .2, .2]],
    "far": [[0, 0, .05, .05], [.9, .9, 1, 1]],
    "pad_middle": [[.1, .1, .2, .2], [0, 0, 0, 0], [.3, .1, .4, .2]],
    "sum_zero_is_pad": [[-.5, -.5, .5, .5], [.1, .1, .2, .2]],
}


def adversarial_boxes(rs, n):
    """Centres on a k/16 grid: |dx|==|dy|, dx==0, dy==0 pairs sit exactly on octant boundaries."""
    cx = rs.randint(2, 15, n) / 16.0
    cy = rs.randint(2, 15, n) / 16.0
    hw = rs.randint(1, 4, n) / 32.0
    hh = rs.randint(1, 4, n) / 32.0
    b = np.stack([cx - hw, cy - hh, cx + hw, cy + hh], 1).astype(np.float32)
    b[rs.rand(n) < 0.1] = 0
    return b.astype(np.float64)


def ref_heads(S, shared, context):
    m = S.torch_broadcast_adj_matrix(torch.from_numpy(shared["1"]))
    for c in (3, 5, 7, 9):
        if c > context:
            break
        for k in ("%d1" % c, "%d2" % c):
            m = torch.max(m, S.torch_broadcast_adj_matrix(torch.from_numpy(shared[k])))
    return m.numpy()


def ref_graph_fn(S):
    def fn(boxes):
        return np.stack([S.build_graph_using_normalized_boxes(b)["1"] for b in boxes])
    return fn


def make_graph_golden(S):
    out = {}
    rs = np.random.RandomState(1234)
    sets = {("kat_" + k): np.array(v, dtype=np.float64) for k, v in KAT_BOXES.items()}
    for i in range(3):
        sets["rand_%d" % i] = synth.make_boxes(rs, 1, 48)[0, :, :4].astype(np.float64)
        sets["rand_%d" % i][40 + i:] = 0
        sets["grid_%d" % i] = adversarial_boxes(rs, 48)
    for name, b in sets.items():
        shared = S.build_graph_using_normalized_boxes(b)
        out[name + "/boxes"] = b
        for k, v in shared.items():
            out[name + "/m" + k] = v
        for c in (1, 3, 5):
            out[name + "/heads%d" % c] = ref_heads(S, shared, c)
    np.savez_compressed(os.path.join(GOLD, "graph_kat.npz"), **out)
    print("graph_kat.npz:", len(sets), "box sets")


def build_ref_model(M, mmt, tb, seed=0):
    model = M.SAM4C(M.BertConfig.from_dict(mmt), M.BertConfig.from_dict(tb))
    # weights come from synth.seeded_state (keyed by parameter name), so no checkpoint has to
    # be shipped: every parity test regenerates the same state_dict from names + shapes.
    sd = model.state_dict()
    model.load_state_dict(synth.seeded_state([(k, v.shape) for k, v in sd.items()], seed), strict=True)
    return model


def make_sam4c_golden(M, S, registry):
    V = 500
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    mmt, tb = c3_config(layer_type_list=["s"], mix_list=["share3"], hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = build_ref_model(M, mmt, tb)
    batch = synth.make_batch(4, T=20, O=36, R=50, D=12, V=V, seed=0, contexts=(1, 3),
                             graph_fn=ref_graph_fn(S))
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out["batch/" + k] = v.numpy()
    out["batch/adj1"] = batch["spatial_adj_matrices"]["1"].numpy()
    out["batch/adj3"] = batch["spatial_adj_matrices"]["3"].numpy()

    # teacher-forced pass + loss + grads (train() with every dropout prob at 0)
    model.train()
    bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    scores = model(bd)["textvqa_scores"]
    losses = torch.nn.functional.binary_cross_entropy_with_logits(scores, batch["targets"], reduction="none")
    loss = (losses * batch["train_loss_mask"].unsqueeze(-1)).sum() / batch["train_loss_mask"].sum().clamp(min=1)
    loss.backward()
    out["tf/scores"] = scores.detach().numpy()
    out["tf/loss"] = loss.detach().numpy()
    for k in ("obj_mmt_in", "ocr_mmt_in", "text_bert_emb", "mmt_seq_output"):
        out["tf/" + k] = bd[k].detach().numpy()
    grads = dict((n, p.grad) for n, p in model.named_parameters() if p.grad is not None)
    for n in ["classifier.weight", "classifier.bias", "ocr_ptr_net.query.weight", "ocr_ptr_net.key.bias",
              "mmt.encoder.spatial_layers.0.attention.self.query.weight",
              "mmt.encoder.spatial_layers.0.attention.self.value.bias",
              "mmt.encoder.spatial_layers.0.attention.output.LayerNorm.weight",
              "mmt.encoder.spatial_layers.0.intermediate.dense.bias",
              "mmt.encoder.spatial_layers.0.output.dense.weight",
              "mmt.prev_pred_embeddings.emb_layer_norm.weight",
              "mmt.prev_pred_embeddings.position_embeddings.weight",
              "text_bert.encoder.layer.0.attention.self.key.weight",
              "text_bert.embeddings.LayerNorm.bias",
              "linear_obj_feat_to_mmt_in.bias", "linear_ocr_bbox_to_mmt_in.weight",
              "obj_feat_layer_norm.weight", "ocr_bbox_layer_norm.bias"]:
        g = grads[n]
        out["grad/" + n] = g.numpy() if g.numel() <= 70000 else g.flatten()[:: max(1, g.numel() // 4096)].numpy()
    out["grad_norm_total"] = np.array(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item())

    # greedy decoding (eval mode)
    model.eval()
    bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    with torch.no_grad():
        scores = model(bd)["textvqa_scores"]
    out["greedy/scores"] = scores.numpy()
    out["greedy/prev_inds"] = bd["train_prev_inds"].numpy()

    np.savez_compressed(os.path.join(GOLD, "sam4c_cfg1.npz"), **out)
    print("sam4c_cfg1.npz loss", float(loss), "greedy tokens", bd["train_prev_inds"][0].tolist())


def make_c3_golden(M, S, registry):
    V = 500
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    mmt, tb = c3_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = build_ref_model(M, mmt, tb, seed=1).train()
    batch = synth.make_batch(3, V=V, seed=5, contexts=(1, 3), graph_fn=ref_graph_fn(S))
    bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    scores = model(bd)["textvqa_scores"]
    losses = torch.nn.functional.binary_cross_entropy_with_logits(scores, batch["targets"], reduction="none")
    loss = (losses * batch["train_loss_mask"].unsqueeze(-1)).sum() / batch["train_loss_mask"].sum().clamp(min=1)
    loss.backward()
    out = {"types": batch["spatial_types"].numpy(), "tf/scores": scores.detach().numpy(), "tf/loss": loss.detach().numpy()}
    grads = dict((n, p.grad) for n, p in model.named_parameters() if p.grad is not None)
    for n in ["classifier.bias", "mmt.encoder.normal_layers.0.attention.self.query.weight",
              "mmt.encoder.normal_layers.1.output.dense.bias",
              "mmt.encoder.spatial_layers.0.attention.self.key.weight",
              "mmt.encoder.spatial_layers.3.intermediate.dense.weight",
              "text_bert.encoder.layer.2.attention.output.LayerNorm.weight",
              "linear_obj_feat_to_mmt_in.weight"]:
        g = grads[n]
        out["grad/" + n] = g.numpy() if g.numel() <= 70000 else g.flatten()[:: max(1, g.numel() // 4096)].numpy()
    np.savez_compressed(os.path.join(GOLD, "sam4c_c3.npz"), **out)
    print("sam4c_c3.npz loss", float(loss))


def make_usebias_golden(M, S, registry):
    V = 500
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    mmt, tb = c3_config(layer_type_list=["s", "s"], mix_list=["share3", "none"], use_bias=True,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = build_ref_model(M, mmt, tb, seed=4).train()
    batch = synth.make_batch(2, T=20, O=36, R=50, D=12, V=V, seed=3, contexts=(1, 3), graph_fn=ref_graph_fn(S))
    bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    scores = model(bd)["textvqa_scores"]
    losses = torch.nn.functional.binary_cross_entropy_with_logits(scores, batch["targets"], reduction="none")
    loss = (losses * batch["train_loss_mask"].unsqueeze(-1)).sum() / batch["train_loss_mask"].sum().clamp(min=1)
    loss.backward()
    out = {"types": batch["spatial_types"].numpy(), "tf/scores": scores.detach().numpy(), "tf/loss": loss.detach().numpy()}
    grads = dict((n, p.grad) for n, p in model.named_parameters() if p.grad is not None)
    for n in ["mmt.encoder.spatial_layers.0.attention.self.biases.weight",
              "mmt.encoder.spatial_layers.1.attention.self.biases.weight",
              "mmt.encoder.spatial_layers.0.attention.output.dense.bias",
              "mmt.encoder.spatial_layers.1.attention.output.dense.weight",
              "mmt.encoder.spatial_layers.0.attention.self.value.weight"]:
        g = grads[n]
        out["grad/" + n] = g.numpy() if g.numel() <= 70000 else g.flatten()[:: max(1, g.numel() // 4096)].numpy()
    np.savez_compressed(os.path.join(GOLD, "sam4c_usebias.npz"), **out)
    print("sam4c_usebias.npz loss", float(loss))


def make_attn_golden(M):
    mmt, _ = c3_config(attention_probs_dropout_prob=0.0)
    cfg = M.BertConfig.from_dict(mmt)
    attn = M.SpatialBertSelfAttention(cfg).eval()
    attn.load_state_dict(synth.seeded_state([(k, v.shape) for k, v in attn.state_dict().items()], 3))
    torch.manual_seed(3)
    B, T, A, D = 2, 20, 30, 12
    L = T + A + D
    rs = np.random.RandomState(5)
    hidden = torch.randn(B, L, 768)
    types = torch.from_numpy(rs.randint(0, 13, (B, A, A)).astype(np.int8))
    types[0, 5] = 0  # an entity row with no relation at all
    adj = synth.expand_types_to_heads(types, 3)
    valid = torch.ones(B, L, dtype=torch.long)
    valid[:, T + A:] = 0
    valid[0, 15:T] = 0
    valid[1, T + 20:T + A] = 0
    ext = valid[:, None, None, :].repeat(1, 1, L, 1).float()
    ext[:, :, -D:, -D:] = torch.tril(torch.ones(D, D))
    add = (1.0 - ext) * -10000.0
    with torch.no_grad():
        ctx = attn(hidden, add, adj)[0]
    out = {"hidden": hidden.numpy().astype(np.float32), "types": types.numpy(), "adj": adj.numpy(), "valid": valid.numpy(),
           "ctx": ctx.numpy(), "T": np.array(T), "A": np.array(A), "D": np.array(D)}
    np.savez_compressed(os.path.join(GOLD, "attn_unit.npz"), **out)
    print("attn_unit.npz ctx", tuple(ctx.shape))


def main():
    os.makedirs(GOLD, exist_ok=True)
    M, S, registry = load_reference()
    make_graph_golden(S)
    make_attn_golden(M)
    make_sam4c_golden(M, S, registry)
    make_usebias_golden(M, S, registry)
    make_c3_golden(M, S, registry)


if __name__ == "__main__":
    main()
