"""Oracle pinning: functional SA-M4C restatement vs goldens produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import sam4c_oracle as O
from sam_textvqa_b200 import synth
from tests._util import cfg1, golden_batch, load_golden, rel_err, sam4c_state_shapes, usebias_case


@pytest.fixture(scope="module")
def setup():
    g = load_golden("sam4c_cfg1.npz")
    mmt, tb = cfg1()
    P = synth.seeded_state(sam4c_state_shapes(mmt, tb, V=500), 0)
    return g, mmt, tb, P, golden_batch(g)


def test_teacher_forced_scores_loss_and_grads(setup):
    g, mmt, tb, P, batch = setup
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    scores, _, seq = O.forward(P, batch, mmt, tb, train=True)
    assert rel_err(scores, g["tf/scores"]) < 2e-6
    assert rel_err(seq, g["tf/mmt_seq_output"]) < 2e-5
    loss = O.bce_with_mask_loss(scores, batch["targets"], batch["train_loss_mask"])
    assert abs(loss.item() - float(g["tf/loss"])) < 1e-4 * float(g["tf/loss"])
    loss.backward()
    for k in g.files:
        if k.startswith("grad/"):
            got = P[k[5:]].grad
            ref = torch.from_numpy(g[k])
            if got.numel() > 70000:
                got = got.flatten()[:: max(1, got.numel() // 4096)]
            assert rel_err(got, ref) < 2e-4, k
    # padded OCR slots are raw - 10000 (sa_m4c.py:879,893)
    pad = batch["pad_ocr_mask"] == 0
    assert (scores[:, :, 500:][pad[:, None, :].expand(-1, 12, -1)] < -9000).all()


def test_greedy_decode_tokens(setup):
    g, mmt, tb, P, batch = setup
    with torch.no_grad():
        scores, prev, _ = O.forward(P, batch, mmt, tb, train=False)
    assert np.array_equal(prev.numpy(), g["greedy/prev_inds"])
    assert rel_err(scores, g["greedy/scores"]) < 2e-5


def test_text_rows_of_spatial_layer_are_dead(setup):
    """Known-answer fact: with quadrants [1,2] every text query row is fully masked."""
    g = load_golden("attn_unit.npz")
    T = int(g["T"])
    assert np.abs(g["ctx"][:, :T]).max() == 0.0


def test_context_biases_use_bias_true():
    """`use_bias: true` (sa_m4c.py:439-443, 600-603): two spatial layers, logits / loss / gradients of the unmodified
    reference."""
    g, mmt, tb, P, batch = usebias_case()
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    scores, _, _ = O.forward(P, batch, mmt, tb, train=True)
    assert rel_err(scores, g["tf/scores"]) < 2e-6
    loss = O.bce_with_mask_loss(scores, batch["targets"], batch["train_loss_mask"])
    assert abs(loss.item() - float(g["tf/loss"])) < 1e-4 * float(g["tf/loss"])
    loss.backward()
    for k in g.files:
        if k.startswith("grad/"):
            got = P[k[5:]].grad
            if got.numel() > 70000:
                got = got.flatten()[:: max(1, got.numel() // 4096)]
            assert rel_err(got, torch.from_numpy(g[k])) < 2e-4, k


def test_shipped_c3_layer_schedule_full_stack():
    """n,n,s,s,s,s with 100 objects (L = 182) against the unmodified reference (tests/golden/sam4c_c3.npz): the same
    model (seed-1 weights) and batch (synth.make_batch(3, seed=5)) as the GPU test of the full c3 stack, so what that
    test compares against is pinned here."""
    from oracle import graph_oracle
    from sam_textvqa_b200.config import c3_config
    g = load_golden("sam4c_c3.npz")
    mmt, tb = c3_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    P = synth.seeded_state(sam4c_state_shapes(mmt, tb, 500), 1)
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    graph_fn = lambda boxes: np.stack([graph_oracle.build_graph(b)["1"] for b in np.asarray(boxes)])
    batch = synth.make_batch(3, V=500, seed=5, contexts=(1, 3), graph_fn=graph_fn)
    assert np.array_equal(batch["spatial_types"].numpy(), g["types"])
    scores, _, _ = O.forward(P, batch, mmt, tb, train=True)
    assert rel_err(scores, g["tf/scores"]) < 5e-6
    assert torch.equal(scores.argmax(-1), torch.from_numpy(g["tf/scores"]).argmax(-1))
    loss = O.bce_with_mask_loss(scores, batch["targets"], batch["train_loss_mask"])
    assert abs(loss.item() - float(g["tf/loss"])) < 1e-4 * float(g["tf/loss"])
    loss.backward()
    for k in g.files:
        if k.startswith("grad/"):
            got = P[k[5:]].grad
            if got.numel() > 70000:
                got = got.flatten()[:: max(1, got.numel() // 4096)]
            assert rel_err(got, torch.from_numpy(g[k])) < 5e-4, k
