"""GPU parity: spatial-graph kernel vs the reference goldens and the numpy oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as G
from sam_textvqa_b200 import synth
from tests._util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from sam_textvqa_b200 import spatial_utils
    return spatial_utils


def test_golden_box_sets_all_nine_matrices_and_head_masks(S):
    g = load_golden("graph_kat.npz")
    for name in sorted({k.split("/")[0] for k in g.files}):
        boxes = g[name + "/boxes"]
        out = S.build_graph_using_normalized_boxes(boxes)
        for key in G.SHARED_KEYS:
            assert out[key].dtype == np.int8 and np.array_equal(out[key], g[name + "/m" + key]), (name, key)
        for c in (1, 3, 5):
            types, _, bits = S.build_graph_batch(boxes[None], 0.5, context=c)
            heads = S.expand_context(types, c)[0].cpu().numpy()
            assert np.array_equal(heads, g[name + "/heads%d" % c]), (name, c)
            hb = bits[0].cpu().numpy().astype(np.uint16)
            assert np.array_equal(((hb[..., None] >> np.arange(12)) & 1).astype(np.int8), heads)
        bc = S.torch_broadcast_adj_matrix(torch.from_numpy(g[name + "/m1"]))
        assert np.array_equal(bc.numpy(), g[name + "/heads1"])


def test_f32_and_f64_paths_agree_with_oracle_full_size(S):
    rs = np.random.RandomState(7)
    B, N = 32, 150
    bx = synth.make_boxes(rs, B, N)[..., :4]
    bx[:, 137:] = 0
    bx[3] = 0                                              # all-pad sample
    ref = np.stack([G.build_graph(b.astype(np.float64))["1"] for b in bx])
    t32 = S.build_graph_batch(bx.astype(np.float32))[0].cpu().numpy()
    t64 = S.build_graph_batch(bx.astype(np.float64))[0].cpu().numpy()
    assert np.array_equal(t32, ref) and np.array_equal(t64, ref)
    assert (ref[3] == 0).all()


def test_adversarial_grid_aligned_boxes(S):
    rs = np.random.RandomState(11)
    for _ in range(4):
        n = 96
        cx, cy = rs.randint(2, 15, n) / 16.0, rs.randint(2, 15, n) / 16.0
        hw, hh = rs.randint(0, 4, n) / 32.0, rs.randint(0, 4, n) / 32.0       # includes zero-area boxes
        b = np.stack([cx - hw, cy - hh, cx + hw, cy + hh], 1).astype(np.float32).astype(np.float64)
        ref = G.build_graph(b)
        out = S.build_graph_using_normalized_boxes(b)
        for key in G.SHARED_KEYS:
            assert np.array_equal(out[key], ref[key]), key


def test_empty_and_size_independent_properties(S):
    types, _, _ = S.build_graph_batch(np.zeros((0, 5, 4)))
    assert types.shape == (0, 5, 5)
    rs = np.random.RandomState(3)
    bx = synth.make_boxes(rs, 128, 150)[..., :4]
    t = S.build_graph_batch(bx.astype(np.float32))[0]
    diag = torch.diagonal(t, dim1=1, dim2=2)
    assert (diag == 12).all()
    off = t.clone()
    off.diagonal(dim1=1, dim2=2).zero_()
    assert int(off.max()) <= 11 and int(off.min()) >= 0
    # relation present one way <=> present the other way; containment types pair up as (1,2)
    assert ((off != 0) == (off.transpose(1, 2) != 0)).all()
    assert ((off == 1) == (off.transpose(1, 2) == 2)).all()
    assert ((off == 3) == (off.transpose(1, 2) == 3)).all()
    # permuting the boxes permutes the directional structure consistently for non-boundary pairs
    perm = torch.randperm(150)
    tp = S.build_graph_batch(bx[:, perm.numpy()].astype(np.float32))[0]
    same = (tp != 0) == (t[:, perm][:, :, perm] != 0)
    assert same.all()
