"""Oracle pinning: numpy graph restatement vs the reference's own outputs."""
import numpy as np
import pytest

from oracle import graph_oracle as G
from oracle.ref_loader import load_reference, reference_available
from sam_textvqa_b200 import synth
from tests._util import load_golden


def _sets(g):
    return sorted({k.split("/")[0] for k in g.files})


def test_oracle_matches_golden_graphs():
    g = load_golden("graph_kat.npz")
    for name in _sets(g):
        got = G.build_graph(g[name + "/boxes"])
        for key in G.SHARED_KEYS:
            assert np.array_equal(got[key], g[name + "/m" + key]), (name, key)
        for c in (1, 3, 5):
            assert np.array_equal(G.expand_context(got, c), g[name + "/heads%d" % c]), (name, c)


def test_closed_form_head_bits_equal_reference_chain():
    g = load_golden("graph_kat.npz")
    for name in _sets(g):
        types = g[name + "/m1"]
        for c in (1, 3, 5):
            bits = G.head_bits_closed_form(types, c)
            heads = ((bits[..., None] >> np.arange(12)) & 1).astype(np.int8)
            assert np.array_equal(heads, g[name + "/heads%d" % c]), (name, c)
            lut = synth.head_bits_for_context(c)
            assert np.array_equal(lut[types.astype(np.int64)], bits)


def test_known_answers_appendix_a():
    g = load_golden("graph_kat.npz")
    expect = {
        "identical": [[12, 3], [3, 12]], "same_centre_cross": [[12, 4], [4, 12]],
        "strict_containment": [[12, 1], [2, 12]], "touching_edge": [[12, 4], [8, 12]],
        "dy0_right": [[12, 3], [7, 12]], "dy0_left": [[12, 7], [11, 12]],
        "dx0_below": [[12, 5], [9, 12]], "diag": [[12, 4], [8, 12]], "far": [[12, 0], [0, 12]],
        "pad_middle": [[12, 0, 7], [0, 0, 0], [11, 0, 12]],
    }
    for k, v in expect.items():
        assert g["kat_" + k + "/m1"].tolist() == v, k
        assert G.build_graph(g["kat_" + k + "/boxes"])["1"].tolist() == v, k


def test_zero_union_pair_takes_nan_branch():
    b = np.array([[.5, .5, .5, .5], [.5, .5, .5, .5]])
    assert G.build_graph(b)["1"].tolist() == [[12, 4], [4, 12]]


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference_on_fresh_boxes():
    _, S, _ = load_reference()
    rs = np.random.RandomState(99)
    for trial in range(4):
        b = synth.make_boxes(rs, 1, 40)[0, :, :4].astype(np.float64)
        if trial % 2:
            b = (np.round(b * 16) / 16).astype(np.float32).astype(np.float64)
            b[:, 2:] = np.maximum(b[:, 2:], b[:, :2] + 1 / 32)
        b[36:] = 0
        ref = S.build_graph_using_normalized_boxes(b)
        got = G.build_graph(b)
        for key in G.SHARED_KEYS:
            assert np.array_equal(ref[key], got[key]), (trial, key)
