"""Pins oracle/preprocess.py against golden vectors minted from the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import preprocess as op

WHOLE = ["dummy", "c1", "c2", "sim5", "sim5_trivial"]


@pytest.mark.parametrize("case", ["trivial_c1", "trivial_sim"])
def test_remove_trivial_cases(golden, case):
    g = golden(case)
    q, t, b = op.remove_trivial_cases(g["unfiltered/q"], g["unfiltered/t"], g["unfiltered/bits"],
                                      g["genome_of"])
    o = op.canonical_order(q, t)
    assert np.array_equal(np.stack((q[o], t[o])), g["filtered/edge_index"])     # bit-exact
    assert np.array_equal(b[o], g["filtered/bits"])
    assert q.size < g["unfiltered/q"].size                                       # filter did something


@pytest.mark.parametrize("case", WHOLE)
def test_normalize_sim_scores(golden, case):
    g = golden(case)
    src, dst, w = op.normalize_sim_scores(g["raw/q"], g["raw/t"], g["raw/bits"], g["genome_of"],
                                          temp=0.8, epsilon=1e-8, pseudo_count=1.0)
    assert np.array_equal(np.stack((src, dst)), g["norm/edge_index"])            # bit-exact index
    np.testing.assert_allclose(w, g["norm/w64"], rtol=1e-9, atol=0)
    # value-range contract of src/preprocessing.py:541
    assert w.min() >= 1.0 and w.max() <= 81.0 + 1e-9


@pytest.mark.parametrize("case", WHOLE)
def test_whole_graph_assembly(golden, case):
    g = golden(case)
    src, dst, w = op.normalize_sim_scores(g["raw/q"], g["raw/t"], g["raw/bits"], g["genome_of"])
    assert np.array_equal(np.stack((src, dst)), g["graph/edge_index"])
    assert np.array_equal(w.astype(np.float32), g["graph/edge_attr"])            # fp64 -> .float()
    y = op.map_labels(src, dst, g["group_of"])
    assert np.array_equal(y, g["graph/y"])
    assert op.class_balance(y) == pytest.approx(float(g["graph/class_balance"]), rel=1e-6)
    nb = op.neighbour_band(int(g["num_genes"]), int(g["neighbours"]))
    assert np.array_equal(nb, g["graph/neighbour_edge_index"])
    n, N = int(g["neighbours"]), int(g["num_genes"])
    assert nb.shape[1] == (2 * n + 1) * N - n * (n + 1)


@pytest.mark.parametrize("case", WHOLE)
def test_baseline_labels(golden, case):
    g = golden(case)
    src, dst, w = op.normalize_sim_scores(g["raw/q"], g["raw/t"], g["raw/bits"], g["genome_of"])
    bl = op.baseline_labels(src, dst, w, g["genome_of"])
    assert np.array_equal(bl, g["graph/base_labels"])
    # raw baseline scans the raw dict incl. self hits (src/helper.py:470-475)
    q, t, b = g["raw/q"], g["raw/t"], g["raw/bits"]
    blr_all = op.baseline_labels(q, t, b, g["genome_of"])
    key = {(int(a), int(c)): int(v) for a, c, v in zip(q, t, blr_all)}
    blr = np.asarray([key[(int(a), int(c))] for a, c in zip(src, dst)])
    assert np.array_equal(blr, g["graph/base_labels_raw"])


@pytest.mark.parametrize("variant,case", [("union_skip", "c1"), ("union_n4", "sim5")])
def test_union_assembly(golden, variant, case):
    g = golden(case)
    n = {"union_skip": 3, "union_n4": 4}[variant]
    nb = op.neighbour_band(int(g["num_genes"]), n)
    ei, w = op.union_whole_graph(g["graph/edge_index"], g["graph/edge_attr"], nb)
    assert np.array_equal(ei, g[f"model/{variant}/union_edge_index"])
    assert np.array_equal(w, g[f"model/{variant}/edge_attr"])


def test_dedupe_and_centre():
    q = np.array([0, 0, 1, 0]); t = np.array([1, 2, 0, 1]); b = np.array([5., 6., 7., 9.])
    q2, t2, b2 = op.dedupe_last(q, t, b)
    assert sorted(zip(q2, t2, b2)) == [(0, 1, 9.), (0, 2, 6.), (1, 0, 7.)]
    assert np.array_equal(op.center_scores(np.array([10., 12., 30.])), np.array([1., 3., 21.]))


def test_collate_offsets():
    g1 = dict(x=np.ones((3, 1), np.float32), edge_index=np.array([[0, 1], [1, 2]]),
              edge_attr=np.array([2., 3.], np.float32), y=np.array([1., 0.], np.float32),
              neighbour_edge_index=np.array([[0], [2]]))
    g2 = dict(x=np.ones((2, 1), np.float32), edge_index=np.array([[1], [0]]),
              edge_attr=np.array([5.], np.float32), y=np.array([1.], np.float32),
              neighbour_edge_index=np.array([[0, 1], [1, 0]]))
    b = op.collate([g1, g2])
    assert b["x"].shape == (5, 1)
    assert np.array_equal(b["edge_index"], np.array([[0, 1, 4], [1, 2, 3]]))
    assert np.array_equal(b["neighbour_edge_index"], np.array([[0, 3, 4], [2, 4, 3]]))
    assert np.array_equal(b["edge_attr"], np.array([2., 3., 5.], np.float32))
    assert np.array_equal(b["batch"], np.array([0, 0, 0, 1, 1]))
    assert np.array_equal(b["ptr"], np.array([0, 3, 5]))
