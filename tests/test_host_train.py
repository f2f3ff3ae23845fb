"""Host-side logic of the re-hosted training loop (pangnn_b200/train.py, setup.py): metric formulas, the
Youden threshold of --dynamic_binary_threshold, flags that are accepted but inert.  No GPU needed."""
import logging

import numpy as np
import pytest
import torch


def test_prf_never_divides_by_zero():
    """tp == 0 with fp > 0 and fn > 0 used to raise in evaluate() (eps = 0, src/predict.py:114-121)."""
    from pangnn_b200.train import prf
    assert prf(10, 3, 4, 0, eps=0.0) == (0.0, 0.0, 0.0)
    assert prf(10, 0, 0, 0, eps=0.0) == (0.0, 0.0, 0.0)
    assert prf(1, 1, 1, 1, eps=0.0) == (0.5, 0.5, 0.5)
    p, r, f = prf(5, 1, 3, 4)                                   # training-loop epsilons, pangnn.py:291-294
    assert p == pytest.approx(4 / (5 + 1e-10)) and r == pytest.approx(4 / (7 + 1e-10))
    assert f == pytest.approx(2 * p * r / (p + r + 1e-10))


@pytest.mark.parametrize("n", [10, 200, 5000])
def test_youden_threshold_matches_sklearn(n):
    from sklearn.metrics import roc_curve
    from pangnn_b200.train import youden_threshold
    g = torch.Generator().manual_seed(n)
    y = (torch.rand(n, generator=g) < 0.3).float()
    p = (torch.sigmoid(torch.randn(n, generator=g) + 2 * y) * 50).round() / 50      # many ties
    fpr, tpr, th = roc_curve(y.numpy(), p.numpy())
    assert youden_threshold(p, y) == pytest.approx(float(th[np.argmax(tpr - fpr)]))
    assert youden_threshold(p, torch.zeros(n)) is None          # one class only: keep the old threshold


def test_inert_reference_flags_are_reported(caplog):
    from pangnn_b200 import setup
    try:
        with caplog.at_level(logging.WARNING, logger="pangnn"):
            a = setup.parse(["--mixed_precision", "bf16", "--to_pickle", "x.pkl", "--cpus", "3"])
        assert a.cpus == 3
        assert set(setup.inert_flags(a)) == {"mixed_precision", "to_pickle"}
        assert sum("has no effect" in r.message for r in caplog.records) == 2
    finally:
        setup.reset()
    assert setup.inert_flags(setup.args) == []


def test_scored_chunk_runs_cover_every_edge_once():
    """ops.scored_chunk_runs (the chunked scorer's split of the source rows): runs tile the edge range in order, stay
    within the bound except for single over-long rows, skip empty stretches."""
    import numpy as np
    from pangnn_b200.ops import scored_chunk_runs
    rng = np.random.RandomState(0)
    for trial in range(20):
        N = int(rng.randint(1, 400))
        deg = rng.randint(0, 12, size=N)
        deg[rng.rand(N) < 0.3] = 0
        if trial % 3 == 0:
            deg[rng.randint(0, N)] = 500                         # one row longer than any bound below
        rowptr = np.concatenate(([0], np.cumsum(deg))).astype(np.int64)
        for bound in (1, 7, 64, 10**6):
            runs = scored_chunk_runs(rowptr, bound)
            assert sum(c1 - c0 for _, _, c0, c1 in runs) == rowptr[-1]
            pos = 0
            for r0, r1, c0, c1 in runs:
                assert r0 < r1 and c0 == rowptr[r0] and c1 == rowptr[r1] and c1 > c0
                assert c0 >= pos                                   # in order, no overlap (gaps = empty rows only)
                assert (rowptr[r0:r1 + 1][-1] - rowptr[r0]) <= bound or r1 - r0 == 1
                pos = c1
            covered = np.zeros(N, bool)
            for r0, r1, _, _ in runs:
                covered[r0:r1] = True
            assert not deg[~covered].any()                         # rows outside every run have no edges
    assert scored_chunk_runs(np.zeros(1, np.int64), 5) == [] and scored_chunk_runs(np.zeros(6, np.int64), 5) == []


def test_graphed_batch_padding_on_host_tensors():
    """GraphedBatchStep's bucket key and padding (pure tensor logic, no capture): pad nodes isolated, pad edges are
    self loops on the last pad node, labels / weights padded, mask and 1 / E set."""
    import torch
    from pangnn_b200.data import Data
    from pangnn_b200.graphs import GraphedBatchStep
    st = object.__new__(GraphedBatchStep)
    st.nq, st.eq = 64, 128
    n, E, Enb = 70, 130, 200
    b = Data(torch.ones(n, 1), torch.randint(0, n, (2, E)), torch.rand(E) + 1, (torch.rand(E) < 0.3).float())
    b.neighbour_edge_index = torch.randint(0, n, (2, Enb))
    b.node_id = torch.arange(n)
    b.batch = torch.zeros(n, dtype=torch.long)
    key = st._bucket_key(b)
    assert dict(key) == {"#nodes": 128, "edge_index": 256, "edge_attr": 256, "y": 256, "neighbour_edge_index": 256}
    g = st._alloc(key, b)
    st._load(g, key, b)
    assert g.x.shape == (128, 1) and bool((g.x == 1).all()) and g.node_id.shape == (128,)
    assert torch.equal(g.edge_index[:, :E], b.edge_index) and bool((g.edge_index[:, E:] == 127).all())
    assert torch.equal(g.neighbour_edge_index[:, :Enb], b.neighbour_edge_index)
    assert bool((g.neighbour_edge_index[:, Enb:] == 127).all())
    assert torch.equal(g.edge_attr[:E], b.edge_attr) and bool((g.edge_attr[E:] == 1).all())
    assert torch.equal(g.y[:E], b.y) and bool((g.y[E:] == 0).all())
    assert float(g._mask.sum()) == E and bool((g._mask[:E] == 1).all()) and abs(float(g._inv) - 1 / E) < 1e-9
    # a smaller batch into the same buffers: the previous batch's tail is overwritten by padding
    b2 = Data(torch.ones(66, 1), torch.randint(0, 66, (2, 129)), torch.rand(129) + 1, torch.zeros(129))
    b2.neighbour_edge_index, b2.node_id = torch.randint(0, 66, (2, 150)), torch.arange(66)
    assert st._bucket_key(b2) == key
    st._load(g, key, b2)
    assert bool((g.edge_index[:, 129:] == 127).all()) and float(g._mask.sum()) == 129


def test_local_numbering_matches_the_plain_formulation():
    """dist.local_numbering (in-place, memory-lean) == mask + unique + searchsorted + where written out plainly;
    the input list is left untouched."""
    import torch
    from pangnn_b200.dist import local_numbering

    def plain(edge_index, lo, hi, anchor):
        a = edge_index[1] if anchor == "dst" else edge_index[0]
        keep = (a >= lo) & (a < hi)
        ei = edge_index[:, keep]
        other = ei[0] if anchor == "dst" else ei[1]
        halo = torch.unique(other[(other < lo) | (other >= hi)])

        def loc(ids):
            own = (ids >= lo) & (ids < hi)
            pos = torch.searchsorted(halo, ids) if halo.numel() else torch.zeros_like(ids)
            return torch.where(own, ids - lo, (hi - lo) + pos)
        return keep, halo, torch.stack((loc(ei[0]), loc(ei[1])))
    g = torch.Generator().manual_seed(0)
    for _ in range(60):
        N = int(torch.randint(1, 300, (1,), generator=g))
        E = int(torch.randint(0, 1500, (1,), generator=g))
        ei = torch.randint(0, N, (2, E), generator=g)
        lo = int(torch.randint(0, N, (1,), generator=g))
        hi = int(torch.randint(lo, N + 1, (1,), generator=g))
        for anchor in ("dst", "src"):
            before = ei.clone()
            k1, h1, e1 = local_numbering(ei, lo, hi, anchor)
            k2, h2, e2 = plain(ei, lo, hi, anchor)
            assert torch.equal(ei, before)
            assert torch.equal(k1, k2) and torch.equal(h1, h2) and torch.equal(e1, e2)
