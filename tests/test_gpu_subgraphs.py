"""Device sub-graph extraction (a9, a10, sub-graph a11; pangnn_b200/subgraphs.py) against the reference-run golden
vectors (global ids, as a multiset) and, bit for bit including the canonical local numbering and the order of
the sub-graphs, against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import preprocess as op
from oracle import subgraphs as osg
from tests.test_oracle_subgraphs import check_against_golden, golden_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def arena_graphs(arena, union):
    """Arena -> the oracle's dict form (local ids)."""
    out = []
    for i in range(arena.num_graphs):
        g = arena.graph(i)
        d = dict(order=g.node_id.cpu().numpy(), sim_ei=g.edge_index.cpu().numpy(), y=g.y.cpu().numpy())
        e = d["sim_ei"].shape[1]
        if union:
            u = g.union_edge_index.cpu().numpy()
            d["nb_ei"], d["w"] = u[:, :u.shape[1] - e], g.edge_attr.cpu().numpy()[u.shape[1] - e:]
            assert np.array_equal(u[:, u.shape[1] - e:], d["sim_ei"])
            assert np.all(g.edge_attr.cpu().numpy()[:u.shape[1] - e] == 1.0)
        else:
            d["nb_ei"], d["w"] = g.neighbour_edge_index.cpu().numpy(), g.edge_attr.cpu().numpy()
        assert tuple(g.x.shape) == (d["order"].size, 1) and bool((g.x == 1).all())
        out.append(d)
    return out


def run_extract(src, dst, w, y, N, n, groups, union, subset, chunks=1):
    from pangnn_b200 import subgraphs
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt, device=DEV)
    return subgraphs.extract(t(src, torch.int32), t(dst, torch.int32), t(w, torch.float32), t(y, torch.float32),
                             N, n, groups, union=union, gff_is_subset=subset, chunks=chunks)


@pytest.mark.parametrize("union", [False, True])
@pytest.mark.parametrize("case", ["c1_sub", "sim5_sub"])
def test_sub_graphs_match_reference_goldens(golden, case, union):
    g = golden(case)
    src, dst, w, y, N, n, groups = golden_inputs(g)
    arena = run_extract(src, dst, w, y, N, n, groups, union, True)
    check_against_golden(g, arena_graphs(arena, union))


@pytest.mark.parametrize("n", [1, 2, 3])
@pytest.mark.parametrize("sim", [(40, 3, 0.5, 5, 2), (300, 4, 0.7, 10, 3), (2000, 2, 0.5, 10, 3)])
def test_sub_graphs_equal_oracle_bit_for_bit(sim, n):
    """Simulated pan-genomes through the oracle preprocessing; every sub-graph identical to the oracle's:
    order of sub-graphs, local numbering, edges, weights, labels, class balance."""
    from pangnn_b200.simulate import simulate_hits
    s = simulate_hits(*sim, seed=3)
    q, t, b = op.dedupe_last(s["q"].astype(np.int64), s["t"].astype(np.int64), s["bits"])
    q, t, b = op.remove_trivial_cases(q, t, b, s["genome_of"])
    src, dst, w = op.normalize_sim_scores(q, t, b, s["genome_of"])
    o = np.lexsort((dst, src))
    src, dst, w = src[o], dst[o], w[o].astype(np.float32)
    y = op.map_labels(src, dst, s["group_of"]).astype(np.float32)
    N = s["num_genes"]
    order = np.argsort(s["group_of"], kind="stable")
    groups = np.split(order, np.cumsum(np.bincount(s["group_of"]))[:-1])
    ref, cb = osg.sub_graphs(src, dst, w, y, N, n, groups, gff_is_subset=True)
    arena = run_extract(src, dst, w, y, N, n, groups, True, True)
    got = arena_graphs(arena, True)
    assert len(got) == len(ref) and len(ref) > 0
    for a, r in zip(got, ref):
        for k in ("order", "sim_ei", "w", "y", "nb_ei"):
            assert np.array_equal(a[k], r[k]), k
    assert arena.class_balance == pytest.approx(cb, rel=1e-12)
    # the reference's --cpus worker chunks: mean of the per-chunk ratios
    for chunks in (2, 3):
        _, cbc = osg.sub_graphs(src, dst, w, y, N, n, groups, gff_is_subset=True, chunks=chunks)
        got_c = run_extract(src, dst, w, y, N, n, groups, True, True, chunks=chunks).class_balance
        assert got_c == pytest.approx(cbc, rel=1e-12)


def test_graph_list_and_loader_over_the_arena():
    """GraphList is the list view the reference's API exposes; DeviceLoader collates straight from the arena and
    yields what the host collate of the same graphs yields."""
    from pangnn_b200.data import DeviceLoader, collate
    from pangnn_b200.simulate import simulate_hits
    from pangnn_b200.subgraphs import GraphList
    s = simulate_hits(200, 3, 0.5, 10, 3, seed=1)
    q, t, b = op.dedupe_last(s["q"].astype(np.int64), s["t"].astype(np.int64), s["bits"])
    q, t, b = op.remove_trivial_cases(q, t, b, s["genome_of"])
    src, dst, w = op.normalize_sim_scores(q, t, b, s["genome_of"])
    o = np.lexsort((dst, src))
    src, dst, w = src[o], dst[o], w[o].astype(np.float32)
    y = op.map_labels(src, dst, s["group_of"]).astype(np.float32)
    order = np.argsort(s["group_of"], kind="stable")
    groups = np.split(order, np.cumsum(np.bincount(s["group_of"]))[:-1])
    arena = run_extract(src, dst, w, y, s["num_genes"], 2, groups, True, True)
    gl = GraphList(arena, np.arange(arena.num_graphs))
    sub = gl[10:50]
    assert len(sub) == 40 and len(gl[[3, 1, 2]]) == 3
    assert torch.equal(sub[0].node_id, gl[10].node_id)
    loader = DeviceLoader(sub, batch_size=16, shuffle=True, device=DEV, seed=4)
    import random
    perm = list(range(40))
    random.Random(4).shuffle(perm)
    batches = list(loader)
    assert len(batches) == 3 and batches[-1].num_graphs == 8
    for bi, b in enumerate(batches):
        ref = collate([sub[i] for i in perm[bi * 16:(bi + 1) * 16]])
        for k in ref.keys():
            va, vb = getattr(ref, k), getattr(b, k)
            assert (torch.equal(va.to(DEV), vb) if torch.is_tensor(va) else va == vb), k
