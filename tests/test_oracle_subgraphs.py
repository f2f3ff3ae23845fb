"""The oracle's sub-graph extraction (a9, a10) against the reference-run golden vectors: for every ortholog
group the reference turned into a training sub-graph — node set, similarity edges with weights and labels and
neighbour edges, compared in GLOBAL ids (the reference's local numbering is CPython set order)."""
import numpy as np
import pytest

from oracle import preprocess as op
from oracle import subgraphs as osg


def golden_inputs(g):
    """Normalised hit table + labels of the golden case and ALL its ortholog groups (from ``group_of``; the
    golden's own seed lists are matched to sub-graphs by a subset test and can be off by one group)."""
    ei, w64 = g["norm/edge_index"], g["norm/w64"]
    y = op.map_labels(ei[0], ei[1], g["group_of"]).astype(np.float32)
    group_of = g["group_of"]
    groups = [np.flatnonzero(group_of == gid) for gid in np.unique(group_of[group_of >= 0])]
    return ei[0], ei[1], w64.astype(np.float32), y, int(g["num_genes"]), int(g["neighbours"]), groups


def canon(s, t, *cols):
    o = np.lexsort((t, s))
    return (s[o], t[o]) + tuple(c[o] for c in cols)


def _fingerprint(nodes, sim, nb):
    """Hashable canonical form of one sub-graph in global ids (weights as fp32 bit patterns)."""
    return (tuple(np.sort(nodes).tolist()),) + tuple(np.ascontiguousarray(a).tobytes() for a in sim + nb)


def golden_sub_graphs(g):
    """Multiset of the reference's sub-graphs (two groups can close over the same gene set)."""
    from collections import Counter
    out = Counter()
    for i in range(int(g["num_sub_graphs"])):
        a, b = g["sub/node_ptr"][i:i + 2]
        c, d = g["sub/sim_ptr"][i:i + 2]
        e, f = g["sub/nb_ptr"][i:i + 2]
        out[_fingerprint(g["sub/nodes"][a:b],
                         (g["sub/sim_src"][c:d], g["sub/sim_dst"][c:d], g["sub/sim_w"][c:d].astype(np.float32),
                          g["sub/sim_y"][c:d].astype(np.float32)),
                         (g["sub/nb_src"][e:f], g["sub/nb_dst"][e:f]))] += 1
    return out


def check_against_golden(g, subs):
    """``subs``: dicts with order / sim_ei / w / y / nb_ei (local ids) -> compared in global ids, as a multiset
    (the order of the sub-graphs follows the group list, which the golden file does not carry)."""
    from collections import Counter
    got = Counter()
    for sg in subs:
        order = np.asarray(sg["order"]).astype(np.int64)
        sim = canon(order[np.asarray(sg["sim_ei"][0])], order[np.asarray(sg["sim_ei"][1])],
                    np.asarray(sg["w"], dtype=np.float32), np.asarray(sg["y"], dtype=np.float32))
        nb = canon(order[np.asarray(sg["nb_ei"][0])], order[np.asarray(sg["nb_ei"][1])])
        got[_fingerprint(order, sim, nb)] += 1
    assert sum(got.values()) == int(g["num_sub_graphs"])
    assert got == golden_sub_graphs(g)


@pytest.mark.parametrize("case", ["c1_sub", "sim5_sub"])
def test_sub_graphs_match_reference(golden, case):
    g = golden(case)
    src, dst, w, y, N, n, groups = golden_inputs(g)
    subs, _ = osg.sub_graphs(src, dst, w, y, N, n, groups, gff_is_subset=True)
    check_against_golden(g, subs)


def test_local_numbering_is_canonical():
    # 6 genes in a row, one group {1, 4}, edges 1 -> 4, 4 -> 1, 4 -> 5; n = 1
    src, dst = np.array([1, 4, 4]), np.array([4, 1, 5])
    w, y = np.array([3.0, 2.0, 1.0], np.float32), np.array([1, 1, 0], np.float32)
    subs, cb = osg.sub_graphs(src, dst, w, y, 6, 1, [[1, 4]])
    (sg,) = subs
    # connected (1 hop): {1, 4, 5} ascending, then window genes in first-encounter order: 0, 2 (of 1), 3 (of 4)
    assert sg["order"].tolist() == [1, 4, 5, 0, 2, 3]
    assert sg["sim_ei"].tolist() == [[0, 1, 1], [1, 0, 2]] and sg["w"].tolist() == [3.0, 2.0, 1.0]
    # window pairs: 1-0, 1-2, 4-3, 4-5, 5-4 -> both directions, unique, sorted
    assert sg["nb_ei"].tolist() == [[0, 0, 1, 1, 2, 3, 4, 5], [3, 4, 2, 5, 1, 0, 0, 1]]
    assert cb == 0.5
    ei, uw = osg.union_sub_graph(sg)
    assert ei.shape == (2, 11) and uw.tolist() == [1.0] * 8 + [3.0, 2.0, 1.0]
