"""TEST INFRASTRUCTURE — CPU restatement of the reference's training-mode sub-graph extraction
(SURVEY.md §8 a9, a10 and the sub-graph half of a11).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this; the product path is pangnn_b200.subgraphs (device).

Pinned against the reference-run golden vectors tests/golden/{c1_sub,sim5_sub}.npz (node sets, sim edges with
weights and labels, neighbour edges, all in global ids; tests/test_oracle_subgraphs.py).

One sub-graph per ortholog group (``src/dataset.py:222-322``):
  * ``get_connected_nodes`` (``src/helper.py:327-362``): the group plus its ``n``-hop closure over the OUT-edges
    of the normalised similarity dict;
  * ``get_neighbour_graph`` (``src/helper.py:366-417``): for every sub-graph gene, the genes within +-n positions
    in the global order (self excluded); genes not yet in the sub-graph get new local ids; both directions are
    appended and duplicates removed (``src/helper.py:420-433``);
  * ``build_edge_index`` (``src/preprocessing.py:73-118``): similarity edges whose BOTH endpoints are sub-graph
    genes (window-added genes included);
  * union assembly ``[nb ; sim]`` with weights ``[1... ; w]`` (``src/dataset.py:287-303``).
The reference's local numbering is CPython set order; the canonical order used here and by the
device path: connected genes ascending, then window-added genes in first-encounter order; edges by (src, dst).
"""
import numpy as np


def sub_graph(group, rowptr, dst, w, y, num_genes, n):
    """-> dict(order [k] global ids in local order, sim_ei [2,e] local, w, y, nb_ei [2,m] local) or None when
    no gene of the sub-graph has an outgoing similarity edge."""
    group = np.asarray(group, dtype=np.int64)
    connected = set(group.tolist())
    frontier = list(connected)
    for _ in range(n):
        nxt = set()
        for g in frontier:
            nxt.update(dst[rowptr[g]:rowptr[g + 1]].tolist())
        nxt -= connected
        if not nxt:
            break
        connected |= nxt
        frontier = list(nxt)
    order = sorted(connected)
    local = {g: i for i, g in enumerate(order)}
    nb = set()
    for g in list(order):
        for j in range(g - n, g + n + 1):
            if j < 0 or j >= num_genes or j == g:
                continue
            if j not in local:
                local[j] = len(order)
                order.append(j)
            nb.add((local[g], local[j]))
            nb.add((local[j], local[g]))
    if not any(rowptr[g + 1] > rowptr[g] for g in order):
        return None
    sim = []
    for g in order:
        for e in range(rowptr[g], rowptr[g + 1]):
            t = int(dst[e])
            if t in local:
                sim.append((local[g], local[t], w[e], y[e]))
    sim.sort(key=lambda r: (r[0], r[1]))
    nb = sorted(nb)
    return dict(order=np.asarray(order, dtype=np.int64),
                sim_ei=np.asarray([[r[0] for r in sim], [r[1] for r in sim]], dtype=np.int64).reshape(2, -1),
                w=np.asarray([r[2] for r in sim], dtype=np.float32), y=np.asarray([r[3] for r in sim], dtype=np.float32),
                nb_ei=np.asarray([[a for a, _ in nb], [b for _, b in nb]], dtype=np.int64).reshape(2, -1))


def sub_graphs(src, dst, w, y, num_genes, n, groups, gff_is_subset=False, chunks=1):
    """All groups with more than one gene (``src/dataset.py:230``); (src, dst) sorted by (src, dst).  Returns the
    list of sub-graphs and the class balance: the mean over the ``chunks`` (= ``--cpus``) worker chunks
    ``groups[i::chunks]`` of each chunk's neg / pos (``src/dataset.py:128,141-142,319``)."""
    src = np.asarray(src, dtype=np.int64)
    rowptr = np.zeros(num_genes + 1, dtype=np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    groups = list(groups)
    out, pos, tot = [], [0.0] * chunks, [0.0] * chunks
    for gi, group in enumerate(groups):
        if len(group) <= 1:
            continue
        g = sub_graph(group, rowptr, np.asarray(dst, dtype=np.int64), w, y, num_genes, n)
        if g is None:
            continue
        if g["sim_ei"].shape[1] < len(group):
            if gff_is_subset:
                continue
            raise AssertionError("fewer similarity edges than genes in the origin family (src/dataset.py:279)")
        pos[gi % chunks] += float(g["y"].sum()); tot[gi % chunks] += g["y"].size
        out.append(g)
    live = range(min(chunks, len(groups)))
    ratios = [((tot[c] - pos[c]) / pos[c] if pos[c] else float("inf")) for c in live]
    return out, (sum(ratios) / len(ratios) if ratios else float("inf"))


def union_sub_graph(g):
    """``[nb ; sim]`` / ``[1... ; w]`` (``src/dataset.py:287-303``)."""
    return (np.concatenate((g["nb_ei"], g["sim_ei"]), axis=1),
            np.concatenate((np.ones(g["nb_ei"].shape[1], dtype=np.float32), g["w"])))
