"""Oracle restatement of panGNN's candidate normalisation and graph assembly (numpy, fp64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference works on a dict-of-dicts keyed
by gene-id strings; the restatement works on the equivalent *hit table*: three parallel arrays
``q`` (query node id), ``t`` (target node id), ``bits`` (score), unique in ``(q, t)``, plus
``genome_of[node]`` (the reference derives the genome from the id prefix before ``'_'``,
``src/preprocessing.py:375,463``; node ids are genome-major positions, ``src/dataset.py:72,101``)
and ``group_of[node]`` (RIBAP / ortholog group row, -1 = none; ``src/preprocessing.py:159-193``,
``src/simulate.py:148-152``).

Edge ORDER in the reference is CPython ``set`` iteration order (``src/helper.py:420-433``), i.e.
arbitrary; every function here returns the canonical ``(src, dst)``-sorted form
and the tests compare the reference's output after the same canonicalisation.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------
# hit-table helpers
# ----------------------------------------------------------------------------------------------
def dict_to_hits(sim_score_dict, gene_pos):
    """dict[q][t] -> (q, t, bits) arrays, dict iteration order; targets unknown to ``gene_pos``
    are dropped (``src/preprocessing.py:110``); queries are always known in the reference."""
    q, t, b = [], [], []
    for qs, cands in sim_score_dict.items():
        if qs not in gene_pos:
            continue
        for ts, s in cands.items():
            if ts in gene_pos:
                q.append(gene_pos[qs]); t.append(gene_pos[ts]); b.append(float(s))
    return (np.asarray(q, dtype=np.int64), np.asarray(t, dtype=np.int64),
            np.asarray(b, dtype=np.float64))


def dedupe_last(q, t, bits):
    """``dict(zip(targets, bits))`` keeps the LAST row of a repeated (query, target) pair
    (``src/preprocessing.py:413-416``)."""
    n = q.size
    order = np.lexsort((np.arange(n), t, q))
    qs, ts = q[order], t[order]
    last = np.ones(n, dtype=bool)
    if n > 1:
        last[:-1] = (qs[1:] != qs[:-1]) | (ts[1:] != ts[:-1])
    keep = order[last]
    return q[keep], t[keep], bits[keep]


def center_scores(bits):
    """``bits - min(bits) + 1`` over the filtered frame, ``src/preprocessing.py:403-405``."""
    return bits - bits.min() + 1 if bits.size else bits


def _segment_ids(q, g):
    """Sort by (query, target genome); returns order, segment start flags, per-element segment id."""
    order = np.lexsort((g, q))
    qs, gs = q[order], g[order]
    head = np.ones(q.size, dtype=bool)
    if q.size > 1:
        head[1:] = (qs[1:] != qs[:-1]) | (gs[1:] != gs[:-1])
    seg = np.cumsum(head) - 1
    return order, head, seg


# ----------------------------------------------------------------------------------------------
# a1  remove_trivial_cases  (src/preprocessing.py:370-385)
# ----------------------------------------------------------------------------------------------
def remove_trivial_cases(q, t, bits, genome_of):
    """Drop candidates whose genome occurs exactly once among a query's candidates.  The self hit
    COUNTS towards the genome's multiplicity (it is still in the dict at this point)."""
    if q.size == 0:
        return q, t, bits
    g = genome_of[t]
    order, head, seg = _segment_ids(q, g)
    counts = np.bincount(seg)
    keep_sorted = counts[seg] > 1
    keep = np.zeros(q.size, dtype=bool)
    keep[order] = keep_sorted
    return q[keep], t[keep], bits[keep]


# ----------------------------------------------------------------------------------------------
# a3  normalize_sim_scores + softmax_with_temperature  (src/preprocessing.py:430-443,454-548)
# ----------------------------------------------------------------------------------------------
def normalize_sim_scores(q, t, bits, genome_of, temp=0.8, epsilon=1e-8, pseudo_count=1.0):
    """Per (query, target genome) segment, excluding the self hit:
    ``p = softmax(x / temp)`` (logsumexp, fp64) if the segment has >= 2 members else ``p = 1``;
    ``w = -10 log10(clip(1 - p, eps, 1 - eps)) + pseudo_count`` (NaN p -> ``-10 log10(1-eps)``).
    Returns ``(src, dst, w)`` (fp64) sorted by ``(src, dst)``."""
    notself = q != t                                          # src/preprocessing.py:474
    q, t, bits = q[notself], t[notself], bits[notself]
    if q.size == 0:
        return q, t, bits
    g = genome_of[t]
    order, head, seg = _segment_ids(q, g)
    x = bits[order] / temp
    nseg = int(seg[-1]) + 1
    counts = np.bincount(seg, minlength=nseg)
    # logsumexp per segment, scipy style: subtract the max, sum exps, log, add back.
    mx = np.full(nseg, -np.inf)
    np.maximum.at(mx, seg, x)
    ssum = np.bincount(seg, weights=np.exp(x - mx[seg]), minlength=nseg)
    with np.errstate(divide="ignore", invalid="ignore"):
        lse = np.log(ssum) + mx
        p = np.exp(x - lse[seg])
    p = np.where(counts[seg] > 1, p, 1.0)                     # src/preprocessing.py:491
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(np.isnan(p), -10 * np.log10(1 - epsilon),
                     -10 * np.log10(np.clip(1 - p, epsilon, 1 - epsilon)))   # :492
    w = w + pseudo_count                                      # :494
    src, dst = q[order], t[order]
    canon = np.lexsort((dst, src))
    return src[canon], dst[canon], w[canon]


# ----------------------------------------------------------------------------------------------
# a4-a7  edge index / weights / labels  (src/preprocessing.py:73-118,264-325,122-156)
# ----------------------------------------------------------------------------------------------
def canonical_order(src, dst):
    return np.lexsort((dst, src))


def map_labels(src, dst, group_of):
    """y = 1 iff dst is listed among src's co-orthologs or vice versa: same RIBAP row, both in a
    row, and not the same gene (``src/preprocessing.py:146-149,181``)."""
    gs, gd = group_of[src], group_of[dst]
    return ((gs >= 0) & (gs == gd) & (src != dst)).astype(np.float32)


def class_balance(y):
    """``(y == 0).sum() / y.sum()``, ``src/dataset.py:346``."""
    pos = float(y.sum())
    return float((y == 0).sum()) / pos


# ----------------------------------------------------------------------------------------------
# a8  neighbour band of the whole graph  (src/dataset.py:351-366)
# ----------------------------------------------------------------------------------------------
def neighbour_band(num_genes, n):
    """Edges i -> j for j in [i-n, i+n] ∩ [0, N), self loop included, in the reference's loop
    order (i ascending, j ascending); crosses genome seams."""
    i = np.repeat(np.arange(num_genes, dtype=np.int64), 2 * n + 1)
    j = i + np.tile(np.arange(-n, n + 1, dtype=np.int64), num_genes)
    ok = (j >= 0) & (j < num_genes)
    return np.stack((i[ok], j[ok]))


# ----------------------------------------------------------------------------------------------
# a11 union assembly of the whole graph  (src/dataset.py:373-381): [sim ; nb], weights [w ; 1...]
# ----------------------------------------------------------------------------------------------
def union_whole_graph(sim_edge_index, sim_w, nb_edge_index):
    ei = np.concatenate((sim_edge_index, nb_edge_index), axis=1)
    w = np.concatenate((sim_w.astype(np.float32),
                        np.ones(nb_edge_index.shape[1], dtype=np.float32)))
    return ei, w


# ----------------------------------------------------------------------------------------------
# a12 PyG Batch collation  (call sites pangnn.py:121,152-153)
# ----------------------------------------------------------------------------------------------
def collate(graphs):
    """dict-of-arrays graphs -> one dict.  Keys containing 'index' are concatenated on the last
    axis and offset by the cumulative node count; everything else is concatenated on axis 0."""
    out, off, batch, ptr = {}, 0, [], [0]
    keys = [k for k in graphs[0] if graphs[0][k] is not None]
    parts = {k: [] for k in keys}
    for gi, g in enumerate(graphs):
        n = g["x"].shape[0]
        for k in keys:
            parts[k].append(g[k] + off if "index" in k else g[k])
        batch.append(np.full(n, gi, dtype=np.int64))
        off += n
        ptr.append(off)
    for k in keys:
        out[k] = np.concatenate(parts[k], axis=-1 if "index" in k else 0)
    out["batch"] = np.concatenate(batch)
    out["ptr"] = np.asarray(ptr, dtype=np.int64)
    return out


# ----------------------------------------------------------------------------------------------
# f1  max-candidate baseline  (src/helper.py:437-485) — "next" row, kept for the parity tests
# ----------------------------------------------------------------------------------------------
def baseline_labels(src, dst, score, genome_of):
    """1 iff no candidate of the same (query, target-genome) segment has a strictly larger score.
    ``src``/``dst``/``score`` must hold EVERY entry of the dict the reference scans (for the raw
    baseline that includes self hits, ``src/helper.py:470-475``)."""
    g = genome_of[dst]
    order, head, seg = _segment_ids(src, g)
    mx = np.full(int(seg[-1]) + 1 if seg.size else 0, -np.inf)
    np.maximum.at(mx, seg, score[order])
    out = np.zeros(src.size, dtype=np.int64)
    out[order] = (score[order] >= mx[seg]).astype(np.int64)
    return out
