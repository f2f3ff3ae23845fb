"""Synthetic pan-genome generator (host, vectorised numpy): the ``--simulate_dataset n G f frags shuf``
input source of the reference, ``src/simulate.py:83-230``, re-expressed as a HIT TABLE in node ids.

Distribution-faithful, not stream-faithful: the reference draws from unseeded ``random`` / numpy
Mersenne-Twister streams through Python dict insertion, which cannot be replayed; what is kept is
  * positives: every pair of the G genes at one position, both directions, score
    ``int(Gamma(k = mu^2 / 1e4, theta = 1e4 / mu))``, mu = 500            (``:11-17,156-168``)
  * negatives: for every source gene of genomes 0..G-2, ``k ~ clip(NegBin(0.2, 0.2/(m+0.2)), 1, n)``
    distinct uniform positions of the NEXT genome, mu = 200, both directions (``:131-132,170-190``);
    a negative that lands on the ortholog position overwrites the positive score (dict semantics)
  * ``m = floor((E_pos/f - E_pos) / (n G))``                                (``:120-129``)
  * synteny shuffle: ``shuf`` of the ``floor(n/frags)``-sized blocks of every genome permuted among
    themselves                                                              (``:202-230``)
The reference's own Python generator cannot get past ~6e4 genes.

Every random stream is keyed by ``(seed, genome pair)`` / ``(seed, genome)``, so any rank of a
genome-partitioned run can generate exactly the slab it owns (``genomes=(lo, hi)``) and two ranks
agree on the hits that cross their seam.  ``adjacent_only`` skips the positives between
non-adjacent genomes: with the default trivial-case filter those (query, genome) segments hold the
ortholog alone and are dropped anyway, so the filtered graph is identical while the
generated table grows with G instead of G^2.
"""
import math

import numpy as np


def _gamma_scores(rng, mean, dispersion, size):
    return np.floor(rng.gamma(mean * mean / dispersion, dispersion / mean, size=size))


def negatives_mean(n, G, frac_pos):
    e_pos = (G * (G - 1)) // 2 * n
    return (math.floor(e_pos / frac_pos) - e_pos) // (n * G)


def _pair_hits(n, g1, g2, seed, pos_mean, dispersion):
    """Positives between genomes g1 < g2 (one direction; the caller mirrors)."""
    rng = np.random.default_rng([seed, 1, g1, g2])
    p = np.arange(n, dtype=np.int64)
    return g1 * n + p, g2 * n + p, _gamma_scores(rng, pos_mean, dispersion, n)


def _negatives(n, g, m, seed, neg_mean, dispersion):
    """Negatives from the sources of genome g to distinct positions of genome g + 1."""
    rng = np.random.default_rng([seed, 2, g])
    k = rng.negative_binomial(0.2, 0.2 / (m + 0.2), size=n) if m > 0 else np.zeros(n, dtype=np.int64)
    k = np.clip(k, 1, n)
    owner = np.repeat(np.arange(n, dtype=np.int64), k)
    pos = rng.integers(0, n, size=owner.size)
    for _ in range(64):                                   # redraw within-source duplicates
        order = np.lexsort((pos, owner))
        so, sp = owner[order], pos[order]
        dup = np.zeros(owner.size, dtype=bool)
        dup[order[1:]] = (so[1:] == so[:-1]) & (sp[1:] == sp[:-1])
        nd = int(dup.sum())
        if nd == 0:
            break
        pos[dup] = rng.integers(0, n, size=nd)
    return g * n + owner, (g + 1) * n + pos, _gamma_scores(rng, neg_mean, dispersion, owner.size)


def synteny_permutation(n, g, num_fragments, num_frags_to_shuffle, seed):
    """new position (within the genome) of every old position of genome g."""
    new_of_old = np.arange(n, dtype=np.int64)
    frag = int(math.floor(n / num_fragments)) if num_fragments else n
    shuf = int(num_frags_to_shuffle)
    if shuf > 1 and frag > 0:
        rng = np.random.default_rng([seed, 3, g])
        nfrag = (n + frag - 1) // frag
        sel = rng.choice(nfrag, size=min(shuf, nfrag), replace=False)
        perm = rng.permutation(sel)
        blocks = [np.arange(i * frag, min((i + 1) * frag, n)) for i in range(nfrag)]
        new_blocks = list(blocks)
        for dst_i, src_i in zip(sel, perm):
            new_blocks[dst_i] = blocks[src_i]
        order_g = np.concatenate(new_blocks)               # old position at each new position
        new_of_old[order_g] = np.arange(n)
    return new_of_old


def simulate_hits(n, G, frac_pos, num_fragments=10, num_frags_to_shuffle=3, score_means=(200, 500),
                  dispersion=1e4, seed=0, genomes=None, adjacent_only=False):
    """-> dict(q, t, bits, genome_of, group_of, num_genes, ...).  ``q``/``t`` are int32 GLOBAL node ids
    in the genome-major order AFTER the synteny shuffle; rows are ordered so that "last row wins"
    equals the reference's dict-overwrite order.  With ``genomes=(lo, hi)`` only hits whose QUERY
    lies in genomes [lo, hi) are emitted (their candidate sets are complete)."""
    n, G = int(n), int(G)
    neg_mean, pos_mean = score_means
    m = negatives_mean(n, G, frac_pos)
    N = n * G
    lo, hi = (0, G) if genomes is None else (max(int(genomes[0]), 0), min(int(genomes[1]), G))
    qs, ts, bs = [], [], []
    # ---- positives (both directions); a query of genome g meets every other genome (or g +- 1)
    for g1 in range(G):
        for g2 in range(g1 + 1, G):
            if adjacent_only and g2 != g1 + 1:
                continue
            if not ((lo <= g1 < hi) or (lo <= g2 < hi)):
                continue
            a, b, s = _pair_hits(n, g1, g2, seed, pos_mean, dispersion)
            if lo <= g1 < hi:
                qs.append(a); ts.append(b); bs.append(s)
            if lo <= g2 < hi:
                qs.append(b); ts.append(a); bs.append(s)
    # ---- negatives g -> g+1 (both directions), after the positives so that they win collisions
    for g in range(G - 1):
        if not ((lo <= g < hi) or (lo <= g + 1 < hi)):
            continue
        a, b, s = _negatives(n, g, m, seed, neg_mean, dispersion)
        if lo <= g < hi:
            qs.append(a); ts.append(b); bs.append(s)
        if lo <= g + 1 < hi:
            qs.append(b); ts.append(a); bs.append(s)
    q = np.concatenate(qs) if qs else np.zeros(0, np.int64)
    t = np.concatenate(ts) if ts else np.zeros(0, np.int64)
    bits = np.concatenate(bs) if bs else np.zeros(0, np.float64)
    # ---- synteny shuffle -> new node id of every old node; ortholog group = position before shuffle
    new_of_old = np.empty(N, dtype=np.int64)
    for g in range(G):
        new_of_old[g * n:(g + 1) * n] = g * n + synteny_permutation(n, g, num_fragments, num_frags_to_shuffle, seed)
    group_of = np.empty(N, dtype=np.int32)
    group_of[new_of_old] = np.tile(np.arange(n, dtype=np.int32), G)
    return dict(q=new_of_old[q].astype(np.int32), t=new_of_old[t].astype(np.int32),
                bits=bits.astype(np.float64), genome_of=np.repeat(np.arange(G, dtype=np.int32), n),
                group_of=group_of, num_genes=N, neg_mean_per_gene=m, genes_per_genome=n, num_genomes=G)


def simulate_hits_device(n, G, frac_pos, num_fragments=10, num_frags_to_shuffle=3, score_means=(200, 500),
                         dispersion=1e4, seed=0, genomes=None, device="cuda", group_of=None):
    """Device form of ``simulate_hits(..., adjacent_only=True)`` (``pangnn_simulate_edges``: Philox streams keyed by
    (seed, genome, gene, draw); same distributions, different stream).  ``q, t, bits`` are device tensors; the
    small per-gene maps stay numpy.  ``genomes=(lo, hi)``: only hits whose QUERY lies in genomes [lo, hi)."""
    import torch
    from . import _abi, ops
    lib = _abi.load()
    n, G = int(n), int(G)
    neg_mean, pos_mean = score_means
    m = negatives_mean(n, G, frac_pos)
    N = n * G
    lo, hi = (0, G) if genomes is None else (max(int(genomes[0]), 0), min(int(genomes[1]), G))
    dev = torch.device(device)
    g_first, g_end = max(lo - 1, 0), min(hi, G - 1)                 # source genomes g: pair (g, g + 1)
    ng = max(g_end - g_first, 0)
    # synteny permutation of the genomes touched (host: O(n) per genome), as new GLOBAL ids
    perm_lo, perm_hi = g_first, min(g_end + 1, G)
    new_of_old = np.concatenate([g * n + synteny_permutation(n, g, num_fragments, num_frags_to_shuffle, seed)
                                 for g in range(perm_lo, perm_hi)] or [np.zeros(0, np.int64)]).astype(np.int32)
    nod = torch.from_numpy(new_of_old).to(dev)
    k = torch.empty(ng * n, dtype=torch.int32, device=dev)
    st = ops._stream()
    _abi.check(lib.pangnn_simulate_neg_counts(int(seed), n, int(m), g_first, ng, ops._p(k), st), "simulate_neg_counts")
    gsrc = g_first + torch.arange(ng * n, device=dev) // n
    fwd, rev = (gsrc >= lo) & (gsrc < hi), (gsrc + 1 >= lo) & (gsrc + 1 < hi)
    kf, kr = k.long() * fwd, k.long() * rev
    fwd_off, rev_off = torch.cumsum(kf, 0) - kf, torch.cumsum(kr, 0) - kr
    pair = torch.arange(g_first, g_end)
    n_pos = int((((pair >= lo) & (pair < hi)).long() + ((pair + 1 >= lo) & (pair + 1 < hi)).long()).sum()) * n
    n_fwd, n_rev = int(kf.sum().item()), int(kr.sum().item())
    rows = n_pos + n_fwd + n_rev
    q = torch.empty(rows, dtype=torch.int32, device=dev)
    t = torch.empty(rows, dtype=torch.int32, device=dev)
    bits = torch.empty(rows, dtype=torch.float64, device=dev)
    _abi.check(lib.pangnn_simulate_edges(int(seed), n, G, lo, hi, float(neg_mean), float(pos_mean), float(dispersion),
                                         g_first, ng, ops._p(k), ops._p(fwd_off), ops._p(rev_off), n_pos, n_fwd,
                                         ops._p(nod), perm_lo, ops._p(q), ops._p(t), ops._p(bits), st), "simulate_edges")
    ops.LAUNCHES["count"] += 3
    # group = position before the shuffle (all genomes: the labels of halo targets are needed too); a caller that
    # generates slab by slab passes the map of its first call back in
    if group_of is None:
        group_of = np.empty(N, dtype=np.int32)
        for g in range(G):
            group_of[g * n + synteny_permutation(n, g, num_fragments, num_frags_to_shuffle, seed)] = np.arange(n, dtype=np.int32)
    return dict(q=q, t=t, bits=bits, genome_of=np.repeat(np.arange(G, dtype=np.int32), n), group_of=group_of,
                num_genes=N, neg_mean_per_gene=m, genes_per_genome=n, num_genomes=G, neg_counts=k,
                rows=dict(pos=n_pos, neg_fwd=n_fwd, neg_rev=n_rev))
