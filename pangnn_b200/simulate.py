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
"""
import math

import numpy as np


def _gamma_scores(rng, mean, dispersion, size):
    return np.floor(rng.gamma(mean * mean / dispersion, dispersion / mean, size=size))


def negatives_mean(n, G, frac_pos):
    e_pos = (G * (G - 1)) // 2 * n
    return (math.floor(e_pos / frac_pos) - e_pos) // (n * G)


def simulate_hits(n, G, frac_pos, num_fragments=10, num_frags_to_shuffle=3, score_means=(200, 500),
                  dispersion=1e4, seed=0):
    """-> dict(q, t, bits, genome_of, group_of, num_genes).  ``q``/``t`` are int32 node ids in the
    genome-major order AFTER the synteny shuffle; rows are ordered so that "last row wins" equals
    the reference's dict-overwrite order."""
    rng = np.random.default_rng(seed)
    n, G = int(n), int(G)
    neg_mean, pos_mean = score_means
    m = negatives_mean(n, G, frac_pos)
    N = n * G
    # ---- positives: all ordered genome pairs at every position
    g1, g2 = np.triu_indices(G, k=1)
    p = np.arange(n, dtype=np.int64)
    a = (g1[None, :] * n + p[:, None]).ravel()
    b = (g2[None, :] * n + p[:, None]).ravel()
    s = _gamma_scores(rng, pos_mean, dispersion, a.size)
    pq, pt, pb = np.concatenate((a, b)), np.concatenate((b, a)), np.concatenate((s, s))
    # ---- negatives: source (g, p), g < G-1 -> k distinct positions of genome g+1
    k_all = rng.negative_binomial(0.2, 0.2 / (m + 0.2), size=N) if m > 0 else np.zeros(N, dtype=np.int64)
    k_all = np.clip(k_all, 1, n)
    src_g = np.repeat(np.arange(G - 1, dtype=np.int64), n)
    src_p = np.tile(p, G - 1)
    k = k_all[: src_g.size]
    owner = np.repeat(np.arange(src_g.size, dtype=np.int64), k)
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
    ns = src_g[owner] * n + src_p[owner]
    nt = (src_g[owner] + 1) * n + pos
    nb = _gamma_scores(rng, neg_mean, dispersion, ns.size)
    q = np.concatenate((pq, ns, nt))
    t = np.concatenate((pt, nt, ns))
    bits = np.concatenate((pb, nb, nb))
    group_old = np.tile(p, G)                              # ortholog group = position before shuffle
    # ---- synteny shuffle -> new node id of every old node
    new_of_old = np.arange(N, dtype=np.int64)
    frag = int(math.floor(n / num_fragments)) if num_fragments else n
    shuf = int(num_frags_to_shuffle)
    if shuf > 1 and frag > 0:
        nfrag = (n + frag - 1) // frag
        for g in range(G):
            sel = rng.choice(nfrag, size=min(shuf, nfrag), replace=False)
            perm = rng.permutation(sel)
            blocks = [np.arange(i * frag, min((i + 1) * frag, n)) for i in range(nfrag)]
            new_blocks = list(blocks)
            for dst_i, src_i in zip(sel, perm):
                new_blocks[dst_i] = blocks[src_i]
            order_g = np.concatenate(new_blocks)           # old position at each new position
            new_of_old[g * n + order_g] = g * n + np.arange(n)
    group_of = np.empty(N, dtype=np.int32)
    group_of[new_of_old] = group_old
    return dict(q=new_of_old[q].astype(np.int32), t=new_of_old[t].astype(np.int32),
                bits=bits.astype(np.float64), genome_of=np.repeat(np.arange(G, dtype=np.int32), n),
                group_of=group_of, num_genes=N, neg_mean_per_gene=m)
