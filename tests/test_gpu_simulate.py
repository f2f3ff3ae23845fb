"""Device generator of ``--simulate_dataset`` hit tables (``pangnn_simulate_edges``, csrc/simulate.cu) against the
distributions of the reference generator (``src/simulate.py:120-190``; SURVEY §8d: E_pos, E_neg, NegBin mean / var,
score mean / var, label fraction) and against the host generator of this package through the whole preprocessing."""
import numpy as np
import pytest
import torch
from scipy import stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(n, G, f, seed=0, genomes=None):
    from pangnn_b200.simulate import simulate_hits_device
    return simulate_hits_device(n, G, f, 10, 3, seed=seed, genomes=genomes, device=DEV)


def test_row_counts_symmetry_and_distinct_negatives():
    n, G, f = 3000, 4, 0.3
    s = _gen(n, G, f)
    q, t, b = (s[k].cpu().numpy() for k in ("q", "t", "bits"))
    k = s["neg_counts"].cpu().numpy().astype(np.int64)
    r = s["rows"]
    assert r["pos"] == 2 * n * (G - 1)                                   # both directions of every adjacent pair
    assert r["neg_fwd"] == r["neg_rev"] == int(k.sum()) and q.size == r["pos"] + 2 * int(k.sum())
    assert k.min() >= 1 and k.max() <= n
    # every (q, t, score) row has its mirror (t, q, score)  (src/simulate.py:163-164,184-185)
    fwd = np.stack((q, t, b.astype(np.int64)))
    rev = np.stack((t, q, b.astype(np.int64)))
    assert np.array_equal(fwd[:, np.lexsort(fwd[::-1])], rev[:, np.lexsort(rev[::-1])])
    # negatives of one source are DISTINCT positions of the next genome (random.sample, :174)
    nq, nt = q[r["pos"]:r["pos"] + r["neg_fwd"]], t[r["pos"]:r["pos"] + r["neg_fwd"]]
    assert np.unique(nq.astype(np.int64) * (n * G) + nt).size == nq.size
    assert np.array_equal(nt // n, nq // n + 1)
    assert np.array_equal(np.bincount(nq, minlength=n * G)[: n * (G - 1)].sum(), k.sum())
    # scores are whole numbers (int(...) of the gamma draw, :16)
    assert np.array_equal(b, np.floor(b)) and b.min() >= 0


def test_negative_counts_follow_the_clipped_negative_binomial():
    from pangnn_b200.simulate import negatives_mean
    n, G, f = 200_000, 2, 0.02
    m = negatives_mean(n, G, f)
    assert m >= 20
    k = _gen(n, G, f)["neg_counts"].cpu().numpy().astype(np.int64)
    p = 0.2 / (m + 0.2)
    ref = np.clip(np.random.default_rng(0).negative_binomial(0.2, p, size=4_000_000), 1, n)
    assert abs(k.mean() - ref.mean()) < 5 * ref.std() / np.sqrt(k.size)
    assert abs(k.var() - ref.var()) < 0.08 * ref.var()                   # heavy tail: var of var is large
    # the bulk of the distribution, bin by bin, against the exact pmf
    edges = np.array([1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 10**9])
    cdf = stats.nbinom.cdf(edges - 1, 0.2, p)
    probs = np.diff(np.concatenate(([0.0], cdf[1:])))                    # P(k <= 1) goes to the first bin (clip at 1)
    obs = np.histogram(k, bins=edges)[0]
    chi2 = ((obs - probs * k.size) ** 2 / (probs * k.size)).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-6, df=len(probs) - 1)


def test_scores_follow_the_floored_gamma_distributions():
    n, G, f = 60_000, 3, 0.3
    s = _gen(n, G, f)
    b, r = s["bits"].cpu().numpy(), s["rows"]
    # positives: rows are pair-major, the n forward rows of a pair followed by its n reverse rows with the SAME scores
    # (src/simulate.py:163-164) -> independent draws = the forward blocks only
    fwd_pos = np.concatenate([b[2 * n * g: 2 * n * g + n] for g in range(G - 1)])
    for vals, mu in ((fwd_pos[: 100_000], 500.0), (b[r["pos"]:r["pos"] + r["neg_fwd"]][: 100_000], 200.0)):
        shape, scale = mu * mu / 1e4, 1e4 / mu
        # floor(X): compare with the gamma cdf at the bin edges (KS on the discretised variable)
        d = np.abs(np.searchsorted(np.sort(vals), np.arange(0, 1500), side="left") / vals.size
                   - stats.gamma.cdf(np.arange(0, 1500), shape, scale=scale)).max()
        assert d < 1.95 / np.sqrt(vals.size)                             # KS critical value at alpha ~ 0.001
        assert abs(vals.mean() - (mu - 0.5)) < 5 * 100.0 / np.sqrt(vals.size)


def test_negative_positions_are_uniform():
    n, G, f = 5000, 2, 0.05
    s = _gen(n, G, f)
    r = s["rows"]
    pos = s["t"].cpu().numpy()[r["pos"]:r["pos"] + r["neg_fwd"]] - n
    # the synteny shuffle permutes positions; uniformity is invariant under a permutation
    obs = np.bincount(pos, minlength=n)
    chi2 = ((obs - pos.size / n) ** 2 / (pos.size / n)).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-6, df=n - 1)


def test_slabs_agree_with_the_whole_table():
    """A rank generates the rows whose QUERY lies in its genomes; they are exactly those rows of the full table."""
    n, G, f = 500, 6, 0.2
    full = _gen(n, G, f, seed=5)
    fq, ft, fb = (full[k].cpu().numpy() for k in ("q", "t", "bits"))
    for lo, hi in ((0, 2), (1, 4), (3, 6), (-1, 3), (4, 7), (2, 3)):
        part = _gen(n, G, f, seed=5, genomes=(lo, hi))
        m = (fq // n >= lo) & (fq // n < hi)
        ref = np.stack((fq[m], ft[m], fb[m].astype(np.int64)))
        got = np.stack([part[k].cpu().numpy().astype(np.int64) for k in ("q", "t", "bits")])
        assert got.shape == ref.shape
        assert np.array_equal(ref[:, np.lexsort(ref[::-1])], got[:, np.lexsort(got[::-1])])
        assert np.array_equal(part["group_of"], full["group_of"])


def test_graph_statistics_match_the_host_generator():
    """Through dedupe + trivial-case filter + normalisation + labels: edge count, positive fraction and the weight
    distribution of the device-generated graph agree with the host generator's (different streams, same law)."""
    from pangnn_b200 import preprocessing as pp
    from pangnn_b200.simulate import simulate_hits
    n, G, f = 20_000, 4, 0.5
    outs = []
    for gen in (lambda: simulate_hits(n, G, f, 10, 3, seed=1, adjacent_only=True), lambda: _gen(n, G, f, seed=1)):
        s = gen()
        src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=n * G,
                                                 device=DEV)
        outs.append((src.numel(), float(y.mean()), float((w == 81.0).float().mean()), float((w < 1.5).float().mean())))
    (e0, p0, s0, l0), (e1, p1, s1, l1) = outs
    assert abs(e1 - e0) < 0.01 * e0
    assert abs(p1 - p0) < 0.01 and abs(s1 - s0) < 0.01 and abs(l1 - l0) < 0.01
