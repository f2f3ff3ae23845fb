"""GPU parity: device candidate normalisation (sort/dedupe, trivial filter, segmented softmax +
Q-score, edge/label emission) vs golden vectors minted from the reference's own preprocessing."""
import numpy as np
import pytest
import torch

from oracle import preprocess as op

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
WHOLE = ["dummy", "c1", "c2", "sim5", "sim5_trivial"]


def dev(a, dt):
    return torch.as_tensor(np.asarray(a), dtype=dt, device=DEV)


@pytest.mark.parametrize("case", WHOLE)
def test_normalize_matches_reference(golden, case):
    from pangnn_b200 import ops
    g = golden(case)
    N = int(g["num_genes"])
    rng = np.random.RandomState(0)
    shuffle = rng.permutation(g["raw/q"].size)             # the device path must not rely on input order
    q, t, b = g["raw/q"][shuffle], g["raw/t"][shuffle], g["raw/bits"][shuffle]
    qs, ts, bs = ops.hits_sort_unique(dev(q, torch.int32), dev(t, torch.int32), dev(b, torch.float64), N)
    assert qs.numel() == q.size
    # the stored raw table is what the reference normalised: already trivial-filtered unless the case
    # ran with --include_trivial, so on filtered tables the device filter must be idempotent
    include_trivial = "--include_trivial" in str(g["argv"])
    for drop in ((False,) if include_trivial else (True, False)):
        src, dst, w, y = ops.hits_normalize(qs, ts, bs, dev(g["genome_of"], torch.int32),
                                            dev(g["group_of"], torch.int32), temp=0.8,
                                            drop_trivial=drop)
        ei = np.stack((src.cpu().numpy(), dst.cpu().numpy())).astype(np.int64)
        assert np.array_equal(ei, g["graph/edge_index"])                          # bit-exact
        assert np.array_equal(y.cpu().numpy(), g["graph/y"])                      # bit-exact
        ref = g["graph/edge_attr"]
        got = w.cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=0)
        sat_hi, sat_lo = ref == np.float32(81.0), ref == np.float32(1.0)
        assert np.array_equal(got[sat_hi], ref[sat_hi]) and np.array_equal(got[sat_lo], ref[sat_lo])


@pytest.mark.parametrize("case", ["trivial_c1", "trivial_sim"])
def test_trivial_filter_and_self_hits(golden, case):
    from pangnn_b200 import ops
    g = golden(case)
    genome_of = g["genome_of"]
    N = genome_of.size
    q, t, b = g["unfiltered/q"], g["unfiltered/t"], g["unfiltered/bits"]
    # oracle end-to-end on the unfiltered table
    fq, ft, fb = op.remove_trivial_cases(q, t, b, genome_of)
    rs, rd, rw = op.normalize_sim_scores(fq, ft, fb, genome_of)
    qs, ts, bs = ops.hits_sort_unique(dev(q, torch.int32), dev(t, torch.int32), dev(b, torch.float64), N)
    src, dst, w, y = ops.hits_normalize(qs, ts, bs, dev(genome_of, torch.int32), None, drop_trivial=True)
    assert np.array_equal(src.cpu().numpy(), rs) and np.array_equal(dst.cpu().numpy(), rd)
    np.testing.assert_allclose(w.cpu().numpy(), rw.astype(np.float32), rtol=1e-5)
    assert float(y.abs().sum()) == 0.0
    # --include_trivial
    rs2, rd2, rw2 = op.normalize_sim_scores(q, t, b, genome_of)
    src2, dst2, w2, _ = ops.hits_normalize(qs, ts, bs, dev(genome_of, torch.int32), None, drop_trivial=False)
    assert np.array_equal(src2.cpu().numpy(), rs2) and np.array_equal(dst2.cpu().numpy(), rd2)
    np.testing.assert_allclose(w2.cpu().numpy(), rw2.astype(np.float32), rtol=1e-5)
    assert src2.numel() > src.numel()


def test_duplicates_keep_last_and_empty_input():
    from pangnn_b200 import ops
    q = np.array([3, 0, 0, 3, 0, 2]); t = np.array([1, 1, 2, 1, 1, 2]); b = np.array([5., 6., 7., 9., 8., 1.])
    qs, ts, bs = ops.hits_sort_unique(dev(q, torch.int32), dev(t, torch.int32), dev(b, torch.float64), 4)
    assert qs.tolist() == [0, 0, 2, 3] and ts.tolist() == [1, 2, 2, 1] and bs.tolist() == [8., 7., 1., 9.]
    e = torch.zeros(0, dtype=torch.int32, device=DEV)
    qs, ts, bs = ops.hits_sort_unique(e, e, torch.zeros(0, dtype=torch.float64, device=DEV), 4)
    assert qs.numel() == 0
    src, dst, w, y = ops.hits_normalize(qs, ts, bs, torch.zeros(4, dtype=torch.int32, device=DEV))
    assert src.numel() == 0


def test_large_segments_and_skew():
    """A NegBin-like tail: a few (query, genome) segments with thousands of members."""
    from pangnn_b200 import ops
    rng = np.random.RandomState(3)
    n_per, G = 5000, 3
    N = n_per * G
    genome_of = np.repeat(np.arange(G), n_per)
    q_l, t_l = [], []
    for qn, k in ((0, 4000), (1, 1), (2, 33), (n_per + 7, 2500), (n_per + 8, 32), (2 * n_per + 1, 31)):
        for gtarget in range(G):
            if gtarget == genome_of[qn]:
                continue
            q_l.append(np.full(k, qn)); t_l.append(gtarget * n_per + rng.choice(n_per, k, replace=False))
    q = np.concatenate(q_l); t = np.concatenate(t_l)
    b = np.floor(rng.gamma(4.0, 50.0, size=q.size))
    rs, rd, rw = op.normalize_sim_scores(q, t, b, genome_of)
    qs, ts, bs = ops.hits_sort_unique(dev(q, torch.int32), dev(t, torch.int32), dev(b, torch.float64), N)
    src, dst, w, _ = ops.hits_normalize(qs, ts, bs, dev(genome_of, torch.int32), None, drop_trivial=False)
    assert np.array_equal(src.cpu().numpy(), rs) and np.array_equal(dst.cpu().numpy(), rd)
    np.testing.assert_allclose(w.cpu().numpy(), rw.astype(np.float32), rtol=1e-5)


@pytest.mark.parametrize("case", WHOLE)
def test_baseline_labels_match_reference(golden, case):
    """Segmented arg-max baselines (src/helper.py:437-485) bit-exact against the goldens minted from the
    reference: on the normalised Q-scores, and on the raw bit scores (scanning the raw table incl. self hits)."""
    from pangnn_b200 import ops, preprocessing as pp
    g = golden(case)
    gen = dev(g["genome_of"], torch.int32)
    ei, w = g["graph/edge_index"], g["graph/edge_attr"]
    bl = pp.baseline_labels(dev(ei[0], torch.int32), dev(ei[1], torch.int32), dev(w, torch.float32), gen)
    assert np.array_equal(bl.cpu().numpy(), g["graph/base_labels"])
    N = int(g["num_genes"])
    qs, ts, bs = ops.hits_sort_unique(dev(g["raw/q"], torch.int32), dev(g["raw/t"], torch.int32),
                                      dev(g["raw/bits"], torch.float64), N)
    raw = pp.baseline_labels(qs, ts, bs, gen).cpu().numpy()
    key = {(int(a), int(c)): int(v) for a, c, v in zip(qs.cpu().numpy(), ts.cpu().numpy(), raw)}
    blr = np.asarray([key[(int(a), int(c))] for a, c in zip(ei[0], ei[1])])
    assert np.array_equal(blr, g["graph/base_labels_raw"])


def test_baseline_labels_long_segments_and_ties():
    """Segments longer than the thread-per-entry limit take the cooperative kernel; ties are all labelled 1."""
    from pangnn_b200 import ops
    rng = np.random.RandomState(3)
    n_genes, G = 500, 3
    genome_of = np.repeat(np.arange(G), n_genes).astype(np.int32)
    rows = []
    for q in (0, 7, 400):                                   # three queries with 1, 33 and 300 candidates per genome
        for gi, cnt in zip(range(G), (1, 33, 300)):
            ts = np.sort(rng.choice(n_genes, cnt, replace=False)) + gi * n_genes
            rows += [(q, int(tt)) for tt in ts]
    q = np.asarray([r[0] for r in rows], np.int32); t = np.asarray([r[1] for r in rows], np.int32)
    score = rng.randint(0, 40, size=q.size).astype(np.float64)          # many ties
    ref = op.baseline_labels(q.astype(np.int64), t.astype(np.int64), score, genome_of)
    for dt in (torch.float64, torch.float32):
        got = ops.segment_max_labels(dev(q, torch.int32), dev(t, torch.int32), dev(score, dt), dev(genome_of, torch.int32))
        assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("N,n", [(1, 0), (1, 3), (2, 3), (3, 3), (7, 3), (8, 4), (50, 1), (1000, 3), (4097, 7),
                                 (20000, 3)])
def test_neighbour_band_kernel_matches_oracle(N, n):
    """a8 (src/dataset.py:351-366): closed-form band kernel, bit-exact incl. order, also for N <= n."""
    from pangnn_b200 import ops
    ref = op.neighbour_band(N, n)
    got = ops.neighbour_band(N, n, DEV)
    assert got.dtype == torch.int64 and tuple(got.shape) == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)


def test_union_assembly_on_device_matches_oracle():
    """a11 (src/dataset.py:373-381): [sim ; nb], weights [w ; 1...], band written into the tail in place."""
    from pangnn_b200 import ops
    rng = np.random.default_rng(3)
    N, n, E = 500, 3, 1234
    ei = rng.integers(0, N, size=(2, E)).astype(np.int64)
    w = rng.random(E).astype(np.float32)
    ref_ei, ref_w = op.union_whole_graph(ei, w, op.neighbour_band(N, n))
    union = ops.union_index(torch.from_numpy(ei).to(DEV), N, n)
    uw = ops.union_weights(torch.from_numpy(w).to(DEV), union.size(1))
    assert np.array_equal(union.cpu().numpy(), ref_ei)
    assert np.array_equal(uw.cpu().numpy(), ref_w.astype(np.float32))
    # the band of a graph at the bench's size: count, first / last rows, sortedness by (src, dst)
    big = ops.neighbour_band(10 ** 6, 3, DEV)
    assert big.size(1) == 7 * 10 ** 6 - 12
    key = big[0] * 10 ** 6 + big[1]
    assert bool((key[1:] > key[:-1]).all())
    assert big[:, :4].t().tolist() == [[0, 0], [0, 1], [0, 2], [0, 3]]
    assert big[:, -1].tolist() == [10 ** 6 - 1, 10 ** 6 - 1]
    assert bool(((big[1] - big[0]).abs() <= 3).all())


def _write_tsv(path, rows, crlf=False, trailing_newline=True, comments=True):
    nl = "\r\n" if crlf else "\n"
    lines = []
    if comments:
        lines.append("# produced by a test")
    for i, (a, b, sc) in enumerate(rows):
        lines.append("\t".join([a, b, "0.963", "82", "3", "0", "1", "82", "82", "1", "82", "82", "1.000", "1.000",
                                "4.700E-48", sc]))
        if comments and i == 3:
            lines += ["", "#another comment"]
    with open(path, "w", newline="") as f:
        f.write(nl.join(lines) + (nl if trailing_newline else ""))


@pytest.mark.parametrize("crlf,trailing", [(False, True), (False, False), (True, True)])
def test_device_tsv_parser_matches_pandas(tmp_path, crlf, trailing):
    """MMseqs2 TSV -> (q, t, bits) on the device == the pandas loader: ids known / unknown,
    integer, decimal and exponent scores, comment and blank lines, CRLF, missing final newline."""
    from pangnn_b200 import preprocessing as pp
    rng = np.random.default_rng(5)
    ids = [f"{p}_{i:05d}" for p in ("FFOKMCCD", "KCMFMKKO", "AB") for i in range(1, 400)]
    pos = {g: i for i, g in enumerate(ids)}
    pool = ids + [f"ZZUNKNOWN_{i:05d}" for i in range(50)]
    scores = ["134", "165", "7", "900", "0", "12.5", "1.25e2", "3E+1", "0.5", "99999", "123456789.25", "-4"]
    rows = [(pool[int(rng.integers(len(pool)))], pool[int(rng.integers(len(pool)))], scores[int(rng.integers(len(scores)))])
            for _ in range(5000)]
    path = str(tmp_path / "hits.tsv")
    _write_tsv(path, rows, crlf=crlf, trailing_newline=trailing, comments=not crlf)
    for center in (True, False):
        rq, rt, rb = pp.load_similarity_score(path, pos, center_scores=center)
        q, t, b = pp.load_similarity_score_device(path, ids, center_scores=center, device=DEV)
        assert q.dtype == torch.int32 and b.dtype == torch.float64
        assert np.array_equal(q.cpu().numpy(), rq) and np.array_equal(t.cpu().numpy(), rt)
        assert np.array_equal(b.cpu().numpy(), rb)


def test_device_tsv_parser_scale_and_hash_table():
    """2e5 lines through the parser; the id hash table rejects duplicate ids."""
    from pangnn_b200 import _abi, ops
    ids = [f"GENOME{g:02d}_{i:06d}" for g in range(4) for i in range(5000)]
    rng = np.random.default_rng(1)
    qa, ta = rng.integers(0, len(ids), 200000), rng.integers(0, len(ids), 200000)
    sc = rng.integers(20, 2000, 200000)
    text = "".join(f"{ids[a]}\t{ids[b]}\t1\t2\t3\t4\t5\t6\t7\t8\t9\t10\t11\t12\t1e-9\t{s}\n" for a, b, s in zip(qa, ta, sc))
    dev_text = torch.frombuffer(bytearray(text.encode()), dtype=torch.uint8).to(DEV)
    q, t, b = ops.parse_hits_tsv(dev_text, ops.GeneIdTable(ids, DEV))
    assert np.array_equal(q.cpu().numpy(), qa) and np.array_equal(t.cpu().numpy(), ta)
    assert np.array_equal(b.cpu().numpy(), sc.astype(np.float64))
    with pytest.raises(_abi.PangnnError):
        ops.GeneIdTable(["A_1", "B_2", "A_1"], DEV)
