"""GPU parity: radix sort / scan / CSR build through the C ABI vs numpy (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n", [0, 1, 31, 2047, 2048, 2049, 4095, 4096, 4097, 100_003, 5_000_000])
def test_exclusive_scan(n):
    from pangnn_b200 import ops
    rng = np.random.RandomState(n % 97)
    x = rng.randint(0, 5, size=n).astype(np.int32)
    out, total = ops.exclusive_scan_u32(torch.from_numpy(x).to(DEV))
    ref = np.concatenate(([0], np.cumsum(x)[:-1])) if n else np.zeros(0)
    assert np.array_equal(out.cpu().numpy(), ref.astype(np.int32))
    assert int(total.item()) == int(x.sum())


@pytest.mark.parametrize("n,bits", [(1, 8), (33, 16), (4096, 20), (4097, 40), (70_001, 30),
                                    (1_000_003, 40), (3_000_000, 64)])
def test_sort_pairs_stable(n, bits):
    from pangnn_b200 import ops
    rng = np.random.RandomState(n % 89)
    hi = min(bits, 62)
    keys = rng.randint(0, 2 ** min(hi, 31), size=n).astype(np.int64)
    if hi > 31:
        keys = (keys << (hi - 31)) | rng.randint(0, 2 ** (hi - 31), size=n).astype(np.int64)
    keys[::7] = keys[0]                                   # plenty of duplicates: stability matters
    k, v = ops.sort_pairs_u64(torch.from_numpy(keys).to(DEV), None, key_bits=bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k.cpu().numpy(), keys[order])
    assert np.array_equal(v.cpu().numpy().astype(np.int64), order)


def _ref_csr(ei, N, by_dst):
    src, dst = ei
    row, col = (dst, src) if by_dst else (src, dst)
    order = np.lexsort((np.arange(row.size), col, row))
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    return np.cumsum(rowptr), col[order].astype(np.int32), order


@pytest.mark.parametrize("N,E", [(1, 0), (5, 0), (1, 3), (12, 16), (1895, 2017), (20000, 96350),
                                 (1000, 200_000), (300_000, 2_000_000)])
@pytest.mark.parametrize("by_dst", [True, False])
def test_csr_build(N, E, by_dst):
    from pangnn_b200 import ops
    rng = np.random.RandomState(N % 83 + E % 7)
    ei = rng.randint(0, N, size=(2, E)).astype(np.int64)
    if E > 10:
        ei[:, :5] = ei[:, 5:10]                            # duplicate edges are kept (no coalescing)
        ei[1, 10:20] = ei[0, 10:20]                        # self loops
    csr = ops.csr_build(torch.from_numpy(ei).to(DEV), N, by_dst=by_dst)
    rowptr, col, perm = _ref_csr(ei, N, by_dst)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(csr.col.cpu().numpy(), col)
    assert np.array_equal(csr.perm.cpu().numpy().astype(np.int64), perm)


def test_csr_build_golden_canonical(golden):
    """The canonical (src,dst)-sorted edge list of the reference graph is a fixed point: building
    the by-source CSR of it yields the identity permutation (bit-exact indexing)."""
    from pangnn_b200 import ops
    g = golden("c2")
    ei = torch.from_numpy(g["graph/edge_index"]).to(DEV)
    N = int(g["num_genes"])
    csr = ops.csr_build(ei, N, by_dst=False)
    assert np.array_equal(csr.perm.cpu().numpy(), np.arange(ei.size(1)))
    assert np.array_equal(csr.col.cpu().numpy(), g["graph/edge_index"][1])
    # property at full size: rowptr is monotone, ends at E, and col within a row is sorted
    rp = csr.rowptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == ei.size(1) and np.all(np.diff(rp) >= 0)


def _same_csr(a, b):
    return (torch.equal(a.rowptr, b.rowptr) and torch.equal(a.col, b.col) and torch.equal(a.perm, b.perm)
            and a.by_dst == b.by_dst and a.num_edges == b.num_edges)


@pytest.mark.parametrize("N,E", [(1, 0), (1, 7), (5, 40), (300, 5000), (70000, 200000), (1 << 16, 100000),
                                 ((1 << 16) + 1, 100000), (1000, 300000)])
def test_csr_transpose_equals_direct_build(N, E):
    """pangnn_csr_transpose: the other orientation from an existing CSR in half the radix passes, bit-exact
    (rowptr, col AND perm) with a direct build -- duplicates, empty rows, heavy rows included."""
    from pangnn_b200 import ops
    g = torch.Generator().manual_seed(N * 7 + E)
    ei = torch.randint(0, N, (2, E), generator=g)
    if E > 100:
        ei[:, : E // 10] = ei[:, E // 10: 2 * (E // 10)]        # duplicated edges: ties resolved by edge id
        ei[1, E // 2: E // 2 + E // 20] = 0                     # one heavy destination row
    ei = ei.to(DEV)
    for by_dst in (True, False):
        a = ops.csr_build(ei, N, by_dst=by_dst)
        assert _same_csr(ops.csr_transpose(a), ops.csr_build(ei, N, by_dst=not by_dst))


@pytest.mark.parametrize("N,n,E", [(1, 3, 0), (2, 3, 5), (3, 3, 10), (7, 3, 0), (50, 1, 300), (50, 0, 300),
                                   (5000, 3, 20000), (5000, 4, 60000), (100000, 3, 1000000)])
def test_csr_merge_band_equals_build_of_the_union_list(N, n, E):
    """pangnn_csr_merge_band: CSR of [sim ; band] from the sim CSR by row-wise merge == sort of the union
    list, incl. sim edges that coincide with band edges, duplicates, rows without sim edges, N <= n."""
    from pangnn_b200 import ops
    g = torch.Generator().manual_seed(N + 13 * E + n)
    ei = torch.randint(0, N, (2, E), generator=g)
    if E >= 10:
        k = E // 5
        ei[1, :k] = (ei[0, :k] + torch.randint(-n - 1, n + 2, (k,), generator=g)).clamp(0, N - 1)   # near the band
        ei[:, k: k + E // 10] = ei[:, : E // 10]                                                  # duplicates
    ei = ei.to(DEV)
    union = ops.union_index(ei, N, n)
    assert union.size(1) == E + ops.neighbour_band(N, n, DEV).size(1)
    for by_dst in (True, False):
        ref = ops.csr_build(union, N, by_dst=by_dst)
        got = ops.csr_merge_band(ops.csr_build(ei, N, by_dst=by_dst), n)
        assert _same_csr(got, ref)


@pytest.mark.parametrize("N,E", [(300, 5000), (70000, 200000), (1000, 300000)])
def test_sorted_edge_lists_skip_the_sort(N, E):
    """Edge lists in canonical (src, dst) order get their by-source CSR without a sort and the other orientation
    by transpose — both identical to the direct builds; an unsorted list falls back to the sort."""
    from pangnn_b200 import ops
    g = torch.Generator().manual_seed(N + E)
    ei = torch.randint(0, N, (2, E), generator=g)
    ei[:, : E // 10] = ei[:, E // 10: 2 * (E // 10)]                    # duplicates
    key = ei[0] * N + ei[1]
    sorted_ei = ei[:, torch.argsort(key, stable=True)].contiguous().to(DEV)
    assert ops.csr_from_sorted(ei.to(DEV), N) is None
    s = ops.csr_from_sorted(sorted_ei, N)
    assert s is not None and _same_csr(s, ops.csr_build(sorted_ei, N, by_dst=False))
    assert torch.equal(s.perm.long(), torch.arange(E, device=DEV))
    for edges in (sorted_ei, ei.to(DEV)):
        gs = ops.GraphStruct(edges, N)
        assert _same_csr(gs.dst, ops.csr_build(edges, N, by_dst=True))
        assert _same_csr(gs.src, ops.csr_build(edges, N, by_dst=False))


def test_sorted_list_with_large_gaps_and_out_of_range_ids():
    """rowptr of a sorted list whose sources leave long runs of empty rows (and a long empty tail) is
    filled in parallel; ids outside [0, N) raise instead of writing out of bounds."""
    from pangnn_b200 import _abi, ops
    N = 300_000
    src = torch.cat((torch.zeros(3000, dtype=torch.int64), torch.full((4000,), 150_000, dtype=torch.int64),
                     torch.full((2000,), 150_007, dtype=torch.int64)))
    dst = torch.arange(src.numel(), dtype=torch.int64) % N
    ei = torch.stack((src, dst)).to(DEV)
    s = ops.csr_from_sorted(ei, N)
    ref = np.searchsorted(src.numpy(), np.arange(N + 1), side="left")
    assert np.array_equal(s.rowptr.cpu().numpy(), ref)
    assert np.array_equal(s.col.cpu().numpy(), dst.numpy().astype(np.int32))
    bad = ei.clone()
    bad[1, 17] = N                                       # one past the end
    with pytest.raises(_abi.PangnnError):
        ops.csr_from_sorted(bad, N)
    bad = ei.clone()
    bad[0, 0] = -1
    with pytest.raises(_abi.PangnnError):
        ops.GraphStruct(bad, N)
