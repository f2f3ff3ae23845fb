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
