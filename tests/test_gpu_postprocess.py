"""Ortholog groups from predictions: device connected components against scipy."""
import numpy as np
import pytest
import torch

from oracle import postprocess as opp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(src, dst, sel, N):
    from pangnn_b200 import ops
    t = lambda a: torch.as_tensor(np.asarray(a), device=DEV)
    got = ops.connected_components(t(src), t(dst), N, select=None if sel is None else t(sel))
    ref = opp.component_labels(src, dst, sel, N)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("N,E,p", [(1, 0, 1.0), (10, 0, 1.0), (50, 40, 1.0), (1000, 700, 0.5), (200000, 150000, 0.7),
                                   (200000, 400000, 0.3), (1000000, 3000000, 0.2)])
def test_components_random(N, E, p):
    rng = np.random.default_rng(N + E)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    sel = (rng.random(E) < p).astype(np.int32)
    _check(src, dst, sel, N)
    _check(src, dst, None, N)


def test_components_long_chains_and_stars():
    N = 300000
    chain = np.arange(N - 1)
    rng = np.random.default_rng(0)
    o = rng.permutation(N - 1)                                  # one path through every node, edges in random order
    _check(chain[o], chain[o] + 1, None, N)
    _check(chain[::-1].copy() + 1, chain[::-1].copy(), None, N)  # descending order
    hub = np.zeros(N - 1, dtype=np.int64)
    _check(hub + (N - 1), chain, None, N)                       # star around the LARGEST id


def test_groups_and_table(tmp_path):
    from pangnn_b200 import postprocessing as post
    ei = torch.tensor([[0, 1, 5, 7, 8, 2], [1, 2, 6, 7, 9, 0]], device=DEV)
    pred = torch.tensor([1, 1, 1, 1, 0, 1], device=DEV)
    labels, groups = post.ortholog_groups(ei, pred, 10)
    assert labels.tolist() == [0, 0, 0, 3, 4, 5, 5, 7, 8, 9]
    assert groups == [[0, 1, 2], [5, 6]] == opp.groups(labels.cpu().numpy())
    names = [f"G_{i:03d}" for i in range(10)]
    path = post.write_groups_file(groups, names, str(tmp_path / "out" / "groups.csv"))
    assert open(path).read() == "group_0, G_000, G_001, G_002\ngroup_1, G_005, G_006\n"
