"""Pins oracle/model.py + oracle/gcn.py.  The model goldens come from the reference's own
src/gnn.py AlternateGCN (run under the shim, GCNConv = the restatement); the GCNConv restatement
itself is checked on hand-computed known answers (the reference holds none: parity unpinned)."""
import numpy as np
import pytest
import torch

from oracle import gcn
from oracle.model import bce_closed_form, bce_with_logits, forward_backward
from tests.helpers import golden_graph, oracle_model, rel_err

CASES = [("minimal", "base"), ("minimal", "default"), ("dummy", "default"), ("dummy", "base"),
         ("dummy", "union_skip"), ("c1", "default"), ("c1", "base"), ("c1", "union_skip"),
         ("c1", "cosine"), ("c1", "union_n4"), ("c2", "default"), ("c2", "union_skip"),
         ("sim5", "default"), ("sim5", "union_skip"), ("sim5", "union_n4"),
         ("sim5_trivial", "default")]


def test_gcn_norm_known_answer():
    # 3 nodes; edges (src->dst, w): 0->1 (4), 2->1 (5), 1->0 (9); node 2 has no in-edge.
    ei = torch.tensor([[0, 2, 1], [1, 1, 0]])
    w = torch.tensor([4.0, 5.0, 9.0])
    norm = gcn.gcn_norm(ei, w, 3)
    # deg (by target): [9, 9, 0] -> dis = [1/3, 1/3, 0]
    np.testing.assert_allclose(norm.numpy(), [4 / 9, 0.0, 1.0], rtol=1e-6)
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    out = gcn.gcn_propagate(x, ei, norm)
    np.testing.assert_allclose(out.numpy(), [[3.0, 4.0], [4 / 9, 8 / 9], [0, 0]], rtol=1e-6)


def test_gcn_conv_unweighted_and_state_dict_order():
    conv = gcn.GCNConv(2, 3, add_self_loops=False)
    assert list(conv.state_dict().keys()) == ["bias", "lin.weight"]
    assert conv.state_dict()["lin.weight"].shape == (3, 2)
    ei = torch.tensor([[0, 1, 1], [1, 0, 1]])
    x = torch.randn(2, 2)
    a = conv(x, ei)                                                        # None -> ones
    b = conv(x, ei, torch.ones(3))
    assert torch.equal(a, b)


@pytest.mark.parametrize("case,variant", CASES)
def test_model_matches_reference(golden, case, variant):
    g = golden(case)
    model, flags = oracle_model(variant)
    assert ",".join(model.state_dict().keys()) == str(g[f"model/{variant}/state_dict_keys"])
    graph = golden_graph(g, variant)
    pw = float(g[f"model/{variant}/pos_weight"])
    logits, loss, grads = forward_backward(model, graph, pw)
    assert rel_err(logits.numpy(), g[f"model/{variant}/logits"]) < 1e-6
    assert abs(loss.item() - float(g[f"model/{variant}/loss"])) <= 1e-6 * abs(float(g[f"model/{variant}/loss"]))
    for k, v in grads.items():
        key = f"model/{variant}/grad/{k}"
        if v is None:
            assert key not in g.files
            continue
        assert rel_err(v.numpy(), g[key]) < 1e-5, k


def test_bce_closed_form_matches_torch():
    torch.manual_seed(0)
    z = torch.randn(1000) * 5
    y = (torch.rand(1000) < 0.3).float()
    for pw in (1.0, 4.8, 0.3):
        a = bce_with_logits(z, y, pw).double()
        b = bce_closed_form(z, y, pw)
        assert abs(a - b) < 1e-6 * abs(b)


@pytest.mark.parametrize("case,variant", [("c1", "union_skip"), ("c2", "default"), ("sim5", "union_n4")])
def test_rank2_identity_of_the_first_layer(golden, case, variant):
    """The algebra behind pangnn_b200's folded first layer (rank1.cu), stated with the ORACLE's modules in fp64:
    for scalar node features x, GCNConv(Linear(1, D)(x)) = a u^T + c v^T + b with a = A_hat x, c = A_hat 1,
    u = W w_e, v = W b_e — for the reference's x = ones and for arbitrary x."""
    g = golden(case)
    model, flags = oracle_model(variant)
    model = model.double()
    graph = golden_graph(g, variant)
    ei = graph.union_edge_index if flags.union_edge_weights else graph.edge_index
    w = graph.edge_attr.double()
    N = graph.x.size(0)
    norm = gcn.gcn_norm(ei, w, N, dtype=torch.float64)
    for x in (graph.x.double(), torch.randn(N, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(1))):
        with torch.no_grad():
            ref = model.conv_in(model.embedding(x), ei, w)
            a = gcn.gcn_propagate(x, ei, norm).squeeze(1)
            c = gcn.gcn_propagate(torch.ones(N, 1, dtype=torch.float64), ei, norm).squeeze(1)
            W, b = model.conv_in.lin.weight, model.conv_in.bias
            u, v = W @ model.embedding.weight.squeeze(1), W @ model.embedding.bias
            got = a[:, None] * u[None, :] + c[:, None] * v[None, :] + b
        assert rel_err(got.numpy(), ref.numpy()) < 1e-12
