"""GPU: the torch.library operators (``torch.ops.pangnn.*``) pass ``torch.library.opcheck`` and the model run through
them is bit-identical to the ``autograd.Function`` spelling and equal to the reference goldens."""
import numpy as np
import pytest
import torch

from oracle.params import make_state_dict
from tests.helpers import GRAD_TOL, VARIANT_FLAGS, golden_graph, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model_run(golden, case, variant, layer):
    from pangnn_b200 import gnn, ops, setup
    setup.reset()
    ops.clear_cache()
    for k, v in VARIANT_FLAGS[variant].items():
        setattr(setup.args, k, v)
    fl = setup.args
    g = golden(case)
    graph = golden_graph(g, variant, device=DEV)
    model = gnn.AlternateGCN(DEV, None, False, dims=[fl.node_dim, fl.hidden_dim])
    model.load_state_dict(make_state_dict(fl.node_dim, fl.hidden_dim, fl.skip_connections, seed=1234))
    model = model.to(DEV)
    gnn.set_operator_layer(layer)
    try:
        model.fuse_embedding = False                      # same composition in both spellings
        loss, logits = model.forward_loss(graph, float(g[f"model/{variant}/pos_weight"]))
        loss.backward()
    finally:
        gnn.set_operator_layer("function")
    return (loss.item(), logits.cpu().numpy(), {k: p.grad.cpu().numpy() for k, p in model.named_parameters()
                                                if p.grad is not None})


@pytest.mark.parametrize("case,variant", [("sim5", "default"), ("c2", "union_skip"), ("c1", "union_n4")])
def test_library_layer_is_bit_identical_and_matches_goldens(golden, case, variant):
    lf, zf, gf = _model_run(golden, case, variant, "function")
    ll, zl, gl = _model_run(golden, case, variant, "library")
    assert lf == ll and np.array_equal(zf, zl)
    assert sorted(gf) == sorted(gl)
    for k in gf:
        assert np.array_equal(gf[k], gl[k]), k
    g = golden(case)
    key = f"model/{variant}"
    assert rel_err(zl, g[f"{key}/logits"]) < 1e-5
    assert abs(ll - float(g[f"{key}/loss"])) <= 1e-5 * abs(float(g[f"{key}/loss"]))
    for k in gl:
        if f"{key}/grad/{k}" in g.files:
            assert rel_err(gl[k], g[f"{key}/grad/{k}"]) < GRAD_TOL, k


def test_opcheck():
    """Schema, fake-tensor and autograd-registration checks of torch.library on real inputs."""
    from pangnn_b200 import ops, torch_ops  # noqa: F401
    torch.manual_seed(0)
    N, E = 300, 2000
    ei = torch.randint(0, N, (2, E), device=DEV)
    gs = ops.GraphStruct(ei, N)
    ent = gs.norm(None, need_src=True)
    x = torch.randn(N, 128, device=DEV, requires_grad=True)
    w = torch.randn(64, 128, device=DEV, requires_grad=True)
    b = torch.randn(64, device=DEV, requires_grad=True)
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.pangnn.node_linear.default, (x, w, b, ops.ACT_ELU, False), test_utils=tests)
    h = torch.randn(N, 64, device=DEV, requires_grad=True)
    d, s = gs.dst, gs.src
    torch.library.opcheck(torch.ops.pangnn.gcn_propagate.default,
                          (h, b, d.rowptr, d.col, ent["dst"], s.rowptr, s.col, ent["src"], N, ops.ACT_ELU), test_utils=tests)
    s32, d32 = gs.endpoints32
    pq = torch.randn(N, 128, device=DEV, requires_grad=True)
    vec = lambda: torch.randn(64, device=DEV, requires_grad=True)
    y = (torch.rand(E, device=DEV) < 0.3).float()
    torch.library.opcheck(torch.ops.pangnn.edge_score_bce.default,
                          (pq, s32, d32, None, None, vec(), torch.randn(64, 64, device=DEV, requires_grad=True), vec(),
                           torch.randn(1, 64, device=DEV, requires_grad=True), torch.randn(1, device=DEV, requires_grad=True),
                           y, 2.0, 1.0 / E, s.rowptr, s.perm, d.rowptr, d.perm), test_utils=tests)
