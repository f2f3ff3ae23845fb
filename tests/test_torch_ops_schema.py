"""The torch.library layer (pangnn_b200/torch_ops.py): every operator is registered under ``torch.ops.pangnn`` with
the documented schema and a fake (meta) implementation that propagates shapes without touching the GPU."""
import torch


def test_operators_are_registered_with_schemas():
    from pangnn_b200 import torch_ops  # noqa: F401  (registration happens at import)
    ns = torch.ops.pangnn
    for name in ("node_linear", "act_bwd_bias", "gemm_tn", "gcn_norm", "gcn_aggregate", "gcn_propagate", "edge_score",
                 "edge_score_bce"):
        op = getattr(ns, name).default
        assert str(op._schema).startswith(f"pangnn::{name}(")
    assert "Tensor? bias" in str(ns.node_linear.default._schema)
    assert "-> (Tensor, Tensor, Tensor, Tensor)" in str(ns.edge_score_bce.default._schema)


def test_fake_implementations_propagate_shapes():
    from torch._subclasses.fake_tensor import FakeTensorMode
    from pangnn_b200 import ops, torch_ops  # noqa: F401
    P = torch.ops.pangnn
    with FakeTensorMode():
        N, E, F = 1000, 5000, 128
        x = torch.empty(N, F)
        w = torch.empty(64, F)
        y = P.node_linear(x, w, torch.empty(64), ops.ACT_ELU, False)
        assert tuple(y.shape) == (N, 64) and y.dtype == torch.float32
        rp, col = torch.empty(N + 1, dtype=torch.int64), torch.empty(E, dtype=torch.int32)
        val = torch.empty(E)
        z = P.gcn_propagate(y, None, rp, col, val, rp, col, val, N, ops.ACT_NONE)
        assert tuple(z.shape) == (N, 64)
        dis, v = P.gcn_norm(rp, col, col, None, N)
        assert tuple(dis.shape) == (N,) and tuple(v.shape) == (E,)
        pq = torch.empty(N, 128)
        vec = torch.empty(64)
        loss, logits, da1, grads = P.edge_score_bce(pq, col, col, None, None, vec, torch.empty(64, 64), vec,
                                                    torch.empty(1, 64), torch.empty(1), torch.empty(E), 2.0, 1.0 / E,
                                                    rp, col, rp, col)
        assert loss.dim() == 0 and tuple(logits.shape) == (E,) and tuple(da1.shape) == (E, 64)
        assert grads.numel() == ops.NGRADS
        assert tuple(P.gemm_tn(x, y).shape) == (F, 64)
