"""Shared helpers for the parity tests."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle.model import AlternateGCN as OracleGCN, Flags
from oracle.params import make_state_dict

VARIANT_FLAGS = {
    "default": dict(),
    "base": dict(base_model=True),
    "union_skip": dict(union_edge_weights=True, neighbours=3, skip_connections=True),
    "union_n4": dict(union_edge_weights=True, neighbours=4),
    "cosine": dict(decoder="cosine"),
}


def golden_graph(g, variant, device="cpu"):
    """Rebuild the graph object the golden model variant ran on."""
    t = lambda a, dt=None: torch.as_tensor(np.asarray(a), dtype=dt, device=device)
    graph = SimpleNamespace()
    graph.x = t(g["graph/x"], torch.float32)
    graph.edge_index = t(g["graph/edge_index"], torch.long)
    graph.y = t(g["graph/y"], torch.float32)
    key = f"model/{variant}"
    if f"{key}/union_edge_index" in g.files:
        graph.union_edge_index = t(g[f"{key}/union_edge_index"], torch.long)
        graph.edge_attr = t(g[f"{key}/edge_attr"], torch.float32)
    else:
        graph.edge_attr = t(g["graph/edge_attr"], torch.float32)
        nb = f"{key}/neighbour_edge_index"
        graph.neighbour_edge_index = t(g[nb] if nb in g.files else g["graph/neighbour_edge_index"],
                                       torch.long)
    return graph


def oracle_model(variant, seed=1234, **kw):
    flags = Flags(**VARIANT_FLAGS[variant], **kw)
    m = OracleGCN(flags)
    m.load_state_dict(make_state_dict(flags.node_dim, flags.hidden_dim, flags.skip_connections,
                                      seed=seed), strict=True)
    return m, flags


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): the 1e-5 'relative (fp32)' bar of BASELINE.json is applied to
    the tensor's scale, since individual elements can be arbitrarily close to zero."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max() / scale) if b.size else 0.0


def assert_close(name, got, ref, tol):
    """Compact failure message (never dumps the arrays)."""
    err = rel_err(got, ref)
    if not err < tol:
        raise AssertionError(f"{name}: rel_err {err:.3e} >= {tol:.1e} (shape {np.shape(ref)})")


# Parameter / input gradients are long fp32 reductions (over all nodes / edges) of terms that cancel,
# and they pass through ReLU / ELU kinks where a last-ulp difference in a pre-activation legitimately
# flips a derivative.  The 1e-5 bar of BASELINE.json is stated for embeddings, probabilities and
# loss; gradients are held to 1e-4 of the tensor's scale against the fp32 reference.
GRAD_TOL = 1e-4
