"""Oracle restatement of ``torch_geometric.nn.GCNConv(in, out, add_self_loops=False)``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: PyG is a third-party,
un-vendored, un-pinned dependency of the reference; this file restates its
published algorithm.  Reference call sites: ``src/gnn.py:100-102`` (construction) and
``src/gnn.py:129,135,138,147,158,165`` (forward with / without ``edge_weight``).

PyG defaults that matter: improved=False, cached=False, normalize=True, bias=True, aggr='add',
flow='source_to_target' (row = source j, col = target i; degree is taken over TARGETS only).
"""
import math

import torch
from torch import nn


def gcn_norm(edge_index, edge_weight, num_nodes, dtype=torch.float32):
    """``gcn_norm(..., add_self_loops=False)``: deg by target, ``deg^-1/2`` (inf -> 0)."""
    row, col = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones(row.numel(), dtype=dtype, device=row.device)
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype, device=row.device)
    deg.scatter_add_(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0.0)
    return dis[row] * edge_weight * dis[col]


def gcn_propagate(x_lin, edge_index, norm):
    """out[i] = sum_{e: col_e = i} norm_e * x_lin[row_e]  (gather by source, scatter-add by target)."""
    row, col = edge_index[0], edge_index[1]
    out = torch.zeros_like(x_lin)
    out.index_add_(0, col, norm.unsqueeze(1) * x_lin.index_select(0, row))
    return out


class _Lin(nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)

    def forward(self, x):
        return x @ self.weight.t()


class GCNConv(nn.Module):
    """state_dict keys: ``bias`` then ``lin.weight``."""

    def __init__(self, in_channels, out_channels, add_self_loops=False, **_):
        super().__init__()
        assert not add_self_loops, "the reference only ever passes add_self_loops=False"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Lin(in_channels, out_channels)

    def forward(self, x, edge_index, edge_weight=None):
        norm = gcn_norm(edge_index, edge_weight, x.size(0), x.dtype)
        return gcn_propagate(self.lin(x), edge_index, norm) + self.bias
