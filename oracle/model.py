"""Oracle restatement of the panGNN model forward, loss and (through torch-CPU autograd) backward.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows ``src/gnn.py:84-207`` (AlternateGCN)
and ``pangnn.py:98,203`` (loss).  Flags the reference reads from the global ``args`` inside
``forward`` (``src/gnn.py:128,132,143,171-172``) are explicit keyword arguments here.
"""
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn

from .gcn import GCNConv


class Flags(SimpleNamespace):
    """The subset of ``src/setup.py:15-50`` flags that reach the model."""

    def __init__(self, **kw):
        d = dict(union_edge_weights=False, base_model=False, skip_connections=False,
                 decoder="mlp", neighbours=1, node_dim=64, hidden_dim=128,
                 categorical_node=False)
        d.update(kw)
        super().__init__(**d)


class AlternateGCN(nn.Module):
    """Same construction order, parameter names and shapes as ``src/gnn.py:84-118``."""

    def __init__(self, flags, num_nodes=None):
        super().__init__()
        D, H = flags.node_dim, flags.hidden_dim
        self.flags = flags
        if flags.categorical_node:
            self.embedding = nn.Embedding(num_nodes, D)          # src/gnn.py:91-93
        else:
            self.embedding = nn.Linear(1, D)                     # src/gnn.py:97
        self.conv_in = GCNConv(D, H, add_self_loops=False)       # src/gnn.py:100
        self.conv_hidden = GCNConv(H, H, add_self_loops=False)   # src/gnn.py:101
        self.conv_out = GCNConv(H, D, add_self_loops=False)      # src/gnn.py:102
        self.linear_out = nn.Linear(H, D)                        # src/gnn.py:104
        self.activation_fct = nn.ELU()                           # src/gnn.py:108
        self.mlp = nn.Sequential(                                # src/gnn.py:110-116
            nn.Linear(2 * D + (1 if flags.skip_connections else 0), D), nn.ReLU(),
            nn.Linear(D, D), nn.ReLU(), nn.Linear(D, 1))

    def embed(self, graph):
        fl = self.flags
        node_embeddings = self.embedding(graph.x)                # src/gnn.py:125
        act = self.activation_fct
        if fl.union_edge_weights:                                # src/gnn.py:128-139
            nodes = act(self.conv_in(node_embeddings, graph.union_edge_index, graph.edge_attr))
            for _ in range(max(fl.neighbours - 2, 1)):
                nodes = act(self.conv_hidden(nodes, graph.union_edge_index, graph.edge_attr))
            nodes = act(self.conv_out(nodes, graph.union_edge_index))
        elif fl.base_model:                                      # src/gnn.py:143-150
            nodes = act(self.conv_in(node_embeddings, graph.edge_index, graph.edge_attr))
            nodes = act(self.linear_out(nodes))
        else:                                                    # src/gnn.py:153-166
            nodes = act(self.conv_in(node_embeddings, graph.edge_index, graph.edge_attr))
            nodes = act(self.conv_out(nodes, graph.neighbour_edge_index))
        return nodes

    def forward(self, graph):
        fl = self.flags
        nodes = self.embed(graph)
        src, dst = graph.edge_index[0], graph.edge_index[1]
        if "mlp" in fl.decoder:                                  # src/gnn.py:171-177
            if fl.skip_connections:
                cat = torch.cat((nodes[src], nodes[dst],
                                 graph.edge_attr[:src.numel()].unsqueeze(1)), dim=1)
            else:
                cat = torch.cat((nodes[src], nodes[dst]), dim=1)
            out = self.mlp(cat).squeeze(-1)
        if "cosine" in fl.decoder:                               # src/gnn.py:179,206-207
            out = F.cosine_similarity(nodes[src], nodes[dst], dim=1)
        if "dot" in fl.decoder:
            # src/gnn.py:180,202-204 is shape-broken ([E,D] @ [E,D]); the intended row-wise dot
            # is the one in MyGCN.decode, src/gnn.py:77-79 (documented deviation).
            out = (nodes[src] * nodes[dst]).sum(dim=1)
        return out


def bce_with_logits(logits, y, pos_weight):
    """``torch.nn.BCEWithLogitsLoss(pos_weight=pw)`` mean reduction, ``pangnn.py:98,203``."""
    return F.binary_cross_entropy_with_logits(
        logits, y, pos_weight=torch.as_tensor(pos_weight, dtype=logits.dtype))


def bce_closed_form(logits, y, pos_weight):
    """SURVEY.md A.7 closed form; used to pin the fused kernel's formula against torch's."""
    z = logits.double()
    y = y.double()
    sp = torch.clamp(-z, min=0) + torch.log1p(torch.exp(-z.abs()))
    return ((1 - y) * z + (1 + (pos_weight - 1) * y) * sp).mean()


def forward_backward(model, graph, pos_weight):
    """One reference training step's fwd+bwd (``pangnn.py:194-207``) -> logits, loss, grads."""
    model.zero_grad(set_to_none=True)
    logits = model(graph)
    loss = bce_with_logits(logits, graph.y, pos_weight)
    loss.backward()
    grads = OrderedDict((k, p.grad.detach().clone() if p.grad is not None else None)
                        for k, p in model.named_parameters())
    return logits.detach(), loss.detach(), grads
