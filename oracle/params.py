"""Deterministic ``AlternateGCN`` state dicts for parity tests (numpy RNG, so the fixture does not
depend on torch's RNG stream).  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Key order, names and shapes are the ones ``AlternateGCN.state_dict()`` produces
(``src/gnn.py:91-116``).  Value ranges follow A.8 (Kaiming-uniform / Glorot) but
conv biases are drawn non-zero so that they are exercised.
"""
from collections import OrderedDict

import numpy as np
import torch


def make_state_dict(node_dim=64, hidden_dim=128, skip_connections=False, seed=1234,
                    categorical_nodes=0):
    rng = np.random.RandomState(seed)
    D, H = node_dim, hidden_dim

    def u(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))

    sd = OrderedDict()
    if categorical_nodes:
        sd["embedding.weight"] = torch.from_numpy(
            rng.standard_normal((categorical_nodes, D)).astype(np.float32))
    else:
        sd["embedding.weight"] = u((D, 1), 1.0)
        sd["embedding.bias"] = u((D,), 1.0)
    for name, fin, fout in (("conv_in", D, H), ("conv_hidden", H, H), ("conv_out", H, D)):
        sd[f"{name}.bias"] = u((fout,), 0.1)
        sd[f"{name}.lin.weight"] = u((fout, fin), float(np.sqrt(6.0 / (fin + fout))))
    sd["linear_out.weight"] = u((D, H), 1.0 / np.sqrt(H))
    sd["linear_out.bias"] = u((D,), 1.0 / np.sqrt(H))
    fin0 = 2 * D + (1 if skip_connections else 0)
    sd["mlp.0.weight"] = u((D, fin0), 1.0 / np.sqrt(fin0))
    sd["mlp.0.bias"] = u((D,), 1.0 / np.sqrt(fin0))
    sd["mlp.2.weight"] = u((D, D), 1.0 / np.sqrt(D))
    sd["mlp.2.bias"] = u((D,), 1.0 / np.sqrt(D))
    sd["mlp.4.weight"] = u((1, D), 1.0 / np.sqrt(D))
    sd["mlp.4.bias"] = u((1,), 1.0 / np.sqrt(D))
    return sd
