#!/usr/bin/env python3
"""rank1_aggregate forward / backward alone, C3-like sizes (N = 1e6, 1.7e7 edges, F = 128), against the gather-based
aggregation they replace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangnn_b200 import ops
dev = "cuda:0"
N, E, F = 1_000_000, 17_000_000, 128
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.stack((torch.randint(0, N, (E,), device=dev, generator=g), torch.randint(0, N, (E,), device=dev, generator=g)))
w = torch.rand(E, device=dev, generator=g) + 0.5
gs = ops.graph_struct(ei, N)
ent = gs.norm(w, need_src=True)
x = torch.ones(N, 1, device=dev)
a, c = ops.rank1_vectors(gs.dst, ent["dst"], x)
w_e = torch.randn(64, 1, device=dev, generator=g); b_e = torch.randn(64, device=dev, generator=g)
W = (torch.randn(F, 64, device=dev, generator=g) * 0.1).requires_grad_(True); b = torch.randn(F, device=dev, generator=g) * 0.1
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
z = ops.RankOneAggFn.apply(a, c, w_e, b_e, W, b, ops.ACT_ELU, gs.dst, ent["dst"], N)
dz = torch.randn(N, F, device=dev, generator=g)
fwd = timed(lambda: ops.RankOneAggFn.apply(a, c, w_e, b_e, W, b, ops.ACT_ELU, gs.dst, ent["dst"], N))
def bwd():
    W.grad = None
    z.backward(dz, retain_graph=True)
tb = timed(bwd)
h1 = ops.RankOneFn.apply(a, c, w_e, b_e, W, b, ops.ACT_ELU).detach()
out = torch.empty(N, F, device=dev)
ga = timed(lambda: ops.gcn_aggregate(gs.dst.rowptr, gs.dst.col, ent["dst"], h1, N, out=out))
ref = out
err = float((z.detach() - ref).abs().max() / ref.abs().max())
print(f"rank1_aggregate fwd {fwd:.3f} ms, bwd {tb:.3f} ms; gather aggregation of the same rows {ga:.3f} ms; fwd rel err {err:.2e}")
