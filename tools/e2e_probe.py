#!/usr/bin/env python3
"""Where does the end-to-end step spend its time?  H2D per tensor, CSR builds, gcn_norm, on the C3 batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangnn_b200 import ops, setup, preprocessing as pp
from pangnn_b200.simulate import simulate_hits
dev = torch.device("cuda:0")
s = simulate_hits(100000, 10, 0.5, 50, 10, seed=0); N = s["num_genes"]
src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, device=dev)
ei = torch.stack((src.long(), dst.long())); nb = pp.neighbour_band(N, 3, dev)
uei = torch.cat((ei, nb), 1); uw = torch.cat((w, torch.ones(nb.size(1), device=dev)))
host = {k: v.cpu().pin_memory() for k, v in dict(edge_index=ei, union_edge_index=uei, edge_attr=uw, y=y).items()}
def t(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for k, v in host.items():
    ms = t(lambda: v.to(dev, non_blocking=True))
    print(f"H2D {k}: {v.numel() * v.element_size() / 1e6:.0f} MB in {ms:.2f} ms = {v.numel() * v.element_size() / ms / 1e6:.1f} GB/s")
for name, e in (("scored", ei), ("union", uei)):
    for by in (True, False):
        print(f"csr_build {name} by_dst={by}: {t(lambda: ops.csr_build(e, N, by_dst=by)):.2f} ms")
gs = ops.GraphStruct(uei, N)
print("gcn_norm (deg+val) + apply:", round(t(lambda: (ops.gcn_norm(gs.dst, uw), ops.gcn_norm_apply(gs.src, uw, ops.gcn_norm(gs.dst, uw)[0]))), 2), "ms (includes one extra gcn_norm)")
print("int32 endpoints:", round(t(lambda: ei.to(torch.int32).contiguous()), 2), "ms")
