#!/usr/bin/env python3
"""Which kernels of a C3 step are NOT ours, and which torch op launches them (torch.profiler, 2 steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pangnn_b200 import ops, setup, preprocessing as pp
from pangnn_b200.data import Data
from pangnn_b200.gnn import AlternateGCN
from pangnn_b200.simulate import simulate_hits
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
setup.reset()
for k, v in wl["flags"].items():
    setattr(setup.args, k, v)
fl = setup.args
dev = "cuda:0"
n, G, f, frags, shuf = wl["sim"]
s = simulate_hits(n, G, f, frags, shuf, seed=0)
N = s["num_genes"]
src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, t_norm=0.8,
                                         include_trivial=False, device=dev)
ei = torch.stack((src.long(), dst.long()))
g = Data(torch.ones(N, 1, device=dev), ei, None, y)
if fl.union_edge_weights:
    g.union_edge_index = ops.union_index(ei, N, fl.neighbours)
    g.edge_attr = ops.union_weights(w, g.union_edge_index.size(1))
else:
    g.edge_attr, g.neighbour_edge_index = w, pp.neighbour_band(N, fl.neighbours, dev)
pw = float(((y == 0).sum() / y.sum()).item())
model = AlternateGCN(dev, None, False).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
def step():
    opt.zero_grad(set_to_none=False)
    loss, _ = model.forward_loss(g, pw)
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ev = prof.events()
# map every device kernel to the innermost CPU op that encloses its launch
cpu = [e for e in ev if e.device_type == torch.autograd.DeviceType.CPU]
rows = {}
for e in ev:
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = e.name
    if "pangnn" in name:
        continue
    key = name[:90]
    r = rows.setdefault(key, [0, 0.0])
    r[0] += 1; r[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
print("non-pangnn device kernels over 2 steps:")
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / 2:9.1f} us/step {c / 2:5.1f} launches/step  {k}")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
