#!/usr/bin/env python3
"""Tuning aid: DistModel (partitioned path, world 1) vs AlternateGCN on the same C3 graph, with a
kernel-level breakdown of one DistModel step from the torch profiler."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pangnn_b200 import dist as pd, ops, setup
from pangnn_b200.gnn import AlternateGCN
dev = torch.device("cuda:0")
setup.reset()
setup.args.union_edge_weights, setup.args.neighbours, setup.args.skip_connections = True, 3, True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pg = pd.PartitionedGraph.from_simulation(n, 10, 0.5, 50, 10, 0, 1, dev, seed=0)
torch.manual_seed(0)
model = AlternateGCN(dev, None, False).to(dev)
dm = pd.DistModel(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
def step():
    opt.zero_grad(set_to_none=False)
    loss, _ = dm.forward_loss(pg, pg.class_balance)
    loss.backward(); dm.allreduce_grads(); opt.step()
def timeit(f, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("DistModel world-1 step ms:", timeit(step))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
