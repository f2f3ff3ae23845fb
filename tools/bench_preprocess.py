#!/usr/bin/env python3
"""Tuning aid: device candidate normalisation (sort/unique + segmented softmax/Q + compaction) on
the C3 hit table, with the per-kernel table from the torch profiler."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pangnn_b200 import ops
from pangnn_b200.simulate import simulate_hits
dev = torch.device("cuda:0")
s = simulate_hits(100000, 10, 0.5, 50, 10, seed=0)
N = s["num_genes"]
q, t = (torch.from_numpy(s[k]).to(dev) for k in ("q", "t"))
bits = torch.from_numpy(s["bits"]).to(dev)
genome_of, group_of = (torch.from_numpy(s[k]).to(dev) for k in ("genome_of", "group_of"))
n = q.numel()
def run():
    qs, ts, bs = ops.hits_sort_unique(q, t, bits, N)
    return ops.hits_normalize(qs, ts, bs, genome_of, group_of)
for _ in range(2): out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = run()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"hits": n, "edges_out": int(out[0].numel()), "ms_total": e0.elapsed_time(e1) / 5}))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]:
    print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:3d}  {e.key[:90]}")
