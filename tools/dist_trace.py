#!/usr/bin/env python3
"""Development aid: per-kernel device time of the partitioned C3 step on rank 0 (torch.profiler, 3 steps), to see
what the partitioned path adds to the single-GPU step.  torchrun --nproc-per-node N tools/dist_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
from pangnn_b200 import dist as pdist, setup
from pangnn_b200.gnn import AlternateGCN

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS["c3"]
setup.reset()
for k, v in wl["flags"].items():
    setattr(setup.args, k, v)
n, G0, f0, frags, shuf = wl["sim"]
G = G0 * world
f = bench.weak_scaling_fraction(n, G0, f0, G)
pg = pdist.PartitionedGraph.from_simulation(n, G, f, frags, shuf, rank, world, dev, seed=0)
torch.manual_seed(0)
model = AlternateGCN(dev, None, False).to(dev)
for p in model.parameters():
    dist.broadcast(p.data, 0)
dm = pdist.DistModel(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
pw = pg.class_balance


def step():
    opt.zero_grad(set_to_none=False)
    loss, _ = dm.forward_loss(pg, pw)
    loss.backward()
    dm.allreduce_grads()
    opt.step()


for _ in range(4):
    step()
dist.barrier(); torch.cuda.synchronize()
K = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(K):
        step()
    torch.cuda.synchronize()
if rank == 0:
    rows = [(e.key, e.device_time_total / K / 1e3, e.count / K) for e in prof.key_averages() if e.device_time_total > 0
            and e.device_type.name == "CUDA"]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"world {world}: device time per step {tot:.3f} ms over {sum(r[2] for r in rows):.0f} kernels")
    for k, ms, cnt in rows[:45]:
        print(f"  {ms:8.3f} ms  {cnt:5.1f}x  {k[:110]}")
dist.destroy_process_group()
