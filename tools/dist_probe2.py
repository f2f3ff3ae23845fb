#!/usr/bin/env python3
"""2-rank profile of one partitioned training step (rank 0 prints the kernel table)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from pangnn_b200 import dist as pd, ops, setup
from pangnn_b200.gnn import AlternateGCN
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
setup.reset()
setup.args.union_edge_weights, setup.args.neighbours, setup.args.skip_connections = True, 3, True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pg = pd.PartitionedGraph.from_simulation(n, 10 * world, 0.6786 if world == 2 else 0.5, 50, 10, rank, world, dev, seed=0)
torch.manual_seed(0)
model = AlternateGCN(dev, None, False).to(dev)
dm = pd.DistModel(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
def step():
    opt.zero_grad(set_to_none=False)
    loss, _ = dm.forward_loss(pg, pg.class_balance)
    loss.backward(); dm.allreduce_grads(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
if rank == 0: print("step ms", e0.elapsed_time(e1) / 5, flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
dist.barrier(); dist.destroy_process_group()
