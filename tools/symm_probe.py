#!/usr/bin/env python3
"""2-rank probe: torch symmetric memory (peer pointers + device barrier) and a P2P gather through our own
aggregation kernel reading the PEER's buffer."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, F = 1_000_000, 128
t = symm.empty(N, F, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "attrs", [a for a in dir(hdl) if not a.startswith("_")][:25], flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer(1 - rank, (N, F), torch.float32)
print(rank, "peer mean", float(peer[:1000].mean()), flush=True)
# our kernel gathering from the peer's memory
from pangnn_b200 import ops
E = 2_000_000
src = torch.randint(0, N, (E,), device=dev); dst = torch.sort(torch.randint(0, N, (E,), device=dev)).values
gs = ops.GraphStruct(torch.stack((src, dst)), N)
out = torch.empty(N, F, device=dev)
def run(x):
    return ops.gcn_aggregate(gs.dst.rowptr, gs.dst.col, None, x, N, out=out)
for name, x in (("local", t), ("peer", peer)):
    for _ in range(2): run(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, name, "aggregate ms", round(ms, 3), "GB/s gathered", round(E * F * 4 / ms / 1e6, 1), "check", float(out.sum() / (E * F)), flush=True)
hdl.barrier()
dist.destroy_process_group()
