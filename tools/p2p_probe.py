#!/usr/bin/env python3
"""2-rank probe: NCCL send/recv and all_reduce bandwidth + the halo exchange of pangnn_b200.dist."""
import os, sys, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def timed(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mb in (1, 16, 51, 256):
    n = mb * 1024 * 1024 // 4
    a = torch.randn(n, device=dev); b = torch.empty(n, device=dev)
    peer = 1 - rank
    def sr():
        ops = [dist.P2POp(dist.irecv, b, peer), dist.P2POp(dist.isend, a, peer)]
        for w in dist.batch_isend_irecv(ops): w.wait()
    t = timed(sr)
    t2 = timed(lambda: dist.all_reduce(a))
    if rank == 0:
        print(f"{mb} MB: sendrecv {t:.3f} ms = {mb / t:.1f} GB/s per direction; all_reduce {t2:.3f} ms", flush=True)
if rank == 0:
    print("can_access_peer", torch.cuda.can_device_access_peer(0, 1), flush=True)
dist.destroy_process_group()
