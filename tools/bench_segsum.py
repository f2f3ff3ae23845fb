#!/usr/bin/env python3
"""Scorer backward's node reduction: dpq[:, :64] = sum of da1 rows by source, dpq[:, 64:] by target (C3 sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangnn_b200 import ops
dev = "cuda:0"
N, E = 1_000_000, 9_970_000
g = torch.Generator(device=dev).manual_seed(0)
src = torch.sort(torch.randint(0, N, (E,), device=dev, generator=g)).values
dst = torch.randint(0, N, (E,), device=dev, generator=g)
gs = ops.graph_struct(torch.stack((src, dst)), N)
gs.src
da1 = torch.randn(E, 64, device=dev)
dpq = torch.empty(N, 128, device=dev)
def run():
    ops.gcn_aggregate(gs.src.rowptr, gs.src.perm, None, da1, N, out=dpq[:, :64])
    ops.gcn_aggregate(gs.dst.rowptr, gs.dst.perm, None, da1, N, out=dpq[:, 64:])
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"PANGNN_SEGSUM_UNROLL={os.environ.get('PANGNN_SEGSUM_UNROLL', '-')}: {ms:.3f} ms for both ({2 * E * 256 / ms / 1e6:.0f} GB/s of da1 reads), checksum {dpq.double().sum().item():.6e}")
