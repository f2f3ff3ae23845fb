#!/usr/bin/env python3
"""Tuning aid: pangnn_gemm_tn (tcgen05 3xTF32) vs torch.mm(a.t(), b) on the model's shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pangnn_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
R = 1_000_000
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for m, k in ((64, 64), (128, 64), (64, 128), (128, 128)):
    a = torch.randn(R, m, device=dev); b = torch.randn(R, k, device=dev)
    t = timeit(lambda: ops.gemm_tn(a, b)); tl = timeit(lambda: torch.mm(a.t(), b))
    ref = a.double().t() @ b.double()
    err = float((ops.gemm_tn(a, b).double() - ref).abs().max() / ref.abs().max())
    errl = float((torch.mm(a.t(), b).double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({"m": m, "k": k, "ms": round(t, 4), "GBs": round(R * (m + k) * 4 / t / 1e6, 1), "lib_ms": round(tl, 4), "err": err, "err_lib": errl}))
