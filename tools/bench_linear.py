#!/usr/bin/env python3
"""Tuning aid: pangnn_node_linear (tcgen05 3xTF32) vs the library fp32 GEMM on the model's shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pangnn_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000

def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n, k in ((64, 64), (128, 64), (64, 128), (128, 128)):
    for kn in (False, True):
        x = torch.randn(M, k, device=dev)
        w = torch.randn((k, n) if kn else (n, k), device=dev) / k ** 0.5
        b = torch.randn(n, device=dev)
        out = torch.empty(M, n, device=dev)
        t_tc = timeit(lambda: ops.node_linear(x, w, b, ops.ACT_ELU, w_is_kn=kn, out=out))
        t_lib = timeit(lambda: torch.nn.functional.elu_(torch.addmm(b, x, w if kn else w.t())))
        t_mm = timeit(lambda: torch.mm(x, w if kn else w.t(), out=out))
        t_tc0 = timeit(lambda: ops.node_linear(x, w, None, ops.ACT_NONE, w_is_kn=kn, out=out))
        ref = torch.nn.functional.elu(x.double() @ (w.double() if kn else w.double().t()) + b.double())
        err = float((ops.node_linear(x, w, b, ops.ACT_ELU, w_is_kn=kn).double() - ref).abs().max() / ref.abs().max())
        err_lib = float((torch.nn.functional.elu_(torch.addmm(b, x, w if kn else w.t())).double() - ref).abs().max() / ref.abs().max())
        gb = M * (n + k) * 4 / 1e9
        print(json.dumps({"n": n, "k": k, "w_is_kn": kn, "tc_ms": round(t_tc, 4), "tc_noact_ms": round(t_tc0, 4), "tc_GBs": round(gb / t_tc * 1e3, 1),
                          "lib_addmm_elu_ms": round(t_lib, 4), "lib_mm_ms": round(t_mm, 4), "err_tc": err, "err_lib": err_lib}))
