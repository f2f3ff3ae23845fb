#!/usr/bin/env python3
"""Kernel tuning aid (not part of the product path): times pangnn_gcn_aggregate (F = 128 and 64,
both CSR orientations) on the C3 union graph.  During the round-1 sweep the library was built with
several (UNROLL, min-CTAs/SM) instantiations selected by an environment variable; the winners are
now hard-wired in csrc/gcn.cu and the switch is gone.  Results of the sweep: profiles/r01_agg_sweep.md.
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pangnn_b200 import ops, preprocessing as pp
from pangnn_b200.simulate import simulate_hits

dev = torch.device("cuda:0")
variant = int(os.environ.get("PANGNN_AGG_VARIANT", "0"))
cache = "/tmp/tune_agg_graph.pt"
if os.path.exists(cache):
    d = torch.load(cache)
    uei, w, N = d["uei"].to(dev), d["w"].to(dev), d["N"]
else:
    s = simulate_hits(100000, 10, 0.5, 50, 10, seed=0)
    N = s["num_genes"]
    src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, device=dev)
    nb = pp.neighbour_band(N, 3, dev)
    uei = torch.cat((torch.stack((src.long(), dst.long())), nb), dim=1)
    w = torch.cat((w, torch.ones(nb.size(1), device=dev)))
    torch.save({"uei": uei.cpu(), "w": w.cpu(), "N": N}, cache)
gs = ops.graph_struct(uei, N)
ent = gs.norm(w, need_src=True)
E = uei.size(1)
res = {"variant": variant}
torch.manual_seed(0)
for F in (128, 64):
    x = torch.randn(N, F, device=dev)
    bias = torch.randn(F, device=dev)
    out = torch.empty(N, F, device=dev)
    for name, csr, val in (("dst", gs.dst, ent["dst"]), ("src", gs.src, ent["src"])):
        f = lambda: ops.gcn_aggregate(csr.rowptr, csr.col, val, x, N, bias, ops.ACT_ELU, out=out)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ab = E * (8 + 4 * F) + N * 4 * F + 8 * (N + 1)
        ref_path = f"/tmp/tune_agg_ref_{F}_{name}.pt"
        if variant == 9:
            torch.save(out.cpu(), ref_path); err = 0.0
        elif os.path.exists(ref_path):
            ref = torch.load(ref_path).to(dev)
            err = float((out - ref).abs().max() / ref.abs().max())
        else:
            err = None
        res[f"F{F}_{name}"] = {"ms": round(ms, 4), "alg_GBs": round(ab / ms / 1e6, 1), "err_vs_v9": err}
# segment-sum use (val = NULL, col = perm): scorer backward shape, 1e7 edges x 64
print(json.dumps(res))
