#!/usr/bin/env python3
"""Tuning aid: implicit-band aggregation (pangnn_band_aggregate) vs the merged union CSR (pangnn_gcn_aggregate) on
the C3 union graph, F = 128 / 64, both orientations; checks bit-identity on the way."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pangnn_b200 import ops, preprocessing as pp
from pangnn_b200.simulate import simulate_hits

dev = torch.device("cuda:0")
n = int(os.environ.get("NB", 3))
s = simulate_hits(100000, 10, 0.5, 50, 10, seed=0)
N = s["num_genes"]
src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, device=dev)
ei = torch.stack((src.long(), dst.long()))
sim = ops.graph_struct(ei, N)
union = ops.union_index(ei, N, n)
gs = ops.graph_struct_union(union, N, sim, n)
wu = ops.union_weights(w, union.size(1))
ent = gs.norm(wu, need_src=True)
E, Es = union.size(1), ei.size(1)
res = {"N": N, "E_union": E, "E_sim": Es, "n": n}


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
for F in (128, 64):
    x = torch.randn(N, F, device=dev)
    bias = torch.randn(F, device=dev)
    o1, o2 = torch.empty(N, F, device=dev), torch.empty(N, F, device=dev)
    for by_dst in (True, False):
        csr = gs.dst if by_dst else gs.src
        val = ent["dst" if by_dst else "src"]
        ops.BAND_AGG["enabled"] = False
        t_m = timeit(lambda: ops.aggregate(gs, ent, x, by_dst, bias, ops.ACT_ELU, out=o1))
        ops.BAND_AGG["enabled"] = True
        t_b = timeit(lambda: ops.aggregate(gs, ent, x, by_dst, bias, ops.ACT_ELU, out=o2))
        ab = E * (8 + 4 * F) + N * 4 * F + 8 * (N + 1)
        res[f"F{F}_{'dst' if by_dst else 'src'}"] = {"merged_ms": round(t_m, 4), "band_ms": round(t_b, 4),
                                                      "alg_GBs_band": round(ab / t_b / 1e6, 1),
                                                      "bit_identical": bool(torch.equal(o1, o2))}
print(json.dumps(res))
