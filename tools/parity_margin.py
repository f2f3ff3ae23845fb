#!/usr/bin/env python3
"""How much of the 1e-5 budget do we use?  Logits of the CUDA path and of the fp32 oracle, both against
the oracle evaluated in fp64, over several seeds / flag variants of a smoke-sized simulated graph.  Two readings of
"1e-5 relative": the scale-relative one the tests use (max |a - b| / max |ref|) and the ELEMENT-WISE one
(|a - b| / max(|ref|, floor), floor = 1e-3 of the tensor's scale), reported as p50 / p99 / max."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from types import SimpleNamespace
from oracle import preprocess as op
from oracle.model import AlternateGCN as OracleGCN, Flags
from oracle.params import make_state_dict
from pangnn_b200 import ops, setup, preprocessing as pp
from pangnn_b200.data import Data
from pangnn_b200.gnn import AlternateGCN
from pangnn_b200.simulate import simulate_hits
dev = torch.device("cuda:0")
worst = 0.0
for variant in ({}, dict(union_edge_weights=True, neighbours=3, skip_connections=True)):
    for seed in range(6):
        setup.reset(); ops.clear_cache()
        for k, v in variant.items(): setattr(setup.args, k, v)
        fl = Flags(**variant)
        s = simulate_hits(400 + 50 * seed, 4, 0.5, 10, 3, seed=seed)
        N = s["num_genes"]
        src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, device=dev)
        ei = torch.stack((src.long(), dst.long())); nb = pp.neighbour_band(N, fl.neighbours, dev)
        g = Data(torch.ones(N, 1, device=dev), ei, w, y)
        if fl.union_edge_weights:
            g.union_edge_index = torch.cat((ei, nb), 1); g.edge_attr = torch.cat((w, torch.ones(nb.size(1), device=dev)))
        else:
            g.neighbour_edge_index = nb
        sd = make_state_dict(skip_connections=fl.skip_connections, seed=100 + seed)
        m = AlternateGCN(dev, None, False).to(dev); m.load_state_dict(sd)
        with torch.no_grad(): ours = m(g).double().cpu()
        og = SimpleNamespace(**{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in g.__dict__.items()})
        o32 = OracleGCN(fl); o32.load_state_dict(sd)
        with torch.no_grad(): ref32 = o32(og).double()
        o64 = OracleGCN(fl).double(); o64.load_state_dict({k: v.double() for k, v in sd.items()})
        og64 = SimpleNamespace(**{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in og.__dict__.items()})
        with torch.no_grad(): ref64 = o64(og64)
        sc = ref64.abs().max()
        e_ours, e_o32, e_pair = float((ours - ref64).abs().max() / sc), float((ref32 - ref64).abs().max() / sc), float((ours - ref32).abs().max() / sc)
        worst = max(worst, e_pair)
        ew = ((ours - ref64).abs() / ref64.abs().clamp_min(1e-3 * float(sc))).numpy()
        ew32 = ((ref32 - ref64).abs() / ref64.abs().clamp_min(1e-3 * float(sc))).numpy()
        print(f"variant={'union' if variant else 'default'} seed={seed} N={N} E={ei.size(1)}: ours-vs-fp64 {e_ours:.2e}  oracle32-vs-fp64 {e_o32:.2e}  ours-vs-oracle32 {e_pair:.2e}"
              f"  | element-wise ours p50/p99/max {np.percentile(ew, 50):.1e}/{np.percentile(ew, 99):.1e}/{ew.max():.1e}"
              f"  oracle32 {np.percentile(ew32, 50):.1e}/{np.percentile(ew32, 99):.1e}/{ew32.max():.1e}")
print("worst ours-vs-oracle32", worst)
