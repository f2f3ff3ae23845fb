"""Parity gate of the genome-partitioned arm of bench.py: before anything is timed at world N, the SAME
``DistModel`` / transport the bench is about to measure must reproduce

  1. the reference-derived golden vectors (``tests/golden/*.npz``: logits, loss, all-reduced gradients of the
     reference's ``AlternateGCN`` on graphs the unmodified reference built) when the golden graph is partitioned
     over the N ranks (``PartitionedGraph.from_global``), and
  2. the single-GPU whole-graph path of this package on a small simulated pan-genome built the way the bench builds
     its workload (``PartitionedGraph.from_simulation``: whole genomes per rank, halo = the neighbouring genomes,
     i.e. the fused GEMM-epilogue push over peer memory when that transport is active).

Checker only: reads committed fixtures, never ``oracle/``.  Tolerances: logits / loss 1e-5, gradients 1e-4,
relative to the tensor's scale (tests/helpers.py)."""
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
VARIANT_FLAGS = {"default": dict(), "union_skip": dict(union_edge_weights=True, neighbours=3, skip_connections=True)}
GOLDEN_CASES = (("c2", "union_skip"), ("sim5", "default"))
TOL, GRAD_TOL = 1e-5, 1e-4


def _rel(a, b, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = max(float(np.abs(b).max()) if scale is None and b.size else (scale or 0.0), 1e-30)
    return float(np.abs(a - b).max() / s) if b.size else 0.0


def _golden_graph(g, variant, device):
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt, device=device)
    graph = SimpleNamespace(x=t(g["graph/x"], torch.float32), edge_index=t(g["graph/edge_index"], torch.long),
                            y=t(g["graph/y"], torch.float32))
    key = f"model/{variant}"
    if f"{key}/union_edge_index" in g.files:
        graph.union_edge_index = t(g[f"{key}/union_edge_index"], torch.long)
        graph.edge_attr = t(g[f"{key}/edge_attr"], torch.float32)
    else:
        graph.edge_attr = t(g["graph/edge_attr"], torch.float32)
        nb = f"{key}/neighbour_edge_index"
        graph.neighbour_edge_index = t(g[nb] if nb in g.files else g["graph/neighbour_edge_index"], torch.long)
    return graph


def _set_flags(flags):
    from pangnn_b200 import ops, setup
    setup.reset()
    ops.clear_cache()
    for k, v in flags.items():
        setattr(setup.args, k, v)
    return setup.args


def _state_dict(skip, device):
    p = np.load(os.path.join(GOLD, "params_seed1234.npz"))
    pre = f"skip{int(bool(skip))}/"
    return {k[len(pre):]: torch.from_numpy(p[k]).to(device) for k in p.files if k.startswith(pre)}


def _max_over_ranks(x, dev):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _golden_case(case, variant, rank, world, dev):
    from pangnn_b200 import dist as pd
    from pangnn_b200.gnn import AlternateGCN
    fl = _set_flags(VARIANT_FLAGS[variant])
    g = np.load(os.path.join(GOLD, f"{case}.npz"))
    graph = _golden_graph(g, variant, dev)
    model = AlternateGCN(dev, None, False, dims=[fl.node_dim, fl.hidden_dim]).to(dev)
    model.load_state_dict(_state_dict(fl.skip_connections, dev), strict=True)
    pg = pd.PartitionedGraph.from_global(graph, graph.x.size(0), rank, world)
    dm = pd.DistModel(model)
    key = f"model/{variant}"
    loss, logits = dm.forward_loss(pg, float(g[f"{key}/pos_weight"]))
    loss.backward()
    dm.allreduce_grads()
    infer = dm(pg)
    ref_logits = g[f"{key}/logits"]
    ids = pg.scored_edge_ids.cpu().numpy()
    scale = float(np.abs(ref_logits).max())
    e_logit = max(_rel(logits.cpu().numpy(), ref_logits[ids], scale), _rel(infer.cpu().numpy(), ref_logits[ids], scale))
    total = loss.detach().double().clone()
    dist.all_reduce(total)
    cover = torch.tensor([float(ids.size)], dtype=torch.float64, device=dev)
    dist.all_reduce(cover)
    ref_loss = float(g[f"{key}/loss"])
    e_loss = abs(float(total.item()) - ref_loss) / abs(ref_loss)
    e_grad = 0.0
    for k, p in model.named_parameters():
        gk = f"{key}/grad/{k}"
        if gk in g.files and p.grad is not None:
            e_grad = max(e_grad, _rel(p.grad.cpu().numpy(), g[gk]))
    transport = "nvlink peer memory" if pg.conv.plan.p2p is not None else "nccl send/recv"
    ok_cover = int(cover.item()) == ref_logits.size                       # every scored edge owned exactly once
    return dict(case=f"{case}/{variant}", logits=_max_over_ranks(e_logit, dev), loss=e_loss,
                grads=_max_over_ranks(e_grad, dev), cover=ok_cover, transport=transport)


def _simulated_case(rank, world, dev):
    """Partitioned build + step vs the whole-graph single-GPU path on the same simulated pan-genome."""
    from pangnn_b200 import dist as pd, ops
    from pangnn_b200 import preprocessing as pp
    from pangnn_b200.data import Data
    from pangnn_b200.gnn import AlternateGCN
    from pangnn_b200.simulate import simulate_hits
    fl = _set_flags(VARIANT_FLAGS["union_skip"])
    n, G = 400, 2 * world
    s = simulate_hits(n, G, 0.5, 10, 3, seed=11)
    N = n * G
    src, dst, w, y = pp.normalize_sim_scores(s["q"], s["t"], s["bits"], s["genome_of"], s["group_of"], num_nodes=N, device=dev)
    ei = torch.stack((src.long(), dst.long()))
    union = ops.union_index(ei, N, fl.neighbours)
    whole = Data(torch.ones(N, 1, device=dev), ei, ops.union_weights(w, union.size(1)), y)
    whole.union_edge_index = union
    pw = float(((y == 0).sum() / y.sum()).item())
    sd = _state_dict(True, dev)
    ref_model = AlternateGCN(dev, None, False).to(dev)
    ref_model.load_state_dict(sd, strict=True)
    rloss, rlogits = ref_model.forward_loss(whole, pw)
    rloss.backward()
    ops.clear_cache()
    model = AlternateGCN(dev, None, False).to(dev)
    model.load_state_dict(sd, strict=True)
    pg = pd.PartitionedGraph.from_simulation(n, G, 0.5, 10, 3, rank, world, dev, seed=11)
    dm = pd.DistModel(model)
    loss, logits = dm.forward_loss(pg, pw)
    loss.backward()
    dm.allreduce_grads()
    # locally scored edges in GLOBAL ids -> their position in the whole graph's (src, dst)-sorted list
    lg, lo = pg.scored, pg.bounds[rank]
    glob = torch.cat((torch.arange(lo, lo + lg.n_own, device=dev), lg.halo_ids.to(dev)))
    key = glob[lg.edge_index[0]] * N + glob[lg.edge_index[1]]
    pos = torch.searchsorted(ei[0] * N + ei[1], key)
    ok_cover = bool((pos < ei.size(1)).all()) and bool(((ei[0] * N + ei[1])[pos.clamp_max(ei.size(1) - 1)] == key).all())
    scale = float(rlogits.abs().max())
    e_logit = _rel(logits.cpu().numpy(), rlogits[pos.clamp_max(ei.size(1) - 1)].cpu().numpy(), scale)
    total = loss.detach().double().clone()
    dist.all_reduce(total)
    e_loss = abs(float(total.item()) - float(rloss.item())) / abs(float(rloss.item()))
    e_grad = 0.0
    for (k, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        if p.grad is not None and q.grad is not None:
            e_grad = max(e_grad, _rel(p.grad.cpu().numpy(), q.grad.cpu().numpy()))
    cnt = torch.tensor([float(pg.y.numel())], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt)
    ok_cover = ok_cover and int(cnt.item()) == ei.size(1)
    transport = "nvlink peer memory" if pg.conv.plan.p2p is not None else "nccl send/recv"
    fused = bool(pg.conv.plan.p2p is not None and pg.scored.plan.push_maps() is not None)
    return dict(case=f"simulated {n}x{G} union_skip (from_simulation vs whole graph on one GPU)",
                logits=_max_over_ranks(e_logit, dev), loss=e_loss, grads=_max_over_ranks(e_grad, dev),
                cover=bool(_max_over_ranks(0.0 if ok_cover else 1.0, dev) == 0.0), transport=transport,
                fused_epilogue_push=fused)


def run(rank, world, dev):
    """-> the ``parity_gate`` object of the bench line; raises SystemExit(3) on every rank when a case fails."""
    cases = [_golden_case(c, v, rank, world, dev) for c, v in GOLDEN_CASES]
    cases.append(_simulated_case(rank, world, dev))
    worst = {k: max(c[k] for c in cases) for k in ("logits", "loss", "grads")}
    ok = all(c["cover"] for c in cases) and worst["logits"] < TOL and worst["loss"] < TOL and worst["grads"] < GRAD_TOL
    gate = {"passed": bool(ok), "world": world, "cases": cases, "max_rel_err": worst,
            "tolerance": {"logits": TOL, "loss": TOL, "grads": GRAD_TOL, "norm": "max |a-b| / max |ref|"},
            "transport": cases[-1]["transport"]}
    if not ok:
        if rank == 0:
            import json, sys
            print("PARITY GATE FAILED: " + json.dumps(gate), file=sys.stderr, flush=True)
        raise SystemExit(3)
    return gate
