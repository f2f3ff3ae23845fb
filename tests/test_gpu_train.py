"""The re-hosted pangnn.py flow (pangnn_b200/train.py) end to end on a small simulated pan-genome:
train -> model.pkl -> inference branch, and the saved parameters evaluated by the ORACLE model give the
same predictions / F1 at the same threshold (BASELINE.json north_star)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def cli(tmp_path):
    from pangnn_b200 import ops, setup
    setup.reset()
    ops.clear_cache()
    out = str(tmp_path / "runs")
    argv = ["--simulate_dataset", "300", "3", "0.5", "10", "3", "--train", "-e", "3", "-b", "32", "-o", out,
            "-m", str(tmp_path / "absent.pkl"), "--seed", "1"]
    yield setup, argv, out
    setup.reset()
    ops.clear_cache()


def test_train_save_reload_and_oracle_agreement(cli):
    from pangnn_b200 import train
    from oracle.model import AlternateGCN as OracleGCN, Flags
    setup, argv, out = cli
    res = train.run(setup.parse(argv), device="cuda:0")
    hist = res["history"]
    assert len(hist) == 3 and all(np.isfinite(h["train_loss"]) and np.isfinite(h["val_loss"]) for h in hist)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]
    path = res["model_path"]
    assert os.path.exists(path)
    sd = torch.load(path, map_location="cpu")
    # model.pkl layout of the reference: keys, order and shapes
    om = OracleGCN(Flags())
    assert list(sd.keys()) == list(om.state_dict().keys())
    assert all(sd[k].shape == v.shape for k, v in om.state_dict().items())
    # ---- the oracle, loaded from the same model.pkl, on the same whole graph
    om.load_state_dict(sd, strict=True)
    g = res["dataset"].test[0]
    og = SimpleNamespace(x=g.x.cpu(), edge_index=g.edge_index.cpu(), edge_attr=g.edge_attr.cpu(), y=g.y.cpu(),
                         neighbour_edge_index=g.neighbour_edge_index.cpu())
    with torch.no_grad():
        ref_logits = om(og)
    ref_prob = torch.sigmoid(ref_logits)
    test = res["test"][0]
    assert float((test["logits"].cpu() - ref_logits).abs().max() / ref_logits.abs().max()) < 1e-5
    clear = (ref_prob - 0.5).abs() > 1e-5
    ref_pred = (ref_prob >= 0.5).long()
    assert torch.equal(test["pred"].cpu().long()[clear], ref_pred[clear])
    y = og.y.long()
    tp = int((ref_pred * y).sum()); fp = int((ref_pred * (1 - y)).sum()); fn = int(((1 - ref_pred) * y).sum())
    f1_ref = 2 * tp / max(2 * tp + fp + fn, 1)
    assert abs(test["f1"] - f1_ref) < 1e-3
    # ---- inference branch: same flags with an existing model file and no --train
    from pangnn_b200 import ops
    setup.reset(); ops.clear_cache()
    argv2 = ["--simulate_dataset", "300", "3", "0.5", "10", "3", "-m", path, "--seed", "1", "-o", out]
    res2 = train.run(setup.parse(argv2), device="cuda:0")
    assert res2["history"] == []
    assert torch.equal(res2["test"][0]["pred"], test["pred"])
    assert res2["test"][0]["f1"] == pytest.approx(test["f1"], abs=1e-12)
    # ---- output side: ortholog groups = connected components of the predicted edges, written as a table
    from oracle import postprocess as opp
    table = os.path.join(out, "holiest_of_all_tables.csv")
    assert os.path.exists(table)
    ei = res2["dataset"].test[0].edge_index.cpu().numpy()
    ref_labels = opp.component_labels(ei[0], ei[1], res2["test"][0]["pred"].cpu().numpy(), int(res2["dataset"].num_genes))
    assert np.array_equal(res2["test"][0]["group_labels"].cpu().numpy(), ref_labels)
    assert res2["test"][0]["groups"] == opp.groups(ref_labels)
    assert len(open(table).read().splitlines()) == len(res2["test"][0]["groups"])



@pytest.mark.parametrize("extra", [[], ["--union_edge_weights", "--neighbours", "3", "--skip_connections"]])
def test_graphed_batch_step_equals_the_eager_step(cli, extra):
    """pangnn_b200.graphs.GraphedBatchStep (the reference's -b 32 sub-graph regime replayed as one CUDA graph per size
    bucket, batches padded with isolated nodes / masked edges) == the eager step: same per-batch losses and the same
    parameters after an epoch, up to the summation order of the loss."""
    import copy
    from pangnn_b200 import train, ops
    from pangnn_b200.data import DeviceLoader
    from pangnn_b200.graphs import GraphedBatchStep
    setup, argv, out = cli
    argv = [a for a in argv] + extra
    argv[argv.index("-e") + 1] = "0"                            # dataset + model only
    res = train.run(setup.parse(argv), device="cuda:0")
    ds, m_eager = res["dataset"], res["model"]
    m_graph = copy.deepcopy(m_eager)
    pw = float(ds.class_balance)
    o_eager = torch.optim.Adam(m_eager.parameters(), lr=1e-3, capturable=True)
    o_graph = torch.optim.Adam(m_graph.parameters(), lr=1e-3, capturable=True)
    stepper = GraphedBatchStep(m_graph, o_graph, pw)
    le, lg = [], []
    bs = 2 if extra else 32                                     # 3-hop union lists of many sub-graphs exceed the single-launch CSR build
    for packed, ids, ids_dev in DeviceLoader(ds.train, batch_size=bs, shuffle=True, device="cuda:0", seed=3).iter_ids():
        batch = packed.collate(ids, ids_dev)
        o_eager.zero_grad(set_to_none=True)
        loss, logits = m_eager.forward_loss(batch, pw)
        loss.backward()
        o_eager.step()
        le.append(float(loss.detach()))
        del loss
        # default flags: collated straight into the bucket's static buffers; union flags: from a collated batch
        loss_g, logits_g = stepper(batch) if extra else stepper.step_ids(packed, ids, ids_dev)
        lg.append(float(loss_g.detach()))
        assert logits_g.shape == logits.shape
        if len(le) >= 24 and stepper.stats["replays"] >= 5:          # (Adam amplifies rounding differences over long runs)
            break
    assert stepper.stats["replays"] + stepper.stats["eager"] == len(le)
    assert stepper.stats["replays"] >= 5 and (extra or stepper.stats["eager"] == 0)   # (lists over 4096 edges: eager)
    assert extra or stepper.stats["captures"] < stepper.stats["replays"]           # buckets are reused
    assert np.allclose(le, lg, rtol=2e-5, atol=1e-7)
    for (k, a), b in zip(m_eager.state_dict().items(), m_graph.state_dict().values()):
        assert float((a - b).abs().max()) <= 2e-5 * max(float(a.abs().max()), 1e-3), k


def test_training_loop_with_cuda_graphs_follows_the_eager_loop(cli):
    """`--cuda_graphs`: the re-hosted pangnn.py loop with every batch step replayed as a CUDA graph (learning rate in
    device memory for the scheduler) reaches the same losses as the eager loop."""
    from pangnn_b200 import train, ops
    setup, argv, out = cli
    res_e = train.run(setup.parse(list(argv)), device="cuda:0")
    setup.reset(); ops.clear_cache()
    res_g = train.run(setup.parse(list(argv) + ["--cuda_graphs"]), device="cuda:0")
    he, hg = res_e["history"], res_g["history"]
    assert len(he) == len(hg) == 3
    for a, b in zip(he, hg):
        assert abs(a["train_loss"] - b["train_loss"]) < 1e-3 * abs(a["train_loss"])
        assert abs(a["val_loss"] - b["val_loss"]) < 1e-3 * abs(a["val_loss"])
    assert abs(res_e["test"][0]["f1"] - res_g["test"][0]["f1"]) < 2e-2
