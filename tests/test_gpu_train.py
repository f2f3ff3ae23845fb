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

