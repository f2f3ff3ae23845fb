"""§8f rank 4: the per-edge output table ``q_score_vs_logit.csv`` (``src/plot.py:473-504``) — byte-identical to the
file the unmodified reference writer produced from the same inputs (``tests/golden/make_golden_csv.py``)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs():
    return np.load(os.path.join(GOLD, "q_score_vs_logit_c1.npz"))


def test_writer_is_byte_identical_to_the_reference_file(tmp_path):
    from pangnn_b200 import postprocessing as post
    g = _inputs()
    path = post.write_q_score_vs_logit(torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_attr"]),
                                       torch.from_numpy(g["logits"]), torch.from_numpy(g["y"]), list(g["genes"]),
                                       g["base_labels"].tolist(), g["base_labels_raw"].tolist(),
                                       g["logit_baseline"].tolist(), str(tmp_path / "out" / "q_score_vs_logit.csv"))
    with open(path, "rb") as a, open(os.path.join(GOLD, "q_score_vs_logit_c1.csv"), "rb") as b:
        assert a.read() == b.read()


def test_writer_without_string_ids_and_baselines(tmp_path):
    """Simulated data has no gene strings (ids are positions); absent baselines are written as -1."""
    from pangnn_b200 import postprocessing as post
    ei = torch.tensor([[0, 2, 2], [1, 0, 1]])
    path = post.write_q_score_vs_logit(ei, torch.tensor([81.0, 1.0, 1.5, 1.0, 1.0]), torch.tensor([0.25, -1.0, 3.0]),
                                       torch.tensor([1.0, 0.0, 0.0]), path=str(tmp_path / "t.csv"))
    rows = open(path).read().splitlines()
    assert rows[0] == ",".join(post.Q_SCORE_VS_LOGIT_COLUMNS)
    assert rows[1] == "0,1,0,1,81.0,0.25,1,-1,-1,-1" and rows[3] == "2,1,2,1,1.5,3.0,0,-1,-1,-1" and len(rows) == 4


@pytest.mark.gpu
def test_device_logit_baseline_equals_the_reference(tmp_path):
    """The max-logit-candidate column: the segmented arg-max kernel (``pangnn_segment_max_labels``) on the
    reference's edge set against ``calculate_logit_baseline_labels`` run by the reference itself."""
    from pangnn_b200 import preprocessing as pp
    g = _inputs()
    ei, logits = g["edge_index"], g["logits"]
    order = np.lexsort((ei[1], ei[0]))                          # the kernel wants (src, dst) order
    genes = list(g["genes"])
    prefixes = {}
    genome_of = np.asarray([prefixes.setdefault(s.split("_")[0], len(prefixes)) for s in genes], dtype=np.int32)
    dev = "cuda:0"
    lab = pp.baseline_labels(torch.from_numpy(ei[0][order].astype(np.int32)).to(dev),
                             torch.from_numpy(ei[1][order].astype(np.int32)).to(dev),
                             torch.from_numpy(logits[order]).to(dev), torch.from_numpy(genome_of).to(dev))
    assert np.array_equal(lab.cpu().numpy(), g["logit_baseline"][order])
