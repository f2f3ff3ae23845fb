"""Host-side logic of the re-hosted training loop (pangnn_b200/train.py, setup.py): metric formulas, the
Youden threshold of --dynamic_binary_threshold, flags that are accepted but inert.  No GPU needed."""
import logging

import numpy as np
import pytest
import torch


def test_prf_never_divides_by_zero():
    """tp == 0 with fp > 0 and fn > 0 used to raise in evaluate() (eps = 0, src/predict.py:114-121)."""
    from pangnn_b200.train import prf
    assert prf(10, 3, 4, 0, eps=0.0) == (0.0, 0.0, 0.0)
    assert prf(10, 0, 0, 0, eps=0.0) == (0.0, 0.0, 0.0)
    assert prf(1, 1, 1, 1, eps=0.0) == (0.5, 0.5, 0.5)
    p, r, f = prf(5, 1, 3, 4)                                   # training-loop epsilons, pangnn.py:291-294
    assert p == pytest.approx(4 / (5 + 1e-10)) and r == pytest.approx(4 / (7 + 1e-10))
    assert f == pytest.approx(2 * p * r / (p + r + 1e-10))


@pytest.mark.parametrize("n", [10, 200, 5000])
def test_youden_threshold_matches_sklearn(n):
    from sklearn.metrics import roc_curve
    from pangnn_b200.train import youden_threshold
    g = torch.Generator().manual_seed(n)
    y = (torch.rand(n, generator=g) < 0.3).float()
    p = (torch.sigmoid(torch.randn(n, generator=g) + 2 * y) * 50).round() / 50      # many ties
    fpr, tpr, th = roc_curve(y.numpy(), p.numpy())
    assert youden_threshold(p, y) == pytest.approx(float(th[np.argmax(tpr - fpr)]))
    assert youden_threshold(p, torch.zeros(n)) is None          # one class only: keep the old threshold


def test_inert_reference_flags_are_reported(caplog):
    from pangnn_b200 import setup
    try:
        with caplog.at_level(logging.WARNING, logger="pangnn"):
            a = setup.parse(["--mixed_precision", "bf16", "--to_pickle", "x.pkl", "--cpus", "3"])
        assert a.cpus == 3
        assert set(setup.inert_flags(a)) == {"mixed_precision", "to_pickle"}
        assert sum("has no effect" in r.message for r in caplog.records) == 2
    finally:
        setup.reset()
    assert setup.inert_flags(setup.args) == []
