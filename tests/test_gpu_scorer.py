"""GPU parity: fused edge scorer (fwd, bwd, fused BCE) vs the oracle's torch-CPU autograd."""
import numpy as np
import pytest
import torch

from oracle.model import bce_with_logits
from tests.helpers import GRAD_TOL, assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
D = 64


def make(N, E, skip, seed):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(N, D, generator=g)
    ei = torch.randint(0, N, (2, E), generator=g)
    fin = 2 * D + (1 if skip else 0)
    P = dict(w1=torch.randn(D, fin, generator=g) / fin ** 0.5, b1=torch.randn(D, generator=g) * 0.1,
             w2=torch.randn(D, D, generator=g) / 8, b2=torch.randn(D, generator=g) * 0.1,
             w3=torch.randn(1, D, generator=g) / 8, b3=torch.randn(1, generator=g) * 0.1)
    sk = (torch.rand(E, generator=g) * 80 + 1) if skip else None
    # keep the test inputs away from the ReLU kinks: where a pre-activation is within 1e-4 of zero an
    # fp32 reassociation legitimately flips the derivative, which is not what this test measures
    parts = (h[ei[0]], h[ei[1]]) + ((sk.unsqueeze(1),) if skip else ())
    a1 = torch.cat(parts, 1).double() @ P["w1"].double().t() + P["b1"].double()
    a2 = torch.relu(a1) @ P["w2"].double().t() + P["b2"].double()
    ok = (a1.abs().min(dim=1).values > 1e-4) & (a2.abs().min(dim=1).values > 1e-4)
    if E > 1000:
        ei = ei[:, ok].contiguous()
        sk = sk[ok].contiguous() if skip else None
    y = (torch.rand(ei.size(1), generator=g) < 0.3).float()
    return h, ei, P, sk, y


def oracle_logits(h, ei, P, sk):
    parts = (h[ei[0]], h[ei[1]]) + ((sk.unsqueeze(1),) if sk is not None else ())
    a1 = torch.cat(parts, 1) @ P["w1"].t() + P["b1"]
    a2 = torch.relu(a1) @ P["w2"].t() + P["b2"]
    return (torch.relu(a2) @ P["w3"].t() + P["b3"]).squeeze(-1)


NAMES = ["w1", "b1", "w2", "b2", "w3", "b3"]


@pytest.mark.parametrize("N,E", [(50, 1), (50, 127), (300, 128), (300, 129), (2000, 40_000)])
@pytest.mark.parametrize("skip", [False, True])
def test_scorer_forward_backward(N, E, skip):
    from pangnn_b200 import ops
    h, ei, P, sk, y = make(N, E, skip, E)
    E = ei.size(1)
    ho = h.clone().requires_grad_(True)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    zo = oracle_logits(ho, ei, Po, sk)
    dz = torch.randn(E, generator=torch.Generator().manual_seed(1))
    zo.backward(dz)

    hd = h.clone().to(DEV).requires_grad_(True)
    Pd = {k: v.clone().to(DEV).requires_grad_(True) for k, v in P.items()}
    gs = ops.GraphStruct(ei.to(DEV), N)
    zd = ops.EdgeScoreFn.apply(hd, *[Pd[k] for k in NAMES], gs, sk.to(DEV) if skip else None)
    zd.backward(dz.to(DEV))
    assert_close("logits", zd.detach().cpu().numpy(), zo.detach().numpy(), TOL)
    assert_close("dh", hd.grad.cpu().numpy(), ho.grad.numpy(), GRAD_TOL)
    for k in NAMES:
        assert_close(f"d{k}", Pd[k].grad.cpu().numpy(), Po[k].grad.numpy(), GRAD_TOL)


@pytest.mark.parametrize("E,pw", [(1, 1.0), (200, 4.8), (33_333, 0.37)])
@pytest.mark.parametrize("skip", [False, True])
def test_scorer_fused_bce(E, pw, skip):
    from pangnn_b200 import ops
    N = 700
    h, ei, P, sk, y = make(N, E, skip, E + 5)
    E = ei.size(1)
    ho = h.clone().requires_grad_(True)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    zo = oracle_logits(ho, ei, Po, sk)
    lo = bce_with_logits(zo, y, pw)
    (lo * 1.7).backward()                                    # non-trivial upstream gradient

    hd = h.clone().to(DEV).requires_grad_(True)
    Pd = {k: v.clone().to(DEV).requires_grad_(True) for k, v in P.items()}
    gs = ops.GraphStruct(ei.to(DEV), N)
    ld, zd = ops.EdgeScoreBCEFn.apply(hd, *[Pd[k] for k in NAMES], gs, sk.to(DEV) if skip else None,
                                      y.to(DEV), pw)
    (ld * 1.7).backward()
    assert abs(ld.item() - lo.item()) <= TOL * abs(lo.item())
    assert_close("logits", zd.cpu().numpy(), zo.detach().numpy(), TOL)
    assert_close("dh", hd.grad.cpu().numpy(), ho.grad.numpy(), GRAD_TOL)
    for k in NAMES:
        assert_close(f"d{k}", Pd[k].grad.cpu().numpy(), Po[k].grad.numpy(), GRAD_TOL)


def test_scorer_extreme_logits_are_finite():
    """Saturated logits (|z| ~ 1e3) must give finite loss / gradients, as torch's formula does."""
    from pangnn_b200 import ops
    N, E = 100, 1000
    h, ei, P, sk, y = make(N, E, False, 9)
    P["w3"] = P["w3"] * 2000
    zo = oracle_logits(h, ei, P, None)
    lo = bce_with_logits(zo, y, 2.0)
    gs = ops.GraphStruct(ei.to(DEV), N)
    Pd = {k: v.to(DEV).requires_grad_(True) for k, v in P.items()}
    ld, zd = ops.EdgeScoreBCEFn.apply(h.to(DEV).requires_grad_(True), *[Pd[k] for k in NAMES], gs,
                                      None, y.to(DEV), 2.0)
    ld.backward()
    assert torch.isfinite(ld) and abs(ld.item() - lo.item()) <= 1e-5 * abs(lo.item())
    assert all(torch.isfinite(Pd[k].grad).all() for k in NAMES)


@pytest.mark.parametrize("mode", [0, 1])
def test_pair_score(mode):
    from pangnn_b200 import ops
    N, E = 400, 5000
    h, ei, *_ = make(N, E, False, 2)
    gs = ops.GraphStruct(ei.to(DEV), N)
    got = ops.edge_pair_score(h.to(DEV), gs, mode).cpu()
    ref = (torch.nn.functional.cosine_similarity(h[ei[0]], h[ei[1]], dim=1) if mode == 0
           else (h[ei[0]] * h[ei[1]]).sum(1))
    assert rel_err(got.numpy(), ref.numpy()) < TOL
