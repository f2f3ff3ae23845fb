"""GPU parity of the tcgen05 3xTF32 node linear transform (pangnn_node_linear) against an fp64
reference: the kernel must deliver fp32-grade results (the model-level bar is 1e-5 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# fp32 GEMM-grade: relative to the output scale.  Plain TF32 would sit at ~5e-4.
TOL = 2e-6


def _ref(x, w, bias, act, w_is_kn):
    y = x.double() @ (w.double() if w_is_kn else w.double().t())
    if bias is not None:
        y = y + bias.double()
    if act:
        y = torch.nn.functional.elu(y)
    return y


def _err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("n,k", [(64, 64), (128, 64), (64, 128), (128, 128)])
@pytest.mark.parametrize("m", [1, 127, 128, 129, 1000, 70001])
@pytest.mark.parametrize("w_is_kn", [False, True])
def test_node_linear_matches_fp64(n, k, m, w_is_kn):
    from pangnn_b200 import ops
    torch.manual_seed(n * 7 + k + m)
    x = torch.randn(m, k, device=DEV) * torch.rand(m, 1, device=DEV) * 3
    w = torch.randn((k, n) if w_is_kn else (n, k), device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    for b, act in ((None, ops.ACT_NONE), (bias, ops.ACT_ELU)):
        l0 = ops.LAUNCHES["count"]
        y = ops.node_linear(x, w, b, act, w_is_kn=w_is_kn)
        assert ops.LAUNCHES["count"] == l0 + 1, "the tensor-core kernel must be the path taken"
        assert _err(y, _ref(x, w, b, act, w_is_kn)) < TOL


def test_node_linear_strided_views_and_determinism():
    from pangnn_b200 import ops
    torch.manual_seed(3)
    big = torch.randn(5000, 256, device=DEV)
    x = big[:, 64:192]                                   # row stride 256, 16 B aligned
    w = torch.randn(64, 128, device=DEV) / 11
    out = torch.zeros(5000, 128, device=DEV)
    ops.node_linear(x, w, out=out[:, 64:])               # strided output
    assert _err(out[:, 64:], _ref(x, w, None, 0, False)) < TOL
    assert float(out[:, :64].abs().max()) == 0.0
    a = ops.node_linear(x, w)
    b = ops.node_linear(x, w)
    assert torch.equal(a, b)


def test_node_linear_extreme_values():
    """Large dynamic range and exact cancellation: the hi/lo split must not lose the small terms."""
    from pangnn_b200 import ops
    torch.manual_seed(5)
    m = 4096
    x = torch.randn(m, 64, device=DEV) * torch.logspace(-6, 6, m, device=DEV).unsqueeze(1)
    w = torch.randn(64, 64, device=DEV)
    y = ops.node_linear(x, w)
    ref = _ref(x, w, None, 0, False)
    row_scale = ref.abs().max(dim=1, keepdim=True).values.clamp_min(1e-30)
    assert float(((y.double() - ref).abs() / row_scale).max()) < 5e-6
    # identity weight reproduces the input to fp32 rounding of hi + lo
    eye = torch.eye(64, device=DEV)
    z = ops.node_linear(x, eye)
    assert float(((z - x).abs() / x.abs().clamp_min(1e-30)).max()) < 1e-6


def test_node_linear_falls_back_to_library_gemm_for_other_shapes():
    from pangnn_b200 import ops
    x = torch.randn(100, 1, device=DEV)
    w = torch.randn(64, 1, device=DEV)
    y = ops.node_linear(x, w)
    assert torch.allclose(y, x @ w.t())
