"""Atrous scan / merge (csrc/cross.cu) against the reference's own tensor-op formulation
(CrossMamba/FusionMamba/models/cross.py:139-190 EfficientScan, :34-92 EfficientMerge), restated here with the same slices.
Pure data movement: bit-exact, forward and backward, odd sizes included (zero padding); in fp32 also against the numpy oracle
(oracle.atrous_scan_ref / atrous_merge_ref, pinned on CPU against the reference's own classes in tests/test_oracle_atrous.py)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_scan(x, s=2):                     # models/cross.py:143-169
    B, C, H, W = x.shape
    if W % s:
        x = F.pad(x, (0, s - W % s, 0, 0))
    if H % s:
        x = F.pad(x, (0, 0, 0, s - H % s))
    xs = torch.stack([x[:, :, ::s, ::s].reshape(B, C, -1),
                      x.transpose(2, 3)[:, :, ::s, 1::s].reshape(B, C, -1),
                      x[:, :, ::s, 1::s].reshape(B, C, -1),
                      x.transpose(2, 3)[:, :, 1::s, 1::s].reshape(B, C, -1)], dim=1)
    return xs


def ref_merge(ys, H0, W0, s=2):           # models/cross.py:33-56
    B, K, C, L = ys.shape
    H, W = math.ceil(H0 / s), math.ceil(W0 / s)
    y = ys.new_zeros((B, C, H * s, W * s))
    y[:, :, ::s, ::s] = ys[:, 0].reshape(B, C, H, W)
    y[:, :, 1::s, ::s] = ys[:, 1].reshape(B, C, W, H).transpose(2, 3)
    y[:, :, ::s, 1::s] = ys[:, 2].reshape(B, C, H, W)
    y[:, :, 1::s, 1::s] = ys[:, 3].reshape(B, C, W, H).transpose(2, 3)
    return y[:, :, :H0, :W0].reshape(B, C, -1)


@pytest.mark.parametrize("shape", [(2, 3, 8, 8), (1, 5, 7, 9), (2, 4, 14, 14), (1, 2, 56, 56), (3, 2, 1, 6), (1, 1, 64, 63)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_atrous_scan_and_merge_bit_exact(shape, dtype):
    from medical_image_classification_b200.atrous import EfficientMerge, EfficientScan
    B, C, H, W = shape
    torch.manual_seed(H * W)
    x = torch.randn(B, C, H, W, device="cuda", dtype=dtype).requires_grad_()
    xs = EfficientScan.apply(x, 2)
    want = ref_scan(x)
    assert xs.shape == want.shape and torch.equal(xs, want)
    g = torch.randn_like(xs)
    xs.backward(g)
    got = x.grad.clone(); x.grad = None
    want.backward(g)
    assert torch.equal(got, x.grad)
    ys = torch.randn(B, 4, C, xs.shape[-1], device="cuda", dtype=dtype).requires_grad_()
    y = EfficientMerge.apply(ys, H, W, 2)
    want = ref_merge(ys, H, W)
    assert y.shape == want.shape and torch.equal(y, want)
    g = torch.randn_like(y)
    y.backward(g)
    got = ys.grad.clone(); ys.grad = None
    want.backward(g)
    assert torch.equal(got, ys.grad)
    if dtype == torch.float32:
        import oracle
        assert np.array_equal(xs.detach().cpu().numpy(), oracle.atrous_scan_ref(x.detach().cpu().numpy()))
        assert np.array_equal(y.detach().cpu().numpy(), oracle.atrous_merge_ref(ys.detach().cpu().numpy(), H, W))
    # the two maps are inverse to each other on the image
    assert torch.equal(EfficientMerge.apply(EfficientScan.apply(x.detach(), 2), H, W, 2).view(B, C, H, W), x.detach())


def test_cross_selective_scan_new_matches_tensor_op_formulation():
    """The whole atrous SS2D core (models/cross.py:193-262) against the same data flow built from the reference's slices around
    the (oracle-checked) operator: forward and all parameter / input gradients agree norm-wise to 1e-5 or better."""
    from medical_image_classification_b200.atrous import cross_selective_scan_new
    from medical_image_classification_b200.selective_scan_interface import selective_scan_fn
    torch.manual_seed(3)
    B, D, H, W, N, R, K = 2, 8, 9, 10, 16, 2, 4
    dev = "cuda"
    x = torch.randn(B, D, H, W, device=dev).requires_grad_()
    xw = (0.3 * torch.randn(K, R + 2 * N, D, device=dev)).requires_grad_()
    dw = (0.5 * torch.randn(K, D, R, device=dev)).requires_grad_()
    db = (0.1 * torch.randn(K, D, device=dev)).requires_grad_()
    A_logs = torch.log(torch.arange(1, N + 1, device=dev, dtype=torch.float32)).repeat(K * D, 1).requires_grad_()
    Ds = torch.ones(K * D, device=dev).requires_grad_()
    y = cross_selective_scan_new(x, xw, None, dw, db, A_logs, Ds, out_norm=None)
    # slice formulation + the operator on the same tensors
    xs = ref_scan(x)
    L = xs.shape[-1]
    x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs, xw)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("b k r l, k d r -> b k d l", dts, dw)
    ys = selective_scan_fn(xs.reshape(B, -1, L), dts.reshape(B, -1, L), -torch.exp(A_logs), Bs, Cs, Ds, None, db.reshape(-1), True)
    want = ref_merge(ys.view(B, K, -1, L), H, W).transpose(1, 2).reshape(B, H, W, -1)
    assert float((y - want).abs().max() / want.abs().max()) < 1e-6
    params = (x, xw, dw, db, A_logs, Ds)
    g = torch.randn_like(y)
    got = torch.autograd.grad(y, params, g, retain_graph=True)
    ref = torch.autograd.grad(want, params, g)
    for a, b_ in zip(got, ref):
        assert float((a - b_).abs().max() / b_.abs().max().clamp_min(1e-30)) < 1e-5


import glob  # noqa: E402
import os  # noqa: E402

ATROUS_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "atrous_*.npz")))


@pytest.mark.parametrize("path", ATROUS_GOLDEN, ids=[os.path.basename(p)[7:-4] for p in ATROUS_GOLDEN])
def test_atrous_kernels_match_reference_vectors(path):
    """Vectors produced by the reference's own EfficientScan / EfficientMerge classes (oracle/make_golden.py --atrous-only)."""
    from medical_image_classification_b200.atrous import EfficientMerge, EfficientScan
    g = np.load(path)
    B, C, H, W = g["x"].shape
    x = torch.tensor(g["x"]).cuda().requires_grad_()
    xs = EfficientScan.apply(x, 2)
    assert np.array_equal(xs.detach().cpu().numpy(), g["xs"])
    xs.backward(torch.tensor(g["gxs"]).cuda())
    assert np.array_equal(x.grad.cpu().numpy(), g["dx"])
    ys = torch.tensor(g["ys"]).cuda().requires_grad_()
    y = EfficientMerge.apply(ys, H, W, 2)
    assert np.array_equal(y.detach().cpu().numpy(), g["y"])
    y.backward(torch.tensor(g["gy"]).cuda())
    assert np.array_equal(ys.grad.cpu().numpy(), g["dys"])

