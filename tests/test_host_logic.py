"""CPU tests of the host-side logic of the fused SS2D core (no GPU, no library calls): the split-K weight gradient through
strided bmm views, the one-pass backward of the chunk(2, -1) splits, the folded dt projection identity and the layout
algebra of the (rows, B*L) buffers (medical_image_classification_b200/cross.py::SS2DCoreFn, ss2d.py)."""
import torch

from medical_image_classification_b200.cross import _weight_grad_splitk
from medical_image_classification_b200.ss2d import SS2D, SplitHalvesFn


def test_weight_grad_splitk_equals_plain_bmm():
    torch.manual_seed(0)
    for (M2, D, BL) in [(256, 96, 64 * 784), (70, 24, 4096), (448, 192, 8 * 196), (40, 8, 49 * 3)]:
        g = torch.randn(2, M2, BL, dtype=torch.float64)
        x = torch.randn(2, D, BL, dtype=torch.float64)
        want = torch.bmm(g, x.transpose(1, 2))
        got = _weight_grad_splitk(g, x)
        assert got.shape == want.shape
        assert float((got - want).abs().max() / want.abs().max()) < 1e-12


def test_split_halves_backward_is_the_cat_of_the_two_gradients():
    torch.manual_seed(1)
    t = torch.randn(2, 3, 5, 12, requires_grad=True)
    a, b = SplitHalvesFn.apply(t)
    assert torch.equal(a, t[..., :6]) and torch.equal(b, t[..., 6:])
    ga, gb = torch.randn_like(a), torch.randn_like(b)
    (a * ga).sum().add((b * gb).sum()).backward()
    got = t.grad.clone()
    t.grad = None
    c, d = t.chunk(2, dim=-1)
    (c * ga).sum().add((d * gb).sum()).backward()
    assert torch.equal(got, t.grad)


def test_folded_dt_projection_is_the_reference_composition():
    """delta = dt_projs_weight @ (x_proj_weight[:R] @ x) (MedMamba.py:397-400) == (dt_projs_weight @ x_proj_weight[:R]) @ x, and the
    stacked W_all rows reproduce x_dbl's B / C rows; internal direction order = reference order with the middle two swapped."""
    torch.manual_seed(2)
    m = SS2D(d_model=8, d_state=16).double()
    K, D, N, R = 4, m.d_inner, m.d_state, m.dt_rank
    Wx, Wdt, bias, As, Ds = m._dir_params()
    W_all = torch.cat([Wx[:, R:], torch.matmul(Wdt, Wx[:, :R])], dim=1)
    perm = [0, 2, 1, 3]                                         # cross.DIR_PERM
    xs = torch.randn(3, K, D, 11, dtype=torch.float64)          # per-direction inputs, reference order
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, m.x_proj_weight.double())
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts, m.dt_projs_weight.double())
    big = torch.einsum("bkdl,kmd->bkml", xs[:, perm], W_all.double())
    assert torch.allclose(big[:, :, :N], Bs[:, perm], rtol=0, atol=1e-12)
    assert torch.allclose(big[:, :, N:2 * N], Cs[:, perm], rtol=0, atol=1e-12)
    assert torch.allclose(big[:, :, 2 * N:], dts[:, perm], rtol=0, atol=2e-6)      # W_dt @ W_x,dt is formed in fp32
    assert torch.allclose(As.double().view(K, D, N), (-torch.exp(m.A_logs.double())).view(K, D, N)[perm], rtol=1e-6, atol=0)
    assert torch.equal(bias.double().view(K, D), m.dt_projs_bias.double()[perm])
    assert torch.equal(Ds.double().view(K, D), m.Ds.double().view(K, D)[perm])


def test_rows_by_batch_layout_views():
    """big (2, 2M, B*L) viewed as (B, 4, M, L) and x2 (2, D, B, L) viewed as (B, 2, D, L): the strides the scan kernels are given."""
    B, D, L, N = 3, 5, 8, 2
    M = 2 * N + D
    big = torch.arange(2 * 2 * M * B * L, dtype=torch.float32).view(2, 2 * M, B * L)
    big4 = big.view(4, M, B, L).permute(2, 0, 1, 3)
    assert big4.shape == (B, 4, M, L) and big4.stride() == (L, M * B * L, B * L, 1)
    for b in range(B):
        for k in range(4):
            assert torch.equal(big4[b, k], big[k // 2, (k % 2) * M:(k % 2 + 1) * M, b * L:(b + 1) * L])
    x2 = torch.arange(2 * D * B * L, dtype=torch.float32).view(2, D, B, L)
    u4 = x2.permute(2, 0, 1, 3)
    assert u4.stride() == (L, D * B * L, B * L, 1)


def test_dense_layout_helpers():
    """train_step._dense / ssd_combined._empty_like_layout: any dimension order of a dense tensor is recognised (channels_last
    parameters, the models' sequence-contiguous views), gaps and overlaps are not."""
    from medical_image_classification_b200.ssd_combined import _empty_like_layout
    from medical_image_classification_b200.train_step import _dense
    a = torch.empty(4, 6, 5, 3)
    assert _dense(a) and _dense(a.permute(0, 2, 3, 1)) and _dense(a.contiguous(memory_format=torch.channels_last))
    assert not _dense(a[:, :3]) and not _dense(a.expand(2, 4, 6, 5, 3)[0].expand(4, 6, 5, 3)[:, ::2])
    assert _dense(torch.empty(1, 7, 1, 3).permute(2, 0, 3, 1))          # size-1 dimensions carry any stride
    view = torch.empty(2, 8, 10).permute(0, 2, 1).unflatten(2, (2, 4))  # (b, l, h, p) view of channel-major storage
    g = _empty_like_layout(view)
    assert g.shape == view.shape and g.stride() == view.stride() and g.dtype == torch.float32
    sliced = torch.empty(2, 10, 2, 8)[..., :4]                          # not dense: falls back to a contiguous tensor
    g2 = _empty_like_layout(sliced)
    assert g2.shape == sliced.shape and g2.is_contiguous()


def test_chunk_size_mapping_and_wide_state_slices_are_host_side():
    """mamba_chunk_scan_combined maps any chunk_size to a tiled length and selective_scan_fn slices wide state spaces before any
    kernel call: both raise the no-CPU-path error (not a shape error) on CPU tensors."""
    import pytest
    from medical_image_classification_b200.selective_scan_interface import MAX_DSTATE, selective_scan_fn
    u = torch.randn(1, 4, 8)
    with pytest.raises(RuntimeError, match="dstate"):
        selective_scan_fn(u, u, torch.randn(4, MAX_DSTATE + 1), torch.randn(1, MAX_DSTATE + 1, 8), torch.randn(1, MAX_DSTATE + 1, 8))
    with pytest.raises(RuntimeError):   # 40 states: sliced, then the first slice hits require_cuda
        selective_scan_fn(u, u, torch.randn(4, 40), torch.randn(1, 40, 8), torch.randn(1, 40, 8))
