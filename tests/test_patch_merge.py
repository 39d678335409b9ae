"""GPU: the one-pass PatchMerging2D gather (csrc/glue.cu::patch_merge_kernel, reference MedMamba.py:186-204) is bit-exact against the
reference's slice + cat formulation, forward and backward, fp32 and bf16, even and odd H / W (the reference drops the last row / column)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (3, 56, 56, 96), (2, 7, 9, 16), (1, 14, 14, 384)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patch_merge_gather_bit_exact(shape, dtype):
    from medical_image_classification_b200.models import PatchMergeGatherFn
    B, H, W, C = shape
    torch.manual_seed(0)
    x = torch.randn(shape, device="cuda").to(dtype).requires_grad_()
    h2, w2 = H // 2, W // 2
    ref = torch.cat([x[:, i::2, j::2, :][:, :h2, :w2, :] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))], dim=-1)
    g = torch.randn_like(ref)
    (dref,) = torch.autograd.grad(ref, x, g)
    out = PatchMergeGatherFn.apply(x)
    (dx,) = torch.autograd.grad(out, x, g)
    assert out.shape == ref.shape and torch.equal(out, ref)
    assert torch.equal(dx, dref)
