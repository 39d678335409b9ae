"""Block tail (csrc/glue.cu): channel_shuffle(cat(left, x), 2) + input against the reference's
own tensor ops (MedMamba.py:486-499, 533-538).  Pure data movement plus one fp32 add: bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref(left, x, inp):
    left = left.permute(0, 2, 3, 1)
    out = torch.cat((left, x.to(left.dtype)), dim=-1)
    B, H, W, C = out.shape
    out = out.view(B, H, W, 2, C // 2).transpose(3, 4).reshape(B, H, W, C)
    return out + inp


@pytest.mark.parametrize("shape", [(2, 48, 56, 56), (3, 96, 28, 28), (2, 384, 7, 7), (1, 20, 5, 9), (2, 33, 3, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("stream_dtype", [torch.float32, torch.bfloat16])
def test_shuffle_cat_add_bit_exact(shape, dtype, channels_last, stream_dtype):
    from medical_image_classification_b200.models import ShuffleCatAddFn
    B, c, H, W = shape
    dev = "cuda"
    torch.manual_seed(c)
    left = torch.randn(B, c, H, W, device=dev, dtype=dtype)
    if channels_last:
        left = left.contiguous(memory_format=torch.channels_last)
    left.requires_grad_()
    x = torch.randn(B, H, W, c, device=dev, dtype=dtype).requires_grad_()
    inp = torch.randn(B, H, W, 2 * c, device=dev, dtype=stream_dtype).requires_grad_()   # bf16: stages 1-3 under autocast
    out = ShuffleCatAddFn.apply(left, x, inp)
    g = torch.randn_like(out)
    out.backward(g)
    got = [t.grad.clone() for t in (left, x, inp)]
    for t in (left, x, inp):
        t.grad = None
    o2 = ref(left, x, inp)
    o2.backward(g)
    assert out.dtype == o2.dtype and torch.equal(out, o2)   # torch's promoted dtype: bf16 only when all three are bf16
    for a, t in zip(got, (left, x, inp)):
        assert torch.equal(a, t.grad)
