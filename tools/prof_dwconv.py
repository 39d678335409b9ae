"""Driver for ncu / timing of the depthwise conv3x3 + SiLU kernels at a MedMamba-T stage shape.
    python tools/prof_dwconv.py [stage 0..3] [batch] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200.ss2d import DwConvSiluFn

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H, D = [(56, 96), (28, 192), (14, 384), (7, 768)][stage]
torch.manual_seed(0)
xz = torch.randn(B, H, H, 2 * D, device="cuda", dtype=torch.bfloat16).requires_grad_()
w = torch.randn(D, 1, 3, 3, device="cuda").requires_grad_()
b = torch.randn(D, device="cuda").requires_grad_()
g = torch.randn(B, D, H, H, device="cuda")
for it in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = DwConvSiluFn.apply(xz[..., :D], w, b)
    e[1].record()
    out.backward(g)
    e[2].record()
    torch.cuda.synchronize()
    print(f"stage {stage} B={B}: fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd (incl. autograd glue) {e[1].elapsed_time(e[2]):.3f} ms", flush=True)
