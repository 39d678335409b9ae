"""One SSD (mamba_chunk_scan_combined) forward + backward at a MedSSD stage shape, for ncu / timing.
    python tools/prof_ssd.py [stage 0..3] [batch] [precision 0 = 3xTF32 | 1 = TF32] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200 import ssd_combined
from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L, nh = [(3136, 2), (784, 4), (196, 8), (49, 16)][stage]
Q, P, dev = 256, 64, "cuda"
H, N = 4 * nh, 512
torch.manual_seed(0)
x = torch.randn(batch, H * P, L, device=dev).permute(0, 2, 1).unflatten(2, (H, P)).requires_grad_()
Bm = torch.randn(batch, N, L, device=dev).permute(0, 2, 1).unflatten(2, (1, N)).requires_grad_()
Cm = torch.randn(batch, N, L, device=dev).permute(0, 2, 1).unflatten(2, (1, N)).requires_grad_()
dt = (0.5 * torch.rand(batch, H, L, device=dev)).permute(0, 2, 1).requires_grad_()
A = (-0.5 - torch.rand(H, device=dev)).requires_grad_()
D = torch.randn(H, device=dev).requires_grad_()
bias = (0.3 * torch.rand(H, device=dev)).requires_grad_()
g = torch.randn(batch, L, H, P, device=dev)
ssd_combined.set_precision(prec)
for it in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = mamba_chunk_scan_combined(x, dt, A, Bm, Cm, Q, D=D, dt_bias=bias, dt_softplus=True)
    e[1].record()
    torch.autograd.grad(y, (x, dt, A, Bm, Cm, D, bias), g)
    e[2].record()
    torch.cuda.synchronize()
print(f"stage {stage} B={batch} L={L} H={H} N={N} prec={prec}: fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms (host-launched, incl. autograd glue)")
