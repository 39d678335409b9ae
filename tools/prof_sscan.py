"""Small driver for ncu / timing: selective_scan_fn fwd+bwd at one MedMamba-T stage shape.
    python tools/prof_sscan.py [stage 0..3] [batch] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200.selective_scan_interface import selective_scan_fn

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, D = [(3136, 96), (784, 192), (196, 384), (49, 768)][stage]
dim, N, G = 4 * D, 16, 4
dev = "cuda"
torch.manual_seed(0)
u = torch.randn(batch, dim, L, device=dev, requires_grad=True)
delta = (0.5 * torch.rand(batch, dim, L, device=dev)).requires_grad_()
A = (-0.5 * torch.rand(dim, N, device=dev)).requires_grad_()
Bm = torch.randn(batch, G, N, L, device=dev, requires_grad=True)
Cm = torch.randn(batch, G, N, L, device=dev, requires_grad=True)
Dp = torch.randn(dim, device=dev, requires_grad=True)
bias = (0.5 * torch.rand(dim, device=dev)).requires_grad_()
g = torch.randn(batch, dim, L, device=dev)
E = batch * dim * L
Ebc = batch * G * N * L
bf = 4 * (3 * E + 2 * Ebc)
bb = 4 * (5 * E + 2 * Ebc) + 8 * Ebc
for it in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = selective_scan_fn(u, delta, A, Bm, Cm, Dp, delta_bias=bias, delta_softplus=True)
    e[1].record()
    out.backward(g)
    e[2].record()
    torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(f"stage {stage} B={batch} it {it}: fwd {tf:.3f} ms ({bf / tf / 1e6:.0f} GB/s)  bwd(+autograd glue) {tb:.3f} ms ({bb / tb / 1e6:.0f} GB/s)")
