import sys, torch
sys.path.insert(0, ".")
from medical_image_classification_b200.selective_scan_interface import selective_scan_fn
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
dev = "cuda"
b, N, G = 2, 16, 4
u = torch.randn(b, dim, L, device=dev, requires_grad=True)
delta = (0.5 * torch.rand(b, dim, L, device=dev)).requires_grad_()
A = (-0.5 * torch.rand(dim, N, device=dev)).requires_grad_()
Bm = torch.randn(b, G, N, L, device=dev, requires_grad=True)
Cm = torch.randn(b, G, N, L, device=dev, requires_grad=True)
Dp = torch.randn(dim, device=dev, requires_grad=True)
bias = (0.5 * torch.rand(dim, device=dev)).requires_grad_()
out = selective_scan_fn(u, delta, A, Bm, Cm, Dp, delta_bias=bias, delta_softplus=True)
torch.cuda.synchronize(); print("fwd ok", float(out.abs().sum()))
out.backward(torch.randn_like(out))
torch.cuda.synchronize(); print("bwd ok", float(u.grad.abs().sum()))
