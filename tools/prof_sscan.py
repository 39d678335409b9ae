"""Small driver for ncu / timing: selective_scan_fn fwd+bwd at the MedMamba-T stage shapes; kernel-only times from CUDA
events recorded immediately around the C-ABI launches (the hook bench.py uses).
    python tools/prof_sscan.py [stage 0..3 | all] [batch] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200 import selective_scan_interface as ssi
from medical_image_classification_b200.selective_scan_interface import selective_scan_fn


class Hook:
    def __init__(self): self.rec = []
    def begin(self):
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    def end(self, e0, kind, u, delta, Bm, algo_len=None):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); self.rec.append((kind, e0, e1))


stages = [0, 1, 2, 3] if (len(sys.argv) > 1 and sys.argv[1] == "all") else [int(sys.argv[1]) if len(sys.argv) > 1 else 0]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
hook = Hook()
ssi.set_profiler(hook)
tot = {"fwd": 0.0, "bwd": 0.0}
calls = [2, 2, 4, 2]
for stage in stages:
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
        hook.rec.clear()
        out = selective_scan_fn(u, delta, A, Bm, Cm, Dp, delta_bias=bias, delta_softplus=(len(sys.argv) <= 4))
        out.backward(g)
        torch.cuda.synchronize()
        t = {k: e0.elapsed_time(e1) for k, e0, e1 in hook.rec}
    print(f"stage {stage} B={batch} L={L} dim={dim}: fwd {t['fwd']:.3f} ms ({bf / t['fwd'] / 1e6:.0f} GB/s)  "
          f"bwd {t['bwd']:.3f} ms ({bb / t['bwd'] / 1e6:.0f} GB/s)", flush=True)
    tot["fwd"] += calls[stage] * t["fwd"]; tot["bwd"] += calls[stage] * t["bwd"]
if len(stages) == 4:
    print(f"MedMamba-T step (10 calls): fwd {tot['fwd']:.2f} ms + bwd {tot['bwd']:.2f} ms = {tot['fwd'] + tot['bwd']:.2f} ms "
          f"-> {11.342 / (tot['fwd'] + tot['bwd']) * 1e3:.0f} GB/s of 6545")
