"""bwd kernel time at stage-0 shape under the SS2D launch variants (rev_mask / shared u / shared dout)."""
import sys, torch
sys.path.insert(0, ".")
from medical_image_classification_b200 import selective_scan_interface as ssi
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, D = 3136, 96
dim, N, G = 4 * D, 16, 4
dev = "cuda"
torch.manual_seed(0)
class Hook:
    def __init__(self): self.rec = []
    def begin(self):
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    def end(self, e0, kind, u, delta, Bm, algo_len=None):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); self.rec.append((kind, e0, e1))
hook = Hook(); ssi.set_profiler(hook)
delta = 0.5 * torch.rand(batch, dim, L, device=dev)
A = -0.5 * torch.rand(dim, N, device=dev)
xdbl = torch.randn(batch, G, 3 + 2 * N, L, device=dev)
Bm, Cm = xdbl[:, :, 3:3 + N], xdbl[:, :, 3 + N:]
Dp = torch.randn(dim, device=dev); bias = 0.5 * torch.rand(dim, device=dev)
for name, rev, ud, dd in [("plain", 0, 1, 1), ("rev", 0b1010, 1, 1), ("shared-u", 0, 2, 1), ("shared-dout", 0, 1, 2), ("ss2d", 0b1010, 2, 2)]:
    u = torch.randn(batch, dim // ud, L, device=dev)
    g = torch.randn(batch, dim // dd, L, device=dev)
    for it in range(3):
        hook.rec.clear()
        out, _, ck = ssi.launch_fwd(u, delta, A, Bm, Cm, Dp, None, bias, True, rev, ud, want_ckpt=True)
        ssi.launch_bwd(u, delta, A, Bm, Cm, Dp, None, bias, True, ck, g, rev, ud, dd)
        torch.cuda.synchronize()
        t = {k: e0.elapsed_time(e1) for k, e0, e1 in hook.rec}
    print(f"{name:12s} fwd {t['fwd']:.3f} ms  bwd {t['bwd']:.3f} ms", flush=True)
