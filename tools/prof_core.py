"""Driver for ncu / timing of one SS2D block (MedMamba-T stage shape, bf16 autocast) -- the scan kernels then see the
layouts the model gives them (B, C, delta as strided views of the fused projection output, SS2DCoreFn).
    python tools/prof_core.py [stage 0..3] [batch] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200 import selective_scan_interface as ssi
from medical_image_classification_b200.ss2d import SS2D


class Hook:
    def __init__(self): self.rec = []
    def begin(self):
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    def end(self, e0, kind, u, delta, Bm, algo_len=None):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); self.rec.append((kind, e0, e1))


stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H, C = [(56, 48), (28, 96), (14, 192), (7, 384)][stage]
torch.manual_seed(0)
m = SS2D(d_model=C, d_state=16).cuda()
x = torch.randn(B, H, H, C, device="cuda", requires_grad=True)
hook = Hook()
ssi.set_profiler(hook)
for it in range(iters):
    hook.rec.clear()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    y.float().square().mean().backward()
    torch.cuda.synchronize()
    t = {k: e0.elapsed_time(e1) for k, e0, e1 in hook.rec}
    print(f"stage {stage} B={B} SS2D block: sscan fwd {t['fwd']:.3f} ms  bwd {t['bwd']:.3f} ms", flush=True)
