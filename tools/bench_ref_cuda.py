"""Time the REFERENCE's own CUDA selective scan (baseline/_ref/selective_scan_cuda.so, built unmodified for sm_100a by
tools/build_ref_cuda.py) next to libb200ssm at the MedMamba-T stage shapes: the "kernel to beat" of SURVEY.md 2.2.
    python tools/bench_ref_cuda.py [batch] [iters]
Prints a markdown table: per stage, reference fwd / bwd ms vs this repo's fwd / bwd ms (CUDA events around the calls, L2 flushed
between iterations, median)."""
import importlib.util
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medical_image_classification_b200 import selective_scan_interface as ssi  # noqa: E402

so = os.path.join(ROOT, "baseline", "_ref", "selective_scan_cuda.so")
if not os.path.exists(so):
    print("baseline/_ref/selective_scan_cuda.so is missing (run tools/build_ref_cuda.py in the build container)")
    sys.exit(0)
spec = importlib.util.spec_from_file_location("selective_scan_cuda", so)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
flush = torch.empty(128 << 20, dtype=torch.float32, device="cuda")


def timed(fn):
    ts = []
    for it in range(iters + 2):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), out


print(f"| stage (B={batch}) | KD | L | reference fwd ms | this repo fwd ms | x | reference bwd ms | this repo bwd ms | x | max rel diff out |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
tot = [0.0, 0.0, 0.0, 0.0]
calls = [2, 2, 4, 2]
for stage, (L, D) in enumerate([(3136, 96), (784, 192), (196, 384), (49, 768)]):
    dim, N, G = 4 * D, 16, 4
    torch.manual_seed(0)
    u = torch.randn(batch, dim, L, device="cuda")
    delta = 0.5 * torch.rand(batch, dim, L, device="cuda")
    A = -0.5 * torch.rand(dim, N, device="cuda")
    Bm = torch.randn(batch, G, N, L, device="cuda")
    Cm = torch.randn(batch, G, N, L, device="cuda")
    Dp = torch.randn(dim, device="cuda")
    bias = 0.5 * torch.rand(dim, device="cuda")
    g = torch.randn(batch, dim, L, device="cuda")
    t_rf, res = timed(lambda: ref.fwd(u, delta, A, Bm, Cm, Dp, None, bias, True))
    out_ref, x = res[0], res[1]
    t_rb, _ = timed(lambda: ref.bwd(u, delta, A, Bm, Cm, Dp, None, bias, g, x, None, None, True, False))
    t_f, res2 = timed(lambda: ssi.launch_fwd(u, delta, A, Bm, Cm, Dp, None, bias, True, want_ckpt=True))
    out, _, ckpt = res2
    t_b, _ = timed(lambda: ssi.launch_bwd(u, delta, A, Bm, Cm, Dp, None, bias, True, ckpt, g))
    diff = float((out - out_ref).abs().max() / out_ref.abs().max())
    print(f"| {stage} | {dim} | {L} | {t_rf:.3f} | {t_f:.3f} | {t_rf / t_f:.2f} | {t_rb:.3f} | {t_b:.3f} | {t_rb / t_b:.2f} | {diff:.1e} |", flush=True)
    for k, t in enumerate((t_rf, t_f, t_rb, t_b)):
        tot[k] += calls[stage] * t
print(f"| MedMamba-T step (10 calls) | | | {tot[0]:.2f} | {tot[1]:.2f} | {tot[0] / tot[1]:.2f} | {tot[2]:.2f} | {tot[3]:.2f} | {tot[2] / tot[3]:.2f} | |")
