"""selective_scan_fn fwd+bwd at the API boundary, swept over the MedMamba-T stage shapes,
batch sizes and I/O dtypes; achieved GB/s = algorithmic bytes (s = bytes per I/O element) / kernel time.

Kernel times come from CUDA events recorded immediately around the C-ABI launches (the hook bench.py uses); L2 is flushed
with a 512 MB write between iterations, the median of `iters` is reported.
    python tools/microbench_sscan.py [iters] > gpurun_out/sscan_microbench.md"""
import json
import statistics
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


def peak_gbs():
    try:
        with open("MEASURED_PEAKS.json") as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6544.7


def run(batch, L, d_inner, dtype, a_init, iters, flush):
    dim, N, G = 4 * d_inner, 16, 4
    dev = "cuda"
    torch.manual_seed(0)
    u = torch.randn(batch, dim, L, device=dev).to(dtype).requires_grad_()
    delta = (0.5 * torch.rand(batch, dim, L, device=dev)).to(dtype).requires_grad_()
    if a_init == "test":
        A = (-0.5 * torch.rand(dim, N, device=dev)).requires_grad_()
    else:   # the model's init, A[d, n] = -(n + 1)   (MedMamba.py:351-363)
        A = (-torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(dim, 1)).requires_grad_()
    Bm = torch.randn(batch, G, N, L, device=dev).to(dtype).requires_grad_()
    Cm = torch.randn(batch, G, N, L, device=dev).to(dtype).requires_grad_()
    Dp = torch.randn(dim, device=dev, requires_grad=True)
    bias = (0.5 * torch.rand(dim, device=dev)).requires_grad_()
    g = torch.randn(batch, dim, L, device=dev).to(dtype)
    s = u.element_size()
    E, Ebc = batch * dim * L, batch * G * N * L
    bf = s * (3 * E + 2 * Ebc) + 4 * (dim * N + 2 * dim)
    bb = s * (5 * E + 2 * Ebc) + 4 * 2 * Ebc + 4 * (2 * dim * N + 4 * dim)
    tf, tb = [], []
    for it in range(iters + 2):
        flush.fill_(it)
        hook.rec.clear()
        out = selective_scan_fn(u, delta, A, Bm, Cm, Dp, delta_bias=bias, delta_softplus=True)
        out.backward(g)
        torch.cuda.synchronize()
        if it >= 2:
            t = {k: e0.elapsed_time(e1) for k, e0, e1 in hook.rec}
            tf.append(t["fwd"]); tb.append(t["bwd"])
        u.grad = delta.grad = A.grad = Bm.grad = Cm.grad = Dp.grad = bias.grad = None
    mf, mb = statistics.median(tf), statistics.median(tb)
    return mf, mb, bf / mf / 1e6, bb / mb / 1e6


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    hook = Hook()
    ssi.set_profiler(hook)
    flush = torch.empty(128 << 20, dtype=torch.float32, device="cuda")
    peak = peak_gbs()
    print(f"# selective_scan_fn microbench (config 5): N=16, G=4, dim=4*d_inner, delta_softplus, D, delta_bias; peak {peak:.1f} GB/s\n")
    print("| L | d_inner | B | I/O | A | fwd ms | fwd GB/s | frac | bwd ms | bwd GB/s | frac |")
    print("|---:|---:|---:|---|---|---:|---:|---:|---:|---:|---:|")
    shapes = [(3136, 96), (784, 192), (196, 384), (49, 768), (3136, 192), (784, 384), (196, 768), (49, 1536)]
    for L, d_inner in shapes:
        for batch in (1, 8, 64, 256):
            for dtype, a_init in ((torch.float32, "test"), (torch.float32, "model"), (torch.bfloat16, "test")):
                if a_init == "model" and batch != 64:
                    continue
                if batch * 4 * d_inner * L * 4 > (5 << 30):     # keep one operand under 5 GB
                    continue
                mf, mb, gf, gb = run(batch, L, d_inner, dtype, a_init, iters, flush)
                print(f"| {L} | {d_inner} | {batch} | {'f32' if dtype == torch.float32 else 'bf16'} | {a_init} | {mf:.3f} | {gf:.0f} | "
                      f"{gf / peak:.3f} | {mb:.3f} | {gb:.0f} | {gb / peak:.3f} |", flush=True)
