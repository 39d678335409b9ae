"""SSD microbench at the MedSSD stage shapes (x (B,L,H',64), B/C (B,L,1,N'), chunk 256) + MedSSD step.
    python tools/bench_ssd.py [batch] [d_state] [--model]
Prints one JSON line per (stage, precision): ms fwd / bwd (CUDA events), TFLOP/s with the SURVEY.md 8(d) FLOP formula."""
import json
import sys

import torch

sys.path.insert(0, ".")
from medical_image_classification_b200 import ssd_combined  # noqa: E402
from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
d_state = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 128
Q, P = 256, 64
dev = "cuda"


def flops(b, L, H, N, G=1):
    f = bw = 0
    for c0 in range(0, L, Q):
        q = min(Q, L - c0)
        f += 2 * q * q * N * G + H * (2 * q * q * P + 4 * q * N * P)
        bw += 6 * q * q * N * G + H * (4 * q * q * P + 10 * q * N * P)
    return b * f, b * bw


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stage(L, nheads):
    H, N = 4 * nheads, 4 * d_state
    # channel-major storage, (b, l, .) views with L stride 1 -- as the reference passes them
    x = torch.randn(batch, H * P, L, device=dev).permute(0, 2, 1).unflatten(2, (H, P)).requires_grad_()
    Bm = torch.randn(batch, N, L, device=dev).permute(0, 2, 1).unflatten(2, (1, N)).requires_grad_()
    Cm = torch.randn(batch, N, L, device=dev).permute(0, 2, 1).unflatten(2, (1, N)).requires_grad_()
    dt = (0.5 * torch.rand(batch, H, L, device=dev)).permute(0, 2, 1).requires_grad_()
    A = (-0.5 - torch.rand(H, device=dev)).requires_grad_()
    D = torch.randn(H, device=dev).requires_grad_()
    bias = (0.3 * torch.rand(H, device=dev)).requires_grad_()
    g = torch.randn(batch, L, H, P, device=dev)
    ff, fb = flops(batch, L, H, N)
    for prec in (0, 1):
        ssd_combined.set_precision(prec)
        fwd = lambda: mamba_chunk_scan_combined(x, dt, A, Bm, Cm, Q, D=D, dt_bias=bias, dt_softplus=True)
        with torch.no_grad():
            t_f = timeit(fwd)

        def both():
            y = fwd()
            torch.autograd.grad(y, (x, dt, A, Bm, Cm, D, bias), g)
        t_fb = timeit(both)
        t_b = t_fb - t_f
        print(json.dumps({"L": L, "H": H, "N": N, "batch": batch, "precision": ["3xTF32", "TF32"][prec],
                          "fwd_ms": round(t_f, 3), "bwd_ms": round(t_b, 3),
                          "fwd_tflops": round(ff / t_f / 1e9, 1), "bwd_tflops": round(fb / t_b / 1e9, 1)}), flush=True)
    ssd_combined.set_precision(0)


if "--model" not in sys.argv:
    for L, nh in ((3136, 2), (784, 4), (196, 8), (49, 16)):
        stage(L, nh * (1 if d_state else 1))
else:
    from medical_image_classification_b200.models import medssd
    torch.backends.cudnn.benchmark = True
    net = medssd(num_classes=6, d_state=d_state).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    x = torch.randn(batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 6, (batch,), device=dev)
    for prec in (0, 1):
        ssd_combined.set_precision(prec)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(net(x).float(), y)
            loss.backward()
            opt.step()
        t = timeit(step, 5)
        print(json.dumps({"model": "MedSSD", "d_state": d_state, "batch": batch, "precision": ["3xTF32", "TF32"][prec],
                          "ms_per_step": round(t, 2), "images_per_s": round(batch / t * 1e3, 1)}), flush=True)
