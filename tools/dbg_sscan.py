"""Debug driver: selective_scan_dirs_fn fwd(+bwd) vs the oracle on one small case.
    python tools/dbg_sscan.py L rev_mask dim [bwd=1]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from medical_image_classification_b200.selective_scan_interface import selective_scan_dirs_fn
L, rev, dim = int(sys.argv[1]), int(sys.argv[2], 0), int(sys.argv[3])
do_bwd = len(sys.argv) <= 4 or sys.argv[4] == "1"
r = np.random.RandomState(3)
batch, N, G = 2, 16, 4
f = lambda *s: r.randn(*s).astype(np.float32)
u, delta = f(batch, dim, L), (0.5 * r.rand(batch, dim, L)).astype(np.float32)
A, Bm, Cm, D = (-0.5 * r.rand(dim, N)).astype(np.float32), f(batch, G, N, L), f(batch, G, N, L), f(dim)
bias, g = (0.5 * r.rand(dim)).astype(np.float32), f(batch, dim, L)
T = lambda a: torch.tensor(a, device="cuda", requires_grad=True)
tu, td, tA, tB, tC, tD, tb = map(T, (u, delta, A, Bm, Cm, D, bias))
out = selective_scan_dirs_fn(tu, td, tA, tB, tC, tD, tb, True, rev_mask=rev)
torch.cuda.synchronize()
print("fwd ran", flush=True)
if do_bwd:
    out.backward(torch.tensor(g, device="cuda"))
    torch.cuda.synchronize()
    print("bwd ran", flush=True)
flip = lambda a, k: np.ascontiguousarray(a[..., ::-1]) if k else a
rpg = dim // G
ref = np.empty_like(u)
grads = {k: np.zeros_like(v) for k, v in (("du", u), ("ddelta", delta), ("dB", Bm), ("dC", Cm), ("dA", A), ("dD", D), ("ddelta_bias", bias))}
for gi in range(G):
    rows = slice(gi * rpg, (gi + 1) * rpg)
    k = (rev >> gi) & 1
    args = (flip(u[:, rows], k), flip(delta[:, rows], k), A[rows], flip(Bm[:, gi:gi + 1], k), flip(Cm[:, gi:gi + 1], k))
    kw = dict(D=D[rows], delta_bias=bias[rows], delta_softplus=True, precision="f64")
    o, _ = oracle.sscan_fwd(*args, **kw)
    ref[:, rows] = flip(o, k)
    if do_bwd:
        gg = oracle.sscan_bwd(*args, dout=flip(g[:, rows], k), **kw)
        grads["du"][:, rows] = flip(gg["du"], k); grads["ddelta"][:, rows] = flip(gg["ddelta"], k)
        grads["dB"][:, gi:gi + 1] = flip(gg["dB"], k); grads["dC"][:, gi:gi + 1] = flip(gg["dC"], k)
        grads["dA"][rows] = gg["dA"]; grads["dD"][rows] = gg["dD"]; grads["ddelta_bias"][rows] = gg["ddelta_bias"]
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
print("out", rel(out.detach().cpu().numpy(), ref))
if do_bwd:
    for name, t in (("du", tu), ("ddelta", td), ("dB", tB), ("dC", tC), ("dA", tA), ("dD", tD), ("ddelta_bias", tb)):
        print(name, rel(t.grad.cpu().numpy(), grads[name]))
