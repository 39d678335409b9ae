"""Cross-checks for the SSD oracle (oracle/ssd_oracle.c).  The reference holds no vectors for this
call (PARITY UNPINNED, see the oracle header), so the oracle is checked against
 (1) an independent chunked ("state-space dual") numpy restatement of the published algorithm,
 (2) known answers (A = 0 -> prefix sums of dt*x (outer) B),
 (3) torch float64 autograd through a literal per-step recurrence, for every gradient."""
import numpy as np
import pytest
import torch

import oracle


def chunked_ssd_numpy(x, dt, A, B, C, chunk):
    """Chunked SSD in float64: Y = (C B^T o L) (dt x) + decay * C S_prev ;  S passes chunk to chunk."""
    x, dt, A, B, C = (np.asarray(a, np.float64) for a in (x, dt, A, B, C))
    b, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    hpg = H // G
    y = np.zeros_like(x)
    S = np.zeros((b, H, P, N))
    for c0 in range(0, L, chunk):
        c1 = min(L, c0 + chunk)
        for h in range(H):
            g = h // hpg
            dA = dt[:, c0:c1, h] * A[h]                       # (b, q)
            cs = np.cumsum(dA, axis=1)                         # inclusive
            Lmat = np.exp(cs[:, :, None] - cs[:, None, :])     # (b, q, s)
            Lmat = np.tril(np.ones(Lmat.shape[1:]))[None] * Lmat
            CB = np.einsum("bqn,bsn->bqs", C[:, c0:c1, g], B[:, c0:c1, g])
            xdt = x[:, c0:c1, h] * dt[:, c0:c1, h, None]       # (b, q, p)
            y[:, c0:c1, h] = np.einsum("bqs,bsp->bqp", CB * Lmat, xdt)
            y[:, c0:c1, h] += np.exp(cs)[:, :, None] * np.einsum("bqn,bpn->bqp", C[:, c0:c1, g], S[:, h])
            decay_to_end = np.exp(cs[:, -1:] - cs)             # (b, q)
            S[:, h] = np.exp(cs[:, -1])[:, None, None] * S[:, h] + np.einsum(
                "bq,bqp,bqn->bpn", decay_to_end, xdt, B[:, c0:c1, g])
    return y, S


def make(b=2, L=70, H=4, P=8, G=1, N=12, seed=0):
    r = np.random.RandomState(seed)
    x = r.randn(b, L, H, P).astype(np.float32)
    dt = (0.5 * r.rand(b, L, H)).astype(np.float32)
    A = (-0.5 - r.rand(H)).astype(np.float32)
    B = r.randn(b, L, G, N).astype(np.float32)
    C = r.randn(b, L, G, N).astype(np.float32)
    return x, dt, A, B, C


@pytest.mark.parametrize("chunk", [7, 16, 64, 256])
@pytest.mark.parametrize("G", [1, 2])
def test_sequential_equals_chunked(chunk, G):
    x, dt, A, B, C = make(G=G)
    out, fin = oracle.ssd_fwd(x, dt, A, B, C)
    y, S = chunked_ssd_numpy(x, dt, A, B, C, chunk)
    assert np.abs(out - y).max() / np.abs(y).max() < 1e-6
    assert np.abs(fin - S).max() / np.abs(S).max() < 1e-6


def test_known_answer_A_zero():
    x, dt, A, B, C = make(b=1, L=33, H=2, P=4, N=5)
    A[:] = 0
    out, fin = oracle.ssd_fwd(x, dt, A, B, C)
    S = np.cumsum(np.einsum("blhp,blgn->blhpn", x * dt[..., None], B.astype(np.float64)), axis=1)
    y = np.einsum("blhpn,bln->blhp", S, C[:, :, 0].astype(np.float64))
    assert np.abs(out - y).max() < 1e-5
    assert np.abs(fin - S[:, -1]).max() < 1e-5


def torch_recurrence(x, dt, A, B, C, D, z, dt_bias, softplus, init):
    b, L, H, P = x.shape
    G = B.shape[2]
    hpg = H // G
    dtv = dt + (dt_bias if dt_bias is not None else 0)
    if softplus:
        dtv = torch.where(dtv > 20, dtv, torch.log1p(torch.exp(dtv)))
    S = init if init is not None else x.new_zeros(b, H, P, B.shape[3])
    ys = []
    for t in range(L):
        a = torch.exp(dtv[:, t] * A)                                       # (b, h)
        Bt = B[:, t].repeat_interleave(hpg, dim=1)                          # (b, h, n)
        Ct = C[:, t].repeat_interleave(hpg, dim=1)
        S = a[..., None, None] * S + (dtv[:, t, :, None] * x[:, t])[..., None] * Bt[:, :, None, :]
        y = (S * Ct[:, :, None, :]).sum(-1)
        if D is not None:
            y = y + x[:, t] * (D if D.dim() == 2 else D[:, None])
        ys.append(y)
    y = torch.stack(ys, 1)
    if z is not None:
        y = y * z * torch.sigmoid(z)
    return y, S


@pytest.mark.parametrize("cfg", [dict(), dict(z=True), dict(hdim=True, G=2), dict(init=True, softplus=False)])
def test_backward_matches_torch_autograd(cfg):
    x, dt, A, B, C = make(b=2, L=41, H=4, P=6, G=cfg.get("G", 1), N=7, seed=3)
    r = np.random.RandomState(5)
    H, P = x.shape[2], x.shape[3]
    D = r.randn(H, P).astype(np.float32) if cfg.get("hdim") else r.randn(H).astype(np.float32)
    z = r.randn(*x.shape).astype(np.float32) if cfg.get("z") else None
    dt_bias = (0.3 * r.rand(H)).astype(np.float32)
    init = r.randn(2, H, P, B.shape[3]).astype(np.float32) if cfg.get("init") else None
    softplus = cfg.get("softplus", True)
    dout = r.randn(*x.shape).astype(np.float32)
    out, fin = oracle.ssd_fwd(x, dt, A, B, C, D=D, z=z, dt_bias=dt_bias, dt_softplus=softplus, initial_states=init)
    gr = oracle.ssd_bwd(x, dt, A, B, C, D=D, z=z, dt_bias=dt_bias, dt_softplus=softplus, initial_states=init, dout=dout)
    T = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tx, tdt, tA, tB, tC, tD, tz, tb = map(T, (x, dt, A, B, C, D, z, dt_bias))
    y, S = torch_recurrence(tx, tdt, tA, tB, tC, tD, tz, tb, softplus,
                            None if init is None else torch.tensor(init, dtype=torch.float64))
    assert np.abs(out - y.detach().numpy()).max() < 2e-5
    assert np.abs(fin - S.detach().numpy()).max() < 2e-5
    y.backward(torch.tensor(dout, dtype=torch.float64))
    ref = dict(dx=tx.grad, ddt=tdt.grad, dA=tA.grad, dB=tB.grad, dC=tC.grad, dD=tD.grad, ddt_bias=tb.grad,
               dz=None if tz is None else tz.grad)
    for k, v in ref.items():
        if v is None:
            continue
        v = v.numpy()
        assert np.abs(gr[k] - v).max() / max(np.abs(v).max(), 1e-30) < 1e-6, k
