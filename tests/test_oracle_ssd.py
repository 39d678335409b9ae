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


# ------------------------------------------------------------------------------------------------------------------
# Cross-check against an INDEPENDENT implementation: transformers' pure-PyTorch Mamba-2 mixer
# (transformers/models/mamba2/modeling_mamba2.py::Mamba2Mixer.torch_forward, the "ssd naive implementation" section).
# mamba_ssm 2.2.2 itself is not installable here; this is the closest obtainable statement of what
# mamba_chunk_scan_combined computes that the oracle's author did not write.
# ------------------------------------------------------------------------------------------------------------------
def _hf_mixer(L, heads, headdim, groups, dstate, chunk, seed):
    m2 = pytest.importorskip("transformers.models.mamba2.modeling_mamba2")
    cfg_cls = pytest.importorskip("transformers").Mamba2Config
    cfg = cfg_cls(num_heads=heads, head_dim=headdim, hidden_size=heads * headdim // 2, state_size=dstate, expand=2, n_groups=groups,
                  chunk_size=chunk, conv_kernel=4, num_hidden_layers=1, vocab_size=16, use_conv_bias=True, use_bias=False,
                  time_step_limit=(0.0, float("inf")))
    torch.manual_seed(seed)
    mixer = m2.Mamba2Mixer(cfg, layer_idx=0).double().eval()
    with torch.no_grad():
        mixer.A_log.copy_(torch.log(torch.rand(heads, dtype=torch.float64) * 3 + 0.2))
        mixer.D.copy_(torch.randn(heads, dtype=torch.float64))
        mixer.dt_bias.copy_(torch.randn(heads, dtype=torch.float64) * 0.5)
    return mixer


@pytest.mark.parametrize("L,heads,headdim,groups,dstate,chunk", [(70, 4, 8, 1, 12, 32), (64, 4, 16, 2, 8, 16), (33, 2, 8, 1, 16, 256)])
def test_oracle_matches_transformers_torch_forward(L, heads, headdim, groups, dstate, chunk):
    mixer = _hf_mixer(L, heads, headdim, groups, dstate, chunk, seed=L)
    torch.manual_seed(1)
    hidden = torch.randn(2, L, heads * headdim // 2, dtype=torch.float64)
    captured = {}
    handle = mixer.norm.register_forward_pre_hook(lambda mod, args: captured.setdefault("y", args[0].detach().clone()))
    with torch.no_grad():
        mixer.torch_forward(hidden)
    handle.remove()
    # the operator's inputs, recomputed with the mixer's own layers (steps 1-2 of torch_forward)
    with torch.no_grad():
        proj = mixer.in_proj(hidden)
        inter, gn = mixer.intermediate_size, mixer.n_groups * mixer.ssm_state_size
        d_mlp = (proj.shape[-1] - 2 * inter - 2 * gn - mixer.num_heads) // 2
        _, _, _gate, xBC, dt = proj.split([d_mlp, d_mlp, inter, mixer.conv_dim, mixer.num_heads], dim=-1)
        xBC = mixer.act(mixer.conv1d(xBC.transpose(1, 2))[..., :L].transpose(1, 2))
        x, Bm, Cm = torch.split(xBC, [inter, gn, gn], dim=-1)
    x = x.reshape(2, L, heads, headdim)
    Bm = Bm.reshape(2, L, groups, dstate)
    Cm = Cm.reshape(2, L, groups, dstate)
    A = -torch.exp(mixer.A_log.detach())
    out, _ = oracle.ssd_fwd(x.numpy(), dt.numpy(), A.numpy(), Bm.numpy(), Cm.numpy(), D=mixer.D.detach().numpy(),
                            dt_bias=mixer.dt_bias.detach().numpy(), dt_softplus=True)
    y_hf = captured["y"].reshape(2, L, heads, headdim).numpy()
    err = np.abs(out.astype(np.float64) - y_hf).max() / np.abs(y_hf).max()
    assert err < 2e-6, err      # the oracle returns fp32 outputs of an fp64 recurrence


def test_oracle_grads_match_transformers_autograd():
    """All seven gradients of the operator (dx, ddt, dA, dB, dC, dD, ddt_bias) against torch autograd THROUGH transformers'
    torch_forward: the loss only sees the SSD output (captured at the input of the mixer's gated norm), the gradients are read
    at the operator's inputs (output of the conv activation; the dt columns of in_proj's output; A_log, D, dt_bias)."""
    L, heads, headdim, groups, dstate, chunk = 50, 4, 8, 1, 12, 16
    mixer = _hf_mixer(L, heads, headdim, groups, dstate, chunk, seed=7)
    torch.manual_seed(2)
    hidden = torch.randn(2, L, heads * headdim // 2, dtype=torch.float64)
    cap = {}
    hooks = [mixer.norm.register_forward_pre_hook(lambda mod, args: cap.setdefault("y", args[0])),
             mixer.act.register_forward_hook(lambda mod, args, out: cap.setdefault("xBC", out)),
             mixer.in_proj.register_forward_hook(lambda mod, args, out: cap.setdefault("proj", out))]
    mixer.torch_forward(hidden)
    for h in hooks:
        h.remove()
    y, xBC, proj = cap["y"], cap["xBC"], cap["proj"]
    dout = torch.randn_like(y)
    g_xBC, g_proj, g_Alog, g_D, g_bias = torch.autograd.grad((y * dout).sum(), [xBC, proj, mixer.A_log, mixer.D, mixer.dt_bias])
    inter, gn = mixer.intermediate_size, groups * dstate
    x, Bm, Cm = (t.detach() for t in torch.split(xBC[:, :L], [inter, gn, gn], dim=-1))
    gx, gB, gC = torch.split(g_xBC[:, :L], [inter, gn, gn], dim=-1)
    dt = proj.detach()[..., -heads:]
    A = -torch.exp(mixer.A_log.detach())
    g = oracle.ssd_bwd(x.reshape(2, L, heads, headdim).numpy(), dt.numpy(), A.numpy(), Bm.reshape(2, L, groups, dstate).numpy(),
                       Cm.reshape(2, L, groups, dstate).numpy(), D=mixer.D.detach().numpy(), dt_bias=mixer.dt_bias.detach().numpy(),
                       dt_softplus=True, dout=dout.reshape(2, L, heads, headdim).numpy())

    def rel(a, b):
        b = b.detach().numpy()
        return np.abs(np.asarray(a, np.float64).reshape(b.shape) - b).max() / np.abs(b).max()

    assert rel(g["dx"], gx.reshape(2, L, heads, headdim)) < 2e-6
    assert rel(g["dB"], gB.reshape(2, L, groups, dstate)) < 2e-6
    assert rel(g["dC"], gC.reshape(2, L, groups, dstate)) < 2e-6
    assert rel(g["ddt"], g_proj[..., -heads:]) < 2e-6
    assert rel(g["dA"] * A.numpy(), g_Alog) < 2e-6          # A = -exp(A_log)  =>  dA_log = dA * A
    assert rel(g["dD"], g_D) < 2e-6
    assert rel(g["ddt_bias"], g_bias) < 2e-6
