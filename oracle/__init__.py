"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's hot path, used only as the checker:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product package ``medical_image_classification_b200`` never
imports it and fails loudly when its CUDA library is missing.

* ``sscan_fwd`` / ``sscan_bwd``  -- Mamba-1 selective scan, restating ``selective_scan_ref``
  (reference CrossMamba/FusionMamba/mamba_ssm/ops/selective_scan_interface.py:92-158); pinned
  against the reference itself through tests/golden/sscan_*.npz (oracle/make_golden.py).
* ``ssd_fwd`` / ``ssd_bwd``      -- Mamba-2 SSD ``mamba_chunk_scan_combined`` from the published
  recurrence (call contract SSD/MedSSD.py:344-375).  PARITY UNPINNED: mamba_ssm==2.2.2 is not
  vendored in the reference and the reference has no test of that call.
* ``ss2d_core_ref`` / ``cross_scan_ref`` / ``cross_merge_ref`` -- numpy restatement of the SS2D
  cross-scan / cross-merge index maps (MedMamba.py:393-395, 420-424, 476-477).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile oracle/*.c into oracle/_build/liboracle.so (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def threads() -> int:
    return int(lib().sscan_oracle_threads_f64())


def _c(a, shape=None):
    """float32 C-contiguous numpy array (accepts torch CPU tensors)."""
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().float().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        assert tuple(a.shape) == tuple(shape), (a.shape, shape)
    return a


def _p(a):
    return a.ctypes.data_as(_f32p) if a is not None else _f32p()


# --------------------------------------------------------------------------------------
# Mamba-1 selective scan
# --------------------------------------------------------------------------------------
def _sscan_shapes(u, A, B):
    batch, dim, L = u.shape
    N = A.shape[1]
    if B.ndim == 3:  # (batch, N, L) == one group  (interface.py:37-42)
        B = B[:, None]
    G = B.shape[1]
    assert dim % G == 0
    return batch, dim, L, N, G


def sscan_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
              precision="f64"):
    """Returns (out, last_state) as float32 arrays; internal arithmetic in ``precision``."""
    u, delta, A, B, C = _c(u), _c(delta), _c(A), _c(B), _c(C)
    batch, dim, L, N, G = _sscan_shapes(u, A, B)
    B = B.reshape(batch, G, N, L)
    C = C.reshape(batch, G, N, L)
    D, z, delta_bias = _c(D), _c(z), _c(delta_bias)
    out = np.empty_like(u)
    last = np.empty((batch, dim, N), np.float32)
    fn = getattr(lib(), f"sscan_oracle_fwd_{precision}")
    fn.restype = None
    fn(batch, dim, L, N, G, _p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
       int(bool(delta_softplus)), _p(out), _p(last))
    return out, last


def sscan_bwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, dout=None,
              precision="f64"):
    """Returns dict(du, ddelta, dA, dB, dC, dD, ddelta_bias, dz)."""
    u, delta, A, B, C, dout = _c(u), _c(delta), _c(A), _c(B), _c(C), _c(dout)
    batch, dim, L, N, G = _sscan_shapes(u, A, B)
    bshape = B.shape
    B = B.reshape(batch, G, N, L)
    C = C.reshape(batch, G, N, L)
    D, z, delta_bias = _c(D), _c(z), _c(delta_bias)
    du, ddelta = np.empty_like(u), np.empty_like(u)
    dA = np.empty_like(A)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD = np.empty(dim, np.float32)
    dbias = np.empty(dim, np.float32)
    dz = np.empty_like(u) if z is not None else None
    fn = getattr(lib(), f"sscan_oracle_bwd_{precision}")
    fn.restype = None
    fn(batch, dim, L, N, G, _p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
       int(bool(delta_softplus)), _p(dout), _p(du), _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD),
       _p(dbias), _p(dz))
    return dict(du=du, ddelta=ddelta, dA=dA, dB=dB.reshape(bshape), dC=dC.reshape(bshape),
                dD=dD if D is not None else None,
                ddelta_bias=dbias if delta_bias is not None else None, dz=dz)


# --------------------------------------------------------------------------------------
# Mamba-2 SSD
# --------------------------------------------------------------------------------------
def ssd_fwd(x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False,
            dt_limit=(0.0, float("inf")), initial_states=None):
    """x (b,l,h,p), dt (b,l,h), A (h), B/C (b,l,g,n) -> (out (b,l,h,p), final_states (b,h,p,n))."""
    x, dt, A, B, C = _c(x), _c(dt), _c(A), _c(B), _c(C)
    batch, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    D, z, dt_bias, initial_states = _c(D), _c(z), _c(dt_bias), _c(initial_states)
    d_hdim = int(D is not None and D.ndim == 2)
    out = np.empty_like(x)
    fin = np.empty((batch, H, P, N), np.float32)
    fn = lib().ssd_oracle_fwd
    fn.restype = None
    fn(batch, L, H, P, G, N, _p(x), _p(dt), _p(A), _p(B), _p(C), _p(D), d_hdim, _p(z), _p(dt_bias),
       int(bool(dt_softplus)), ctypes.c_double(dt_limit[0]), ctypes.c_double(dt_limit[1]),
       _p(initial_states), _p(out), _p(fin))
    return out, fin


def ssd_bwd(x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False,
            dt_limit=(0.0, float("inf")), initial_states=None, dout=None):
    x, dt, A, B, C, dout = _c(x), _c(dt), _c(A), _c(B), _c(C), _c(dout)
    batch, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    D, z, dt_bias, initial_states = _c(D), _c(z), _c(dt_bias), _c(initial_states)
    d_hdim = int(D is not None and D.ndim == 2)
    dx, ddt = np.empty_like(x), np.empty_like(dt)
    dA = np.empty_like(A)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD = np.empty_like(D) if D is not None else None
    dbias = np.empty(H, np.float32)
    dz = np.empty_like(x) if z is not None else None
    fn = lib().ssd_oracle_bwd
    fn.restype = None
    fn(batch, L, H, P, G, N, _p(x), _p(dt), _p(A), _p(B), _p(C), _p(D), d_hdim, _p(z), _p(dt_bias),
       int(bool(dt_softplus)), ctypes.c_double(dt_limit[0]), ctypes.c_double(dt_limit[1]),
       _p(initial_states), _p(dout), _p(dx), _p(ddt), _p(dA), _p(dB), _p(dC), _p(dD), _p(dbias), _p(dz))
    return dict(dx=dx, ddt=ddt, dA=dA, dB=dB, dC=dC, dD=dD,
                ddt_bias=dbias if dt_bias is not None else None, dz=dz)


# --------------------------------------------------------------------------------------
# SS2D cross-scan / cross-merge index maps (numpy)
# --------------------------------------------------------------------------------------
def cross_scan_ref(x):
    """x (B, D, H, W) -> xs (B, 4, D, L).  MedMamba.py:393-395.
    k=0 row-major, k=1 column-major, k=2/3 = time reversal of k=0/1."""
    x = np.asarray(x)
    Bn, Dn, H, W = x.shape
    L = H * W
    hw = x.reshape(Bn, Dn, L)
    wh = np.ascontiguousarray(x.transpose(0, 1, 3, 2)).reshape(Bn, Dn, L)
    fwd = np.stack([hw, wh], axis=1)
    return np.concatenate([fwd, fwd[..., ::-1]], axis=1)


def cross_merge_ref(ys, H, W):
    """ys (B, 4, D, L) in scan order -> y (B, H, W, D).  MedMamba.py:420-424, 476-477."""
    ys = np.asarray(ys)
    Bn, K, Dn, L = ys.shape
    inv = ys[:, 2:4, :, ::-1]
    wh = ys[:, 1].reshape(Bn, Dn, W, H).transpose(0, 1, 3, 2).reshape(Bn, Dn, L)
    invwh = inv[:, 1].reshape(Bn, Dn, W, H).transpose(0, 1, 3, 2).reshape(Bn, Dn, L)
    y = ys[:, 0] + inv[:, 0] + wh + invwh
    return np.ascontiguousarray(y.transpose(0, 2, 1)).reshape(Bn, H, W, Dn)


# --------------------------------------------------------------------------------------
# EfficientVMamba atrous scan / merge index maps (numpy), step 2
# --------------------------------------------------------------------------------------
def atrous_scan_ref(x):
    """x (B, C, H, W) -> xs (B, 4, C, ceil(H/2) * ceil(W/2)).  CrossMamba/FusionMamba/models/cross.py:143-169 (EfficientScan.forward):
    sub-lattice k has row parity k & 1 and column parity k >> 1; k even is flattened row-major, k odd column-major; zero padding."""
    x = np.asarray(x)
    Bn, Cn, H, W = x.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    xp = np.zeros((Bn, Cn, 2 * H2, 2 * W2), x.dtype)
    xp[:, :, :H, :W] = x
    out = np.empty((Bn, 4, Cn, H2 * W2), x.dtype)
    for k in range(4):
        sub = xp[:, :, (k & 1)::2, (k >> 1)::2]                       # (B, C, H2, W2)
        out[:, k] = (sub if k % 2 == 0 else sub.transpose(0, 1, 3, 2)).reshape(Bn, Cn, -1)
    return out


def atrous_merge_ref(ys, H, W):
    """ys (B, 4, C, ceil(H/2) * ceil(W/2)) -> y (B, C, H * W).  models/cross.py:33-56 (EfficientMerge.forward)."""
    ys = np.asarray(ys)
    Bn, K, Cn, L2 = ys.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    y = np.zeros((Bn, Cn, 2 * H2, 2 * W2), ys.dtype)
    for k in range(4):
        sub = ys[:, k].reshape(Bn, Cn, H2, W2) if k % 2 == 0 else ys[:, k].reshape(Bn, Cn, W2, H2).transpose(0, 1, 3, 2)
        y[:, :, (k & 1)::2, (k >> 1)::2] = sub
    return np.ascontiguousarray(y[:, :, :H, :W]).reshape(Bn, Cn, H * W)

