"""SS2D cross-scan / selective scan / cross-merge as two autograd Functions over libb200ssm.

Reference semantics (MedMamba.py:386-424, 476-477): four orderings of every channel plane --
k=0 row-major, k=1 column-major, k=2/3 their time reversals -- are scanned independently and the
four outputs are un-permuted and summed into (B, H, W, D).

B200 layout: directions are kept in the order  hw, hw-reversed, wh, wh-reversed  (`DIR_PERM` maps
this internal index to the reference's k), because then
  * the two reversed directions need no data at all (the kernel scans backwards, rev_mask=0b1010),
  * directions (0,1) and (2,3) read the same image (u_group_div=2) and receive the same upstream
    gradient (dout_group_div=2),
so cross-scan is one pack kernel (x -> [x, x^T]) and cross-merge is one kernel.
"""
from __future__ import annotations

import torch

from . import _lib
from .selective_scan_interface import _f32c, launch_bwd, launch_fwd

DIR_PERM = (0, 2, 1, 3)   # internal direction -> reference direction k (self-inverse)
REV_MASK = 0b1010         # internal directions 1 and 3 run backwards in time


def _plane_op(name, src, dst, batch, D, H, W):
    lib = _lib.load()
    with torch.cuda.device(src.device):
        rc = getattr(lib, name)(src.data_ptr(), dst.data_ptr(), batch, D, H, W, _lib.dtype_code(src.dtype),
                                _lib.stream_ptr(src.device))
    _lib.check(rc, name)


class CrossScanPackFn(torch.autograd.Function):
    """x (B, D, H, W) -> x2 (B, 2, D, L): x2[:,0] row-major image, x2[:,1] column-major image."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        x = x.contiguous()
        B, D, H, W = x.shape
        x2 = torch.empty((B, 2, D, H * W), dtype=x.dtype, device=x.device)
        _plane_op("b200_cross_scan_pack", x, x2, B, D, H, W)
        ctx.hw = (H, W)
        return x2

    @staticmethod
    def backward(ctx, dx2):
        H, W = ctx.hw
        dx2 = dx2.contiguous()
        B, _, D, _ = dx2.shape
        dx = torch.empty((B, D, H, W), dtype=dx2.dtype, device=dx2.device)
        _plane_op("b200_cross_scan_pack_bwd", dx2, dx, B, D, H, W)
        return dx


class CrossMergeFn(torch.autograd.Function):
    """ys (B, 4, D, L) at memory positions (internal direction order) -> y (B, L, D)."""

    @staticmethod
    def forward(ctx, ys, H, W):
        _lib.require_cuda(ys)
        ys = ys.contiguous()
        B, K, D, L = ys.shape
        assert K == 4 and L == H * W
        y = torch.empty((B, L, D), dtype=ys.dtype, device=ys.device)
        _plane_op("b200_cross_merge", ys, y, B, D, H, W)
        ctx.hw = (H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        H, W = ctx.hw
        dy = dy.contiguous()
        B, L, D = dy.shape
        d2 = torch.empty((B, 2, D, L), dtype=dy.dtype, device=dy.device)
        _plane_op("b200_cross_merge_bwd", dy, d2, B, D, H, W)
        dys = d2.unsqueeze(2).expand(B, 2, 2, D, L).reshape(B, 4, D, L)
        return dys, None, None


class ScanMergeFn(torch.autograd.Function):
    """Four-direction selective scan + cross-merge in one autograd node.

    x2 (B, 2, D, L) from CrossScanPackFn; delta (B, 4*D, L); A (4*D, N); Bs, Cs (B, 4, N, L);
    Ds, delta_bias (4*D) -- all in the internal direction order.  Returns y (B, L, D) =
    sum of the four directions at their image positions (== (B, H, W, D) flattened).
    """

    @staticmethod
    def forward(ctx, x2, delta, A, Bs, Cs, Ds, delta_bias, H, W):
        _lib.require_cuda(x2, delta, A, Bs, Cs, Ds, delta_bias)
        B, two, D, L = x2.shape
        assert two == 2 and L == H * W and delta.shape == (B, 4 * D, L)
        x2 = x2.contiguous()
        u = x2.view(B, 2 * D, L)
        delta = delta if delta.stride(-1) == 1 else delta.contiguous()
        Bs = Bs if Bs.stride(-1) == 1 else Bs.contiguous()
        Cs = Cs if Cs.stride(-1) == 1 else Cs.contiguous()
        if not (u.dtype == delta.dtype == Bs.dtype == Cs.dtype):
            raise RuntimeError("ScanMergeFn: x2, delta, Bs, Cs must share a dtype")
        A32, D32, b32 = _f32c(A), _f32c(Ds), _f32c(delta_bias)
        need_grad = any(ctx.needs_input_grad)
        ys, _, ckpt = launch_fwd(u, delta, A32, Bs, Cs, D32, None, b32, True, REV_MASK, 2, want_ckpt=need_grad)
        y = torch.empty((B, L, D), dtype=ys.dtype, device=ys.device)
        _plane_op("b200_cross_merge", ys, y, B, D, H, W)
        if need_grad:
            ctx.save_for_backward(u, delta, A32, Bs, Cs, D32, b32, ckpt)
        ctx.hw = (H, W)
        ctx.dtypes = (A.dtype, Ds.dtype, delta_bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        u, delta, A32, Bs, Cs, D32, b32, ckpt = ctx.saved_tensors
        H, W = ctx.hw
        B, KD, L = delta.shape
        D = KD // 4
        dy = dy.contiguous().to(u.dtype)
        d2 = torch.empty((B, 2 * D, L), dtype=dy.dtype, device=dy.device)
        _plane_op("b200_cross_merge_bwd", dy, d2, B, D, H, W)
        du, ddelta, dA, dB, dC, dD, dbias, _ = launch_bwd(u, delta, A32, Bs, Cs, D32, None, b32, True, ckpt, d2,
                                                          REV_MASK, 2, 2, True, True)
        dx2 = du.view(B, 2, 2, D, L).sum(2)
        return (dx2, ddelta, dA.to(ctx.dtypes[0]), dB.to(Bs.dtype), dC.to(Cs.dtype), dD.to(ctx.dtypes[1]),
                dbias.to(ctx.dtypes[2]), None, None)


class CrossScan4Fn(torch.autograd.Function):
    """SSD twin of the cross-scan: x (B, C, H, W) fp32 (a channel slice of a wider NCHW tensor is read in place) ->
    x4 (B, 4, C, L) in the reference's direction order (hw, wh, hw reversed, wh reversed; SSD/MedSSD.py:332-336)."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        lib = _lib.load()
        x = x.float()
        B, C, H, W = x.shape
        if not (x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W):
            x = x.contiguous()
        x4 = torch.empty((B, 4, C, H * W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.b200_cross_scan4(x.data_ptr(), x.stride(0), x4.data_ptr(), B, C, H, W, _lib.stream_ptr(x.device)), "b200_cross_scan4")
        ctx.hw = (H, W)
        return x4

    @staticmethod
    def backward(ctx, dx4):
        lib = _lib.load()
        H, W = ctx.hw
        dx4 = dx4.float().contiguous()
        B, _, C, _ = dx4.shape
        dx = torch.empty((B, C, H, W), dtype=torch.float32, device=dx4.device)
        with torch.cuda.device(dx4.device):
            _lib.check(lib.b200_cross_scan4_bwd(dx4.data_ptr(), dx.data_ptr(), dx.stride(0), B, C, H, W, _lib.stream_ptr(dx4.device)),
                       "b200_cross_scan4_bwd")
        return dx


class SsdMerge4Fn(torch.autograd.Function):
    """SSD twin of the cross-merge: y (B, L, 4, d) fp32 -> (B, L, d), the four directions un-permuted and summed
    (SSD/MedSSD.py:380-391)."""

    @staticmethod
    def forward(ctx, y, H, W):
        _lib.require_cuda(y)
        lib = _lib.load()
        y = y.float().contiguous()
        B, L, K, d = y.shape
        assert K == 4 and L == H * W
        out = torch.empty((B, L, d), dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(lib.b200_ssd_merge4(y.data_ptr(), out.data_ptr(), B, d, H, W, _lib.stream_ptr(y.device)), "b200_ssd_merge4")
        ctx.meta = (B, L, d, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        B, L, d, H, W = ctx.meta
        dout = dout.float().contiguous()
        dy = torch.empty((B, L, 4, d), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            _lib.check(lib.b200_ssd_merge4_bwd(dout.data_ptr(), dy.data_ptr(), B, d, H, W, _lib.stream_ptr(dout.device)), "b200_ssd_merge4_bwd")
        return dy, None, None


def cross_scan4(x):
    return CrossScan4Fn.apply(x)


def ssd_merge4(y, H, W):
    return SsdMerge4Fn.apply(y, H, W)


def cross_scan_pack(x):
    return CrossScanPackFn.apply(x)


def cross_merge(ys, H, W):
    return CrossMergeFn.apply(ys, H, W)


def scan_merge(x2, delta, A, Bs, Cs, Ds, delta_bias, H, W):
    return ScanMergeFn.apply(x2, delta, A, Bs, Cs, Ds, delta_bias, H, W)
