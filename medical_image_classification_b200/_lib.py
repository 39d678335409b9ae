"""ctypes binding of libb200ssm.so (include/b200_ssm.h).

The library is the product: there is NO CPU or PyTorch fallback.  If the shared object is missing
and cannot be built (nvcc absent), importing an op raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200ssm.so")

i32, u32, i64, f32 = C.c_int32, C.c_uint32, C.c_int64, C.c_float
vp = C.c_void_p


class SScanFwdParams(C.Structure):
    """b200_sscan_fwd_params (include/b200_ssm.h)."""
    _fields_ = [
        ("batch", i32), ("dim", i32), ("seqlen", i32), ("dstate", i32), ("n_groups", i32),
        ("io_dtype", i32), ("delta_softplus", i32), ("rev_mask", u32), ("u_group_div", i32), ("ckpt_every", i32),
        ("u_batch_stride", i64), ("u_group_stride", i64), ("u_row_stride", i64),
        ("delta_batch_stride", i64), ("delta_group_stride", i64), ("delta_row_stride", i64),
        ("B_batch_stride", i64), ("B_group_stride", i64), ("B_state_stride", i64),
        ("C_batch_stride", i64), ("C_group_stride", i64), ("C_state_stride", i64),
        ("z_batch_stride", i64), ("z_row_stride", i64),
        ("out_batch_stride", i64), ("out_row_stride", i64),
        ("u", vp), ("delta", vp), ("A", vp), ("B", vp), ("C", vp), ("D", vp), ("z", vp), ("delta_bias", vp),
        ("out", vp), ("last_state", vp), ("ckpt", vp),
    ]


class SScanBwdParams(C.Structure):
    """b200_sscan_bwd_params."""
    _fields_ = [
        ("f", SScanFwdParams),
        ("dout_batch_stride", i64), ("dout_group_stride", i64), ("dout_row_stride", i64), ("dout_group_div", i64),
        ("du_batch_stride", i64), ("du_row_stride", i64),
        ("ddelta_batch_stride", i64), ("ddelta_group_stride", i64), ("ddelta_row_stride", i64),
        ("dB_batch_stride", i64), ("dB_group_stride", i64), ("dB_state_stride", i64),
        ("dC_batch_stride", i64), ("dC_group_stride", i64), ("dC_state_stride", i64),
        ("dz_batch_stride", i64), ("dz_row_stride", i64),
        ("dout", vp), ("du", vp), ("ddelta", vp), ("dz", vp),
        ("dA", vp), ("dB", vp), ("dC", vp), ("dD", vp), ("ddelta_bias", vp),
    ]


class SsdFwdParams(C.Structure):
    """b200_ssd_fwd_params."""
    _fields_ = [
        ("batch", i32), ("seqlen", i32), ("nheads", i32), ("headdim", i32), ("n_groups", i32), ("dstate", i32),
        ("chunk_size", i32), ("io_dtype", i32), ("dt_softplus", i32), ("precision", i32),
        ("dt_min", f32), ("dt_max", f32),
        ("x_stride", i64 * 4), ("dt_stride", i64 * 3), ("B_stride", i64 * 4), ("C_stride", i64 * 4),
        ("out_stride", i64 * 4),
        ("x", vp), ("dt", vp), ("A", vp), ("B", vp), ("C", vp), ("D", vp), ("dt_bias", vp),
        ("initial_states", vp), ("out", vp), ("final_states", vp), ("workspace", vp),
    ]


class SsdBwdParams(C.Structure):
    """b200_ssd_bwd_params."""
    _fields_ = [
        ("f", SsdFwdParams),
        ("dout_stride", i64 * 4),
        ("dout", vp), ("dx", vp), ("ddt", vp), ("dB", vp), ("dC", vp), ("dA", vp), ("dD", vp),
        ("ddt_bias", vp), ("scratch", vp),
        ("dx_stride", i64 * 4), ("ddt_stride", i64 * 3), ("dB_stride", i64 * 4), ("dC_stride", i64 * 4),
    ]


_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}

# every symbol include/b200_ssm.h declares
EXPORTS = [
    "b200_sscan_ckpt_bytes", "b200_sscan_fwd", "b200_sscan_bwd", "b200_sscan_last_variant",
    "b200_cross_scan_pack", "b200_cross_scan_pack_bwd", "b200_cross_merge", "b200_cross_merge_bwd",
    "b200_cross_scan4", "b200_cross_scan4_bwd", "b200_ssd_merge4", "b200_ssd_merge4_bwd",
    "b200_ssd_workspace_bytes", "b200_ssd_bwd_scratch_bytes", "b200_ssd_fwd", "b200_ssd_bwd",
    "b200_rmsnorm_gated_fwd", "b200_rmsnorm_gated_bwd",
    "b200_cross_scan_pack_strided", "b200_cross_scan_unpack4", "b200_atrous_scan", "b200_atrous_merge",
    "b200_ln_gate_grid", "b200_ln_gate_fwd", "b200_ln_gate_bwd",
    "b200_dwconv_silu_fwd", "b200_dwconv_silu_bwd",
    "b200_shuffle_cat_add_fwd", "b200_shuffle_cat_add_bwd", "b200_patch_merge",
    "b200_last_error", "b200_version", "b200_kernel_launches", "b200_sizeof_params",
]

_lib = None


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise RuntimeError(f"libb200ssm: unsupported dtype {dt}") from None


def load() -> C.CDLL:
    """Load (building first if the .so is absent and nvcc is present) -- raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.can_build():
        _build.build()          # no-op when the digest stamp matches; rebuilds a stale .so after csrc edits (under a file lock)
    elif not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing and nvcc is not available to build it: there is no fallback path")
    lib = C.CDLL(LIB_PATH)
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_sscan_ckpt_bytes.restype = C.c_size_t
    lib.b200_sscan_ckpt_bytes.argtypes = [i32] * 6
    lib.b200_ssd_workspace_bytes.restype = C.c_size_t
    lib.b200_ssd_workspace_bytes.argtypes = [i32] * 7
    lib.b200_ssd_bwd_scratch_bytes.restype = C.c_size_t
    lib.b200_ssd_bwd_scratch_bytes.argtypes = [i32] * 7
    lib.b200_sizeof_params.restype = C.c_size_t
    lib.b200_sizeof_params.argtypes = [i32]
    lib.b200_sscan_fwd.argtypes = [C.POINTER(SScanFwdParams), vp]
    lib.b200_sscan_bwd.argtypes = [C.POINTER(SScanBwdParams), vp]
    lib.b200_ssd_fwd.argtypes = [C.POINTER(SsdFwdParams), vp]
    lib.b200_ssd_bwd.argtypes = [C.POINTER(SsdBwdParams), vp]
    for name in ("b200_cross_scan_pack", "b200_cross_scan_pack_bwd", "b200_cross_merge", "b200_cross_merge_bwd"):
        getattr(lib, name).argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.b200_cross_scan4.argtypes = [vp, i64, vp, i32, i32, i32, i32, vp]
    lib.b200_cross_scan4_bwd.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp]
    lib.b200_ssd_merge4.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.b200_ssd_merge4_bwd.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.b200_rmsnorm_gated_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i32, f32, vp]
    lib.b200_rmsnorm_gated_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, vp]
    lib.b200_dwconv_silu_fwd.argtypes = [vp, i64, i32, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.b200_dwconv_silu_bwd.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.b200_shuffle_cat_add_fwd.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, vp]
    lib.b200_shuffle_cat_add_bwd.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp]
    lib.b200_patch_merge.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.b200_atrous_scan.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.b200_atrous_merge.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.b200_cross_scan_pack_strided.argtypes = [vp, vp, i64, i64, i64, i32, i32, i32, i32, vp]
    lib.b200_cross_scan_unpack4.argtypes = [vp, vp, i64, i64, i64, vp, i32, i32, i32, i32, vp]
    lib.b200_ln_gate_grid.argtypes = [i64]
    lib.b200_ln_gate_fwd.argtypes = [vp, i32, i64, vp, i64, i32, vp, vp, vp, i32, vp, vp, i64, i32, f32, vp]
    lib.b200_ln_gate_bwd.argtypes = [vp, vp, i32, i64, vp, i64, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    for which, st in enumerate((SScanFwdParams, SScanBwdParams, SsdFwdParams, SsdBwdParams)):
        if lib.b200_sizeof_params(which) != C.sizeof(st):
            raise RuntimeError(f"libb200ssm ABI mismatch for {st.__name__}: "
                               f"C {lib.b200_sizeof_params(which)} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libb200ssm ops run on CUDA tensors only: there is no CPU fallback")


def no_autocast(fn):
    """Decorator for the forward / backward of every ctypes-backed autograd.Function: the kernels read raw data_ptr()s with a fixed
    dtype, so nothing inside may be re-cast by an enclosing torch.autocast region (loss.backward() called inside the autocast
    block runs the backward under autocast too; a torch.bmm in there would then hand a bf16 buffer to an fp32 kernel)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with torch.autocast("cuda", enabled=False):
            return fn(*args, **kwargs)
    return wrapper


def launches() -> int:
    return int(load().b200_kernel_launches())
