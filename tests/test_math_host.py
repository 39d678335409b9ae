"""Host restatement (numpy float32, operation by operation) of csrc/common.cuh::softplus_sigmoid -- the once-per-element
routine of the selective-scan kernels -- checked against float64 F.softplus(threshold=20) / sigmoid.  What this pins is the
ALGORITHM (range reduction, the shared reciprocal, the truncated atanh series); the MUFU approximations themselves (ex2, rcp:
~1-2 ulp) are covered by the GPU parity tests of the scan."""
import numpy as np

f32 = np.float32


def softplus_sigmoid_f32(x):
    x = x.astype(f32)
    e = np.exp2(-np.abs(x) * f32(1.4426950408889634)).astype(f32)
    t1, t2 = f32(1) + e, f32(2) + e
    r = (f32(1) / (t1 * t2)).astype(f32)
    inv1 = r * t2
    s = e * (r * t1)
    s2 = s * s
    p = s2 * f32(1 / 15) + f32(1 / 13)
    for c in (1 / 11, 1 / 9, 1 / 7, 0.2, 1 / 3, 1.0):
        p = (s2 * p + f32(c)).astype(f32)
    sp = ((s + s) * p + np.maximum(x, f32(0))).astype(f32)
    sig = np.where(x >= 0, inv1, e * inv1).astype(f32)
    big = x > 20
    return np.where(big, x, sp), np.where(big, f32(1), sig)


def test_softplus_sigmoid_algorithm_error():
    x = np.concatenate([np.linspace(-30, 20, 200001), np.linspace(-1e-3, 1e-3, 2001), [20.0, 20.0001, 25.0, 60.0, -87.0, -200.0, 0.0]])
    sp, sig = softplus_sigmoid_f32(x)
    x64 = x.astype(f32).astype(np.float64)
    sp_ref = np.where(x64 > 20, x64, np.log1p(np.exp(np.minimum(x64, 20.0))))
    sig_ref = np.where(x64 > 20, 1.0, 1.0 / (1.0 + np.exp(-x64)))
    rel_sp = np.abs(sp - sp_ref) / np.maximum(np.abs(sp_ref), 1e-30)
    rel_sig = np.abs(sig - sig_ref) / np.maximum(np.abs(sig_ref), 1e-30)
    sel = x64 >= -30          # below that both outputs are < 1e-13: only their absolute error matters
    # e = ex2(-|x| log2 e) carries the rounding of the product: relative error |x| 2^-24 on the (tiny) results for x << 0
    bound = np.maximum(6.5e-7, np.abs(x64) * 1e-7)
    assert np.all(rel_sp[sel] < bound[sel]), (rel_sp[sel] / bound[sel]).max()
    assert np.all(rel_sig[sel] < bound[sel]), (rel_sig[sel] / bound[sel]).max()
    assert np.abs(sp - sp_ref).max() < 1e-6 and np.abs(sig - sig_ref).max() < 2e-7      # absolute, whole range (half an ulp of 20 is 9.5e-7)
    assert np.abs(sp - sp_ref)[~sel].max() < 1e-30 and np.abs(sig - sig_ref)[~sel].max() < 1e-30
    assert np.all(np.isfinite(sp)) and np.all(np.isfinite(sig))
