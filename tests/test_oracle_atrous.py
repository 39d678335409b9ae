"""CPU: the numpy restatement of the atrous scan / merge (oracle.atrous_scan_ref / atrous_merge_ref) against the reference's
own EfficientScan / EfficientMerge code.  models/cross.py cannot be imported whole (it needs mamba_ssm, timm and a compiled
selective_scan_cuda at import time), so the two autograd classes -- pure tensor ops -- are compiled from their source lines
in place (CrossMamba/FusionMamba/models/cross.py:34-90, 139-190), unmodified; when /root/reference is absent (GPU box) the
same slice formulas restated in tests/test_atrous.py pin the oracle instead."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle

REF = "/root/reference/CrossMamba/FusionMamba/models/cross.py"


def _reference_classes():
    lines = open(REF).read().split("\n")
    src = "\n".join(lines[33:90]) + "\n\n" + "\n".join(lines[138:190]) + "\n"      # EfficientMerge (:34-90), EfficientScan (:139-190)
    ns = {"torch": torch, "F": F, "math": math}
    exec(compile(src, REF, "exec"), ns)
    return ns["EfficientScan"], ns["EfficientMerge"]


def _slices_scan(x, s=2):
    B, C, H, W = x.shape
    if W % s:
        x = F.pad(x, (0, s - W % s, 0, 0))
    if H % s:
        x = F.pad(x, (0, 0, 0, s - H % s))
    return torch.stack([x[:, :, ::s, ::s].reshape(B, C, -1), x.transpose(2, 3)[:, :, ::s, 1::s].reshape(B, C, -1),
                        x[:, :, ::s, 1::s].reshape(B, C, -1), x.transpose(2, 3)[:, :, 1::s, 1::s].reshape(B, C, -1)], dim=1)


@pytest.mark.parametrize("shape", [(2, 3, 8, 8), (1, 5, 7, 9), (2, 2, 14, 14), (3, 2, 1, 6), (1, 1, 10, 3)])
def test_atrous_oracle_matches_reference(shape):
    B, C, H, W = shape
    rng = np.random.default_rng(H * 31 + W)
    x = rng.standard_normal(shape).astype(np.float32)
    xs = oracle.atrous_scan_ref(x)
    assert np.array_equal(xs, _slices_scan(torch.tensor(x)).numpy())
    ys = rng.standard_normal(xs.shape).astype(np.float32)
    y = oracle.atrous_merge_ref(ys, H, W)
    # merge is the inverse of scan on the image and its adjoint on gradients
    assert np.array_equal(oracle.atrous_merge_ref(xs, H, W).reshape(B, C, H, W), x)
    assert abs(float((xs * ys).sum()) - float((x.reshape(B, C, -1) * y).sum())) < 1e-3
    if os.path.exists(REF):
        Scan, Merge = _reference_classes()
        assert np.array_equal(xs, Scan.apply(torch.tensor(x), 2).numpy())
        assert np.array_equal(y, Merge.apply(torch.tensor(ys), H, W, 2).numpy())


GOLDENS = sorted(f for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if f.startswith("atrous_"))


@pytest.mark.parametrize("name", GOLDENS)
def test_atrous_oracle_matches_committed_reference_vectors(name):
    """tests/golden/atrous_*.npz were produced by the reference's own EfficientScan / EfficientMerge (oracle/make_golden.py
    --atrous-only); they travel to the GPU box, where /root/reference does not exist."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    B, C, H, W = g["x"].shape
    assert np.array_equal(oracle.atrous_scan_ref(g["x"]), g["xs"])
    assert np.array_equal(oracle.atrous_merge_ref(g["ys"], H, W), g["y"])
    # the reference's autograd: each map is the other's adjoint
    assert np.array_equal(oracle.atrous_merge_ref(g["gxs"], H, W).reshape(B, C, H, W), g["dx"])
    assert np.array_equal(oracle.atrous_scan_ref(g["gy"].reshape(B, C, H, W)), g["dys"])

