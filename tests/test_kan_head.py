"""Host-side KAN head of the MedSSD_kan family (BASELINE.json configs[3]; outside the hot path): the mirror in
medical_image_classification_b200/kan_head.py against golden vectors made by the reference's KansModule
(MedSSD_kan/MedSSD_kan.py:475-501; oracle/make_golden.py::kan_head_case).  CPU."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from medical_image_classification_b200.kan_head import KansModule, bspline_basis


def rel(a, b):
    return float(np.abs(a.detach().numpy() - b).max() / max(np.abs(b).max(), 1e-30))


def test_kans_module_matches_reference():
    g = np.load(os.path.join(GOLDEN, "kan_head.npz"))
    m = KansModule(24, 6)
    m.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    x = torch.tensor(g["x"], requires_grad=True)
    m.train()
    out = m(x)
    assert rel(out, g["out_train"]) < 1e-6
    out.backward(torch.tensor(g["g"]))
    assert rel(x.grad, g["dx"]) < 1e-5
    n = 0
    for k, p in m.named_parameters():
        if "grad." + k in g.files:
            assert rel(p.grad, g["grad." + k]) < 1e-5, k
            n += 1
    assert n >= 6
    m.eval()
    with torch.no_grad():
        assert rel(m(x), g["out_eval"]) < 1e-6


def test_bspline_basis_partition_of_unity():
    """Inside the un-extended grid range the order-k basis functions sum to 1 (known answer, no reference needed)."""
    k, num = 3, 5
    knots = (torch.arange(-k, num + k + 1, dtype=torch.float64) * (2.0 / num) - 1.0)[None, :].repeat(4, 1)
    x = torch.rand(100, 4, dtype=torch.float64) * 1.98 - 0.99
    b = bspline_basis(x, knots, k)
    assert b.shape == (100, 4, num + k)
    assert torch.allclose(b.sum(-1), torch.ones(100, 4, dtype=torch.float64), atol=1e-12)
    assert (b >= 0).all()
