"""mamba_ssm.ops.selective_scan_interface -> the B200 operator (reference MedMamba.py:14 imports both names)."""
from medical_image_classification_b200.selective_scan_interface import (  # noqa: F401
    SelectiveScanFn, selective_scan_fn, selective_scan_ref)
