"""mamba_ssm.ops.triton.ssd_combined -> the B200 SSD operator (reference SSD/MedSSD.py:41-42)."""
from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined  # noqa: F401


def mamba_split_conv1d_scan_combined(*args, **kwargs):
    """Imported by the reference (SSD/MedSSD.py:42) but never called: its use_mem_eff_path branch is dead code there."""
    raise NotImplementedError("mamba_split_conv1d_scan_combined is not part of the B200 hot path (never called by the reference models)")
