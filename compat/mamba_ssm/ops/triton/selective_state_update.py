"""mamba_ssm.ops.triton.selective_state_update: decode-time state update, imported inside try/except (SSD/MedSSD.py:31-34)
and never called by the reference models."""


def selective_state_update(*args, **kwargs):
    raise NotImplementedError("selective_state_update (single-step decoding) is not used by the reference models")
