"""mamba_ssm.distributed.distributed_utils: see tensor_parallel.py (SSD/MedSSD.py:39, only used with a process group)."""


def all_reduce(*args, **kwargs):
    raise NotImplementedError("all_reduce: tensor parallelism is dead code in the reference models")


def reduce_scatter(*args, **kwargs):
    raise NotImplementedError("reduce_scatter: tensor parallelism is dead code in the reference models")
