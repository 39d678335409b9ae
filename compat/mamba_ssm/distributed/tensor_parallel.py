"""mamba_ssm.distributed.tensor_parallel: only reached when a model is built with process_group != None, which no
reference script does (SSD/MedSSD.py:38, 207-222: process_group defaults to None)."""


class _Unavailable:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}: tensor parallelism is dead code in the reference models "
                                  "(process_group is always None); the B200 path shards by batch only")


class ColumnParallelLinear(_Unavailable):
    pass


class RowParallelLinear(_Unavailable):
    pass
