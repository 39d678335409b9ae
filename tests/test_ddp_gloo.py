"""N > 1 host logic on CPU: one process per rank, batch sharded across ranks, no data-path
collective, DDP all-reduce(avg) of the parameter gradients -- world_size 2, gloo backend.  The scan runs through
oracle/cpu_path.py here (there is no GPU in this container); the sharding / reduction logic is what is tested:
the rank-averaged gradients must equal the single-process gradients of the full batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.cpu_path import bind_cpu_core
from medical_image_classification_b200.models import VSSM


def _net():
    torch.manual_seed(0)
    net = VSSM(num_classes=6, depths=[1, 1, 1, 1], dims=[8, 16, 32, 64], drop_path_rate=0.0)
    bind_cpu_core(net)
    net.eval()   # BatchNorm on running statistics: per-rank batch statistics (no SyncBN, as in ddp_train.py) would differ by design
    return net


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 6, (4,), generator=g)


def _worker(rank, world, port, out, impl="torch"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _net()
    x, y = _data()
    per = x.shape[0] // world
    xs, ys = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]   # DistributedSampler-style shard
    if impl == "torch":
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        loss = torch.nn.functional.cross_entropy(ddp(xs), ys)
        loss.backward()
    else:   # the product's reducer (train_step.FlatGradSync): flat gradient buffer, chunked all-reduce from accumulate hooks
        from medical_image_classification_b200.train_step import FlatGradSync
        if rank == 1:   # construction must bring every rank to rank 0's parameters
            with torch.no_grad():
                for p_ in net.parameters():
                    p_.add_(1.0)
        sync = FlatGradSync(net)
        assert len(sync.slices) >= 2 and sum(sync.need) == len(sync.params)
        for _ in range(2):   # twice: begin() must reset the in-place accumulated buffer
            sync.begin()
            loss = torch.nn.functional.cross_entropy(net(xs), ys)
            loss.backward()
            assert any(sync.sent[:-1]) or len(sync.sent) == 1   # early chunks left from the hooks, during backward
            sync.finish()
            assert all(sync.sent)
        assert all(p_.grad.data_ptr() >= sync.flat.data_ptr() for p_ in net.parameters())
    t = torch.tensor([float(loss)])
    dist.all_reduce(t)           # the bench's max/mean-over-ranks plumbing
    if rank == 0:
        grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        torch.save({"grads": grads, "loss_sum": float(t)}, out)
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("impl", ["torch", "flat"])
def test_two_rank_gloo_matches_single_process(tmp_path, impl):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out, impl), nprocs=2, join=True)
    got = torch.load(out)
    net = _net()
    x, y = _data()
    losses = [torch.nn.functional.cross_entropy(net(x[i:i + 2]), y[i:i + 2]) for i in (0, 2)]
    (sum(losses) / 2).backward()
    assert abs(got["loss_sum"] - float(sum(losses))) < 1e-5
    n = 0
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = p.grad.numpy()
        err = np.abs(got["grads"][k].numpy() - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert err < 1e-4, (k, err)
        n += 1
    assert n > 20
