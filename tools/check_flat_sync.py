"""N-rank GPU check of train_step.FlatGradSync (run under torchrun): every rank feeds the same batch to the same weights, so the
rank-averaged gradients must equal the gradients of a plain single-process backward; then two steps of TrainStep (eager and
graph-captured) must leave identical weights on every rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_flat_sync.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from medical_image_classification_b200.models import VSSM
from medical_image_classification_b200.train_step import FlatGradSync, TrainStep

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def net():
    torch.manual_seed(0)
    return VSSM(num_classes=6, depths=[1, 1, 1], dims=[32, 64, 128], drop_path_rate=0.0).to(dev)


g = torch.Generator().manual_seed(1)
x = torch.randn(8, 3, 64, 64, generator=g).to(dev)
y = torch.randint(0, 6, (8,), generator=g).to(dev)
ref = net()
torch.nn.functional.cross_entropy(ref(x), y).backward()
torch.cuda.synchronize()
m = net()
sync = FlatGradSync(m)
for it in range(3):
    sync.begin()
    torch.nn.functional.cross_entropy(m(x), y).backward()
    sync.finish()
torch.cuda.synchronize()
worst = 0.0
for (k, p), q in zip(m.named_parameters(), ref.parameters()):
    err = (p.grad - q.grad).abs().max().item() / max(q.grad.abs().max().item(), 1e-12)
    worst = max(worst, err)
print(f"rank {rank}: FlatGradSync gradients vs single-process backward: max rel diff {worst:.2e} ({len(sync.slices)} chunks)", flush=True)
assert worst < 1e-5, worst
for graph in (False, True):
    m = net()
    step = TrainStep(m, lr=1e-3, autocast=torch.bfloat16, ddp=True, local_rank=local, graph=graph)
    if graph:
        step.warmup(x, y)
        assert step.capture(x, y), step.note
    for _ in range(3):
        loss = step(x, y)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in m.parameters()])
    other = flat.clone()
    dist.broadcast(other, 0)
    d = (flat - other).abs().max().item()
    print(f"rank {rank}: TrainStep graph={graph}: loss {float(loss):.4f}, max |w - w_rank0| = {d:.2e}", flush=True)
    assert d == 0.0 or d < 1e-6
dist.barrier()
os._exit(0)
