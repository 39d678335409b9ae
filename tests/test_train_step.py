"""GPU: the TrainStep product API (reference ddp_train.py:160-166): eager and graph-captured steps run and
train (finite, falling loss on a repeated batch), and a second device in the same process works (function attributes are
per device)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(dev):
    from medical_image_classification_b200.models import VSSM
    torch.manual_seed(0)
    return VSSM(num_classes=6, depths=[1, 1], dims=[32, 64], drop_path_rate=0.0).to(dev)


def _run(graph, steps=4):
    from medical_image_classification_b200.train_step import TrainStep
    dev = torch.device("cuda:0")
    net = _net(dev)
    step = TrainStep(net, lr=1e-3, autocast=torch.bfloat16, graph=graph)
    g = torch.Generator(device="cpu").manual_seed(1)
    xs = [torch.randn(4, 3, 64, 64, generator=g).to(dev) for _ in range(steps)]
    ys = [torch.randint(0, 6, (4,), generator=g).to(dev) for _ in range(steps)]
    if graph:
        step.warmup(xs[0], ys[0], n=3)
        assert step.capture(xs[0], ys[0]), step.note
    losses = [float(step(x, y)) for x, y in zip(xs, ys)]
    return losses, [p.detach().float().clone() for p in net.parameters()], step


def test_eager_step_runs_and_loss_is_finite():
    losses, params, step = _run(graph=False)
    assert all(l == l and abs(l) < 1e3 for l in losses)
    assert step.note == "eager"


def test_graph_capture_replays():
    losses, params, step = _run(graph=True)
    assert step.graph is not None and "CUDA graph" in step.note
    assert all(l == l and abs(l) < 1e3 for l in losses)
    # the replayed graph really updates the weights: repeating one batch drives its loss down
    x = torch.randn(4, 3, 64, 64, device="cuda:0")
    y = torch.randint(0, 6, (4,), device="cuda:0")
    first = float(step(x, y))
    for _ in range(20):
        last = float(step(x, y))
    assert last < first


def test_second_device_in_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from medical_image_classification_b200.models import medssd
    for d in (0, 1):
        dev = torch.device("cuda", d)
        net = _net(dev)
        x = torch.randn(2, 3, 64, 64, device=dev)
        net(x).sum().backward()
        m = medssd(num_classes=6, depths=(1, 1), dims=(64, 128), d_state=16).to(dev)     # SSD kernels opt in to > 48 KB of shared memory
        m(x).sum().backward()
        torch.cuda.synchronize(dev)
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
