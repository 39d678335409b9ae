"""The training step around the hot path -- SURVEY.md 8(f) rank 4: the B200 counterpart of the loop body of the reference's
`ddp_train.py:160-166` (zero_grad, forward, CrossEntropy, backward, `optimizer.step()`) and of its DistributedDataParallel
wrapping (`ddp_train.py:132-134`: one process per GPU, per-rank BatchNorm -- no SyncBN -- and DDP's default per-iteration
buffer broadcast).

    step = TrainStep(net, lr=1e-4, autocast=torch.bfloat16, ddp=world > 1, local_rank=local_rank)
    step.warmup(x, y)            # eager iterations on a side stream (cuDNN autotune, DDP bucket build)
    step.capture(x, y)           # forward + loss + backward (+ the NCCL all-reduces) + fused Adam in ONE CUDA graph
    loss = step(x, y)            # copies the batch into the graph's static inputs and replays

What is B200-specific here:
  * the whole step is one CUDA graph: ~2 900 kernel launches per MedMamba-T step, most of them microsecond-scale at
    stages 2-3, cost nothing on the host (the reference launches them one by one);
  * fused (multi-tensor, capturable) Adam;
  * DDP buckets sized for launch latency and overlap, not for link count (NVSwitch gives every pair full bandwidth): small
    buckets (`bucket_cap_mb`, default 8 MB instead of PyTorch's 25) so that the LAST bucket -- stage-0 / patch-embed
    gradients, ready only when backward ends -- leaves a short exposed all-reduce tail, `gradient_as_bucket_view=True`
    (no gradient copy into the buckets), `static_graph=True`;
  * `broadcast_buffers=False`: DDP's default re-broadcasts rank 0's BatchNorm running statistics before every forward
    (`ddp_train.py:134` keeps that default).  In training mode BatchNorm normalises with batch statistics, so the broadcast never
    changes a gradient or a weight, and rank 0 is never
    overwritten: the validation accuracy it logs and the checkpoint it saves (`ddp_train.py:181-193`, main process only) are bit-identical either way.  Skipping it removes ~90 small broadcasts (one coalesced launch plus
    its dependencies) from every step: 24.64 -> 24.43 ms at N = 2 (`profiles/bench_r02_n2*.json`).  Pass True to get DDP's default
    back (ranks > 0 then evaluate with rank 0's statistics);
  * optional bf16 gradient compression of the all-reduce payload (`grad_bf16=True`: PyTorch's bf16_compress_hook);
    off by default because the reference reduces fp32 gradients.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class TrainStep:
    def __init__(self, net, lr=1e-4, autocast=torch.bfloat16, ddp=False, local_rank=0, graph=True, bucket_cap_mb=8,
                 grad_bf16=False, broadcast_buffers=False, loss_fn=None):
        self.net = net
        self.dev = next(net.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("TrainStep runs on CUDA modules only: the B200 path has no CPU fallback")
        self.autocast = autocast
        self.use_graph = bool(graph)
        self.ddp = bool(ddp)
        self.loss_fn = loss_fn or torch.nn.functional.cross_entropy
        self.side = torch.cuda.Stream(device=self.dev)
        if self.ddp:
            self.side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.side):   # DDP built on the capture-warm-up stream (PyTorch CUDA-graph + DDP recipe)
                self.model = torch.nn.parallel.DistributedDataParallel(
                    net, device_ids=[local_rank], gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb, static_graph=True,
                    broadcast_buffers=broadcast_buffers)
                if grad_bf16:
                    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                    self.model.register_comm_hook(None, default_hooks.bf16_compress_hook)
            torch.cuda.current_stream(self.dev).wait_stream(self.side)
        else:
            self.model = net
        self.local_rank = local_rank
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, fused=True, capturable=self.use_graph)
        self.graph = None
        self.static_x = self.static_y = self.static_loss = None
        self.note = "eager"

    # ---- one step, eagerly -----------------------------------------------------------------------------------------
    def _fwd_bwd_opt(self, x, y):
        if self.autocast is not None:
            with torch.autocast("cuda", dtype=self.autocast):
                loss = self.loss_fn(self.model(x).float(), y)
        else:
            loss = self.loss_fn(self.model(x), y)
        loss.backward()
        self.opt.step()
        return loss

    def eager(self, x, y):
        self.opt.zero_grad(set_to_none=True)
        return self._fwd_bwd_opt(x, y)

    def barrier(self):
        if self.ddp:
            dist.barrier(device_ids=[self.local_rank])
        torch.cuda.synchronize(self.dev)

    def warmup(self, x, y, n=3):
        """n eager steps on a side stream (>= 11 under DDP: its buckets are rebuilt during the first iterations and must be
        final before a capture)."""
        n = max(n, 11 if (self.ddp and self.use_graph) else 1)
        self.side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.side):
            for _ in range(n):
                self.eager(x, y)
        torch.cuda.current_stream(self.dev).wait_stream(self.side)
        self.barrier()

    def capture(self, x, y):
        """Capture forward, loss, backward (with DDP's bucketed all-reduces) and the fused Adam update in one CUDA graph.
        Falls back to eager launches (and says so in .note) if the capture fails."""
        if not self.use_graph:
            return False
        try:
            self.static_x, self.static_y = x.clone(), y.clone()
            self.opt.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._fwd_bwd_opt(self.static_x, self.static_y)
            self.note = "whole step captured in one CUDA graph"
            return True
        except Exception as exc:
            self.graph = None
            self.note = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            torch.cuda.synchronize(self.dev)
            self.opt.zero_grad(set_to_none=True)
            return False

    def __call__(self, x, y):
        if self.graph is None:
            return self.eager(x, y)
        if x is not self.static_x:
            self.static_x.copy_(x, non_blocking=True)
            self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.static_loss
