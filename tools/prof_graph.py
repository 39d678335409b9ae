"""Kernel timeline of one CUDA-graph replay of the training step (CUPTI through torch.profiler): per stream busy time, idle gaps
on the critical (main) stream, top kernels.   python tools/prof_graph.py [model] [--no-overlap]"""
import sys, json, collections
import torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
import bench
from medical_image_classification_b200 import models
from medical_image_classification_b200.train_step import TrainStep
name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "medmamba_t"
models.SS_Conv_SSM.overlap_branches = "--no-overlap" not in sys.argv
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
net = bench.build_model(name).to(dev)
step = TrainStep(net, lr=1e-4, autocast=torch.bfloat16)
x = torch.randn(64, 3, 224, 224, device=dev)
y = torch.randint(0, 6, (64,), device=dev)
step.warmup(x, y, n=3)
assert step.capture(x, y), step.note
for _ in range(3):
    step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, y)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = [e for e in json.load(open("/tmp/trace.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} device activities, span {(t1 - t0) / 1e3:.3f} ms")
by = collections.defaultdict(list)
for e in ev:
    by[e["args"].get("stream")].append(e)
for st, es in sorted(by.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in es)
    gaps = [b["ts"] - (a["ts"] + a["dur"]) for a, b in zip(es, es[1:])]
    pos = [g for g in gaps if g > 0]
    small = [g for g in pos if g < 20]
    print(f"stream {st}: {len(es)} activities, busy {busy / 1e3:.3f} ms, idle between its kernels {sum(pos) / 1e3:.3f} ms "
          f"({len(small)} gaps < 20 us totalling {sum(small) / 1e3:.3f} ms; {len(pos) - len(small)} longer gaps {sum(g for g in pos if g >= 20) / 1e3:.3f} ms)")
main = max(by.values(), key=lambda es: sum(e["dur"] for e in es))
cnt = collections.Counter(); dur = collections.Counter()
for e in main:
    cnt[e["name"][:90]] += 1; dur[e["name"][:90]] += e["dur"]
print("main stream, kernels by launch count:")
for n, c in cnt.most_common(22):
    print(f"  {c:4d} x  {dur[n] / 1e3:7.3f} ms  {n}")
print("main stream, gaps >= 20 us (after -> before):")
for a, b in zip(main, main[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g >= 20:
        print(f"  {g:7.1f} us at t = {(a['ts'] - t0) / 1e3:6.2f} ms   {a['name'][:60]}  ->  {b['name'][:60]}")
agg = collections.Counter()
for e in ev:
    agg[e["name"][:70]] += e["dur"]
for n, d in agg.most_common(14):
    print(f"  {d / 1e3:7.3f} ms  {n}")
