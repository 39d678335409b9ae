"""torch.profiler view of one eager MedMamba-T training step: which ATen ops (and from where) launch the glue kernels.
    python tools/prof_ops.py [model] > gpurun_out/prof_ops.txt"""
import sys
import torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "medmamba_t"
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
net = bench.build_model(name).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
x = torch.randn(64, 3, 224, 224, device=dev)
y = torch.randint(0, 6, (64,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(net(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=60, max_shapes_column_width=90))
print(prof.key_averages(group_by_stack_n=6).table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=50, max_src_column_width=110))
# ---- the glue ops by input shape: where the small copies / fills / sums come from ----
rows = [e for e in prof.key_averages(group_by_input_shape=True)
        if e.key in ("aten::copy_", "aten::fill_", "aten::sum", "aten::add_", "aten::add", "aten::mul", "aten::cat", "aten::zero_")]
rows.sort(key=lambda e: -e.self_device_time_total)
print("glue ops by input shape (self CUDA us, calls, shapes):")
for e in rows[:60]:
    print(f"  {e.key:12s} {e.self_device_time_total:9.1f} us  x{e.count:3d}  {str(e.input_shapes)[:150]}")
