"""torch.profiler breakdown of one MedMamba-T training step (batch 64, bf16 autocast)."""
import sys
import torch
sys.path.insert(0, ".")
from medical_image_classification_b200.models import medmamba_t
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
torch.backends.cudnn.benchmark = True
net = medmamba_t(num_classes=6).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 6, (B,), device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(net(x).float(), y)
    loss.backward()
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
# kernel-only view: device-side events, per step
from torch.autograd import DeviceType
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
ks.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ks)
print(f"\n== kernels: {tot / 3e3:.2f} ms of device time per step ==")
for e in ks[:70]:
    print(f"{e.self_device_time_total / 3e3:8.3f} ms/step {e.count / 3:7.1f} launches/step  {e.key[:150]}")
