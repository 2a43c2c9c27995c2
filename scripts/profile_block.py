import sys, torch
sys.path.insert(0, ".")
from torch.profiler import ProfilerActivity, profile
from mamba_ssm import Mamba
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
m = Mamba(d_model=64, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nframes=5).cuda()
x = torch.randn(B, 20480, 64, device="cuda", requires_grad=True)
gy = torch.randn(B, 20480, 64, device="cuda")
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    y.backward(gy.to(y.dtype))
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted([(e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0], reverse=True)
tot = sum(r[0] for r in rows)
print(f"batch {B}: total {tot:.0f} us, {sum(r[1] for r in rows)} launches")
for t, n, k in rows[:28]:
    print(f"{t:8.1f} us  x{n:<3d} {k[:120]}")
