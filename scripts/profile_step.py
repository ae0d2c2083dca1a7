"""Kernel-time table of one training step (torch.profiler / CUPTI), config 2 by default."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200 as U  # noqa: E402

B = int(os.environ.get("B", 16))
S = int(os.environ.get("S", 512))
torch.manual_seed(0)
net = U.UNet(3, 2).cuda().train()
U.loss.CLASS_NUMBER = 2
opt = U.FusedSGD(net, lr=0.01, momentum=0.9, weight_decay=1e-4)
x = torch.randn(B, 3, S, S, device="cuda")
y = torch.randint(0, 2, (B, S, S), device="cuda").float()


def step():
    out = net(x)
    loss = U.calc_loss(out, y, loss_type="dice_bce_mc")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in rows.values())
print(f"total device time for 2 steps: {tot/1e3:.2f} ms")
for name, (cnt, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/2e3:9.3f} ms/step {100*t/tot:5.1f}%  x{cnt//2:4d}  {name[:110]}")
