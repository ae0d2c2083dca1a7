"""Kernel-time table of one data-parallel training step on rank 0 (torchrun, N >= 2): where the multi-GPU overhead is."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200 as U  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
ctx = U.init_from_env(sync_bn=True)
torch.manual_seed(0)
net = U.UNet(3, 2).cuda().train()
U.loss.CLASS_NUMBER = 2
opt = U.FusedSGD(net, lr=0.01, momentum=0.9, weight_decay=1e-4)
x = torch.randn(16, 3, 512, 512, device="cuda")
y = torch.randint(0, 2, (16, 512, 512), device="cuda").float()


def step():
    out = net(x)
    loss = U.calc_loss(out, y, loss_type="dice_bce_mc")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"eager DP step: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
if rank == 0:
    rows = {}
    tmin, tmax = None, None
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name, [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time
            t0, t1 = ev.time_range.start, ev.time_range.end
            tmin = t0 if tmin is None else min(tmin, t0)
            tmax = t1 if tmax is None else max(tmax, t1)
    tot = sum(v[1] for v in rows.values())
    print(f"summed device time {tot / 2e3:.2f} ms/step; first-to-last kernel span {(tmax - tmin) / 2e3:.2f} ms/step")
    for name, (cnt, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{t / 2e3:9.3f} ms/step  x{cnt // 2:4d}  {name[:100]}")
    for name, (cnt, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        if "nccl" in name.lower() or "nvl" in name.lower():
            print(f"{t / 2e3:9.3f} ms/step  x{cnt // 2:4d}  {name[:100]}")
dist.barrier()
dist.destroy_process_group()
