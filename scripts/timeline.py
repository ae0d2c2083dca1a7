"""In-situ kernel timeline of the bench step (torch.profiler / CUPTI): per-kernel device time inside the running step (warm
caches, real clocks - unlike ncu's serialised replays) and the IDLE time between consecutive kernels, with the largest gaps
named. Answers "where does measured step time - summed kernel time go".

    GRAPHS=1 python scripts/timeline.py > gpurun_out/timeline.txt
    MODEL=UNet_attention python scripts/timeline.py  # the attention-gated network
    torchrun --nproc-per-node N scripts/timeline.py     # data parallel + SyncBN: rank 0's timeline (all streams)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200 as U  # noqa: E402

B = int(os.environ.get("B", 16))
S = int(os.environ.get("S", 512))
GRAPHS = os.environ.get("GRAPHS", "1") != "0"
NSTEP = 3
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    U.init_from_env(sync_bn=True)
    if RANK != 0:
        sys.stdout = open(os.devnull, "w")
torch.manual_seed(0)
net = getattr(U, os.environ.get("MODEL", "UNet"))(3, 2).cuda().train()
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


step()
net.enable_cuda_graphs(GRAPHS, share_grads=GRAPHS)
for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print(f"unprofiled: {e0.elapsed_time(e1) / 10:.3f} ms/step (graphs={GRAPHS})")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tr = ev.time_range
        evs.append((tr.start, tr.end, ev.name))
evs.sort()
t_first, t_last = evs[0][0], max(e[1] for e in evs)
span = (t_last - t_first) / NSTEP
busy = 0.0
cur_end = evs[0][0]
gaps = []
rows = {}
for i, (s, e, name) in enumerate(evs):
    r = rows.setdefault(name, [0, 0.0])
    r[0] += 1
    r[1] += e - s
    if s > cur_end:
        gaps.append((s - cur_end, evs[i - 1][2], name))
    busy += max(0.0, e - max(s, cur_end))
    cur_end = max(cur_end, e)
busy /= NSTEP
print(f"profiled span {span / 1e3:.3f} ms/step, GPU busy {busy / 1e3:.3f} ms/step, idle {(span - busy) / 1e3:.3f} ms/step "
      f"({len(evs) // NSTEP} device activities per step)")
gaps.sort(reverse=True)
print("largest idle gaps (us): before-kernel <- after-kernel")
for g, a, b in gaps[:25]:
    print(f"  {g:8.1f}  {a[:60]}  ->  {b[:60]}")
bucket = {}
for g, a, b in gaps:
    bucket[b[:50]] = bucket.get(b[:50], 0.0) + g
print("idle time by FOLLOWING kernel (us/step):")
for k, v in sorted(bucket.items(), key=lambda kv: -kv[1])[:20]:
    print(f"  {v / NSTEP:8.1f}  {k}")
tot = sum(v[1] for v in rows.values())
print(f"summed device time {tot / NSTEP / 1e3:.3f} ms/step")
comm = {k: v for k, v in rows.items() if "nccl" in k.lower() or "nvl_" in k.lower()}
if comm:
    print("communication kernels (rank 0): " + "; ".join(f"{k[:50]} x{v[0] // NSTEP} {v[1] / NSTEP / 1e3:.3f} ms/step" for k, v in comm.items()))
for name, (cnt, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / NSTEP / 1e3:9.3f} ms/step {100 * t / tot:5.1f}%  x{cnt // NSTEP:4d}  {name[:120]}")

if WORLD > 1:
    import torch.distributed as dist

    # per-rank split: who waits for whom? A rank that arrives late at a SyncBN exchange spends ~no time in nvl_*; the others
    # spend their lead there. compute = everything except the exchange / NCCL kernels.
    mine = {"rank": RANK, "span_ms": span / 1e3, "busy_ms": busy / 1e3,
            "nvl_ms": sum(v[1] for k, v in rows.items() if "nvl_" in k) / NSTEP / 1e3,
            "nccl_ms": sum(v[1] for k, v in rows.items() if "nccldevkernel" in k.lower()) / NSTEP / 1e3,
            "compute_ms": sum(v[1] for k, v in rows.items() if "nvl_" not in k and "nccl" not in k.lower()) / NSTEP / 1e3}
    allr = [None] * WORLD
    dist.all_gather_object(allr, mine)
    print("per-rank (ms/step): rank span busy compute nvl_exchange nccl")
    for r in allr:
        print(f"  {r['rank']}  {r['span_ms']:.3f}  {r['busy_ms']:.3f}  {r['compute_ms']:.3f}  {r['nvl_ms']:.3f}  {r['nccl_ms']:.3f}")
    U.DataParallelContext.disable()
    dist.destroy_process_group()
