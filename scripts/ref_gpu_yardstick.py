"""Context number: the UNMODIFIED reference modules (oracle/_ref: Model.UNet + loss.calc_loss) on
cuda:0 through PyTorch's own cuDNN path, same training step as bench.py (config 2: 16 x 3 x 512^2, dice_bce_mc, SGD), in the
modes a user of the reference could run: fp32 as shipped (TF32 convs), fp32 without TF32, bf16 autocast NCHW, bf16 autocast
channels_last. Not a bench value of this repository - it only tells what the stock stack reaches on the same box.

    python scripts/ref_gpu_yardstick.py [B] [S]   -> gpurun_out/ref_yardstick.json (+ a profiler table of the cuDNN kernels)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
STEPS, WARM = 8, 3
RefModel, ref_loss = ref_loader.load()
ref_loss.CLASS_NUMBER = 2
dev = torch.device("cuda", 0)
results = {}


def run(name, tf32, autocast, channels_last):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = RefModel.UNet(3, 2).to(dev).train()
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(B, 3, S, S, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 2, (B, S, S), device=dev).float()

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = net(x)
        loss = ref_loss.calc_loss(out.float(), y, loss_type="dice_bce_mc")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(WARM):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    results[name] = {"ms_per_step": ms, "img_per_s": B / ms * 1e3, "loss": float(loss),
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    print(name, results[name], flush=True)
    if name == "bf16_autocast_channels_last":
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90), flush=True)
    del net, opt, x, y
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


for name, cfg in [("fp32_tf32_as_shipped", (True, False, False)), ("fp32_no_tf32", (False, False, False)),
                  ("bf16_autocast_nchw", (True, True, False)), ("bf16_autocast_channels_last", (True, True, True))]:
    try:
        run(name, *cfg)
    except Exception as e:  # noqa: BLE001 - e.g. out of memory in one mode must not hide the others
        results[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(name, results[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ref_yardstick.json"), "w") as f:
    json.dump({"config": f"reference Model.UNet(3,2) + calc_loss('dice_bce_mc') + SGD, {B}x3x{S}x{S}, torch "
                         f"{torch.__version__}, cuDNN {torch.backends.cudnn.version()}", "results": results}, f, indent=1)
