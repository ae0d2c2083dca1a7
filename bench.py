#!/usr/bin/env python
"""Benchmark of the U-Net hot path (BASELINE.json metric: UNet train img/s @512^2 bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : one step = forward + 'dice_bce_mc' loss + backward + SGD step of UNet(3, 2) on a synthetic
            16 x 3 x 512 x 512 batch per GPU (BASELINE.json configs[1]); N > 1 = data parallel under torchrun
            (batch 16 per GPU, SyncBN, bucketed gradient all-reduce), weak scaling.
            `value`  : images/s with the batch already resident in HBM, CUDA-event timed, max over ranks.
            `e2e`    : the same through the public API with HOST buffers: pinned-host -> device copy of inputs and
                       labels and a device -> host read of the loss inside the timed region, every step.
            `roofline`: tcgen05 conv3x3 implicit-GEMM launches (fprop + dgrad), algorithmic FLOPs / CUDA-event time.
            `cpu_baseline`: the oracle port of the reference's CPU path timed on this box's host cores (rank 0, N=1).
reference : the reference's own CPU implementation of the path (oracle port: the exact torch CPU ops Model.py /
            loss.py call), all host threads, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(n_channels=3, n_classes=2, H=512, W=512, batch_per_gpu=16, loss="dice_bce_mc",
           lr=0.01, momentum=0.9, weight_decay=1e-4)
TRAIN_GFLOP_PER_IMG = 1155.21  # fwd + dgrad + wgrad of config 2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def conv_traffic():
    """DRAM bytes per conv3x3 fprop/dgrad launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches
    of one step) from the committed ncu capture, or None if it has not been taken for this build."""
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons. The process is started before the warm-up (its start-up takes longer than
    a short timed region); only samples that ARRIVE between mark_begin() and mark_end() are reported."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = (self.t1 if self.t1 is not None else time.perf_counter()) + 0.06  # a sample describes the period before it
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, s in self.samples:
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_factory(batch, h, w, seed=0):
    """The reference's CPU path (Model.UNet + calc_loss('dice_bce_mc') forward+backward, fp32, all host threads),
    restated by the oracle port with the exact torch ops the reference modules dispatch to."""
    from oracle import cpu_baseline

    return cpu_baseline.make_step(CFG["n_channels"], CFG["n_classes"], batch, h, w, seed)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 2
    step = cpu_reference_step_factory(batch, CFG["H"], CFG["W"])
    for _ in range(max(args.warmup, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = batch / dt
    sample = f"{args.steps} fwd+bwd steps of batch {batch} x 3x512x512 fp32 on {cores} host threads (oracle port, torch CPU ops)"
    line = {
        "impl": "reference", "metric": "unet_train_img_per_s_512", "value": v, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus, batch_override=None, graphs=False, torch_optim=False):
    b = batch_override or CFG["batch_per_gpu"]
    return {
        "workload": "BASELINE configs[1]: UNet(3,2) training step (fwd + dice_bce_mc loss + bwd + SGD), 3x512x512, "
                    f"batch {b} per GPU" + (", data parallel + SyncBN (configs[2] per-GPU batch)" if n_gpus > 1 else ""),
        "global_batch": b * n_gpus, "image": [CFG["n_channels"], CFG["H"], CFG["W"]], "loss": CFG["loss"],
        "optimizer": "SGD(lr=0.01, momentum=0.9, weight_decay=1e-4)" + ("" if torch_optim else
                     " as unet_torch_b200.FusedSGD (torch.optim.SGD arithmetic fused with the bf16 operand re-cast)"),
        "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
        "cuda_graphs": ("forward and backward of the network replayed as captured CUDA graphs"
                        + (" (SyncBN NVLink kernels and NCCL gradient all-reduces captured with them)" if n_gpus > 1 else "")
                        + "; loss and optimizer eager") if graphs else "off",
        "l2": "per-step working set (~10 GB of bf16 activations) is far larger than the 126 MB L2; no explicit flush",
    }


# --------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import unet_torch_b200 as U
    from unet_torch_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dp = U.init_from_env(sync_bn=True) if world > 1 else None
    dev = torch.device("cuda", local)

    torch.manual_seed(0)
    net = U.UNet(CFG["n_channels"], CFG["n_classes"]).to(dev).train()
    U.loss.CLASS_NUMBER = CFG["n_classes"]
    if world > 1:
        import torch.distributed as dist

        for p in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(p.data, 0)
    okw = dict(lr=CFG["lr"], momentum=CFG["momentum"], weight_decay=CFG["weight_decay"])
    opt = torch.optim.SGD(net.parameters(), **okw) if args.torch_optim else U.FusedSGD(net, **okw)
    B, H, W = CFG["batch_per_gpu"], CFG["H"], CFG["W"]
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, CFG["n_channels"], H, W, generator=gen).pin_memory()
    y_host = torch.randint(0, CFG["n_classes"], (B, H, W), generator=gen).float().pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        out = net(x)
        loss = U.calc_loss(out, y, loss_type=CFG["loss"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # one eager step: configures the kernels and counts OUR launches per step (graph replays bypass the counter)
    l0 = _lib.query("b200unet_launch_count")
    step(x_dev, y_dev)
    launches_per_step = _lib.query("b200unet_launch_count") - l0
    use_graphs = not args.no_graphs
    net.enable_cuda_graphs(use_graphs)
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    # ---- device-resident timing
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e1.record()
    barrier()
    sampler.mark_end()
    launches = launches_per_step * args.steps
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end: host buffers in, loss out, every step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Double-buffered input pipeline (what a pin_memory DataLoader + non_blocking copies amount to): the pinned host batch
    # of step i+1 crosses PCIe on a copy stream into one of two persistent device buffers while step i computes.
    copy_stream = torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ybuf = [torch.empty_like(y_dev), torch.empty_like(y_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def fetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(done[b])  # the step that last read this buffer has finished
            xbuf[b].copy_(x_host, non_blocking=True)
            ybuf[b].copy_(y_host, non_blocking=True)
            ready[b].record(copy_stream)

    # The loss of every step is read on the host (pinned 4-byte copy + event wait), one step behind its launch, so the
    # host keeps enqueueing step i+1 while step i runs (asynchronous logging); all reads complete inside the timed region.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    torch.cuda.synchronize()
    e0.record()
    last = None
    fetch(0)
    for i in range(args.steps):
        b = i & 1
        if i + 1 < args.steps:
            fetch(i + 1)  # issued BEFORE this step's kernels, so it overlaps them; every step's inputs cross PCIe in the region
        cur.wait_event(ready[b])
        loss = step(xbuf[b], ybuf[b])
        done[b].record(cur)
        loss_host[b].copy_(loss.detach(), non_blocking=True)  # device -> host read of the step's result, every step
        loss_ready[b].record(cur)
        if i >= 1:
            loss_ready[1 - b].synchronize()
            last = float(loss_host[1 - b])
    loss_ready[(args.steps - 1) & 1].synchronize()
    last = float(loss_host[(args.steps - 1) & 1])
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel class (tcgen05 conv3x3 implicit GEMM), CUDA events around each launch
    net.enable_cuda_graphs(False)  # per-launch CUDA events need the eager launch path
    prof = []
    orig = ops.conv3x3

    def timed_conv3x3(x, w_op, out, stats_partial=None):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig(x, w_op, out, stats_partial)
        b.record()
        n, h, w, cin = x.shape
        prof.append((2.0 * n * h * w * out.shape[3] * 9 * cin, a, b))
        return r

    ops.conv3x3 = timed_conv3x3
    try:
        for _ in range(2):
            step(x_dev, y_dev)
        torch.cuda.synchronize()
    finally:
        ops.conv3x3 = orig
    flops = sum(p[0] for p in prof)
    kms = sum(p[1].elapsed_time(p[2]) for p in prof)
    pk = peaks()
    achieved = flops / (kms / 1e3) / 1e12 if kms > 0 else 0.0
    roofline = {"bound": "tensor",
                "kernel": "conv3x3 fprop+dgrad implicit GEMMs (conv3_pair / conv3_res2 / conv3_res kernels, tcgen05; "
                          "34 launches per step, 12.3 of the 18.5 TFLOP of a step)",
                "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                "peak_source": f"{pk['src']} bf16_tflops_sustained", "launches_timed": len(prof),
                "share_of_step": (kms / 2) / ms_per_step, "traffic": conv_traffic()}

    if rank != 0:
        return 0
    line = {
        "metric": "unet_train_img_per_s_512", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world, graphs=use_graphs, torch_optim=args.torch_optim),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": last,
                "how": "pinned host batch -> one of two device buffers on a copy stream, one step ahead; the loss of every step is copied to pinned host memory and read one step behind its launch"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "model_tflops": value * TRAIN_GFLOP_PER_IMG / 1e3 / world,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg()
    print(json.dumps(line), flush=True)
    return 0


def _shutdown_dist():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            try:
                import unet_torch_b200 as U

                U.DataParallelContext.disable()  # drops CUDA graphs that captured NCCL kernels (they block the teardown)
            except Exception:
                pass
            dist.destroy_process_group()
    except Exception:
        pass


def cpu_baseline_leg():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 1
    step = cpu_reference_step_factory(batch, CFG["H"], CFG["W"])
    step()  # warm-up
    n, t0 = 0, time.perf_counter()
    while n < 3 and time.perf_counter() - t0 < 25.0:
        step()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": batch / dt, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} fwd+bwd steps of batch {batch} x 3x512x512 fp32 (oracle port of Model.UNet + calc_loss, torch CPU ops)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-optim", action="store_true", help="step with stock torch.optim.SGD instead of FusedSGD")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly (no CUDA-graph replay)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    try:
        return run_ours(args)
    finally:
        _shutdown_dist()


if __name__ == "__main__":
    sys.exit(main())
