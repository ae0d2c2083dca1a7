"""CUDA-event microbench of the edge kernels (edge.cu) against their algorithmic HBM bytes.
L2 is flushed between repetitions."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_torch_b200 as U  # noqa: E402,F401
from unet_torch_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
print("peaks:", json.load(open(pk)) if os.path.exists(pk) else None)
for n, h, w, c in [(16, 512, 512, 3), (32, 1024, 1024, 3), (8, 768, 768, 3)]:
    img = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ops.znorm_to_chw(img))
    # algorithmic bytes: the image read twice (moments, then apply) + fp32 output written once
    by = n * h * w * c * (1 + 1 + 4)
    print(f"znorm_to_chw {n}x{h}x{w}x{c}: {ms * 1e3:8.1f} us  {by / ms / 1e6:8.1f} GB/s  ({by / 1e6:.1f} MB algorithmic)")
for n, h, w, ncls in [(32, 1024, 1024, 5), (16, 512, 512, 2)]:
    a = torch.randn(n, h, w, 64, device="cuda").to(torch.bfloat16)
    wt = torch.randn(ncls, 64, 1, 1, device="cuda") * 0.1
    b = torch.zeros(ncls, device="cuda")
    logits = torch.empty(n, ncls, h, w, device="cuda")
    ms_f = timeit(lambda: ops.head_mask(a, wt, b))
    ms_u = timeit(lambda: (ops.head_fprop(a, wt, b, logits), ops.softmax_argmax(logits)))
    ms_d = timeit(lambda: ops.head_density(a, wt, b, 200.0))
    px = n * h * w
    print(f"head_mask    {n}x{h}x{w} ncls={ncls}: fused {ms_f * 1e3:8.1f} us ({px * 129 / ms_f / 1e6:7.1f} GB/s of 129 B/px)"
          f"   unfused head_fprop+softmax_argmax {ms_u * 1e3:8.1f} us")
    print(f"head_density {n}x{h}x{w} ncls={ncls}: {ms_d * 1e3:8.1f} us ({px * (128 + 4 * ncls) / ms_d / 1e6:7.1f} GB/s of {128 + 4 * ncls} B/px)")
