"""Inference throughput, BASELINE configs[3]: UNet(3,5).eval(), 32 x 3 x 1024 x 1024 tiles -> class mask
(test_mc3serousv5.py:876-887). CUDA events over whole forward passes (inputs resident in HBM, far larger than L2).
Variants: separate BN-apply pass vs BatchNorm+ReLU folded into the conv epilogues; logits + softmax/argmax kernel vs
the fused OutConv+softmax+argmax+uint8 head. 1541.89 GFLOP / image."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200 as U  # noqa: E402

B = int(os.environ.get("B", 32))
S = int(os.environ.get("S", 1024))
GFLOP = 385.37 * (S / 512) ** 2 + (5 - 2) * 2 * 64 * S * S / 1e9  # config-2 forward scaled to the tile + 5-class head
torch.manual_seed(0)
net = U.UNet(3, 5).cuda().eval()
eng = net._get_engine()
x = torch.randn(B, 3, S, S, device="cuda")
img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    with torch.no_grad():
        fn()
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, fold, fn in [
    ("BN-apply pass, logits + softmax_argmax", False, lambda: U.predict_mask(net(x))),
    ("folded BN,     logits + softmax_argmax", True, lambda: U.predict_mask(net(x))),
    ("folded BN,     fused uint8 mask head  ", True, lambda: net.predict(x)),
    ("uint8 image -> preprocess -> folded BN -> fused mask head", True, lambda: net.predict(U.preprocess(img))),
]:
    eng.fold_eval_bn = fold
    ms = timeit(fn)
    print(f"{name}: {ms:8.2f} ms / {B} tiles  {B / ms * 1e3:7.1f} img/s  {B * GFLOP / ms:7.1f} TFLOP/s")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
