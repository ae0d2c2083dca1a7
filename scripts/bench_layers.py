"""Per-kernel timing on the BASELINE config-2 layer shapes (B=16, 512^2): TFLOP/s for the tensor-core kernels,
GB/s for the bandwidth kernels. CUDA events, L2 flushed between repetitions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200  # noqa: E402
from unet_torch_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
B = int(os.environ.get("B", 16))
S = int(os.environ.get("S", 512))
REPS = int(os.environ.get("REPS", 5))
ONLY = os.environ.get("ONLY", "")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*shape):
    return (torch.randn(*shape, device="cuda") * 0.5).to(BF16)


def want(name):
    return (not ONLY) or any(tok in name for tok in ONLY.split(","))


rows = []


def rec(name, ms, flops=None, bytes_=None):
    s = f"{name:44s} {ms:8.3f} ms"
    if flops:
        s += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    if bytes_:
        s += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    print(s, flush=True)
    rows.append((name, ms))


# ---- conv3x3 layers (fprop shapes; dgrad = swapped channels; wgrad)
layers = [("inc.conv2", 64, 64, 0), ("down1.conv1", 64, 128, 1), ("down1.conv2", 128, 128, 1), ("down2.conv1", 128, 256, 2),
          ("down2.conv2", 256, 256, 2), ("down3.conv1", 256, 512, 3), ("down3.conv2", 512, 512, 3),
          ("down4.conv1", 512, 1024, 4), ("down4.conv2", 1024, 1024, 4), ("up1.conv1", 1024, 512, 3),
          ("up2.conv1", 512, 256, 2), ("up3.conv1", 256, 128, 1), ("up4.conv1", 128, 64, 0)]
tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
for name, cin, cout, lvl in layers:
    h = S >> lvl
    flops = 2.0 * B * h * h * cout * 9 * cin
    x, dy = rnd(B, h, h, cin), rnd(B, h, h, cout)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, wd = ops.prep_conv3x3_weight(w)
    y = torch.empty(B, h, h, cout, dtype=BF16, device="cuda")
    dx = torch.empty(B, h, h, cin, dtype=BF16, device="cuda")
    st = torch.empty(ops.conv3x3_stat_rows(B, h, h, cin, cout) * 2 * cout, device="cuda")
    dw = torch.empty(cout, cin, 3, 3, device="cuda")
    if want("fprop"):
        ms = timeit(lambda: ops.conv3x3(x, wf, y, st)); rec(f"fprop {name} {cin}->{cout} @{h}", ms, flops); tot["fprop"] += ms
    if want("dgrad"):
        ms = timeit(lambda: ops.conv3x3(dy, wd, dx)); rec(f"dgrad {name} {cout}->{cin} @{h}", ms, flops); tot["dgrad"] += ms
    if want("wgrad"):
        ms = timeit(lambda: ops.conv3x3_wgrad(x, dy, dw)); rec(f"wgrad {name} {cin}->{cout} @{h}", ms, flops); tot["wgrad"] += ms
    del x, dy, y, dx, st
print("totals (ms, one instance of each distinct layer shape):", tot, flush=True)

# ---- conv transpose
if want("convt"):
    for name, cin, lvl in (("up1.up", 1024, 4), ("up2.up", 512, 3), ("up3.up", 256, 2), ("up4.up", 128, 1)):
        h, cup = S >> lvl, cin // 2
        flops = 2.0 * B * h * h * 4 * cup * cin
        x = rnd(B, h, h, cin)
        w = torch.randn(cin, cup, 2, 2, device="cuda") * 0.05
        wf, wd = ops.prep_convt2x2_weight(w)
        cat = torch.empty(B, 2 * h, 2 * h, 2 * cup, dtype=BF16, device="cuda")
        bias = torch.zeros(cup, device="cuda")
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        rec(f"convT fprop {name} {cin}->{cup} @{h}", timeit(lambda: ops.convt2x2(x, wf, bias, cat[..., cup:])), flops)
        rec(f"convT dgrad {name}", timeit(lambda: ops.convt2x2_dgrad(cat[..., cup:], wd, dx)), flops)
        rec(f"convT wgrad {name}", timeit(lambda: ops.convt2x2_wgrad(x, cat[..., cup:], dw)), flops)
        del x, cat, dx

# ---- bandwidth kernels at level 0 (64 ch @ 512^2) and level 1
if want("bn"):
    for c, lvl in ((64, 0), (128, 1), (512, 3)):
        h = S >> lvl
        n_el = B * h * h * c
        y, g = rnd(B, h, h, c), rnd(B, h, h, c)
        a = torch.empty_like(y)
        scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
        mean, rstd, gamma = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda"), torch.ones(c, device="cuda")
        dgm, dbt = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        rec(f"bn_relu_fwd c{c} @{h}", timeit(lambda: ops.bn_relu_fwd(y, scale, shift, a)), bytes_=4.0 * n_el)
        pooled = torch.empty(B, h // 2, h // 2, c, dtype=BF16, device="cuda")
        idx = torch.empty(B, h // 2, h // 2, c, dtype=torch.uint8, device="cuda")
        rec(f"bn_relu_fwd+pool c{c} @{h}", timeit(lambda: ops.bn_relu_fwd(y, scale, shift, a, pooled, idx)), bytes_=4.75 * n_el)
        dy = torch.empty_like(y)
        rec(f"bn_relu_bwd (reduce+apply) c{c} @{h}",
            timeit(lambda: ops.bn_relu_bwd(g, None, None, y, gamma, scale, shift, mean, rstd, dy, dgm, dbt)), bytes_=10.0 * n_el)
        gp = rnd(B, h // 2, h // 2, c)
        rec(f"bn_relu_bwd+pool c{c} @{h}",
            timeit(lambda: ops.bn_relu_bwd(g, gp, idx, y, gamma, scale, shift, mean, rstd, dy, dgm, dbt)), bytes_=11.5 * n_el)
        del y, g, a, dy, pooled, idx, gp

if want("edge"):
    h = S
    x = torch.randn(B, 3, h, h, device="cuda")
    w = torch.randn(64, 3, 3, 3, device="cuda") * 0.2
    y = torch.empty(B, h, h, 64, dtype=BF16, device="cuda")
    st = torch.empty(ops.first_conv_stat_rows(B, h, h) * 128, device="cuda")
    dw = torch.empty_like(w)
    rec("first_fprop 3->64", timeit(lambda: ops.conv3x3_first(x, w, y, st)), bytes_=B * h * h * (128.0 + 12))
    rec("first_wgrad 3->64", timeit(lambda: ops.conv3x3_first_wgrad(x, y, dw)), bytes_=B * h * h * (128.0 + 12))
    col = torch.empty(B, h, h, 64, dtype=BF16, device="cuda")
    w1 = ops.prep_first_weight(w)
    st2 = torch.empty(ops.conv1x1_c64_stat_rows(B, h, h, 64) * 128, device="cuda")
    rec("first_im2col 3->64col", timeit(lambda: ops.first_im2col(x, col)), bytes_=B * h * h * (128.0 + 12))
    rec("first_gemm 64col->64 (+stats)", timeit(lambda: ops.conv1x1_c64(col, w1, y, st2)), bytes_=B * h * h * 256.0)
    rec("first_wgrad_tc", timeit(lambda: ops.conv1x1_c64_wgrad(col, y, dw)), bytes_=B * h * h * 256.0)
    hw_ = torch.randn(2, 64, 1, 1, device="cuda")
    hb = torch.zeros(2, device="cuda")
    z = torch.empty(B, 2, h, h, device="cuda")
    rec("head_fprop 64->2", timeit(lambda: ops.head_fprop(y, hw_, hb, z)), bytes_=B * h * h * (128.0 + 8))
    da = torch.empty_like(y)
    dwh, dbh = torch.empty_like(hw_), torch.empty_like(hb)
    rec("head_bwd", timeit(lambda: ops.head_bwd(z, y, hw_, da, dwh, dbh)), bytes_=B * h * h * (256.0 + 8))
    t = torch.randint(0, 2, (B, h, h), device="cuda").float()
    rec("loss fwd", timeit(lambda: ops.loss_ce_dice_fwd(z, t, 0)), bytes_=B * h * h * 12.0)
    out, sums, err = ops.loss_ce_dice_fwd(z, t, 0)
    go = torch.ones(1, device="cuda")
    rec("loss bwd", timeit(lambda: ops.loss_ce_dice_bwd(z, t, sums, go, 0)), bytes_=B * h * h * 20.0)
