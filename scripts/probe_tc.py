"""GPU bring-up probe for the tcgen05 kernels: structured inputs whose expected outputs are exact, with an error
breakdown per tap / channel chunk / pixel row so a descriptor or layout mistake can be located from the log."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_torch_b200  # noqa: E402
from unet_torch_b200 import ops  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

BF16 = torch.bfloat16


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to("cuda", BF16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def report(tag, got, want):
    d = (got.double() - want.double())
    rel = float(d.norm() / (want.double().norm() + 1e-30))
    print(f"[{tag}] rel_l2={rel:.3e} max_abs={float(d.abs().max()):.3e} want_norm={float(want.norm()):.3e} "
          f"got_norm={float(got.norm()):.3e} nan={int(torch.isnan(got).sum())}", flush=True)
    return rel


def breakdown(got, want):
    # got/want NCHW
    d = (got - want).abs()
    n, c, h, w = d.shape
    per_c = d.amax((0, 2, 3)).view(-1, 8).amax(1)
    print("   max err per 8-channel chunk:", [f"{v:.2g}" for v in per_c.tolist()][:32])
    print("   max err per row h:", [f"{v:.2g}" for v in d.amax((0, 1, 3)).tolist()][:32])
    print("   max err per col w:", [f"{v:.2g}" for v in d.amax((0, 1, 2)).tolist()][:32])


def conv_case(n, h, w, cin, cout, wt, tag):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g).to(BF16).float()
    wf, wd = ops.prep_conv3x3_weight(wt.cuda())
    y = torch.zeros(n, h, w, cout, dtype=BF16, device="cuda")
    rows = ops.conv3x3_stat_rows(n, h, w, cin, cout)
    st = torch.zeros(rows * 2 * cout, device="cuda")
    ops.conv3x3(nhwc(x), wf, y, st)
    torch.cuda.synchronize()
    want = O.conv3x3(x, wt)
    got = nchw(y)
    r = report(tag, got, want)
    if r > 1e-2:
        breakdown(got, want)
    return r


def main():
    print("device:", torch.cuda.get_device_name(0), "lib version", unet_torch_b200._lib.query("b200unet_version"), flush=True)
    ok = True
    try:
        # 1. centre tap identity: y == x
        for cin, cout in ((64, 64), (128, 128), (256, 256)):
            wt = torch.zeros(cout, cin, 3, 3)
            for k in range(min(cin, cout)):
                wt[k, k, 1, 1] = 1.0
            ok &= conv_case(1, 8, 16, cin, cout, wt, f"identity c{cin}->{cout} 1 tile") < 1e-6
        ok &= conv_case(2, 16, 32, 64, 64, wt[:64, :64].clone(), "identity multi-tile") < 1e-6
        # 2. single shifted taps
        for r in range(3):
            for s in range(3):
                wt = torch.zeros(64, 64, 3, 3)
                for k in range(64):
                    wt[k, k, r, s] = 1.0
                ok &= conv_case(1, 16, 32, 64, 64, wt, f"shift tap r={r} s={s}") < 1e-6
        # 3. channel permutation through the centre tap (checks K-chunk / swizzle addressing)
        wt = torch.zeros(64, 64, 3, 3)
        perm = torch.randperm(64, generator=torch.Generator().manual_seed(3))
        for k in range(64):
            wt[k, perm[k], 1, 1] = 1.0
        ok &= conv_case(1, 8, 16, 64, 64, wt, "channel permutation") < 1e-6
        # 4. random
        g = torch.Generator().manual_seed(2)
        for (n, h, w, cin, cout) in ((1, 8, 16, 64, 64), (2, 16, 32, 128, 128), (1, 24, 40, 256, 256), (1, 16, 16, 128, 64)):
            wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).to(BF16).float()
            ok &= conv_case(n, h, w, cin, cout, wt, f"random n{n} {h}x{w} c{cin}->{cout}") < 1e-2
    except Exception:
        traceback.print_exc()
        ok = False
    # 5. wgrad
    try:
        g = torch.Generator().manual_seed(5)
        for (n, h, w, cin, cout) in ((1, 8, 16, 64, 64), (1, 8, 16, 128, 128), (2, 16, 32, 128, 256), (1, 24, 40, 64, 128)):
            x = torch.randn(n, cin, h, w, generator=g).to(BF16).float()
            dy = torch.randn(n, cout, h, w, generator=g).to(BF16).float()
            wv = torch.zeros(cout, cin, 3, 3, requires_grad=True)
            (O.conv3x3(x, wv) * dy).sum().backward()
            dw = torch.zeros(cout, cin, 3, 3, device="cuda")
            ops.conv3x3_wgrad(nhwc(x), nhwc(dy), dw)
            torch.cuda.synchronize()
            r = report(f"wgrad n{n} {h}x{w} c{cin}->{cout}", dw.cpu(), wv.grad)
            if r > 1e-3:
                d = (dw.cpu() - wv.grad).abs()
                print("   max err per tap:", [f"{v:.2g}" for v in d.amax((0, 1)).flatten().tolist()])
                print("   max err per cout chunk(8):", [f"{v:.2g}" for v in d.amax((1, 2, 3)).view(-1, 8).amax(1).tolist()][:32])
                print("   max err per cin chunk(8):", [f"{v:.2g}" for v in d.amax((0, 2, 3)).view(-1, 8).amax(1).tolist()][:32])
                ok = False
    except Exception:
        traceback.print_exc()
        ok = False
    # 6. conv transpose
    try:
        g = torch.Generator().manual_seed(6)
        for (n, h, w, cin, cup) in ((1, 8, 16, 128, 64), (2, 8, 16, 256, 128)):
            xv = torch.randn(n, cin, h, w, generator=g).to(BF16).float().requires_grad_(True)
            wv = (torch.randn(cin, cup, 2, 2, generator=g) * (1.0 / cin) ** 0.5).to(BF16).float().requires_grad_(True)
            b = torch.randn(cup, generator=g)
            want = O.conv_transpose2x2(xv, wv, b)
            du = torch.randn(want.shape, generator=g).to(BF16).float()
            (want * du).sum().backward()
            wf, wd = ops.prep_convt2x2_weight(wv.detach().cuda())
            cat = torch.zeros(n, 2 * h, 2 * w, 2 * cup, dtype=BF16, device="cuda")
            ops.convt2x2(nhwc(xv.detach()), wf, b.cuda(), cat[..., cup:])
            torch.cuda.synchronize()
            ok &= report(f"convT fprop c{cin}->{cup}", nchw(cat[..., cup:]), want.detach()) < 1e-2
            dcat = torch.zeros(n, 2 * h, 2 * w, 2 * cup, dtype=BF16, device="cuda")
            dcat[..., cup:] = nhwc(du)
            dx = torch.zeros(n, h, w, cin, dtype=BF16, device="cuda")
            ops.convt2x2_dgrad(dcat[..., cup:], wd, dx)
            torch.cuda.synchronize()
            ok &= report(f"convT dgrad c{cin}->{cup}", nchw(dx), xv.grad) < 1e-2
            dw = torch.zeros(cin, cup, 2, 2, device="cuda")
            ops.convt2x2_wgrad(nhwc(xv.detach()), dcat[..., cup:], dw)
            torch.cuda.synchronize()
            ok &= report(f"convT wgrad c{cin}->{cup}", dw.cpu(), wv.grad) < 1e-3
    except Exception:
        traceback.print_exc()
        ok = False
    print("PROBE", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
