import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from unet_torch_b200 import ops
BF16 = torch.bfloat16
torch.manual_seed(0)
for (n, cin, h, w) in [(1, 1, 16, 8), (1, 3, 16, 8), (1, 3, 32, 16)]:
    cout = 64
    x = torch.randn(n, cin, h, w).cuda()
    wt = torch.randn(cout, cin, 3, 3).cuda() * 0.3
    w1 = ops.prep_first_weight(wt)
    col = torch.zeros(n, h, w, 64, dtype=BF16, device="cuda")
    ops.first_im2col(x, col)
    y = torch.empty(n, h, w, cout, dtype=BF16, device="cuda")
    ops.conv1x1_c64(col, w1, y, None)
    y2 = torch.empty(n, h, w, cout, dtype=BF16, device="cuda")
    ops.conv3x3_first_tc(x, w1, y2, None)
    torch.cuda.synchronize()
    d = (y2.float() - y.float()).abs()
    print((n, cin, h, w), "max diff", float(d.max()), "frac mismatched", float((d > 0).float().mean()))
    # is y2 the conv of a permuted / partial im2col? test: y2 vs GEMM of variants
    colf = col.float().view(-1, 64)
    wf = w1.float()
    want = colf @ wf.t()
    print("  ref check", float((y.float().view(-1, 64) - want).abs().max()))
    # per-pixel: which pixel's row did y2 use? find best matching row for pixel 0..3
    for p in (0, 1, 9, 100):
        if p >= colf.shape[0]: continue
        errs = ((want - y2.float().view(-1, 64)[p][None]).abs().sum(1))
        print("  pixel", p, "best matching ref row", int(errs.argmin()), float(errs.min()))
    # try partial-K hypothesis: only first 8/16/24/32 columns used
    for kk in (8, 16, 24, 32):
        wk = colf[:, :kk] @ wf[:, :kk].t()
        print("  K<", kk, float((wk - y2.float().view(-1, 64)).abs().max()))
