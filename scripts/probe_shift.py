"""Does the tensor core apply the 128B swizzle on absolute smem address bits, so that a K-major operand may start at
any 128-byte row of a swizzled tile and its 8-row groups may be `sbo` bytes apart with sbo not a multiple of 1024?
Prints the error for row shifts x {sbo = 1024, 1280 (conv3_res.cu: 10-pixel pitch of the halo tile)}."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_torch_b200 import _lib  # noqa: E402

g = torch.Generator().manual_seed(0)
a = torch.randn(256, 64, generator=g).to(torch.bfloat16).cuda()
b = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
bad = 0
for sbo in (1024, 1280, 2304):
    for shift in (0, 1, 2, 3, 7, 8, 10, 11, 12, 20, 21, 22):
        out = torch.zeros(128, 64, device="cuda")
        _lib.call("b200unet_probe_shift", a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, 0, sbo,
                  torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        rows = torch.tensor([shift + (m // 8) * (sbo // 128) + (m % 8) for m in range(128)])
        if int(rows.max()) >= 256:
            continue
        want = a[rows.cuda()].float() @ b.float().t()
        err = float((out - want).abs().max())
        ok = err < 1e-3
        bad += not ok
        print(f"sbo={sbo} shift={shift:2d} max_abs_err={err:.3e} {'OK' if ok else 'MISMATCH'}")
print("PROBE", "PASS" if bad == 0 else f"FAIL ({bad})")
