"""Does the tensor core apply the 128B swizzle on absolute smem address bits (so a K-major operand may start at any
128-byte row of a swizzled tile)? Prints the error for every row shift, with and without the descriptor
base_offset field."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_torch_b200 import _lib  # noqa: E402

g = torch.Generator().manual_seed(0)
a = torch.randn(160, 64, generator=g).to(torch.bfloat16).cuda()
b = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
for ubo in (0, 1):
    for shift in range(0, 33):
        out = torch.zeros(128, 64, device="cuda")
        _lib.call("b200unet_probe_shift", a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, ubo,
                  torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = a[shift:shift + 128].float() @ b.float().t()
        err = float((out - want).abs().max())
        print(f"base_offset={ubo} shift={shift:2d} max_abs_err={err:.3e} {'OK' if err < 1e-3 else 'MISMATCH'}")
