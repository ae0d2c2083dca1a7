"""Fused optimizer for the B200 UNet.

The reference builds `torch.optim.SGD(model.parameters(), lr, momentum, weight_decay)` (train.py:341-347) and steps
it once per iteration (Trainer.py:719-725). `FusedSGD` is a drop-in for that object (same constructor arguments,
`param_groups` / `state_dict()` layout with `momentum_buffer`, so Trainer's poly-LR writes to
`param_group['lr']` keep working) whose `step()` runs the identical arithmetic in ONE pass per conv / convT weight
and writes the bf16 GEMM operands of the tensor-core kernels in the same pass; all small tensors share one launch.
Stock `torch.optim` optimizers keep working with the UNet (operands are then re-cast lazily on the next forward).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, net, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        from .model import UNet

        if not isinstance(net, UNet):
            raise TypeError("FusedSGD takes the B200 UNet module itself (it updates the kernels' bf16 operands in place)")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self.net = net
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(list(net.parameters()), defaults)

    def _plan(self):
        """(big, small): big = [(param, holder, kind)] for tensor-core operands, small = every other parameter."""
        big = {}
        if not self.net._fast_supported() or self.net._check_fp32:
            return big  # generic fp32 engine: no bf16 operands to maintain, every tensor takes the plain update
        eng = self.net._get_engine()
        for c1, c2 in eng.enc + eng.dec:
            for cb in (c1, c2):
                if not cb.first:
                    big[cb.conv.weight] = (cb, "conv3")
        for u in eng.ups:
            big[u.up.weight] = (u, "convt")
        return big

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        big = self._plan()
        stream = torch.cuda.current_stream().cuda_stream
        for group in self.param_groups:
            lr, mom, damp = float(group["lr"]), float(group["momentum"]), float(group["dampening"])
            wd, nest = float(group["weight_decay"]), int(bool(group["nesterov"]))
            small = {True: [], False: []}  # keyed by first_step
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedSGD: parameters must be contiguous CUDA fp32 tensors")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                st = self.state[p]
                first = 0
                buf_ptr = None
                if mom != 0.0:
                    if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                        st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                        first = 1
                    buf_ptr = st["momentum_buffer"].data_ptr()
                if p in big:
                    holder, kind = big[p]
                    wf, wd_op = holder.operands()  # allocates the operand tensors on first use
                    if kind == "conv3":
                        k, c = p.shape[0], p.shape[1]
                        _lib.call("b200unet_sgd_conv3x3_weight", p.data_ptr(), g.data_ptr(), buf_ptr, wf.data_ptr(),
                                  wd_op.data_ptr(), k, c, lr, mom, damp, wd, nest, first, stream)
                    else:
                        cin, cup = p.shape[0], p.shape[1]
                        _lib.call("b200unet_sgd_convt2x2_weight", p.data_ptr(), g.data_ptr(), buf_ptr, wf.data_ptr(),
                                  wd_op.data_ptr(), cin, cup, lr, mom, damp, wd, nest, first, stream)
                    torch.autograd.graph.increment_version(p)
                    holder._ver = (p._version, p.data_ptr())  # operands are current: no lazy re-cast
                else:
                    small[bool(first)].append((p, g, buf_ptr))
            for first, items in small.items():
                if not items:
                    continue
                n = len(items)
                PtrArr, IntArr = ctypes.c_void_p * n, ctypes.c_int * n
                w_arr = PtrArr(*[p.data_ptr() for p, _, _ in items])
                g_arr = PtrArr(*[g.data_ptr() for _, g, _ in items])
                b_arr = PtrArr(*[b for _, _, b in items]) if mom != 0.0 else None
                n_arr = IntArr(*[p.numel() for p, _, _ in items])
                _lib.call("b200unet_sgd_small", w_arr, g_arr, b_arr, n_arr, n, lr, mom, damp, wd, nest, int(first), stream)
                for p, _, _ in items:
                    torch.autograd.graph.increment_version(p)
        return loss
