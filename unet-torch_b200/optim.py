"""Fused optimizer for the B200 UNet.

The reference builds `torch.optim.SGD(model.parameters(), lr, momentum, weight_decay)` (train.py:341-347) and steps
it once per iteration (Trainer.py:719-725). `FusedSGD` is a drop-in for that object (same constructor arguments,
`param_groups` / `state_dict()` layout with `momentum_buffer`, so Trainer's poly-LR writes to
`param_group['lr']` keep working) whose `step()` runs the identical arithmetic in ONE pass per conv / convT weight
and writes the bf16 GEMM operands of the tensor-core kernels in the same pass; all small tensors share one launch.
`FusedAdam` is the same for `torch.optim.Adam(model.parameters(), lr, weight_decay)` (train.py:341-343; the optimizer of
configseros.yml:15): identical arithmetic and `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`).
Stock `torch.optim` optimizers keep working with the UNet (operands are then re-cast lazily on the next forward).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _launch_weights(entry, kind, items, hyper, stream):
    """One launch for all conv3x3 (kind "conv3") or ConvTranspose2d (kind "convt") weights: update + both bf16 operands.
    items: (param, grad, buf1_ptr, buf2_ptr or None, w_fprop, w_dgrad, holder)."""
    n = len(items)
    PtrArr, IntArr = ctypes.c_void_p * n, ctypes.c_int * n
    args = [0 if kind == "conv3" else 1,
            PtrArr(*[it[0].data_ptr() for it in items]), PtrArr(*[it[1].data_ptr() for it in items]),
            PtrArr(*[it[2] for it in items]) if items[0][2] is not None else None]
    if entry == "b200unet_adam_weights":
        args.append(PtrArr(*[it[3] for it in items]))
    args += [PtrArr(*[it[4].data_ptr() for it in items]), PtrArr(*[it[5].data_ptr() for it in items]),
             IntArr(*[it[0].shape[0] for it in items]), IntArr(*[it[0].shape[1] for it in items]), n, *hyper, stream]
    _lib.call(entry, *args)
    for it in items:
        p, holder = it[0], it[6]
        torch.autograd.graph.increment_version(p)
        holder._ver = (p._version, p.data_ptr())  # operands are current: no lazy re-cast


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
            bigs = {}                      # (kind, first_step) -> conv / convT weights, ONE launch per key
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
                    bigs.setdefault((kind, first), []).append((p, g, buf_ptr, None, wf, wd_op, holder))
                else:
                    small[bool(first)].append((p, g, buf_ptr))
            for (kind, first), items in bigs.items():
                _launch_weights("b200unet_sgd_weights", kind, items, (lr, mom, damp, wd, nest, int(first)), stream)
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


class FusedAdam(FusedSGD):
    """Drop-in for `torch.optim.Adam(net.parameters(), lr, betas, eps, weight_decay)` (no amsgrad / maximize): one fused pass
    per conv / convT weight that also writes the bf16 GEMM operands, one launch for all small tensors."""

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        from .model import UNet

        if not isinstance(net, UNet):
            raise TypeError("FusedAdam takes the B200 UNet module itself (it updates the kernels' bf16 operands in place)")
        if amsgrad:
            raise NotImplementedError("FusedAdam: amsgrad is not implemented (the reference never sets it, train.py:341-343)")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or lr < 0 or eps < 0 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        self.net = net
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False)
        torch.optim.Optimizer.__init__(self, list(net.parameters()), defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        big = self._plan()
        stream = torch.cuda.current_stream().cuda_stream
        for group in self.param_groups:
            lr, (b1, b2), eps, wd = float(group["lr"]), group["betas"], float(group["eps"]), float(group["weight_decay"])
            small = {}  # (step_size, inv_sqrt_bc2, first) -> items: parameters of one group normally share their step count
            bigs = {}   # (kind, step_size, inv_sqrt_bc2, first) -> conv / convT weights, ONE launch per key
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous CUDA fp32 tensors")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                st = self.state[p]
                first = 0
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)  # host scalar, like torch's default (non-capturable) Adam
                    st["exp_avg"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    first = 1
                st["step"] += 1
                t = float(st["step"])
                step_size = lr / (1.0 - b1 ** t)
                inv_sqrt_bc2 = 1.0 / (1.0 - b2 ** t) ** 0.5
                m_ptr, v_ptr = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                if p in big:
                    holder, kind = big[p]
                    wf, wd_op = holder.operands()
                    bigs.setdefault((kind, step_size, inv_sqrt_bc2, first), []).append((p, g, m_ptr, v_ptr, wf, wd_op, holder))
                else:
                    small.setdefault((step_size, inv_sqrt_bc2, first), []).append((p, g, m_ptr, v_ptr))
            for (kind, step_size, inv_sqrt_bc2, first), items in bigs.items():
                _launch_weights("b200unet_adam_weights", kind, items, (b1, b2, eps, wd, step_size, inv_sqrt_bc2, int(first)), stream)
            for (step_size, inv_sqrt_bc2, first), items in small.items():
                n = len(items)
                PtrArr, IntArr = ctypes.c_void_p * n, ctypes.c_int * n
                _lib.call("b200unet_adam_small", PtrArr(*[p.data_ptr() for p, _, _, _ in items]),
                          PtrArr(*[g.data_ptr() for _, g, _, _ in items]), PtrArr(*[m for _, _, m, _ in items]),
                          PtrArr(*[v for _, _, _, v in items]), IntArr(*[p.numel() for p, _, _, _ in items]), n, b1, b2, eps, wd,
                          step_size, inv_sqrt_bc2, int(first), stream)
                for p, _, _, _ in items:
                    torch.autograd.graph.increment_version(p)
        return loss
