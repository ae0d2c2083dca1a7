"""Helpers for the -m gpu parity tests: layout conversion between the reference's NCHW fp32 world and the
kernels' NHWC bf16 world, and error metrics."""
import torch

BF16 = torch.bfloat16


def to_nhwc_bf16(x_nchw, device="cuda"):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(device=device, dtype=BF16)


def from_nhwc(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()


def bf16_round(x):
    return x.to(BF16).float()


def rel_l2(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return float((got - want).norm() / (want.norm() + 1e-30))


def max_abs(got, want):
    return float((got.double().cpu() - want.double().cpu()).abs().max())
