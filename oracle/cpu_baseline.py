"""ORACLE / CPU BASELINE (test and measurement infrastructure, not product code).

The reference's CPU path for the hot path is `Model.UNet(...)(x)` + `loss.calc_loss(..., 'dice_bce_mc')` +
`.backward()` in fp32. /root/reference cannot travel to the GPU box, so this port issues the
SAME torch CPU library ops the reference modules dispatch to (F.conv2d / F.batch_norm / F.relu / F.max_pool2d /
F.conv_transpose2d / F.pad / torch.cat: Model.py:15-22, 36, 56-57, 69-79, 89) in the same order with the same
shapes, so its host-core throughput is the reference's. tests/test_oracle.py checks it against the explicit-algebra
oracle (unet_oracle.py), which is itself pinned to the reference's golden vectors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import unet_oracle as O


class _RoundBF16(torch.autograd.Function):
    """bf16 storage point: the value is rounded to bf16 on the way forward AND its gradient on the way back (the B200
    engine stores activations, pre-BatchNorm conv outputs and their gradients as bf16 tensors)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _st(t, emulate):
    return _RoundBF16.apply(t) if emulate else t


def _w(t, emulate):
    """GEMM weight operand: bf16 copy of the fp32 master parameter (gradient flows to the master unrounded, fp32)."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach() if emulate else t


def _double_conv(p, prefix, x, training=True, emulate=False, fold=False):
    for ci, bi in ((0, 1), (3, 4)):
        x = F.conv2d(_st(x, emulate), _w(p[f"{prefix}.{ci}.weight"], emulate), None, padding=1)
        if not fold:
            x = _st(x, emulate)  # y is stored (bf16) before BatchNorm; eval folds BN + ReLU into the conv epilogue instead
        x = F.batch_norm(x, p[f"{prefix}.{bi}.running_mean"], p[f"{prefix}.{bi}.running_var"], p[f"{prefix}.{bi}.weight"],
                         p[f"{prefix}.{bi}.bias"], training, O.MOMENTUM, O.EPS_BN)
        x = _st(F.relu(x), emulate)
    return x


def unet_forward_torchops(p, x, training=True, emulate_bf16=False):
    """Model.UNet.forward (Model.py:142-153) through the torch CPU ops the reference dispatches to.

    emulate_bf16=True inserts the B200 engine's bf16 STORAGE points (and nothing else) into the same fp32 computation:
    conv / convT inputs and weight operands, the pre-BatchNorm conv output, every activation, and - through
    `_RoundBF16.backward` - the gradients stored at the same places. Arithmetic stays fp32 (the tensor cores accumulate
    bf16 products in fp32), so what is left between this and the CUDA path is summation order: a composition check that is
    two orders of magnitude tighter than comparing a bf16 network with an fp32 one."""
    return _forward(p, x, training, emulate_bf16, [("up{i}.up", "up{i}.conv.double_conv", "outc.conv")])[0]


def unet_multitask_forward_torchops(p, x, training=True, emulate_bf16=False):
    """Model.UNet_multitask.forward (Model.py:232-250): one encoder, two decoders over the same skips -> (logits1, logits2)."""
    return tuple(_forward(p, x, training, emulate_bf16,
                          [(f"up{{i}}_decod{d}.up", f"up{{i}}_decod{d}.conv.double_conv", f"outc_decod{d}.conv") for d in (1, 2)]))


def _forward(p, x, training, e, decoders):
    fold = e and not training
    x1 = _double_conv(p, "inc.double_conv", x, training, e, fold)
    skips, cur = [x1], x1
    for i in range(1, 5):
        cur = _double_conv(p, f"down{i}.maxpool_conv.1.double_conv", F.max_pool2d(cur, 2), training, e, fold)
        skips.append(cur)
    outs = []
    for up_name, dc_name, head_name in decoders:
        cur = skips[4]
        for i in range(1, 5):
            up = F.conv_transpose2d(_st(cur, e), _w(p[up_name.format(i=i) + ".weight"], e), p[up_name.format(i=i) + ".bias"], stride=2)
            cur = _double_conv(p, dc_name.format(i=i), O.pad_and_cat(skips[4 - i], _st(up, e)), training, e, fold)
        outs.append(F.conv2d(_st(cur, e), p[head_name + ".weight"], p[head_name + ".bias"]))
    return outs


def init_state(n_channels, n_classes, width=64, seed=0):
    """Random-init state in the reference's state_dict format (kaiming-normal 3x3 / 1x1 convs, Model.py:167-169)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def dc(prefix, cin, cout):
        for ci, bi, a, b in ((0, 1, cin, cout), (3, 4, cout, cout)):
            sd[f"{prefix}.{ci}.weight"] = torch.randn(b, a, 3, 3, generator=g) * (2.0 / (a * 9)) ** 0.5
            sd[f"{prefix}.{bi}.weight"] = torch.ones(b)
            sd[f"{prefix}.{bi}.bias"] = torch.zeros(b)
            sd[f"{prefix}.{bi}.running_mean"] = torch.zeros(b)
            sd[f"{prefix}.{bi}.running_var"] = torch.ones(b)
            sd[f"{prefix}.{bi}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    dc("inc.double_conv", n_channels, width)
    for i in range(1, 5):
        dc(f"down{i}.maxpool_conv.1.double_conv", width << (i - 1), width << i)
    for i in range(1, 5):
        cin = width << (5 - i)
        bound = (1.0 / (cin // 2 * 4)) ** 0.5
        sd[f"up{i}.up.weight"] = (torch.rand(cin, cin // 2, 2, 2, generator=g) * 2 - 1) * bound
        sd[f"up{i}.up.bias"] = (torch.rand(cin // 2, generator=g) * 2 - 1) * bound
        dc(f"up{i}.conv.double_conv", cin, cin // 2)
    sd["outc.conv.weight"] = torch.randn(n_classes, width, 1, 1, generator=g) * (2.0 / width) ** 0.5
    sd["outc.conv.bias"] = torch.zeros(n_classes)
    return sd


def make_step(n_channels, n_classes, batch, h, w, seed=0, loss_type="dice_bce_mc", relu=False, train=True, sgd=None):
    """Returns step() = one step of the reference path on the host CPU: forward + loss (`dice_bce_mc`, loss.py:488-500, or
    relu + `mseMC`, Trainer.py:709-712 / loss.py:476) + backward (+ the torch.optim.SGD update when `sgd` is given), or,
    with train=False, the eval forward + softmax/argmax mask (test_mc3serousv5.py:878-881)."""
    sd = init_state(n_channels, n_classes, 64, seed)
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(batch, n_channels, h, w, generator=g)
    if loss_type == "mseMC":
        y = torch.rand(batch, n_classes, h, w, generator=g) * 200.0 * (torch.rand(batch, n_classes, h, w, generator=g) > 0.9)
    else:
        y = torch.randint(0, n_classes, (batch, h, w), generator=g).float()
    opt = torch.optim.SGD(list(params.values()), **sgd) if (sgd and train) else None

    def step():
        if not train:
            with torch.no_grad():
                out = unet_forward_torchops(sd, x, False)
                return torch.argmax(F.softmax(out, dim=1), dim=1).to(torch.uint8)
        for p in params.values():
            p.grad = None
        out = unet_forward_torchops(sd, x, True)
        if relu:
            out = F.relu(out)
        if loss_type == "mseMC":
            loss = F.mse_loss(out, y)
        else:
            loss = 0.5 * F.cross_entropy(out, y.long()) + 0.5 * O.dice_softmax(out, y, n_classes)
        loss.backward()
        if opt is not None:
            opt.step()
        return float(loss)

    return step
