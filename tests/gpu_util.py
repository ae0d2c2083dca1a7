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


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


def host_step(sd, x, y, ncls, loss_type, relu=False, emulate=False, use_ref_modules=False, dtype=torch.float32,
              training=True, input_grad=False):
    """One training forward + loss + backward of the reference path on the HOST CPU, from a reference-format state_dict.

    use_ref_modules: the unmodified reference (oracle/_ref: Model.UNet + loss.calc_loss); otherwise the oracle port
    `cpu_baseline.unet_forward_torchops` (the torch CPU ops the reference dispatches to), optionally with the B200 engine's
    bf16 storage points inserted (emulate=True, see oracle/cpu_baseline.py). Returns (logits, loss, {name: grad},
    {name: updated BatchNorm buffer})."""
    import os

    import torch.nn.functional as F

    from oracle import cpu_baseline, ref_loader
    from oracle import unet_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    x, y = x.to(dtype), y.to(dtype)
    if use_ref_modules:
        RefModel, ref_loss = ref_loader.load()
        width = sd["inc.double_conv.0.weight"].shape[0]
        net = RefModel.UNet(x.shape[1], ncls, width)
        net.load_state_dict(sd)
        net = net.to(dtype).train(training)
        ref_loss.CLASS_NUMBER = ncls
        if input_grad:
            x = x.clone().requires_grad_(True)
        out = net(x)
        pred = F.relu(out) if relu else out
        loss = ref_loss.calc_loss(pred, y, loss_type=loss_type)
        loss.backward()
        grads = {k: p.grad for k, p in net.named_parameters()}
        if input_grad:
            grads["__input__"] = x.grad
        bufs = {k: v.clone() for k, v in net.state_dict().items() if "running" in k}
        return out.detach(), float(loss.detach()), grads, bufs
    p = {k: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
             else (v.to(dtype).clone() if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    if input_grad:
        x = x.clone().requires_grad_(True)
    out = cpu_baseline.unet_forward_torchops(p, x, training, emulate_bf16=emulate)
    pred = F.relu(out) if relu else out
    if loss_type == "mseMC":
        loss = F.mse_loss(pred, y)
    elif loss_type == "mse":
        loss = F.mse_loss(pred.squeeze(1), y)
    elif loss_type == "CE":
        loss = F.cross_entropy(pred, y.long())
    else:
        loss = 0.5 * F.cross_entropy(pred, y.long()) + 0.5 * O.dice_softmax(pred, y, ncls)
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if v.requires_grad}
    if input_grad:
        grads["__input__"] = x.grad
    bufs = {k: v for k, v in p.items() if "running" in k}  # F.batch_norm updated them in place
    return out.detach(), float(loss.detach()), grads, bufs
