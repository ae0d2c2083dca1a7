"""Drop-in `calc_loss` / `DiceLoss` for the reference's loss.py:215-251, 442-516.

Hot branches ('dice_bce_mc', 'CE', 'mse', 'mseMC') run as fused sm_100a kernels (csrc/loss.cu) behind
torch.autograd.Functions: one read of logits + labels per pass, no per-class `.item()` host syncs
(reference loss.py:249). The remaining live branches of the reference's string dispatch are outside the hot path
 and are composed from stock torch ops purely so that callers keep working; each of them is
checked against the reference's own outputs (tests/golden/ref_loss_branches.pt).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

CLASS_NUMBER = None  # set by the caller exactly like the reference's module global (train.py:163)


def _as_f32(t):
    return t.detach().contiguous().float()


# Out-of-range labels: nn.CrossEntropyLoss raises inside the reference's calc_loss (loss.py:469, 498). The fused kernel
# records them in a device flag; reading it at once would cost a host sync per step, so the flag of every call is copied to
# pinned host memory asynchronously and examined at the NEXT loss call (or by check_pending_label_errors()): the error
# surfaces one step late instead of never. ce_dice_loss(check_labels=True) checks immediately.
_pending = []


def check_pending_label_errors(wait: bool = False):
    """Raise IndexError if an earlier fused CE/Dice call saw a label outside [0, n_classes)."""
    keep = []
    for host, ev in _pending:
        if wait:
            ev.synchronize()
        if ev.query():
            if int(host.item()) != 0:
                _pending.clear()
                raise IndexError("Target out of bounds for the number of classes (seen in an earlier calc_loss call)")
        else:
            keep.append((host, ev))
    _pending[:] = keep


def _watch(err):
    check_pending_label_errors()
    if torch.cuda.is_current_stream_capturing():
        return
    host = torch.empty((1,), dtype=torch.int32).pin_memory()
    host.copy_(err, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _pending.append((host, ev))
    del _pending[:-8]


class _CEDiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mode):
        if not pred.is_cuda:
            raise RuntimeError("the fused loss runs on CUDA (sm_100a) only; there is no CPU fallback")
        logits = _as_f32(pred)
        tgt = _as_f32(target)
        if logits.dim() != 4 or tgt.shape != (logits.shape[0],) + tuple(logits.shape[2:]):
            raise ValueError(f"predict {tuple(pred.shape)} & target {tuple(target.shape)} shape do not match")
        with torch.cuda.device(logits.device):
            out, sums, err = ops.loss_ce_dice_fwd(logits, tgt, mode)
            _watch(err)
        ctx.save_for_backward(logits, tgt, sums)
        ctx.mode = mode
        ctx.err = err
        return out[0].clone(), out[1:].clone(), err

    @staticmethod
    def backward(ctx, g, _g_parts, _g_err):
        logits, tgt, sums = ctx.saved_tensors
        with torch.cuda.device(logits.device):
            dz = ops.loss_ce_dice_bwd(logits, tgt, sums, g.contiguous().float().reshape(1), ctx.mode)
        return dz, None, None


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, relu_input):
        if not pred.is_cuda:
            raise RuntimeError("the fused loss runs on CUDA (sm_100a) only; there is no CPU fallback")
        p, t = _as_f32(pred), _as_f32(target)
        if p.shape != t.shape:
            raise ValueError(f"mse: pred {tuple(p.shape)} and target {tuple(t.shape)} differ")
        ctx.save_for_backward(p, t)
        ctx.relu_input = relu_input
        ctx.shape = pred.shape
        with torch.cuda.device(p.device):
            return ops.mse_fwd(p, t, relu_input)[0].clone()

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        with torch.cuda.device(p.device):
            d = ops.mse_bwd(p, t, g.contiguous().float().reshape(1), ctx.relu_input)
        return d.view(ctx.shape), None, None


def ce_dice_loss(pred, target, ce_only=False, check_labels=False):
    """0.5*CE + 0.5*softmax-Dice (loss.py:498-500) or CE alone (loss.py:469). Returns a 0-dim tensor."""
    loss, _parts, err = _CEDiceFn.apply(pred, target, 1 if ce_only else 0)
    if check_labels and int(err.item()) != 0:  # the reference raises inside nll_loss; opt-in (it costs a sync)
        raise IndexError("Target out of bounds for the number of classes")
    return loss


def relu_mse_loss(out, target):
    """mean((relu(out) - target)^2): the Trainer's F.relu (Trainer.py:709-710) fused with nn.MSELoss."""
    return _MSEFn.apply(out, target, True)


class DiceLoss(nn.Module):
    """Multi-class soft Dice (loss.py:215-251) on the fused kernel when softmax=True and weight is None."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        if softmax and weight is None and inputs.is_cuda and inputs.shape[1] == self.n_classes:
            _loss, parts, _ = _CEDiceFn.apply(inputs, target, 0)
            # parts = (CE, Dice) of the same pass; Dice's gradient comes from the fused backward at weight 0.5,
            # so rebuild it as 2*(L - 0.5*CE) to keep autograd exact
            return 2.0 * (_loss - 0.5 * _CEDiceFn.apply(inputs, target, 1)[0])
        if softmax:
            inputs = torch.softmax(inputs, dim=1)
        onehot = torch.stack([(target == i) for i in range(self.n_classes)], dim=1).float()
        w = [1.0] * self.n_classes if weight is None else weight
        total = 0.0
        for i in range(self.n_classes):
            p, t = inputs[:, i], onehot[:, i]
            total = total + (1 - (2 * (p * t).sum() + 1e-5) / ((p * p).sum() + (t * t).sum() + 1e-5)) * w[i]
        return total / self.n_classes


def calc_loss(pred, target, bce_weight=0.5, loss_type='mse'):
    """Same signature and string dispatch as the reference (loss.py:442)."""
    if loss_type == 'dice_bce_mc':
        if CLASS_NUMBER is not None and pred.shape[1] != CLASS_NUMBER:
            raise AssertionError(f'predict {tuple(pred.shape)} & CLASS_NUMBER {CLASS_NUMBER} do not match')
        return ce_dice_loss(pred, target)
    if loss_type == 'CE':
        return ce_dice_loss(pred, target, ce_only=True)
    if loss_type == 'mse':
        return _MSEFn.apply(pred.squeeze(1), target, False)
    if loss_type == 'mseMC':
        return _MSEFn.apply(pred, target, False)
    # ---- branches outside the hot path: stock torch ops (not part of the B200 claim)
    if loss_type == 'rmse':
        return torch.sqrt(_MSEFn.apply(pred, target, False))
    if loss_type == 'l1loss':
        return F.l1_loss(pred, target)
    if loss_type == 'BCE':
        return F.binary_cross_entropy_with_logits(pred.squeeze(1), target)
    if loss_type == 'dice_bce':  # loss.py:483-486 with BinaryDiceLoss (loss.py:254-306: per-sample, smooth 1, mean)
        p = pred.squeeze(1)
        bce = F.binary_cross_entropy_with_logits(p, target)
        s = torch.sigmoid(p).contiguous().view(p.shape[0], -1)
        t = target.contiguous().view(target.shape[0], -1).float()
        dice = 1 - (2 * (s * t).sum(1) + 1) / ((s.abs() + t.abs()).sum(1) + 1)
        return 0.5 * bce + 0.5 * dice.mean()
    raise NotImplementedError(
        f"calc_loss(loss_type={loss_type!r}) is outside the B200 hot path (SURVEY.md section 8); "
        "supported: dice_bce_mc, CE, mse, mseMC (fused) and rmse, l1loss, BCE, dice_bce (torch ops)")


class MultitaskUncertaintyLoss(nn.Module):
    """Homoscedastic-uncertainty weighting of per-task losses, the reference's loss.py:309-325 (constructed at
    Trainer.py:1007, called at :1065 on the two fused relu+MSE task losses of `UNet_multitask`):
        total = sum_i coeff_i * loss_i + log(std_i),  std_i = exp(log_var_i) ** 0.5,
        coeff_i = 1 / (2 std_i^2) for a regression task (regg_flag[i]) else 1 / std_i^2.
    `log_var_tasks` are the caller's own leaf tensors (CPU `torch.zeros((1,), requires_grad=True)` in the Trainer, stepped
    by its Adam next to the model parameters); they are moved to the loss' device / dtype here exactly as the reference does,
    so autograd carries their gradients back across the device boundary. A handful of one-element tensor ops between the
    fused loss kernels and backward: scalar glue, not a kernel of the path. Returns a tensor of log_var's shape ([1])."""

    def __init__(self):
        super().__init__()

    def forward(self, loss_values, log_var_tasks, regg_flag):
        total_loss = 0
        for i in range(len(loss_values)):
            dtype, device = loss_values[i].dtype, loss_values[i].device
            stds = (torch.exp(log_var_tasks[i]) ** (1 / 2)).to(device).to(dtype)
            coeff = 1 / (2 * (stds ** 2)) if regg_flag[i] else 1 / (stds ** 2)
            total_loss = total_loss + coeff * loss_values[i] + torch.log(stds)
        return total_loss


def MRAccuracy(pred, target):
    """Mean relative counting error of the reference's loss.py:421-440 (validation metric, Trainer.py:382): sigmoid >= 0.5
    mask per image, 8-connected components counted with OpenCV on the HOST, compared with the number of annotated dots.
    CPU post-processing exactly like the reference; the mask itself comes from the device."""
    import cv2
    import numpy as np

    batch_size = target.shape[0]
    target = target.detach().cpu().numpy()
    pred_bin = (torch.sigmoid(pred.detach().squeeze(1)) >= 0.5).to(torch.uint8).cpu().numpy()
    mre = 0
    for b in range(batch_size):
        count_gt = int(np.sum(target[b]))
        count_pred, _ = cv2.connectedComponents(pred_bin[b], connectivity=8)
        if count_gt != 0:
            mre += abs(count_gt - (count_pred - 1)) / count_gt   # minus the background component
        elif count_pred != 1:
            mre += 1
    return mre / batch_size
