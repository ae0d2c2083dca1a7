"""ORACLE (test infrastructure, not product code): CPU restatement of the hot path of caki35/UNet-Torch.

The reference's arithmetic lives in PyTorch library ops (third-party, unpinned by the reference: no requirements
file; this image has torch 2.11.0). Each function below restates, as explicit tensor algebra on plain torch CPU
tensors (fp32 or fp64), what the reference call site computes, citing reference file:line. The restatement
deliberately avoids nn.Module / cuDNN dispatch: convolutions are written as unfold + matmul, BatchNorm, pooling,
the transposed convolution and the losses as their closed forms.

Pinning: the reference has NO tests, golden vectors or fixtures for this path, so the
oracle is pinned against outputs of the reference itself: oracle/make_golden.py imports /root/reference/Model.py
and loss.py in the build container, runs them on seeded inputs and commits the results under tests/golden/;
tests/test_oracle.py checks this file against those vectors (and, when /root/reference is present, live).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS_BN = 1e-5
MOMENTUM = 0.1


# ------------------------------------------------------------------------------------------------ operators
def conv3x3(x, w):
    """nn.Conv2d(k=3, padding=1, bias=False) (Model.py:15-16,19-20): cross-correlation with zero padding.
    x [N,C,H,W], w [K,C,3,3] -> [N,K,H,W]."""
    n, c, h, wd = x.shape
    cols = F.unfold(x, kernel_size=3, padding=1)            # [N, C*9, H*W], ordered (c, r, s)
    y = w.reshape(w.shape[0], c * 9) @ cols                # [N, K, H*W]
    return y.reshape(n, w.shape[0], h, wd)


def batchnorm_train(y, gamma, beta, running_mean=None, running_var=None, count_scale=1):
    """nn.BatchNorm2d in training mode (Model.py:17,21): biased variance for normalisation, unbiased for the
    running estimate, momentum 0.1, eps 1e-5. Returns (out, new_running_mean, new_running_var)."""
    m = y.shape[0] * y.shape[2] * y.shape[3] * count_scale
    mu = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False)
    out = (y - mu[None, :, None, None]) * torch.rsqrt(var + EPS_BN)[None, :, None, None]
    out = out * gamma[None, :, None, None] + beta[None, :, None, None]
    nrm = nrv = None
    if running_mean is not None:
        nrm = (1 - MOMENTUM) * running_mean + MOMENTUM * mu
        nrv = (1 - MOMENTUM) * running_var + MOMENTUM * var * m / (m - 1)
    return out, nrm, nrv


def batchnorm_eval(y, gamma, beta, running_mean, running_var):
    s = gamma * torch.rsqrt(running_var + EPS_BN)
    return y * s[None, :, None, None] + (beta - running_mean * s)[None, :, None, None]


def relu(x):
    """nn.ReLU (Model.py:18,22)."""
    return torch.clamp_min(x, 0)


def maxpool2x2(x):
    """nn.MaxPool2d(2) (Model.py:36,42): floor mode, first maximum in row-major window order.
    Returns (pooled, window position 0..3 = 2*dh+dw, flat int64 index h*W+w as torch reports it)."""
    n, c, h, w = x.shape
    hp, wp = h // 2, w // 2
    win = x[:, :, : 2 * hp, : 2 * wp].reshape(n, c, hp, 2, wp, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, hp, wp, 4)
    best = win[..., 0].clone()
    pos = torch.zeros_like(best, dtype=torch.int64)
    for k in range(1, 4):
        v = win[..., k]
        take = (v > best) | torch.isnan(v)
        best = torch.where(take, v, best)
        pos = torch.where(take, torch.full_like(pos, k), pos)
    hh = torch.arange(hp).view(1, 1, hp, 1) * 2 + pos // 2
    ww = torch.arange(wp).view(1, 1, 1, wp) * 2 + pos % 2
    return best, pos, hh * w + ww


def conv_transpose2x2(x, w, b):
    """nn.ConvTranspose2d(C, C/2, kernel_size=2, stride=2) (Model.py:56-57): non-overlapping upsample.
    x [N,C,H,W], w [C,D,2,2], b [D] -> [N,D,2H,2W]."""
    n, c, h, wd = x.shape
    out = torch.einsum("nchw,cdij->ndhiwj", x, w).reshape(n, w.shape[1], 2 * h, 2 * wd)
    return out + b[None, :, None, None]


def pad_and_cat(skip, up):
    """F.pad + torch.cat([x2, x1], dim=1) (Model.py:69-79): centre pad, skip channels first."""
    dy, dx = skip.shape[2] - up.shape[2], skip.shape[3] - up.shape[3]
    up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return torch.cat([skip, up], dim=1)


def conv1x1(x, w, b):
    """OutConv (Model.py:86-92)."""
    return torch.einsum("nchw,kc->nkhw", x, w.reshape(w.shape[0], -1)) + b[None, :, None, None]


# ------------------------------------------------------------------------------------------------ network
def _double_conv(sd, prefix, x, training, new_buffers):
    for ci, bi in ((0, 1), (3, 4)):
        y = conv3x3(x, sd[f"{prefix}.{ci}.weight"])
        g, b = sd[f"{prefix}.{bi}.weight"], sd[f"{prefix}.{bi}.bias"]
        rm, rv = sd[f"{prefix}.{bi}.running_mean"], sd[f"{prefix}.{bi}.running_var"]
        if training:
            y, nrm, nrv = batchnorm_train(y, g, b, rm, rv)
            new_buffers[f"{prefix}.{bi}.running_mean"] = nrm.detach()
            new_buffers[f"{prefix}.{bi}.running_var"] = nrv.detach()
            new_buffers[f"{prefix}.{bi}.num_batches_tracked"] = sd[f"{prefix}.{bi}.num_batches_tracked"] + 1
        else:
            y = batchnorm_eval(y, g, b, rm, rv)
        x = relu(y)
    return x


def unet_forward(sd, x, training=True, dropout_masks=None):
    """UNet.forward (Model.py:142-153) as a pure function of a reference-format state_dict.
    dropout_masks: None (dropout=False, or eval), or {"down1".."down4", "up1".."up4": multiplier tensors} applied where
    the reference applies nn.Dropout: after the pooling of Down (Model.py:34-39) and after the concat of Up (:81-82).
    Returns (logits, dict of updated BatchNorm buffers)."""
    nb = {}
    dci = 2 if "down1.maxpool_conv.2.double_conv.0.weight" in sd else 1  # dropout=True shifts the DoubleConv index
    x1 = _double_conv(sd, "inc.double_conv", x, training, nb)
    skips = [x1]
    cur = x1
    for i in range(1, 5):
        cur, _, _ = maxpool2x2(cur)
        if dropout_masks is not None:
            cur = cur * dropout_masks[f"down{i}"]
        cur = _double_conv(sd, f"down{i}.maxpool_conv.{dci}.double_conv", cur, training, nb)
        skips.append(cur)
    for i in range(1, 5):
        up = conv_transpose2x2(cur, sd[f"up{i}.up.weight"], sd[f"up{i}.up.bias"])
        cat = pad_and_cat(skips[4 - i], up)
        if dropout_masks is not None:
            cat = cat * dropout_masks[f"up{i}"]
        cur = _double_conv(sd, f"up{i}.conv.double_conv", cat, training, nb)
    return conv1x1(cur, sd["outc.conv.weight"], sd["outc.conv.bias"]), nb


def unet_multitask_forward(sd, x, training=True):
    """UNet_multitask.forward (Model.py:232-250): one encoder, two decoders (`up{i}_decod{d}`, `outc_decod{d}`) over the
    same skips; returns ((logits_decod1, logits_decod2), updated BatchNorm buffers)."""
    nb = {}
    x1 = _double_conv(sd, "inc.double_conv", x, training, nb)
    skips = [x1]
    cur = x1
    for i in range(1, 5):
        cur, _, _ = maxpool2x2(cur)
        cur = _double_conv(sd, f"down{i}.maxpool_conv.1.double_conv", cur, training, nb)
        skips.append(cur)
    outs = []
    for d in (1, 2):
        cur = skips[4]
        for i in range(1, 5):
            up = conv_transpose2x2(cur, sd[f"up{i}_decod{d}.up.weight"], sd[f"up{i}_decod{d}.up.bias"])
            cur = _double_conv(sd, f"up{i}_decod{d}.conv.double_conv", pad_and_cat(skips[4 - i], up), training, nb)
        outs.append(conv1x1(cur, sd[f"outc_decod{d}.conv.weight"], sd[f"outc_decod{d}.conv.bias"]))
    return tuple(outs), nb


def attention_block(sd, prefix, q, x, training, new_buffers):
    """Attention_block.forward(q, x) (Model.py:286-296): q = up(q) (ConvTranspose2d C_q -> C_q), Q1 = BN(W_q q),
    X1 = BN(W_x x) (1x1 convolutions WITH bias, BatchNorm without ReLU), E = relu(Q1 + X1), A = sigmoid(BN(psi E)) with one
    channel, returns x * A (A broadcast over the channels)."""
    def conv_bn(name, t):
        y = conv1x1(t, sd[f"{prefix}.{name}.0.weight"], sd[f"{prefix}.{name}.0.bias"])
        g, b = sd[f"{prefix}.{name}.1.weight"], sd[f"{prefix}.{name}.1.bias"]
        rm, rv = sd[f"{prefix}.{name}.1.running_mean"], sd[f"{prefix}.{name}.1.running_var"]
        if training:
            y, nrm, nrv = batchnorm_train(y, g, b, rm, rv)
            new_buffers[f"{prefix}.{name}.1.running_mean"] = nrm.detach()
            new_buffers[f"{prefix}.{name}.1.running_var"] = nrv.detach()
            new_buffers[f"{prefix}.{name}.1.num_batches_tracked"] = sd[f"{prefix}.{name}.1.num_batches_tracked"] + 1
            return y
        return batchnorm_eval(y, g, b, rm, rv)

    q = conv_transpose2x2(q, sd[f"{prefix}.up.weight"], sd[f"{prefix}.up.bias"])
    e = relu(conv_bn("W_q", q) + conv_bn("W_x", x))
    a = torch.sigmoid(conv_bn("psi", e))
    return x * a


def unet_attention_forward(sd, x, training=True):
    """UNet_attention.forward (Model.py:346-367): the UNet with every skip gated, `x_l_att = attenion_l(q=decoder input,
    x=skip)`; BatchNorm modules run in the reference's order (gate of a level before that level's Up block)."""
    nb = {}
    x1 = _double_conv(sd, "inc.double_conv", x, training, nb)
    skips, cur = [x1], x1
    for i in range(1, 5):
        cur, _, _ = maxpool2x2(cur)
        cur = _double_conv(sd, f"down{i}.maxpool_conv.1.double_conv", cur, training, nb)
        skips.append(cur)
    for i in range(1, 5):
        gated = attention_block(sd, f"attenion{5 - i}", cur, skips[4 - i], training, nb)
        up = conv_transpose2x2(cur, sd[f"up{i}.up.weight"], sd[f"up{i}.up.bias"])
        cur = _double_conv(sd, f"up{i}.conv.double_conv", pad_and_cat(gated, up), training, nb)
    return conv1x1(cur, sd["outc.conv.weight"], sd["outc.conv.bias"]), nb


def multitask_uncertainty_loss(loss_values, log_var_tasks, regg_flag):
    """MultitaskUncertaintyLoss.forward (loss.py:313-325): sum_i c_i * L_i + log(std_i), std_i = exp(log_var_i)^(1/2),
    c_i = 1/(2 std_i^2) for regression tasks, 1/std_i^2 otherwise."""
    total = 0
    for l, lv, reg in zip(loss_values, log_var_tasks, regg_flag):
        std = torch.exp(lv) ** 0.5
        total = total + (1 / (2 * std ** 2) if reg else 1 / std ** 2) * l + torch.log(std)
    return total


def dropout_mask_shapes(n, h, w, width=64):
    """Shapes of the tensors nn.Dropout sees, in the order the reference draws its masks (down1-4, then up1-4)."""
    hs, ws = [h], [w]
    for _ in range(4):
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    out = [(f"down{i}", (n, width << (i - 1), hs[i], ws[i])) for i in range(1, 5)]
    out += [(f"up{i}", (n, width << (5 - i), hs[4 - i], ws[4 - i])) for i in range(1, 5)]
    return out


# ------------------------------------------------------------------------------------------------ losses
def softmax_ce(logits, target):
    """nn.CrossEntropyLoss()(pred, target.long()) (loss.py:469,498): mean over N*H*W of -log softmax[target]."""
    lse = torch.logsumexp(logits, dim=1)
    zt = torch.gather(logits, 1, target.long().unsqueeze(1)).squeeze(1)
    return (lse - zt).mean()


def dice_softmax(logits, target, n_classes):
    """DiceLoss(n_classes)(pred, target, softmax=True) (loss.py:238-251): sums over the WHOLE batch per class."""
    p = torch.softmax(logits, dim=1)
    total = 0.0
    for c in range(n_classes):
        t = (target == c).to(p.dtype)
        inter, y, z = (p[:, c] * t).sum(), (t * t).sum(), (p[:, c] * p[:, c]).sum()
        total = total + (1 - (2 * inter + 1e-5) / (z + y + 1e-5))
    return total / n_classes


def calc_loss(pred, target, loss_type, n_classes=None):
    """The hot branches of calc_loss (loss.py:442-516)."""
    if loss_type == "dice_bce_mc":
        return 0.5 * softmax_ce(pred, target) + 0.5 * dice_softmax(pred, target, n_classes or pred.shape[1])
    if loss_type == "CE":
        return softmax_ce(pred, target)
    if loss_type == "mse":
        return ((pred.squeeze(1) - target) ** 2).mean()
    if loss_type == "mseMC":
        return ((pred - target) ** 2).mean()
    raise ValueError(loss_type)


def softmax_argmax(logits):
    """F.softmax(outputs, 1) then torch.argmax(probs, 1) (test_mc3serousv5.py:880-881) in the logits' dtype:
    p_j = exp(z_j - max) / sum_j exp(z_j - max); first maximum wins."""
    m = logits.max(dim=1, keepdim=True).values
    e = torch.exp(logits - m)
    s = e[:, 0:1].clone()
    for j in range(1, logits.shape[1]):
        s = s + e[:, j:j + 1]
    p = e / s
    best = p[:, 0].clone()
    idx = torch.zeros_like(best, dtype=torch.int64)
    for j in range(1, logits.shape[1]):
        take = p[:, j] > best
        best = torch.where(take, p[:, j], best)
        idx = torch.where(take, torch.full_like(idx, j), idx)
    return idx


# ------------------------------------------------------------------------------------------------ edges of the path
def preprocess(img_org):
    """`preprocess` of test_mc3serousv5.py:100-127 (= DataLoader.py:661-671) without the resize branch: per-channel
    z-normalisation in numpy float64 (np.mean / np.std over axes (0,1)), HWC -> CHW, BGR -> RGB, float32, batch dim."""
    import numpy as np

    img = np.asarray(img_org)
    mean3d = np.mean(img, axis=(0, 1))
    std3d = np.std(img, axis=(0, 1))
    out = (img - mean3d) / std3d
    if img.ndim == 2:
        return torch.from_numpy(out.astype(np.float32)).unsqueeze(0).unsqueeze(0)
    out = out.transpose((2, 0, 1))[::-1]
    return torch.from_numpy(np.ascontiguousarray(out.astype(np.float32))).unsqueeze(0)


def preprocess_crop(img_org, crop_size):
    """`preprocessCrop` of test.py:91-126 (image part): pad H and W up to multiples of crop_size, split (p//2, p - p//2),
    constant 255, then the same z-normalisation / CHW / RGB as `preprocess` on the PADDED image."""
    import numpy as np

    img = np.asarray(img_org)
    ph, pw = (-img.shape[0]) % crop_size, (-img.shape[1]) % crop_size
    pads = ((ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2)) + (((0, 0),) if img.ndim == 3 else ())
    return preprocess(np.pad(img, pads, mode="constant", constant_values=255))


def tiled_masks(forward, x, crop_size, decide):
    """The crop loop of test_single_crop (test.py:432-441): `decide(forward(crop))` for every crop_size x crop_size crop
    of the padded input, stitched back; forward(crop) -> logits, decide(logits) -> [1,h,w] mask."""
    pred = torch.zeros(x.shape[2], x.shape[3], dtype=torch.uint8)
    for i in range(0, x.shape[2], crop_size):
        for j in range(0, x.shape[3], crop_size):
            pred[i:i + crop_size, j:j + crop_size] = decide(forward(x[:, :, i:i + crop_size, j:j + crop_size]))[0]
    return pred


def mask_uint8(logits):
    """np.uint8(torch.argmax(F.softmax(outputs, dim=1), dim=1)) (test_mc3serousv5.py:880-887)."""
    return softmax_argmax(logits).to(torch.uint8)


def sigmoid_mask(logits, threshold=0.5):
    """torch.sigmoid(out) then out[0, 0] >= 0.5 -> 1 else 0 (test.py:393-399), for every image of the batch; the sigmoid is
    evaluated in the logits' dtype as 1 / (1 + exp(-z)) like torch does."""
    s = 1 / (1 + torch.exp(-logits[:, 0]))
    return (s >= threshold).to(torch.uint8)


def density_maps(logits, divisor=200.0):
    """F.relu(model(x)) then the float32 maps / 200 (test_mc3serousv5.py:961-974); also their per-map sums (counts)."""
    d = relu(logits) / torch.tensor(divisor, dtype=logits.dtype)
    return d, d.double().sum(dim=(2, 3))
