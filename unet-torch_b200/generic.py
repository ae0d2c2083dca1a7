"""Generic fp32 engine of the B200 UNet: the same network (reference Model.py:95-153) executed by the plain CUDA-core
kernels of csrc/generic_f32.cu on NCHW fp32 tensors.

Two uses:
  * CHECK MODE (`net.set_check_mode(True)` or B200UNET_CHECK_FP32=1): the reference's own precision, so logits and
    loss agree with the reference to ~1e-6 (north_star asks 1e-4) - separates "is the algorithm right" from bf16 noise.
  * the documented slow-but-correct route for inputs the tensor-core engine does not take: H or W not divisible by
    16 (floor-mode pooling + the F.pad branch, Model.py:69-73), `initial_feature_map` not a multiple of 64, and the
    dropout variants (Model.py:34-39, 81-82). Still CUDA, still this library's kernels: there is no PyTorch fallback.
Dropout masks are drawn with torch's generator (`F.dropout` on a ones tensor of the reference's shape, in the
reference's order), applied and back-propagated by the library's kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, ops
from .dist import DataParallelContext


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _v(t: torch.Tensor):
    """(pointer, batch stride) of an NCHW fp32 view whose (C,H,W) block is dense."""
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 4:
        raise ValueError(f"expected a CUDA fp32 NCHW tensor, got {t.dtype} {tuple(t.shape)}")
    n, c, h, w = t.shape
    if t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w:
        raise ValueError(f"NCHW view is not channel-dense: shape {tuple(t.shape)} strides {t.stride()}")
    return t.data_ptr(), t.stride(0)


def _dc(mod):
    return mod.double_conv


class GenericEngine:
    def __init__(self, net):
        self.net = net
        downs = [net.down1, net.down2, net.down3, net.down4]
        ups = [net.up1, net.up2, net.up3, net.up4]
        enc_dc = [_dc(net.inc)] + [_dc(d.maxpool_conv[-1]) for d in downs]
        self.enc = [((s[0], s[1]), (s[3], s[4])) for s in enc_dc]
        self.ups = [u.up for u in ups]
        self.dec = [((_dc(u.conv)[0], _dc(u.conv)[1]), (_dc(u.conv)[3], _dc(u.conv)[4])) for u in ups]
        self.head = net.outc.conv

    def params_in_backward_order(self):
        out = [self.head.weight, self.head.bias]
        for j in (3, 2, 1, 0):
            (c1, b1), (c2, b2) = self.dec[j]
            out += [b2.weight, b2.bias, c2.weight, b1.weight, b1.bias, c1.weight, self.ups[j].bias, self.ups[j].weight]
        for l in (4, 3, 2, 1, 0):
            (c1, b1), (c2, b2) = self.enc[l]
            out += [b2.weight, b2.bias, c2.weight, b1.weight, b1.bias, c1.weight]
        return out

    def graphed_step(self, x, training, save):
        return None

    def refresh_operands(self):
        pass

    # ------------------------------------------------------------------ pieces
    def _conv_bn_relu(self, conv, bn, inp, a_out, training, dp, save=False):
        n, cin, h, w = inp.shape
        k = conv.weight.shape[0]
        dev = inp.device
        y = torch.empty((n, k, h, w), dtype=torch.float32, device=dev)
        ip, ins = _v(inp)
        _lib.call("b200unet_gen_conv3x3", ip, ins, conv.weight.data_ptr(), y.data_ptr(), y.stride(0), n, cin, k, h, w, 0,
                  _stream())
        scale = torch.empty(k, dtype=torch.float32, device=dev)
        shift = torch.empty(k, dtype=torch.float32, device=dev)
        mean = rstd = None
        count = n * h * w
        if not training:
            ops.bn_eval_affine(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, scale, shift)
            if save:  # eval forward under autograd: the backward kernels need mean / rstd
                mean = torch.empty(k, dtype=torch.float32, device=dev)
                rstd = torch.empty(k, dtype=torch.float32, device=dev)
                ops.bn_eval_stats(bn.running_mean, bn.running_var, bn.eps, mean, rstd)
        else:
            sums = torch.empty(2 * k, dtype=torch.float64, device=dev)
            _lib.call("b200unet_gen_channel_stats", y.data_ptr(), y.stride(0), sums.data_ptr(), n, k, h * w, _stream())
            if dp is not None and dp.sync_bn:
                dp.all_reduce_sum(sums)
                count = count * dp.world_size
            mean = torch.empty(k, dtype=torch.float32, device=dev)
            rstd = torch.empty(k, dtype=torch.float32, device=dev)
            mom = bn.momentum if bn.momentum is not None else 0.1
            track = bn.track_running_stats and bn.running_mean is not None
            ops.bn_finalize(sums, count, bn.weight, bn.bias, bn.eps, mom, bn.running_mean if track else None,
                            bn.running_var if track else None, mean, rstd, scale, shift)
            if track:
                bn.num_batches_tracked += 1
        ap, ans = _v(a_out)
        _lib.call("b200unet_gen_bn_relu_fwd", y.data_ptr(), y.stride(0), scale.data_ptr(), shift.data_ptr(), ap, ans, n, k,
                  h * w, _stream())
        return (inp, y, scale, shift, mean, rstd, count, not training)

    def _dropout(self, t, training):
        """nn.Dropout on `t` (an NCHW view), mask drawn like the reference does; returns the mask or None."""
        net = self.net
        if not (net.dropout and training and net.dropout_p > 0):
            return None
        mask = F.dropout(torch.ones(t.shape, dtype=torch.float32, device=t.device), net.dropout_p, True)
        tp, tns = _v(t)
        _lib.call("b200unet_gen_mul", tp, tns, mask.data_ptr(), t.shape[0], t.shape[1] * t.shape[2] * t.shape[3], _stream())
        return mask

    # ------------------------------------------------------------------ forward
    def forward(self, x, training, save):
        net = self.net
        if x.dim() != 4 or x.shape[1] != net.n_channels:
            raise ValueError(f"UNet expects [B,{net.n_channels},H,W] input, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("the B200 UNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        x = x.contiguous().float()
        n, _, h, w = x.shape
        if h < 16 or w < 16:
            raise ValueError(f"input {h}x{w} is too small for four 2x2 poolings")
        dev = x.device
        dp = DataParallelContext.current() if training else None
        f = net.initial_feature_map
        ch = [f << l for l in range(5)]
        hs, wsz = [h], [w]
        for _ in range(4):
            hs.append(hs[-1] // 2)
            wsz.append(wsz[-1] // 2)
        cat = []
        for l in range(4):
            padded = (2 * hs[l + 1] != hs[l]) or (2 * wsz[l + 1] != wsz[l])
            mk = torch.zeros if padded else torch.empty
            cat.append(mk((n, 2 * ch[l], hs[l], wsz[l]), dtype=torch.float32, device=dev))
        enc_rec, dec_rec = [], []
        inp = x
        for l in range(5):
            (c1, b1), (c2, b2) = self.enc[l]
            a1 = torch.empty((n, ch[l], hs[l], wsz[l]), dtype=torch.float32, device=dev)
            r1 = self._conv_bn_relu(c1, b1, inp, a1, training, dp, save)
            if l < 4:
                a2 = cat[l][:, : ch[l]]
                r2 = self._conv_bn_relu(c2, b2, a1, a2, training, dp, save)
                pooled = torch.empty((n, ch[l], hs[l + 1], wsz[l + 1]), dtype=torch.float32, device=dev)
                idx = torch.empty((n, ch[l], hs[l + 1], wsz[l + 1]), dtype=torch.uint8, device=dev) if save else None
                ap, ans = _v(a2)
                _lib.call("b200unet_gen_maxpool2x2", ap, ans, pooled.data_ptr(), None if idx is None else idx.data_ptr(), n,
                          ch[l], hs[l], wsz[l], _stream())
                pmask = self._dropout(pooled, training)  # Down: MaxPool2d -> Dropout -> DoubleConv (Model.py:34-39)
                enc_rec.append((r1, r2, a2, idx, pmask))
                inp = pooled
            else:
                a2 = torch.empty((n, ch[l], hs[l], wsz[l]), dtype=torch.float32, device=dev)
                r2 = self._conv_bn_relu(c2, b2, a1, a2, training, dp, save)
                enc_rec.append((r1, r2, a2, None, None))
        d_in = enc_rec[4][2]
        for j in range(4):
            l = 3 - j
            up = self.ups[j]
            pt, pl = (hs[l] - 2 * hs[l + 1]) // 2, (wsz[l] - 2 * wsz[l + 1]) // 2  # F.pad (Model.py:69-73)
            dpp, dns = _v(d_in)
            op, ons = _v(cat[l][:, ch[l]:])
            _lib.call("b200unet_gen_convt2x2_fprop", dpp, dns, up.weight.data_ptr(), up.bias.data_ptr(), op, ons, n,
                      ch[l + 1], ch[l], hs[l + 1], wsz[l + 1], hs[l], wsz[l], pt, pl, _stream())
            cmask = self._dropout(cat[l], training)  # Up: cat -> Dropout -> DoubleConv (Model.py:79-83)
            (c1, b1), (c2, b2) = self.dec[j]
            a1 = torch.empty((n, ch[l], hs[l], wsz[l]), dtype=torch.float32, device=dev)
            r1 = self._conv_bn_relu(c1, b1, cat[l], a1, training, dp, save)
            a2 = torch.empty((n, ch[l], hs[l], wsz[l]), dtype=torch.float32, device=dev)
            r2 = self._conv_bn_relu(c2, b2, a1, a2, training, dp, save)
            dec_rec.append((d_in, r1, r2, a2, cmask, (pt, pl)))
            d_in = a2
        logits = torch.empty((n, net.n_classes, h, w), dtype=torch.float32, device=dev)
        hp_, hns = _v(d_in)
        _lib.call("b200unet_gen_conv1x1_fwd", hp_, hns, self.head.weight.data_ptr(), self.head.bias.data_ptr(),
                  logits.data_ptr(), n, ch[0], net.n_classes, h * w, _stream())
        saved = None
        if save:
            saved = dict(enc=enc_rec, dec=dec_rec, head_in=d_in, shapes=(n, ch, hs, wsz), dp=dp)
        return logits, saved

    # ------------------------------------------------------------------ backward
    def backward(self, saved, dlogits):
        n, ch, hs, wsz = saved["shapes"]
        dev = dlogits.device
        dp = saved["dp"]
        grads = {}
        flat = dp.make_flat_grads(self.params_in_backward_order()) if dp is not None else None

        def gbuf(p):
            if flat is not None:
                return flat.view_for(p)
            return torch.empty_like(p, memory_format=torch.contiguous_format)

        def done(*ps):
            if flat is not None:
                flat.mark_ready(ps)

        def bn_conv_bwd(conv, bn, rec, g, need_dx):
            inp, y, scale, shift, mean, rstd, count, frozen = rec
            if mean is None:
                raise RuntimeError("backward of a forward that ran without autograd (no BatchNorm statistics saved)")
            nn_, k, hh, ww = y.shape
            gp, gns = _v(g)
            sums = torch.empty(2 * k, dtype=torch.float64, device=dev)
            _lib.call("b200unet_gen_bn_relu_bwd_reduce", gp, gns, y.data_ptr(), y.stride(0), scale.data_ptr(),
                      shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(), nn_, k, hh * ww, _stream())
            sums_local = None
            if frozen:  # running-statistics BatchNorm: dy = gamma * rstd * da, dgamma / dbeta from the local sums
                sums_local, sums = sums, torch.zeros_like(sums)
            elif dp is not None and dp.sync_bn:
                sums_local = sums.clone()
                dp.all_reduce_sum(sums)
            dgamma, dbeta = gbuf(bn.weight), gbuf(bn.bias)
            _lib.call("b200unet_gen_bn_relu_bwd_apply", gp, gns, y.data_ptr(), y.stride(0), bn.weight.data_ptr(),
                      scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(), float(count),
                      None if sums_local is None else sums_local.data_ptr(), y.data_ptr(), y.stride(0),
                      dgamma.data_ptr(), dbeta.data_ptr(), nn_, k, hh * ww, _stream())
            dy = y  # in place
            dw = gbuf(conv.weight)
            ip, ins = _v(inp)
            cin = inp.shape[1]
            _lib.call("b200unet_gen_conv3x3_wgrad", ip, ins, dy.data_ptr(), dy.stride(0), dw.data_ptr(), nn_, cin, k, hh, ww,
                      _stream())
            grads[bn.weight], grads[bn.bias], grads[conv.weight] = dgamma, dbeta, dw
            done(bn.weight, bn.bias, conv.weight)
            if not need_dx:
                return None
            dx = torch.empty((nn_, cin, hh, ww), dtype=torch.float32, device=dev)
            _lib.call("b200unet_gen_conv3x3", dy.data_ptr(), dy.stride(0), conv.weight.data_ptr(), dx.data_ptr(),
                      dx.stride(0), nn_, k, cin, hh, ww, 1, _stream())
            return dx

        def mul_mask(t, mask):
            if mask is not None:
                tp, tns = _v(t)
                _lib.call("b200unet_gen_mul", tp, tns, mask.data_ptr(), t.shape[0], t.shape[1] * t.shape[2] * t.shape[3],
                          _stream())

        dlogits = dlogits.contiguous().float()
        hw_, hb_ = self.head.weight, self.head.bias
        head_in = saved["head_in"]
        g = torch.empty((n, ch[0], hs[0], wsz[0]), dtype=torch.float32, device=dev)
        dwh, dbh = gbuf(hw_), gbuf(hb_)
        hp_, hns = _v(head_in)
        _lib.call("b200unet_gen_conv1x1_bwd", dlogits.data_ptr(), hp_, hns, hw_.data_ptr(), g.data_ptr(), g.stride(0),
                  dwh.data_ptr(), dbh.data_ptr(), n, ch[0], hw_.shape[0], hs[0] * wsz[0], _stream())
        grads[hw_], grads[hb_] = dwh, dbh
        done(hw_, hb_)
        skip_grads = [None] * 4
        for j in (3, 2, 1, 0):
            l = 3 - j
            d_in, r1, r2, _, cmask, (pt, pl) = saved["dec"][j]
            (c1, b1), (c2, b2) = self.dec[j]
            g = bn_conv_bwd(c2, b2, r2, g, True)
            dcat = bn_conv_bwd(c1, b1, r1, g, True)
            mul_mask(dcat, cmask)
            skip_grads[l] = dcat[:, : ch[l]]
            du = dcat[:, ch[l]:]
            up = self.ups[j]
            db, dwu = gbuf(up.bias), gbuf(up.weight)
            dip, dins = _v(d_in)
            up_, uns = _v(du)
            _lib.call("b200unet_gen_convt2x2_wgrad", dip, dins, up_, uns, dwu.data_ptr(), db.data_ptr(), n, ch[l + 1], ch[l],
                      hs[l + 1], wsz[l + 1], hs[l], wsz[l], pt, pl, _stream())
            grads[up.bias], grads[up.weight] = db, dwu
            done(up.bias, up.weight)
            g = torch.empty((n, ch[l + 1], hs[l + 1], wsz[l + 1]), dtype=torch.float32, device=dev)
            _lib.call("b200unet_gen_convt2x2_dgrad", up_, uns, up.weight.data_ptr(), g.data_ptr(), g.stride(0), n, ch[l + 1],
                      ch[l], hs[l + 1], wsz[l + 1], hs[l], wsz[l], pt, pl, _stream())
        g_pool = None
        for l in (4, 3, 2, 1, 0):
            r1, r2, _, idx, pmask = saved["enc"][l]
            (c1, b1), (c2, b2) = self.enc[l]
            if l == 4:
                g1 = bn_conv_bwd(c2, b2, r2, g, True)
            else:
                gs = skip_grads[l]
                mul_mask(g_pool, pmask)
                sp, sns = _v(gs)
                _lib.call("b200unet_gen_unpool_add", g_pool.data_ptr(), idx.data_ptr(), sp, sns, n, ch[l], hs[l], wsz[l],
                          _stream())
                g1 = bn_conv_bwd(c2, b2, r2, gs, True)
            g_pool = bn_conv_bwd(c1, b1, r1, g1, need_dx=(l > 0))
        if flat is not None:
            flat.finish()
        return grads
