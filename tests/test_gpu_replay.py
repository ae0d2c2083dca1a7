"""-m gpu: operator-by-operator REPLAY of one whole training step of the tensor-core engine (composition check).

A bf16-storage network is chaotic at depth: two correct implementations whose fp32 sums are ordered differently decorrelate
within a few layers (measured: CUDA vs a bit-faithful bf16-storage emulation differ by 1e-2 in the logits and ~18 % in
individual gradients - tests/test_gpu_fullsize.py), so whole-network comparisons cannot be tight. This test is: the engine
records every operator's inputs and outputs (`UNetEngine.trace`), and the host
  1. recomputes EACH operator from the engine's OWN inputs with the torch CPU ops the reference dispatches to
     (Model.py:15-22, 36, 56-57, 79, 89) - one operator deep, so the tolerance is the bf16 output rounding (3e-3) or fp32
     summation (1e-4 ... 1e-3) - forward and backward, including BatchNorm statistics / running buffers, pooling positions,
     the skip + unpool gradient merge, dgamma / dbeta, weight gradients and the convT bias gradient;
  2. checks the WIRING: every operator's input is the very tensor (same storage) an upstream operator produced - skip into
     the concat buffer, upsampled half next to it, pooled tensor into the next level, every gradient edge of the backward
     graph - and every parameter's .grad is the tensor the trace produced for it.
Together: the step is the reference's computation graph, executed correctly operator by operator."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import BF16, rel_l2

pytestmark = pytest.mark.gpu


def nchw(t):  # NHWC bf16 device tensor (possibly a channel slice) -> NCHW fp32 on the host
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def rb(t):  # bf16 storage rounding
    return t.to(BF16).float()


def same_storage(a, b):
    return a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride()


@pytest.mark.parametrize("cls_name,n,h,w", [("UNet", 2, 64, 48), ("UNet", 1, 32, 80), ("UNet_multitask", 2, 32, 32)])
def test_replay_every_operator_and_edge_of_a_training_step(cls_name, n, h, w):
    import unet_torch_b200 as U

    torch.manual_seed(5)
    net = getattr(U, cls_name)(3, 2).cuda().train()
    names = {id(m): k for k, m in net.named_modules()}
    pname = {id(p): k for k, p in net.named_parameters()}
    bufs0 = {k: v.clone().cpu() for k, v in net.state_dict().items() if "running" in k}
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(n, 3, h, w, generator=gen).cuda()
    eng = net._get_engine()
    eng.trace = []
    U.loss.CLASS_NUMBER = 2
    if cls_name == "UNet":
        y = torch.randint(0, 2, (n, h, w), generator=gen).float().cuda()
        out = net(x)
        loss = U.calc_loss(out, y, loss_type="dice_bce_mc")
    else:
        t1, t2 = torch.rand(n, 2, h, w, generator=gen).cuda(), torch.rand(n, 2, h, w, generator=gen).cuda()
        o1, o2 = net(x)
        loss = U.calc_loss(torch.relu(o1), t1, loss_type="mseMC") + U.calc_loss(torch.relu(o2), t2, loss_type="mseMC")
    loss.backward()
    torch.cuda.synchronize()
    trace, eng.trace = eng.trace, None
    fwd = {}     # module name of the conv -> forward record
    produced = []  # forward tensors in production order, for the wiring checks
    n_ops = 0

    # ------------------------------------------------------------------ forward operators
    for kind, r in trace:
        if kind == "conv_bn_relu":
            cb = r["cb"]
            name = names[id(cb.conv)]
            wt = cb.conv.weight.detach().cpu()
            xin = rb(r["x"].cpu()) if cb.first else nchw(r["x"])
            y_cuda = nchw(r["y"])
            assert rel_l2(y_cuda, F.conv2d(xin, rb(wt), None, padding=1)) < 3e-3, name            # Model.py:15-16,19-20
            m = y_cuda.numel() // y_cuda.shape[1]
            assert r["count"] == m
            mean, var = y_cuda.double().mean((0, 2, 3)), y_cuda.double().var((0, 2, 3), unbiased=False)
            rstd = 1.0 / torch.sqrt(var + cb.bn.eps)
            assert torch.allclose(r["mean"].double().cpu(), mean, rtol=1e-4, atol=1e-6), name      # Model.py:17,21
            assert torch.allclose(r["rstd"].double().cpu(), rstd, rtol=1e-4), name
            g_, b_ = cb.bn.weight.detach().double().cpu(), cb.bn.bias.detach().double().cpu()
            assert torch.allclose(r["scale"].double().cpu(), g_ * rstd, rtol=1e-4, atol=1e-7), name
            assert torch.allclose(r["shift"].double().cpu(), b_ - mean * g_ * rstd, rtol=1e-4, atol=1e-5), name
            bn_name = names[id(cb.bn)]
            sd = net.state_dict()
            assert torch.allclose(sd[bn_name + ".running_mean"].cpu().double(), 0.9 * bufs0[bn_name + ".running_mean"].double() + 0.1 * mean,
                                  rtol=1e-4, atol=1e-6), name
            assert torch.allclose(sd[bn_name + ".running_var"].cpu().double(),
                                  0.9 * bufs0[bn_name + ".running_var"].double() + 0.1 * var * m / (m - 1), rtol=1e-4), name
            assert int(sd[bn_name + ".num_batches_tracked"]) == 1
            sc, sh = r["scale"].cpu()[None, :, None, None], r["shift"].cpu()[None, :, None, None]
            a_want = rb(torch.relu(sc * y_cuda + sh))                                              # Model.py:18,22
            a_cuda = nchw(r["a"])
            diff = (a_cuda - a_want).abs()
            assert float((diff > 0).float().mean()) < 2e-3 and float((diff / (a_want.abs() + 1e-3)).max()) < 1e-2, name
            if r["pooled"] is not None:                                                            # Model.py:36,42
                pv, pi = F.max_pool2d(a_cuda, 2, return_indices=True)
                assert torch.equal(nchw(r["pooled"]), pv), name
                hh, ww = a_cuda.shape[2:]
                pos = r["pool_idx"].permute(0, 3, 1, 2).cpu().long()                               # window position 2*dh + dw
                hp = torch.arange(hh // 2).view(1, 1, -1, 1) * 2 + pos // 2
                wp = torch.arange(ww // 2).view(1, 1, 1, -1) * 2 + pos % 2
                assert torch.equal(hp * ww + wp, pi), name                                         # bit-exact argmax indices
            fwd[name] = r
            produced.append((name, r))
            n_ops += 1
        elif kind == "convt":
            up = r["up"].up
            name = names[id(up)]
            want = rb(F.conv_transpose2d(nchw(r["x"]), rb(up.weight.detach().cpu()), up.bias.detach().cpu(), stride=2))
            assert rel_l2(nchw(r["out"]), want) < 3e-3, name                                       # Model.py:56-57,66
            c = r["out"].shape[3]
            assert same_storage(r["cat"][..., c:], r["out"]), name
            fwd[name] = r
            n_ops += 1
        elif kind == "head":
            conv = r["conv"]
            want = F.conv2d(nchw(r["x"]), conv.weight.detach().cpu(), conv.bias.detach().cpu())    # Model.py:89
            assert rel_l2(r["logits"].cpu(), want) < 1e-5
            fwd[names[id(conv)]] = r
            n_ops += 1

    # ------------------------------------------------------------------ forward wiring (Model.py:142-153)
    dec_sets = [("", "")] if cls_name == "UNet" else [("_decod1", "_decod1"), ("_decod2", "_decod2")]
    dc = lambda blk: (fwd[blk + ".double_conv.0"], fwd[blk + ".double_conv.3"])  # noqa: E731
    enc_blocks = ["inc"] + [f"down{i}.maxpool_conv.1" for i in range(1, 5)]
    for i, blk in enumerate(enc_blocks):
        c1, c2 = dc(blk)
        assert same_storage(c2["x"], c1["a"]), blk                         # conv1 -> conv2
        if i > 0:
            assert same_storage(c1["x"], dc(enc_blocks[i - 1])[1]["pooled"]), blk   # MaxPool2d of the previous block feeds it
    for sfx, _ in dec_sets:
        cur = dc(enc_blocks[4])[1]["a"]
        for i in range(1, 5):
            upr = fwd[f"up{i}{sfx}.up"]
            c1, c2 = dc(f"up{i}{sfx}.conv")
            assert same_storage(upr["x"], cur), (i, sfx)                   # x1 of Up.forward
            skip_a = dc(enc_blocks[4 - i])[1]["a"]
            cch = skip_a.shape[3]
            assert same_storage(c1["x"], upr["cat"]), (i, sfx)             # conv reads cat([x2, x1]) (Model.py:79)
            assert torch.equal(upr["cat"][..., :cch], skip_a), (i, sfx)    # skip half FIRST ...
            assert same_storage(upr["cat"][..., cch:], upr["out"]), (i, sfx)  # ... upsampled half second
            if sfx in ("", "_decod1"):
                assert skip_a.data_ptr() == upr["cat"].data_ptr()          # zero-copy: the encoder wrote straight into it
            assert same_storage(c2["x"], c1["a"])
            cur = c2["a"]
        assert same_storage(fwd[f"outc{sfx}.conv"]["x"], cur)

    # ------------------------------------------------------------------ backward operators
    bwd = {}
    for kind, r in trace:
        if kind == "head_bwd":
            conv = r["conv"]
            name = names[id(conv)]
            dz, xin = r["dz"].cpu(), nchw(r["x"])
            wt = conv.weight.detach().cpu()
            assert rel_l2(nchw(r["g"]), rb(torch.einsum("nkhw,kc->nchw", dz, wt.view(wt.shape[0], -1)))) < 3e-3
            assert rel_l2(r["dw"].cpu().view(wt.shape[0], -1), torch.einsum("nkhw,nchw->kc", dz, xin)) < 1e-4
            assert torch.allclose(r["db"].cpu(), dz.sum((0, 2, 3)), rtol=1e-4, atol=1e-7)
            assert conv.weight.grad is not None and torch.equal(conv.weight.grad, r["dw"]) and torch.equal(conv.bias.grad, r["db"])
            bwd[name] = r
            n_ops += 1
        elif kind == "conv_bn_relu_bwd":
            cb = r["cb"]
            name = names[id(cb.conv)]
            f = fwd[name]
            y_f = nchw(f["y"])
            g = nchw(r["g1"]) if r["g1"] is not None else torch.zeros_like(y_f)
            if r["g_pool"] is not None:                                    # gradient through MaxPool2d lands on the argmax
                gp = nchw(r["g_pool"])
                pos = r["pool_idx"].permute(0, 3, 1, 2).cpu().long()
                for k in range(4):
                    g[:, :, k // 2::2, k % 2::2] += gp * (pos == k)
            sc, sh = f["scale"].cpu()[None, :, None, None], f["shift"].cpu()[None, :, None, None]
            da = g * ((sc * y_f + sh) > 0)                                 # ReLU mask = output > 0 (Model.py:18,22)
            mean, rstd = f["mean"].cpu()[None, :, None, None], f["rstd"].cpu()[None, :, None, None]
            xhat = (y_f - mean) * rstd
            s1, s2 = da.double().sum((0, 2, 3)), (da * xhat).double().sum((0, 2, 3))
            scale_of = lambda t: float(t.abs().max()) + 1e-12  # noqa: E731
            assert float((r["dbeta"].cpu().double() - s1).abs().max()) < 2e-3 * scale_of(s1) + 1e-6, name
            assert float((r["dgamma"].cpu().double() - s2).abs().max()) < 2e-3 * scale_of(s2) + 1e-6, name
            m = f["count"]
            gam = cb.bn.weight.detach().cpu()[None, :, None, None]
            dy_want = rb(gam * rstd * (da - (s1 / m).float()[None, :, None, None] - xhat * (s2 / m).float()[None, :, None, None]))
            dy_cuda = nchw(r["dy"])
            assert rel_l2(dy_cuda, dy_want) < 4e-3, name
            wt = cb.conv.weight.detach().cpu()
            xin = rb(fwd[name]["x"].cpu()) if cb.first else nchw(fwd[name]["x"])
            assert rel_l2(r["dw"].cpu(), torch.nn.grad.conv2d_weight(xin, wt.shape, dy_cuda, padding=1)) < 1e-3, name
            if "dx" in r:
                assert rel_l2(nchw(r["dx"]), rb(F.conv_transpose2d(dy_cuda, rb(wt), None, padding=1))) < 3e-3, name
            for p_, t_ in ((cb.conv.weight, r["dw"]), (cb.bn.weight, r["dgamma"]), (cb.bn.bias, r["dbeta"])):
                assert torch.equal(p_.grad, t_), pname[id(p_)]
            bwd[name] = r
            n_ops += 1
        elif kind == "convt_bwd":
            up = r["up"].up
            name = names[id(up)]
            du, xin = nchw(r["du"]), nchw(r["x"])
            wt = up.weight.detach().cpu()
            assert rel_l2(nchw(r["dx"]), rb(F.conv2d(du, rb(wt), None, stride=2))) < 3e-3, name   # dgrad of ConvTranspose2d
            want_dw = torch.einsum("nchw,ndhiwj->cdij", xin, du.view(du.shape[0], du.shape[1], du.shape[2] // 2, 2, du.shape[3] // 2, 2))
            assert rel_l2(r["dw"].cpu(), want_dw) < 1e-3, name
            assert torch.allclose(r["db"].cpu(), du.sum((0, 2, 3)), rtol=2e-3, atol=1e-5), name
            assert torch.equal(up.weight.grad, r["dw"]) and torch.equal(up.bias.grad, r["db"])
            bwd[name] = r
            n_ops += 1

    # ------------------------------------------------------------------ backward wiring (the transposed graph)
    if cls_name == "UNet":
        bdc = lambda blk: (bwd[blk + ".double_conv.0"], bwd[blk + ".double_conv.3"])  # noqa: E731
        g = bwd["outc.conv"]["g"]
        for i in (4, 3, 2, 1):
            b1, b2 = bdc(f"up{i}.conv")
            assert same_storage(b2["g1"], g) and b2["g_pool"] is None
            assert same_storage(b1["g1"], b2["dx"])
            ub = bwd[f"up{i}.up"]
            cch = b1["dx"].shape[3] // 2
            assert same_storage(ub["dcat"], b1["dx"]) and same_storage(ub["du"], b1["dx"][..., cch:])
            g = ub["dx"]
        for i in (4, 3, 2, 1, 0):
            b1, b2 = bdc(enc_blocks[i])
            if i == 4:
                assert same_storage(b2["g1"], g) and b2["g_pool"] is None
            else:
                skip_src = bdc(f"up{4 - i}.conv")[0]["dx"]                  # gradient of cat([x2, x1]): x2 half is the skip's
                assert same_storage(b2["g1"], skip_src[..., : skip_src.shape[3] // 2])
                assert same_storage(b2["g_pool"], bdc(enc_blocks[i + 1])[0]["dx"])
                assert same_storage(b2["pool_idx"], dc(enc_blocks[i])[1]["pool_idx"])
            assert same_storage(b1["g1"], b2["dx"])
        assert "dx" not in bdc("inc")[0]                                     # no input gradient for the first conv
    assert all(p.grad is not None for p in net.parameters())
    n_expected = (18 + 4 + 1) * 2 if cls_name == "UNet" else (10 + 2 * (8 + 4 + 1)) * 2
    assert n_ops == n_expected, (n_ops, n_expected)
    print(f"{cls_name} {n}x3x{h}x{w}: replayed {n_ops} operators of one training step, every edge verified")
