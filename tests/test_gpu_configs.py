"""-m gpu: the other BASELINE.json configs as parity cases (they are not bench lines).

configs[3]  5-class inference on large tiles (test_mc3serousv5.py:878-881): eval-mode forward (running statistics) and
            the fused softmax->argmax mask, against the CPU oracle at a size it finishes in seconds, and at the full
            tile size through properties (determinism, mask == first-maximum of fp32 softmax of the same logits).
configs[4]  regression head, relu + 'mseMC' (Trainer.py:709-712, loss.py:476) at 768^2: loss equals the closed form on
            the device logits, gradients finite, one SGD step lowers the loss.
"""
import statistics

import pytest
import torch

from gpu_util import cos, host_step, rel_l2
from oracle import cpu_baseline
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _trained_buffers(net):
    """Make the running statistics non-trivial (as after training) without running a training loop."""
    g = torch.Generator().manual_seed(99)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def test_config4_eval_forward_and_mask_against_oracle():
    import unet_torch_b200 as U

    torch.manual_seed(35)
    net = U.UNet(3, 5)
    with torch.no_grad():
        _trained_buffers(net)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().eval()
    x = torch.randn(1, 3, 96, 128, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        out = net(x.cuda())
    want, _ = O.unet_forward(sd, x, training=False)
    with torch.no_grad():
        emu = cpu_baseline.unet_forward_torchops(sd, x, False, emulate_bf16=True)  # bf16 storage points only
    e, yard = rel_l2(out, want), rel_l2(emu, want)
    print(f"config4 eval logits rel-L2 vs oracle {e:.3e} (bf16-storage emulation vs oracle {yard:.3e}; CUDA vs emulation "
          f"{rel_l2(out, emu):.3e})")
    assert e <= 1.5 * yard and rel_l2(out, emu) <= yard
    # eval must not touch the buffers
    for k, v in net.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    # mask: bit-exact against the oracle's softmax->argmax on the SAME (device) logits
    mask = U.predict_mask(out)
    assert mask.dtype == torch.int64 and mask.shape == (1, 96, 128)
    assert torch.equal(mask.cpu(), O.softmax_argmax(out.cpu()))
    # and close to the oracle's own mask (differences only where bf16 noise flips a near-tie)
    agree = float((mask.cpu() == O.softmax_argmax(want)).float().mean())
    print(f"config4 mask agreement with the fp32 oracle end to end: {agree:.4f}")
    assert agree > 0.93


def test_config4_full_tile_properties():
    import unet_torch_b200 as U

    torch.manual_seed(1063)
    net = U.UNet(3, 5).cuda().eval()
    x = torch.randn(2, 3, 1024, 1024, device="cuda")
    with torch.no_grad():
        a = net(x)
        b = net(x)
    assert a.shape == (2, 5, 1024, 1024) and torch.isfinite(a).all()
    assert torch.equal(a, b)  # deterministic: no atomics on the path
    mask = U.predict_mask(a)
    p = torch.softmax(a, dim=1)
    assert torch.equal(mask, torch.argmax(p, dim=1))
    assert int(mask.min()) >= 0 and int(mask.max()) < 5
    # images are independent in eval mode: batch of 2 == two batches of 1
    with torch.no_grad():
        a0 = net(x[:1])
    assert torch.equal(a0, a[:1])


def test_config5_regression_mse_768():
    import unet_torch_b200 as U

    torch.manual_seed(0)
    net = U.UNet(3, 2).cuda().train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 768, 768, generator=g).cuda()
    t = (torch.rand(2, 2, 768, 768, generator=g) * 200.0 * (torch.rand(2, 2, 768, 768, generator=g) > 0.9)).cuda()
    losses = []
    for it in range(3):
        out = net(x)
        pred = torch.relu(out)
        loss = U.calc_loss(pred, t, loss_type="mseMC")
        if it == 0:
            want = torch.mean((torch.relu(out.detach().double()) - t.double()) ** 2)
            assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
        opt.step()
        losses.append(float(loss))
    print("config5 losses:", losses)
    assert losses[-1] < losses[0]


def test_eval_folded_bn_matches_unfolded_path(monkeypatch):
    """Inference folds BatchNorm + ReLU into the conv epilogues (14 of 18 layers); B200UNET_FOLD_EVAL_BN=0 keeps the
    separate BN-apply pass. Both must agree to bf16 noise, masks almost everywhere, and the folded path must launch
    fewer kernels."""
    import unet_torch_b200 as U
    from unet_torch_b200 import _lib

    torch.manual_seed(8)
    net = U.UNet(3, 5)
    with torch.no_grad():
        _trained_buffers(net)
    net = net.cuda().eval()
    x = torch.randn(2, 3, 128, 160, device="cuda")
    eng = net._get_engine()
    with torch.no_grad():
        net(x)  # first call prepares the bf16 weight operands (extra launches)
        c0 = _lib.query("b200unet_launch_count")
        folded = net(x)
        c1 = _lib.query("b200unet_launch_count")
        eng.fold_eval_bn = False
        plain = net(x)
        c2 = _lib.query("b200unet_launch_count")
        eng.fold_eval_bn = True
    e = rel_l2(folded, plain)
    agree = float((U.predict_mask(folded) == U.predict_mask(plain)).float().mean())
    print(f"folded vs unfolded eval logits rel-L2 {e:.3e}, mask agreement {agree:.4f}, launches {c1 - c0} vs {c2 - c1}")
    assert e < 1e-2 and agree > 0.97
    assert (c1 - c0) < (c2 - c1)


@pytest.mark.parametrize("n,h,w", [(1, 400, 400), (3, 48, 80)])
def test_training_step_at_config_yml_size_against_oracle(n, h, w):
    """config.yml trains at 400x400 (25x25 at the bottleneck: every level has ragged 16x8 tiles), and an odd batch of
    non-square images: logits, loss and parameter gradients of one training step against the fp32 CPU oracle."""
    import unet_torch_b200 as U

    torch.manual_seed(1063)
    net = U.UNet(3, 2)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, 3, h, w, generator=g)
    y = (torch.rand(n, h, w, generator=g) > 0.6).float()
    net = net.cuda().train()
    U.loss.CLASS_NUMBER = 2
    out = net(x.cuda())
    loss = U.calc_loss(out, y.cuda(), loss_type="dice_bce_mc")
    loss.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    sd_req = dict(sd)
    sd_req.update(params)
    want, _ = O.unet_forward(sd_req, x, training=True)
    want_loss = O.calc_loss(want, y, "dice_bce_mc", 2)
    want_loss.backward()
    # yardstick: the torch CPU ops with ONLY the engine's bf16 storage points inserted (oracle/cpu_baseline.py)
    el, eloss, eg, _ = host_step(sd, x, y, 2, "dice_bce_mc", emulate=True)
    e, yard = rel_l2(out.detach(), want.detach()), rel_l2(el, want.detach())
    e_loss = abs(float(loss.detach()) - float(want_loss.detach())) / abs(float(want_loss.detach()))
    print(f"{n}x{h}x{w}: logits rel {e:.3e} (emulation {yard:.3e}; CUDA vs emulation {rel_l2(out.detach(), el):.3e}), loss rel {e_loss:.3e}")
    assert e <= 1.5 * yard and e_loss < 1e-2
    assert rel_l2(out.detach(), el) <= yard
    grads = dict(net.named_parameters())
    errs = {k: rel_l2(grads[k].grad, p.grad) for k, p in params.items()}
    yards = {k: rel_l2(eg[k], p.grad) for k, p in params.items()}
    vs_emu = {k: (rel_l2(grads[k].grad, eg[k]), cos(grads[k].grad, eg[k])) for k in params}
    # the deepest layers see 25x25 (or 3x5) maps: per-parameter error of a bf16 BatchNorm network vs fp32
    print(f"{n}x{h}x{w}: param-grad rel error vs the fp32 oracle: median {statistics.median(errs.values()):.3e} (emulation "
          f"{statistics.median(yards.values()):.3e}), worst {max(errs.values()):.3e} (emulation {max(yards.values()):.3e}); vs the "
          f"emulation: median {statistics.median(v[0] for v in vs_emu.values()):.3e}, min cos {min(v[1] for v in vs_emu.values()):.4f}")
    assert statistics.median(errs.values()) <= 1.5 * statistics.median(yards.values())
    for k in errs:
        assert errs[k] <= 1.5 * yards[k] + 0.02, (k, errs[k], yards[k])
    assert statistics.median(v[0] for v in vs_emu.values()) <= statistics.median(yards.values())
    assert min(v[1] for v in vs_emu.values()) > 0.9
    gn = torch.sqrt(sum(grads[k].grad.double().norm() ** 2 for k in params))
    wn = torch.sqrt(sum(p.grad.double().norm() ** 2 for p in params.values()))
    assert abs(float(gn / wn) - 1) < 0.1


def test_config2_full_size_training_invariants():
    """BASELINE configs[1] at full size (16 x 3 x 512 x 512, too large for the CPU oracle) through properties of the exact
    gradient that hold for any weights:
      * every 3x3 conv is followed by a train-mode BatchNorm, so the loss does not change when its weight tensor is
        scaled: <dL/dW, W> = 0 - a whole-backward check (dgrad, wgrad, BN backward) of all 18 conv layers;
      * softmax-CE + Dice depend on the logits only through softmax, so the logit gradient sums to zero over the classes at
        every pixel: sum(dL/d outc.bias) = 0;
      * the step is reproducible: same weights and inputs -> same loss and gradients (fp64 statistics are order-dependent
        only below fp32 resolution)."""
    import unet_torch_b200 as U

    torch.manual_seed(35)
    net = U.UNet(3, 2).cuda().train()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    U.loss.CLASS_NUMBER = 2
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(16, 3, 512, 512, device="cuda", generator=g)
    f = torch.nn.functional.interpolate(torch.randn(16, 1, 32, 32, device="cuda", generator=g), size=(512, 512), mode="bilinear")
    y = (f[:, 0] > 0.2).float()

    def step():
        net.load_state_dict(sd)
        net.zero_grad(set_to_none=True)
        out = net(x)
        loss = U.calc_loss(out, y, loss_type="dice_bce_mc")
        loss.backward()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in net.named_parameters()}

    l1, g1 = step()
    l2, g2 = step()
    assert l1 == l1 and abs(l1 - l2) <= 1e-6 * abs(l1)
    worst_rep = max(rel_l2(g2[k], g1[k]) for k in g1)
    params = dict(net.named_parameters())
    cos = {}
    for k, gr in g1.items():
        assert torch.isfinite(gr).all(), k
        if gr.dim() == 4 and gr.shape[-1] == 3:      # 3x3 conv weight (all of them feed a BatchNorm)
            w = params[k].detach().double()
            cos[k] = float((gr.double() * w).sum() / (gr.double().norm() * w.norm() + 1e-300))
    assert len(cos) == 18
    worst_k = max(cos, key=lambda k: abs(cos[k]))
    db = g1["outc.conv.bias"].double()
    print(f"config2 full size: loss {l1:.5f}, repeat rel diff {worst_rep:.2e}, worst |cos(dW, W)| {abs(cos[worst_k]):.2e} ({worst_k}), "
          f"sum(d bias) / |d bias| = {float(db.sum() / db.norm()):.2e}")
    assert worst_rep < 1e-4
    assert abs(cos[worst_k]) < 1e-2      # measured 6e-4
    assert abs(float(db.sum())) < 1e-3 * float(db.norm())


def test_width_128_on_the_tensor_core_engine():
    """initial_feature_map=128 (channel counts 128 ... 2048: the top of the BN / head kernels' range) runs on the tensor-core
    engine; logits, loss and gradients against the fp32 oracle with the bf16-storage emulation as yardstick."""
    import unet_torch_b200 as U

    torch.manual_seed(2)
    net = U.UNet(3, 2, 128)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 32, 48, generator=g)
    y = (torch.rand(2, 32, 48, generator=g) > 0.5).float()
    net = net.cuda().train()
    assert isinstance(net._engine_for(x.cuda()), type(net._get_engine()))
    U.loss.CLASS_NUMBER = 2
    out = net(x.cuda())
    loss = U.calc_loss(out, y.cuda(), loss_type="dice_bce_mc")
    loss.backward()
    rl, rloss, rg, _ = host_step(sd, x, y, 2, "dice_bce_mc")
    el, eloss, eg, _ = host_step(sd, x, y, 2, "dice_bce_mc", emulate=True)
    e, yard = rel_l2(out.detach(), rl), rel_l2(el, rl)
    grads = dict(net.named_parameters())
    errs = {k: rel_l2(grads[k].grad, rg[k]) for k in rg}
    yards = {k: rel_l2(eg[k], rg[k]) for k in rg}
    print(f"width 128: logits {e:.3e} (emulation {yard:.3e}), loss {abs(float(loss) - rloss) / abs(rloss):.2e}, grads median "
          f"{statistics.median(errs.values()):.3e} (emulation {statistics.median(yards.values()):.3e})")
    assert e <= 1.5 * yard and abs(float(loss) - rloss) / abs(rloss) < 1e-2
    assert statistics.median(errs.values()) <= 1.5 * statistics.median(yards.values())
    assert all(errs[k] <= 1.5 * yards[k] + 0.02 for k in errs)
    with pytest.raises(ValueError):
        U.UNet(3, 2, 192)._get_engine()                   # not a power-of-two width: generic fp32 engine only
