"""-m gpu: the generic fp32 engine (csrc/generic_f32.cu + unet-torch_b200/generic.py).

* fp32 CHECK MODE against the unmodified reference's golden outputs (tests/golden/ref_small_nets.pt: narrow nets with
  full state / gradient dumps in fp32 and fp64; ref_full_nets.pt: full-width nets): north_star tolerance 1e-4 for
  logits and loss; gradients of a BatchNorm net are judged with the reference's own fp32-vs-fp64 error as yardstick.
* shapes outside the tensor-core envelope against the oracle (itself pinned to the live reference for these cases in
  tests/test_oracle.py): H, W not divisible by 16 (floor pooling + F.pad, Model.py:69-73) and the dropout variants.
"""
import pytest
import torch

from gpu_util import rel_l2
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

SMALL = ["w4_c1_k2_dicebce", "w4_c3_k5_dicebce", "w4_c3_k3_ce", "w4_c3_k2_msemc"]


@pytest.mark.parametrize("case", SMALL)
def test_check_mode_small_nets_against_reference_golden(golden, case):
    import unet_torch_b200 as U
    from unet_torch_b200.generic import GenericEngine

    g = golden("ref_small_nets.pt")[case]
    ch, ncls, width, n, h, w, seed = g["cfg"]
    net = U.UNet(ch, ncls, width)
    net.load_state_dict(g["sd0"])
    net = net.cuda().train()
    U.loss.CLASS_NUMBER = ncls
    x, y = g["x"].cuda(), g["y"].cuda()
    out = net(x)
    assert isinstance(net._engine_for(x), GenericEngine)  # width 4: outside the tensor-core envelope
    pred = torch.relu(out) if g["loss_type"].startswith("mse") else out
    loss = U.calc_loss(pred, y, loss_type=g["loss_type"])
    loss.backward()
    e_logits = rel_l2(out.detach(), g["logits"])
    e_logits64 = rel_l2(out.detach(), g["logits64"])
    e_loss = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    print(f"{case}: logits rel vs ref fp32 {e_logits:.2e} (vs fp64 {e_logits64:.2e}), loss rel {e_loss:.2e}")
    assert e_logits < 1e-4 and e_logits64 < 1e-4 and e_loss < 1e-4  # north_star fp32 check-mode tolerance
    worst = 0.0
    for k, p in net.named_parameters():
        ours = rel_l2(p.grad, g["grads64"][k])
        ref = rel_l2(g["grads"][k], g["grads64"][k])  # the reference's own fp32 rounding on this gradient
        worst = max(worst, ours)
        assert ours < 5 * ref + 1e-4, (k, ours, ref)
    print(f"{case}: worst gradient rel error vs fp64 reference {worst:.2e}")
    for k, v in g["buffers1"].items():
        got = net.state_dict()[k]
        if "num_batches" in k:
            assert int(got) == int(v)
        else:
            assert rel_l2(got, v) < 1e-5, k
    net.eval()
    with torch.no_grad():
        oe = net(x)
    assert rel_l2(oe, g["logits_eval"]) < 1e-4


@pytest.mark.parametrize("case", ["w64_c3_k2_dicebce", "w64_c3_k5_ce"])
def test_check_mode_full_width(golden, case):
    import unet_torch_b200 as U
    from unet_torch_b200.generic import GenericEngine

    g = golden("ref_full_nets.pt")[case]
    ch, ncls, width, n, h, w, seed = g["cfg"]
    torch.manual_seed(seed)
    net = U.UNet(ch, ncls, width).cuda().train().set_check_mode(True)
    U.loss.CLASS_NUMBER = ncls
    x, y = g["x"].cuda(), g["y"].cuda()
    out = net(x)
    assert isinstance(net._engine_for(x), GenericEngine)
    loss = U.calc_loss(out, y, loss_type=g["loss_type"])
    loss.backward()
    e_logits = rel_l2(out.detach(), g["logits64"])
    e_loss = abs(float(loss) - float(g["loss64"])) / abs(float(g["loss64"]))
    print(f"{case} (check mode): logits rel {e_logits:.2e} loss rel {e_loss:.2e}")
    assert e_logits < 1e-4 and e_loss < 1e-4
    worst = 0.0
    for k, gs in g["grad_small64"].items():
        worst = max(worst, rel_l2(dict(net.named_parameters())[k].grad, gs))
    for k, gs in g["grad_sample64"].items():
        worst = max(worst, rel_l2(dict(net.named_parameters())[k].grad.flatten()[::997], gs))
    print(f"{case} (check mode): worst gradient rel error vs fp64 reference {worst:.2e}")
    assert worst < 5e-2  # the reference's own fp32 gradients differ from fp64 by 0.4-0.8 % on these nets


def _oracle_run(sd, x, y, loss_type, ncls, masks=None):
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.double()
                if v.is_floating_point() else v) for k, v in sd.items()}
    logits, nb = O.unet_forward(sd64, x.double(), training=True, dropout_masks=masks)
    loss = O.calc_loss(logits, y.double(), loss_type, ncls)
    loss.backward()
    grads = {k: v.grad for k, v in sd64.items() if isinstance(v, torch.Tensor) and v.requires_grad}
    return logits.detach(), loss.detach(), grads, nb


@pytest.mark.parametrize("h,w,width", [(37, 51, 8), (40, 56, 64), (33, 16, 8)])
def test_sizes_not_divisible_by_16_against_oracle(h, w, width):
    import unet_torch_b200 as U
    from unet_torch_b200.generic import GenericEngine

    torch.manual_seed(9)
    net = U.UNet(3, 2, width)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    U.loss.CLASS_NUMBER = 2
    gen = torch.Generator().manual_seed(10)
    x = torch.randn(2, 3, h, w, generator=gen)
    y = torch.randint(0, 2, (2, h, w), generator=gen).float()
    out = net(x.cuda())
    assert isinstance(net._engine_for(x.cuda()), GenericEngine)
    loss = U.calc_loss(out, y.cuda(), loss_type="dice_bce_mc")
    loss.backward()
    want, want_loss, want_grads, nb = _oracle_run(sd, x, y, "dice_bce_mc", 2)
    assert out.shape == (2, 2, h, w)
    e = rel_l2(out.detach(), want)
    print(f"{h}x{w} width {width}: logits rel {e:.2e}")
    assert e < 1e-4
    assert abs(float(loss) - float(want_loss)) < 1e-4 * abs(float(want_loss))
    worst = max(rel_l2(p.grad, want_grads[k]) for k, p in net.named_parameters())
    print(f"{h}x{w} width {width}: worst gradient rel error vs fp64 oracle {worst:.2e}")
    assert worst < 5e-2
    for k, v in nb.items():
        if "num_batches" not in k:
            assert rel_l2(net.state_dict()[k], v) < 1e-5, k


@pytest.mark.parametrize("h,w", [(32, 32), (50, 35)])
def test_dropout_variant_against_oracle(h, w):
    """dropout=True (Model.py:34-39, 81-82): masks drawn by the engine are replayed through the oracle."""
    import unet_torch_b200 as U

    torch.manual_seed(12)
    net = U.UNet(3, 2, 8, dropout=True, dropout_p=0.3)
    assert "down1.maxpool_conv.2.double_conv.0.weight" in net.state_dict()  # DoubleConv index shifts, as in the reference
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    gen = torch.Generator().manual_seed(13)
    x = torch.randn(2, 3, h, w, generator=gen)
    y = torch.randint(0, 2, (2, h, w), generator=gen).float()
    eng = net._engine_for(x.cuda())
    logits, saved = eng.forward(x.cuda(), training=True, save=True)
    masks = {f"down{l + 1}": saved["enc"][l][4].double().cpu() for l in range(4)}
    masks.update({f"up{j + 1}": saved["dec"][j][4].double().cpu() for j in range(4)})
    for name, shape in O.dropout_mask_shapes(2, h, w, 8):
        assert tuple(masks[name].shape) == shape
        vals = set(masks[name].unique().tolist())
        assert vals <= {0.0, 1.0 / 0.7} or all(abs(v) < 1e-12 or abs(v - 1 / 0.7) < 1e-6 for v in vals)
    # loss gradient from the oracle on the device logits, then the engine's backward
    lg = logits.detach().cpu().double().requires_grad_(True)
    O.calc_loss(lg, y.double(), "dice_bce_mc", 2).backward()
    grads = eng.backward(saved, lg.grad.float().cuda())
    want, want_loss, want_grads, _ = _oracle_run(sd, x, y, "dice_bce_mc", 2, masks)
    e = rel_l2(logits, want)
    print(f"dropout {h}x{w}: logits rel {e:.2e}")
    assert e < 1e-4
    named = dict(net.named_parameters())
    worst = max(rel_l2(grads[named[k]], want_grads[k]) for k in named)
    print(f"dropout {h}x{w}: worst gradient rel error vs fp64 oracle {worst:.2e}")
    assert worst < 5e-2
    # eval mode: dropout is the identity and nothing is random
    net.eval()
    with torch.no_grad():
        a, b = net(x.cuda()), net(x.cuda())
    assert torch.equal(a, b)
    we, _ = O.unet_forward({k: v.cpu().double() if v.is_floating_point() else v.cpu() for k, v in net.state_dict().items()},
                           x.double(), training=False)
    assert rel_l2(a, we) < 1e-4
