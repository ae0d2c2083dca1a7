"""calc_loss branches next to the fused hot ones, MultitaskUncertaintyLoss and MRAccuracy against the UNMODIFIED reference's
own outputs (tests/golden/ref_loss_branches.pt, oracle/make_golden_losses.py; reference loss.py:309-325, 421-440, 442-486).
The torch-op branches run wherever their tensors live (CPU here); 'rmse' wraps the fused MSE kernel -> -m gpu."""
import pytest
import torch


@pytest.mark.parametrize("lt", ["BCE", "dice_bce", "l1loss"])
def test_torch_op_branches_match_reference(golden, lt):
    import unet_torch_b200 as U

    g = golden("ref_loss_branches.pt")[lt]
    p = g["pred"].clone().requires_grad_(True)
    l = U.calc_loss(p, g["target"], loss_type=lt)
    (gr,) = torch.autograd.grad(l, p)
    assert abs(float(l) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert torch.allclose(gr, g["grad"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", ["uncertainty_reg", "uncertainty_mixed"])
def test_multitask_uncertainty_loss_matches_reference(golden, name):
    """loss.py:309-325 as constructed at Trainer.py:1007 and called at :1065."""
    import loss as shim  # the root drop-in module Trainer.py imports from

    g = golden("ref_loss_branches.pt")[name]
    ls = [v.clone().requires_grad_(True) for v in g["losses"]]
    lv = [v.clone().requires_grad_(True) for v in g["log_vars"]]
    tot = shim.MultitaskUncertaintyLoss()(ls, lv, g["flags"])
    assert tot.shape == g["total"].shape and torch.allclose(tot, g["total"], rtol=1e-6, atol=0)
    grads = torch.autograd.grad(tot.sum(), ls + lv)
    for a, b in zip(grads, g["grads"]):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", ["MRAccuracy", "MRAccuracy_fp"])
def test_mr_accuracy_matches_reference(golden, name):
    import loss as shim

    g = golden("ref_loss_branches.pt")[name]
    assert abs(shim.MRAccuracy(g["pred"], g["target"]) - g["value"]) < 1e-12


def test_unknown_branch_raises():
    import unet_torch_b200 as U

    with pytest.raises(NotImplementedError):
        U.calc_loss(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4), loss_type="HausdorffDTLoss")


@pytest.mark.gpu
def test_rmse_and_uncertainty_on_device(golden):
    import unet_torch_b200 as U

    g = golden("ref_loss_branches.pt")["rmse"]
    p = g["pred"].cuda().requires_grad_(True)
    l = U.calc_loss(p, g["target"].cuda(), loss_type="rmse")
    (gr,) = torch.autograd.grad(l, p)
    assert abs(float(l) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert torch.allclose(gr.cpu(), g["grad"], rtol=1e-4, atol=1e-8)
    # the Trainer's pattern (Trainer.py:1003-1066): CPU log-variance leaves, device losses from the fused relu+MSE kernel
    gen = torch.Generator().manual_seed(3)
    o1 = torch.randn(2, 1, 32, 32, generator=gen).cuda().requires_grad_(True)
    o2 = torch.randn(2, 1, 32, 32, generator=gen).cuda().requires_grad_(True)
    t1, t2 = torch.rand(2, 32, 32, generator=gen).cuda(), torch.rand(2, 32, 32, generator=gen).cuda()
    lv = [torch.zeros((1,), requires_grad=True), torch.full((1,), 0.4, requires_grad=True)]
    l1 = U.calc_loss(torch.relu(o1), t1, loss_type="mse")
    l2 = U.calc_loss(torch.relu(o2), t2, loss_type="mse")
    tot = U.MultitaskUncertaintyLoss()([l1, l2], lv, [True, True])
    tot.backward()
    w1 = torch.mean((torch.relu(o1.detach()).squeeze(1) - t1) ** 2).double().cpu()
    w2 = torch.mean((torch.relu(o2.detach()).squeeze(1) - t2) ** 2).double().cpu()
    import math
    want = w1 / 2 + 0.0 + w2 / (2 * math.exp(0.4)) + 0.2
    assert tot.is_cuda and abs(float(tot) - float(want)) < 1e-5 * abs(float(want))
    assert lv[0].grad is not None and not lv[0].grad.is_cuda and abs(float(lv[0].grad) - float(-w1 / 2 + 0.5)) < 1e-5
    assert o1.grad is not None and float(o1.grad.abs().sum()) > 0


@pytest.mark.gpu
def test_out_of_range_labels_surface_like_the_reference():
    """nn.CrossEntropyLoss raises for a label outside [0, C) (loss.py:469, 498). The fused kernel flags it on the device: an
    immediate IndexError with check_labels=True, otherwise at the next loss call / check_pending_label_errors()."""
    import unet_torch_b200 as U
    from unet_torch_b200 import loss as L

    L.check_pending_label_errors(wait=True)
    z = torch.randn(2, 3, 16, 16, device="cuda")
    t = torch.randint(0, 3, (2, 16, 16), device="cuda").float()
    bad = t.clone()
    bad[0, 3, 3] = 7
    U.loss.CLASS_NUMBER = 3
    with pytest.raises(IndexError):
        L.ce_dice_loss(z, bad, check_labels=True)
    L._pending.clear()
    U.calc_loss(z, bad, loss_type="dice_bce_mc")          # no sync: flagged, not raised yet
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        U.calc_loss(z, t, loss_type="dice_bce_mc")        # the next call reports it
    U.calc_loss(z, t, loss_type="CE")                     # and the state is clean again
    L.check_pending_label_errors(wait=True)
