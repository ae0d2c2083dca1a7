"""-m gpu: UNet_attention (Model.py:257-391) - attention-gated skips - on the library's generic fp32
CUDA engine, against the UNMODIFIED reference's outputs (tests/golden/ref_attention.pt, oracle/make_golden_attention.py):
same seed -> same weights -> logits, loss, every parameter gradient, BatchNorm buffers (three more BatchNorms per gate), eval
logits. fp32 engine => north_star's fp32 tolerance (1e-4) for logits and loss."""
import pytest
import torch

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["w4_c3_k2_dicebce", "w4_c1_k3_msemc"])
def test_attention_forward_backward_against_reference_golden(golden, case):
    import unet_torch_b200 as U

    g = golden("ref_attention.pt")[case]
    ch, ncls, width, n, h, w, seed = g["cfg"]
    torch.manual_seed(seed)
    net = U.UNet_attention(ch, ncls, width)
    assert all(torch.equal(v, g["sd0"][k]) for k, v in net.state_dict().items())
    net = net.cuda().train()
    U.loss.CLASS_NUMBER = ncls
    out = net(g["x"].cuda())
    pred = torch.relu(out) if g["loss_type"].startswith("mse") else out
    loss = U.calc_loss(pred, g["y"].cuda(), loss_type=g["loss_type"])
    loss.backward()
    torch.cuda.synchronize()
    e_logits = rel_l2(out.detach(), g["logits64"])
    e_loss = abs(float(loss) - float(g["loss64"])) / abs(float(g["loss64"]))
    # gradients vs the fp64 reference; yardstick = the reference's own fp32 run; exactly-zero gradients (biases in front of a
    # BatchNorm) are measured against 1e-6 x the largest gradient norm of the net
    floor = 1e-6 * max(float(gr.double().norm()) for gr in g["grads64"].values())
    err = lambda a, b: float((a.double().cpu() - b.double()).norm()) / max(float(b.double().norm()), floor)  # noqa: E731
    grads = {k: p.grad for k, p in net.named_parameters()}
    ours = {k: err(grads[k], g["grads64"][k]) for k in g["grads64"]}
    ref32 = {k: err(g["grads"][k], g["grads64"][k]) for k in g["grads64"]}
    kw = max(ours, key=ours.get)
    print(f"{case}: logits {e_logits:.3e} loss {e_loss:.3e}; grads vs fp64 reference: worst {ours[kw]:.3e} ({kw}); the reference's "
          f"own fp32 run: worst {max(ref32.values()):.3e}")
    assert e_logits < 1e-4 and e_loss < 1e-4
    assert ours[kw] <= max(3 * max(ref32.values()), 2e-3)
    sd1 = net.state_dict()
    for k, v in g["buffers1"].items():
        if "num_batches" in k:
            assert int(sd1[k]) == int(v), k
        else:
            assert rel_l2(sd1[k], v) < 1e-4, k
    net.eval()
    with torch.no_grad():
        oe = net(g["x"].cuda())
    assert rel_l2(oe, g["logits_eval"]) < 1e-4


def test_attention_trains_with_dropout_and_optimizers():
    """The dropout variant (Model.py:320-324, 336-343) and both fused optimizers step the gated network; odd sizes raise."""
    import unet_torch_b200 as U

    torch.manual_seed(0)
    net = U.UNet_attention(3, 2, 8, dropout=True, dropout_p=0.2).cuda().train()
    x = torch.randn(2, 3, 32, 32, device="cuda")
    y = (torch.rand(2, 32, 32, device="cuda") > 0.5).float()
    U.loss.CLASS_NUMBER = 2
    for opt in (U.FusedSGD(net, lr=0.02, momentum=0.9), U.FusedAdam(net, lr=1e-3)):
        losses = []
        for _ in range(4):
            loss = U.calc_loss(net(x), y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
            opt.step()
            losses.append(float(loss))
        assert losses[-1] < losses[0], losses
    with pytest.raises(ValueError):
        net(torch.randn(1, 3, 40, 40, device="cuda"))
