"""-m gpu: UNet_multitask (Model.py:172-250) - one encoder, two decoders - on the B200 engine,
against the UNMODIFIED reference's outputs (tests/golden/ref_multitask.pt, oracle/make_golden_multitask.py): same seed ->
same weights -> both logits, the summed relu+MSE loss of Trainer.py:877-900, BatchNorm buffers and parameter gradients.
Tolerances as in test_gpu_model.py (north_star: 1e-2 relative in bf16)."""
import statistics

import pytest
import torch
import torch.nn.functional as F

from gpu_util import cos, rel_l2, to_nhwc_bf16, from_nhwc
from oracle import cpu_baseline
from test_gpu_model import checksum

pytestmark = pytest.mark.gpu


def test_slice_copy_and_add_kernels():
    from unet_torch_b200 import ops

    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 128, 16, 24, generator=g)
    b = torch.randn(2, 64, 16, 24, generator=g)
    buf_a, buf_b = to_nhwc_bf16(a), to_nhwc_bf16(b)
    dst = torch.zeros_like(buf_a)
    ops.nhwc_copy(buf_a[..., :64], dst[..., 64:])                      # strided slice -> strided slice
    assert torch.equal(dst[..., 64:], buf_a[..., :64]) and not bool(dst[..., :64].any())
    want = (buf_a[..., 64:].float() + buf_b.float()).to(torch.bfloat16)  # fp32 sum, one rounding
    keep = buf_a[..., :64].clone()
    ops.nhwc_add(buf_a[..., 64:], buf_b)                               # in place into a slice
    assert torch.equal(buf_a[..., 64:], want) and torch.equal(buf_a[..., :64], keep)
    assert from_nhwc(buf_a).shape == a.shape
    with pytest.raises(ValueError):
        ops.nhwc_add(buf_a, buf_b)


def test_multitask_forward_backward_against_reference_golden(golden):
    import unet_torch_b200 as U

    g = golden("ref_multitask.pt")["w64_sum"]
    ch, ncls, width, n, h, w, seed = g["cfg"]
    torch.manual_seed(seed)
    net = U.UNet_multitask(ch, ncls, width)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    assert list(sd0.keys()) == list(g["sd0_checksum"].keys()) and len(sd0) == 176
    for k, v in sd0.items():  # identical initial weights as the reference under this seed
        assert torch.allclose(checksum(v), g["sd0_checksum"][k], rtol=1e-12, atol=0), k
    net = net.cuda().train()
    x, t1, t2 = g["x"].cuda(), g["t1"].cuda(), g["t2"].cuda()
    o1, o2 = net(x)
    loss = U.calc_loss(torch.relu(o1), t1, loss_type="mseMC") + U.calc_loss(torch.relu(o2), t2, loss_type="mseMC")
    loss.backward()
    torch.cuda.synchronize()
    e1, e2 = rel_l2(o1.detach(), g["o1_64"]), rel_l2(o2.detach(), g["o2_64"])
    e_loss = abs(float(loss) - float(g["loss64"])) / abs(float(g["loss64"]))
    # yardstick + composition check: the same torch CPU ops with ONLY the engine's bf16 storage points inserted
    # (oracle/cpu_baseline.py; reproduces the reference's own bf16-autocast error to a few percent, tests/test_oracle.py)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd0.items()}
    m1, m2 = cpu_baseline.unet_multitask_forward_torchops(p, g["x"], True, emulate_bf16=True)
    (F.mse_loss(F.relu(m1), g["t1"]) + F.mse_loss(F.relu(m2), g["t2"])).backward()
    y1, y2 = rel_l2(m1.detach(), g["o1_64"]), rel_l2(m2.detach(), g["o2_64"])
    print(f"multitask: logits rel {e1:.3e} / {e2:.3e} (bf16-storage emulation {y1:.3e} / {y2:.3e}), loss rel {e_loss:.3e}; "
          f"vs the emulation {rel_l2(o1.detach(), m1.detach()):.3e} / {rel_l2(o2.detach(), m2.detach()):.3e}")
    assert e1 <= 1.5 * y1 and e2 <= 1.5 * y2 and e_loss < 1e-2
    assert rel_l2(o1.detach(), m1.detach()) <= y1 and rel_l2(o2.detach(), m2.detach()) <= y2
    sd1 = net.state_dict()
    for k, v in g["buffers1_checksum"].items():
        got = checksum(sd1[k])
        if "num_batches" in k:
            assert int(got[0]) == int(v[0]), k
        else:
            assert abs(float(got[1]) - float(v[1])) <= 2e-2 * abs(float(v[1])) + 1e-6, k
    grads = {k: p.grad for k, p in net.named_parameters()}
    assert all(gr is not None for gr in grads.values())
    errs, yard = {}, {}
    for k, gs in list(g["grad_small64"].items()) + list(g["grad_sample64"].items()):
        small = k in g["grad_small64"]
        errs[k] = rel_l2(grads[k] if small else grads[k].flatten()[::997], gs)
        yard[k] = rel_l2(p[k].grad if small else p[k].grad.flatten()[::997], gs)
    med, med_yard = statistics.median(errs.values()), statistics.median(yard.values())
    enc = [k for k in errs if k.startswith(("inc", "down"))]   # encoder gradients are the SUM of both decoders' contributions
    vs_emu = {k: (rel_l2(grads[k], p[k].grad), cos(grads[k], p[k].grad)) for k in errs}
    print(f"multitask: param-grad rel error vs fp64 reference: median {med:.3e} (emulation {med_yard:.3e}), worst "
          f"{max(errs.values()):.3e} (emulation {max(yard.values()):.3e}), encoder worst {max(errs[k] for k in enc):.3e}; vs the "
          f"emulation: median {statistics.median(v[0] for v in vs_emu.values()):.3e}, min cos {min(v[1] for v in vs_emu.values()):.4f}")
    assert med <= 1.5 * med_yard
    for k, e in errs.items():
        assert e <= 1.5 * yard[k] + 0.02, (k, e, yard[k])
    assert statistics.median(v[0] for v in vs_emu.values()) <= med_yard and min(v[1] for v in vs_emu.values()) > 0.9
    # gradient norms: the encoder's must reflect both decoders (a dropped contribution would roughly halve them)
    for k in ("down4.maxpool_conv.1.double_conv.3.weight", "inc.double_conv.3.weight", "down2.maxpool_conv.1.double_conv.0.weight"):
        ratio = float(grads[k].double().norm() / g["grad_norm64"][k])
        assert 0.8 < ratio < 1.25, (k, ratio)
    net.eval()
    with torch.no_grad():
        v1, v2 = net(x)
    assert rel_l2(v1, g["e1"]) <= 1.5 * y1 and rel_l2(v2, g["e2"]) <= 1.5 * y2


def test_multitask_one_output_unused_and_optimizer_step():
    """A loss over decoder 1 only: decoder 2's parameters get zero gradients (autograd passes None for the unused
    output), the encoder's match a run where decoder 2's loss has weight 0; FusedSGD steps the two-decoder net."""
    import unet_torch_b200 as U

    torch.manual_seed(2)
    net = U.UNet_multitask(3, 2).cuda().train()
    x = torch.randn(2, 3, 32, 48, device="cuda")
    t = torch.rand(2, 2, 32, 48, device="cuda")
    o1, o2 = net(x)
    U.calc_loss(torch.relu(o1), t, loss_type="mseMC").backward()
    g_enc = net.inc.double_conv[3].weight.grad.clone()
    assert float(net.up4_decod2.conv.double_conv[0].weight.grad.abs().max()) == 0.0
    assert float(net.up4_decod1.conv.double_conv[0].weight.grad.abs().max()) > 0.0
    net.zero_grad()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    o1, o2 = net(x)
    (U.calc_loss(torch.relu(o1), t, loss_type="mseMC") + 0.0 * U.calc_loss(torch.relu(o2), t, loss_type="mseMC")).backward()
    assert torch.equal(net.inc.double_conv[3].weight.grad, g_enc)
    opt = U.FusedSGD(net, lr=0.01, momentum=0.9, weight_decay=1e-4)
    losses = []
    for _ in range(4):
        o1, o2 = net(x)
        l = U.calc_loss(torch.relu(o1), t, loss_type="mseMC") + U.calc_loss(torch.relu(o2), t, loss_type="mseMC")
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()
        losses.append(float(l))
    assert losses[-1] < losses[0]
    with pytest.raises(ValueError):
        net(torch.randn(1, 3, 30, 30, device="cuda"))


def test_multitask_cuda_graph_replay_matches_eager():
    """enable_cuda_graphs on the two-decoder network: four FusedSGD steps replayed as graphs (one output unused in step 2:
    autograd hands None for it) equal the eager launches."""
    import unet_torch_b200 as U

    x = torch.randn(2, 3, 32, 48, device="cuda")
    t = torch.rand(2, 2, 32, 48, device="cuda")
    states = []
    for graphs in (False, True):
        torch.manual_seed(4)
        net = U.UNet_multitask(3, 2).cuda().train()
        net.enable_cuda_graphs(graphs, share_grads=graphs)
        opt = U.FusedSGD(net, lr=0.01, momentum=0.9, weight_decay=1e-4)
        losses = []
        for it in range(5):
            o1, o2 = net(x)
            l = U.calc_loss(torch.relu(o1), t, loss_type="mseMC")
            if it != 2:
                l = l + U.calc_loss(torch.relu(o2), t, loss_type="mseMC")
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
            losses.append(float(l))
        if graphs:
            assert len(net._get_engine()._graphs) == 1
        states.append((losses, {k: v.clone() for k, v in net.state_dict().items()}))
    (la, sa), (lb, sb) = states
    assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(la, lb)), (la, lb)
    for k in sa:
        if sa[k].is_floating_point():
            assert torch.allclose(sa[k], sb[k], rtol=1e-4, atol=1e-6), k
