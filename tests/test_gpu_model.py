"""-m gpu: whole-network parity of the B200 UNet against the UNMODIFIED reference's outputs (tests/golden/
ref_full_nets.pt, produced by oracle/make_golden.py): same seed -> same initial weights -> logits, loss, BatchNorm
buffers and parameter gradients. Tolerances follow north_star: 1e-2 relative (bf16) for logits and loss; gradients
of a BatchNorm network are judged against the fp64 reference with the reference's own fp32-vs-fp64 and
bf16-autocast error as the yardstick."""
import statistics

import pytest
import torch

from gpu_util import cos, host_step, rel_l2

pytestmark = pytest.mark.gpu

CASES = ["w64_c3_k2_dicebce", "w64_c1_k2_dicebce", "w64_c3_k5_ce", "w64_c3_k2_msemc"]


def build(cfg):
    import unet_torch_b200 as U

    ch, ncls, width, n, h, w, seed = cfg
    torch.manual_seed(seed)
    net = U.UNet(ch, ncls, width)  # CPU init consumes the RNG exactly like the reference constructor
    return net


def checksum(v):
    f = v.double().flatten()
    return torch.tensor([f.sum(), f.abs().sum(), f[0], f[f.numel() // 2], f[-1]], dtype=torch.float64)


@pytest.mark.parametrize("case", CASES)
def test_forward_backward_against_reference_golden(golden, case):
    import unet_torch_b200 as U

    g = golden("ref_full_nets.pt")[case]
    ch, ncls, width, n, h, w, seed = g["cfg"]
    net = build(g["cfg"])
    for k, v in net.state_dict().items():  # identical initial weights as the reference under this seed
        assert torch.allclose(checksum(v), g["sd0_checksum"][k], rtol=1e-12, atol=0), k
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    yard = golden("ref_bf16_yardstick.pt")[case]["bf16_autocast"]  # the reference's OWN bf16-autocast error vs its fp64 run
    net = net.cuda().train()
    U.loss.CLASS_NUMBER = ncls
    x, y = g["x"].cuda(), g["y"].cuda()
    out = net(x)
    assert out.shape == g["logits"].shape and out.dtype == torch.float32
    pred = torch.relu(out) if g["loss_type"].startswith("mse") else out
    loss = U.calc_loss(pred, y, loss_type=g["loss_type"])
    loss.backward()
    torch.cuda.synchronize()
    e_logits = rel_l2(out.detach(), g["logits64"])
    e_loss = abs(float(loss) - float(g["loss64"])) / abs(float(g["loss64"]))
    print(f"{case}: logits rel {e_logits:.3e} (reference bf16-autocast {yard['logits']:.3e}) loss rel {e_loss:.3e} "
          f"(reference bf16-autocast {yard['loss']:.3e})")
    assert e_logits <= 1.5 * yard["logits"]
    assert e_loss < 1e-2
    # BatchNorm buffers after one training step
    sd1 = net.state_dict()
    for k, v in g["buffers1_checksum"].items():
        got = checksum(sd1[k])
        if "num_batches" in k:
            assert int(got[0]) == int(v[0]), k
        else:
            assert abs(float(got[1]) - float(v[1])) <= 2e-2 * abs(float(v[1])) + 1e-6, k
    # gradients: norm and direction vs the fp64 reference
    grads = {k: p.grad for k, p in net.named_parameters()}
    assert all(gr is not None for gr in grads.values())
    errs = {k: rel_l2(grads[k], gs) for k, gs in g["grad_small64"].items()}
    errs.update({k: rel_l2(grads[k].flatten()[::997], gs) for k, gs in g["grad_sample64"].items()})
    assert set(errs) == set(yard["grads"])
    med, med_yard = statistics.median(errs.values()), statistics.median(yard["grads"].values())
    kw = max(errs, key=lambda k: errs[k] / (yard["grads"][k] + 1e-12))
    print(f"{case}: param-grad rel error vs the fp64 reference: median {med:.3e} (reference bf16-autocast {med_yard:.3e}), worst "
          f"{max(errs.values()):.3e} (reference {max(yard['grads'].values()):.3e}), worst ratio to the reference's own error "
          f"{errs[kw] / (yard['grads'][kw] + 1e-12):.2f} ({kw})")
    assert med <= 1.5 * med_yard
    for k, e in errs.items():  # no parameter is worse than 1.5 x what the reference's own bf16 run shows for it
        assert e <= 1.5 * yard["grads"][k] + 0.02, (k, e, yard["grads"][k])
    # composition check: the same torch CPU ops with ONLY the engine's bf16 storage points inserted (oracle/cpu_baseline.py)
    el, eloss, eg, _ = host_step(sd0, g["x"], g["y"], ncls, g["loss_type"], relu=g["loss_type"].startswith("mse"), emulate=True)
    rows = {k: (rel_l2(grads[k], eg[k]), cos(grads[k], eg[k])) for k in eg}
    kw = max(rows, key=lambda k: rows[k][0])
    print(f"{case}: vs the bf16-storage emulation on the host: logits {rel_l2(out.detach(), el):.3e}, loss "
          f"{abs(float(loss) - eloss) / abs(eloss):.2e}, grads median rel {statistics.median(r[0] for r in rows.values()):.3e} "
          f"worst {rows[kw][0]:.3e} ({kw}), min cos {min(r[1] for r in rows.values()):.4f}")
    # two correct bf16-storage runs decorrelate within a few layers (tests/test_gpu_fullsize.py): they must be no further
    # apart than either is from the fp64 truth; the tight composition check is tests/test_gpu_replay.py
    assert rel_l2(out.detach(), el) <= yard["logits"]
    assert statistics.median(r[0] for r in rows.values()) <= med_yard and min(r[1] for r in rows.values()) > 0.9
    # eval mode (running statistics)
    net.eval()
    with torch.no_grad():
        oe = net(x)
    assert oe.shape == g["logits_eval"].shape
    # running statistics after ONE momentum-0.1 update are still close to their init, so eval logits are large-ish
    # and well conditioned: compare values, not only shapes (reference: eval forward right after its own first step)
    e_eval = rel_l2(oe, g["logits_eval"])
    print(f"{case}: eval logits rel {e_eval:.3e}")
    assert e_eval <= 1.5 * yard["logits"]


def test_module_interface_matches_reference():
    import unet_torch_b200 as U

    net = U.UNet(3, 2)
    sd = net.state_dict()
    assert len(sd) == 118
    assert sd["inc.double_conv.0.weight"].shape == (64, 3, 3, 3)
    assert sd["down1.maxpool_conv.1.double_conv.1.running_var"].shape == (128,)
    assert sd["up1.up.weight"].shape == (1024, 512, 2, 2)
    assert sd["outc.conv.bias"].shape == (2,)
    assert sum(p.numel() for p in net.parameters()) == 31037698
    net2 = U.UNet(-1, 2)
    assert net2.n_channels == 1
    net = net.cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(2, 3, 32, 32, device="cuda")
    y = torch.randint(0, 2, (2, 32, 32), device="cuda").float()
    U.loss.CLASS_NUMBER = 2
    losses = []
    for _ in range(3):
        out = net(x)
        l = U.calc_loss(out, y, loss_type="dice_bce_mc")
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(l.item())
    assert losses[-1] < losses[0]  # training on one batch makes progress
    net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})


def test_cuda_graph_replay_is_bit_identical_to_eager():
    """enable_cuda_graphs(): the captured forward/backward must produce the same bits as the eager launches, across
    optimizer steps (weights change -> operands refreshed in place), and refuse a stale backward."""
    import unet_torch_b200 as U

    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 64, 48, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    y = torch.randint(0, 2, (2, 64, 48), device="cuda", generator=torch.Generator("cuda").manual_seed(4)).float()

    def run(graphs):
        torch.manual_seed(11)
        net = U.UNet(3, 2).cuda().train().enable_cuda_graphs(graphs)
        opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
        outs = []
        for _ in range(4):
            out = net(x)
            loss = U.calc_loss(out, y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
            outs.append((out.detach().clone(), float(loss), net.up1.up.weight.grad.clone(),
                         net.inc.double_conv[0].weight.grad.clone()))
            opt.step()
        net.eval()
        with torch.no_grad():
            e1 = net(x).clone()
            e2 = net(x).clone()  # second eval call of this shape replays the captured eval graph
        return net, outs, e1, e2

    net_e, eager, ee1, ee2 = run(False)
    net_g, graph, ge1, ge2 = run(True)
    assert net_g._engine._graphs, "no graph was captured"
    for (o1, l1, g1, h1), (o2, l2, g2, h2) in zip(eager, graph):
        assert torch.equal(o1, o2) and l1 == l2 and torch.equal(g1, g2) and torch.equal(h1, h2)
    assert torch.equal(ee1, ge1) and torch.equal(ee2, ge2) and torch.equal(ge1, ge2)
    for (k, a), (_, b) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
        assert torch.equal(a, b), k
    # stale backward: two forwards of the same shape, then backward of the first
    net_g.train()
    o1 = net_g(x)
    o2 = net_g(x)
    with pytest.raises(RuntimeError, match="ONE set of saved activations"):
        o1.sum().backward()
    o2.sum().backward()


@pytest.mark.parametrize("nesterov,damp,wd", [(False, 0.0, 1e-4), (True, 0.0, 0.0), (False, 0.1, 1e-3)])
def test_fused_sgd_matches_torch_sgd(nesterov, damp, wd):
    """FusedSGD (one pass: update + bf16 operand re-cast) against torch.optim.SGD on identical gradients: parameters
    and momentum buffers to fp32 rounding, operands exactly the bf16 cast of the updated parameter, and the next
    forward identical to one that re-casts lazily."""
    import unet_torch_b200 as U
    from unet_torch_b200 import ops

    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 32, 32, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    y = torch.randint(0, 2, (2, 32, 32), device="cuda", generator=torch.Generator("cuda").manual_seed(8)).float()
    torch.manual_seed(5)
    net_a = U.UNet(3, 2).cuda().train()
    torch.manual_seed(5)
    net_b = U.UNet(3, 2).cuda().train()
    kw = dict(lr=0.05, momentum=0.9, dampening=damp, weight_decay=wd, nesterov=nesterov)
    opt_a = torch.optim.SGD(net_a.parameters(), **kw)
    opt_b = U.FusedSGD(net_b, **kw)
    for it in range(3):
        for net, opt in ((net_a, opt_a), (net_b, opt_b)):
            loss = U.calc_loss(net(x), y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
        # identical gradients go into both optimizers (isolates the optimizer arithmetic from bf16 chaos upstream)
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            pb.grad.copy_(pa.grad)
        opt_a.step()
        opt_b.step()
        for (k, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-6, atol=1e-8), (it, k, float((pa - pb).abs().max()))
            ba, bb = opt_a.state[pa]["momentum_buffer"], opt_b.state[pb]["momentum_buffer"]
            assert torch.allclose(ba, bb, rtol=2e-6, atol=1e-8), (it, k)
            pb.data.copy_(pa.data)  # keep the two runs on identical weights ...
        net_b._engine = None        # ... and let the next forward re-derive b's operands from them
    # operands written by the fused step == lazy re-cast of the same parameter
    torch.manual_seed(5)
    net_c = U.UNet(3, 2).cuda().train()
    opt_c = U.FusedSGD(net_c, **kw)
    loss = U.calc_loss(net_c(x), y, loss_type="dice_bce_mc")
    loss.backward()
    opt_c.step()
    eng = net_c._get_engine()
    for c1, c2 in eng.enc + eng.dec:
        for cb in (c1, c2):
            if cb.first:
                continue
            wf, wd_ = ops.prep_conv3x3_weight(cb.conv.weight.detach())
            assert cb._ver == (cb.conv.weight._version, cb.conv.weight.data_ptr())
            assert torch.equal(wf, cb.wf) and torch.equal(wd_, cb.wd)
    for u in eng.ups:
        wf, wd_ = ops.prep_convt2x2_weight(u.up.weight.detach())
        assert torch.equal(wf, u.wf) and torch.equal(wd_, u.wd)


@pytest.mark.parametrize("check_mode", [False, True])
def test_eval_mode_backward_uses_running_statistics(check_mode):
    """model.eval() WITH autograd (frozen-BatchNorm fine-tuning, saliency maps): BatchNorm normalises with the running
    statistics, so its backward is dy = gamma * rstd * da and dgamma / dbeta use the running mean / rstd (torch's
    batch_norm backward with training=False; Model.py:17,21 under module.eval()). Against the reference modules / oracle
    port on the host in fp32; buffers must stay untouched. check_mode: the fp32 generic engine, tight tolerance."""
    import unet_torch_b200 as U
    from oracle import ref_loader

    torch.manual_seed(3)
    net = U.UNet(3, 2)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 48, 64, generator=g)
    y = torch.randint(0, 2, (2, 48, 64), generator=g).float()
    net = net.cuda().eval().set_check_mode(check_mode)
    U.loss.CLASS_NUMBER = 2
    out = net(x.cuda())
    assert out.requires_grad
    loss = U.calc_loss(out, y.cuda(), loss_type="dice_bce_mc")
    loss.backward()
    torch.cuda.synchronize()
    for k, v in net.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k          # eval: parameters and buffers untouched by forward / backward
    rl, rloss, rg, _ = host_step(sd, x, y, 2, "dice_bce_mc", use_ref_modules=ref_loader.available(), training=False)
    e_logits, e_loss = rel_l2(out.detach(), rl), abs(float(loss) - rloss) / abs(rloss)
    errs = {k: rel_l2(p.grad, rg[k]) for k, p in net.named_parameters()}
    kw = max(errs, key=errs.get)
    print(f"eval backward (check_mode={check_mode}): logits {e_logits:.3e} loss {e_loss:.3e} grads median "
          f"{statistics.median(errs.values()):.3e} worst {errs[kw]:.3e} ({kw})")
    if check_mode:
        assert e_logits < 1e-4 and e_loss < 1e-4 and errs[kw] < 2e-3
    else:
        # no batch statistics in the loop -> no chaotic amplification: plain accumulated bf16 rounding
        el, eloss, eg, _ = host_step(sd, x, y, 2, "dice_bce_mc", emulate=True, training=False)
        yard = {k: rel_l2(eg[k], rg[k]) for k in rg}
        print(f"   bf16-storage emulation vs fp32: logits {rel_l2(el, rl):.3e} grads median {statistics.median(yard.values()):.3e} "
              f"worst {max(yard.values()):.3e}")
        assert e_logits <= 1.5 * rel_l2(el, rl) + 1e-4 and e_loss < 1e-2
        assert statistics.median(errs.values()) <= 1.5 * statistics.median(yard.values())
        assert all(errs[k] <= 1.5 * yard[k] + 0.02 for k in errs)


@pytest.mark.parametrize("wd,betas", [(1e-4, (0.9, 0.999)), (0.0, (0.8, 0.99))])
def test_fused_adam_matches_torch_adam(wd, betas):
    """FusedAdam (update + bf16 operand re-cast in one pass) against torch.optim.Adam (train.py:341-343) on identical
    gradients over three steps: parameters, both moment buffers and the step count; operands == lazy re-cast."""
    import unet_torch_b200 as U
    from unet_torch_b200 import ops

    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 32, 32, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    y = torch.randint(0, 2, (2, 32, 32), device="cuda", generator=torch.Generator("cuda").manual_seed(8)).float()
    torch.manual_seed(5)
    net_a = U.UNet(3, 2).cuda().train()
    torch.manual_seed(5)
    net_b = U.UNet(3, 2).cuda().train()
    kw = dict(lr=2e-3, betas=betas, eps=1e-8, weight_decay=wd)
    opt_a = torch.optim.Adam(net_a.parameters(), **kw)
    opt_b = U.FusedAdam(net_b, **kw)
    for it in range(3):
        for net, opt in ((net_a, opt_a), (net_b, opt_b)):
            loss = U.calc_loss(net(x), y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            pb.grad.copy_(pa.grad)
        opt_a.step()
        opt_b.step()
        for (k, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
            sa, sb = opt_a.state[pa], opt_b.state[pb]
            assert float(sa["step"]) == float(sb["step"]) == it + 1
            assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-5, atol=1e-10), (it, k)
            assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-5, atol=1e-14), (it, k)
            # the update is lr * m / (sqrt(v) + eps) ~ lr in magnitude: compare at that scale
            assert float((pa - pb).abs().max()) <= 2e-3 * 1e-3 + 2e-6 * float(pa.abs().max()), (it, k, float((pa - pb).abs().max()))
            pb.data.copy_(pa.data)
        net_b.refresh_operands(force=True)   # .data writes bypass the version counter
    eng = net_b._get_engine()
    loss = U.calc_loss(net_b(x), y, loss_type="dice_bce_mc")
    opt_b.zero_grad(set_to_none=True)
    loss.backward()
    opt_b.step()
    for c1, c2 in eng.enc + eng.dec:
        for cb in (c1, c2):
            if cb.first:
                continue
            wf, wd_ = ops.prep_conv3x3_weight(cb.conv.weight.detach())
            assert cb._ver == (cb.conv.weight._version, cb.conv.weight.data_ptr())
            assert torch.equal(wf, cb.wf) and torch.equal(wd_, cb.wd)
    for u in eng.ups:
        wf, wd_ = ops.prep_convt2x2_weight(u.up.weight.detach())
        assert torch.equal(wf, u.wf) and torch.equal(wd_, u.wd)
    sd = opt_b.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and sd["param_groups"][0]["betas"] == betas


def test_cuda_graph_share_grads_is_identical_and_copy_free():
    """enable_cuda_graphs(True, share_grads=True): .grad is the graph's own static gradient tensor (no per-step copy), the
    training trajectory is bit-identical to the copying mode, and a foreign gradient already in .grad is accumulated."""
    import unet_torch_b200 as U

    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 32, 48, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    y = torch.randint(0, 2, (2, 32, 48), device="cuda", generator=torch.Generator("cuda").manual_seed(4)).float()

    def run(share):
        torch.manual_seed(11)
        net = U.UNet(3, 2).cuda().train().enable_cuda_graphs(True, share_grads=share)
        opt = U.FusedSGD(net, lr=0.05, momentum=0.9)
        ptrs, losses = [], []
        for _ in range(4):
            loss = U.calc_loss(net(x), y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
            ptrs.append(net.up1.up.weight.grad.data_ptr())
            losses.append(float(loss))
            opt.step()
        return net, losses, ptrs

    net_a, la, pa = run(False)
    net_b, lb, pb = run(True)
    assert la == lb
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert torch.equal(a, b), k
    assert len(set(pb[1:])) == 1            # graph replays: always the same static buffer, never a copy
    # accumulation into an existing foreign gradient
    loss = U.calc_loss(net_b(x), y, loss_type="dice_bce_mc")
    net_b.zero_grad(set_to_none=True)
    w = net_b.outc.conv.bias
    w.grad = torch.ones_like(w)
    loss.backward()
    ref = U.calc_loss(net_a(x), y, loss_type="dice_bce_mc")
    net_a.zero_grad(set_to_none=True)
    ref.backward()
    assert torch.allclose(w.grad, net_a.outc.conv.bias.grad + 1.0)
