"""-m gpu: the attention gates of UNet_attention (Model.py:257-296) on the tensor-core engine.

Operator level (tight, same bf16 inputs on both sides): the 1x1-convolution GEMMs (forward / backward-data / weight
gradient, BatchNorm statistics rows, channel padding), the ConvTranspose2d statistics epilogue, the strided fp32 GEMM of the
weight-side composition, and the five bandwidth-bound gate kernels against torch autograd over the same formulas.
Network level: full-width UNet_attention on the tensor-core engine against (a) the UNMODIFIED reference's fp64 logits / loss /
small-parameter gradients with the reference's own bf16-autocast error as the yardstick
(tests/golden/ref_attention_bf16_yardstick.pt, oracle/make_golden_attention_yardstick.py) and (b) the library's fp32 check
engine (itself pinned to the reference by test_gpu_attention.py) for the large gradients that do not fit a fixture."""
import statistics

import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _rand_bf16(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(BF16)


@pytest.mark.parametrize("n,h,w,cin,cout,creal", [(2, 24, 40, 128, 64, 32), (1, 16, 16, 64, 256, 256), (3, 8, 16, 512, 128, 128)])
def test_conv1x1_gemm_forward_stats_backward(n, h, w, cin, cout, creal):
    from unet_torch_b200 import ops

    torch.manual_seed(cin + cout)
    x = _rand_bf16(n, h, w, cin)
    wt = _rand_bf16(cout, cin, scale=cin ** -0.5)
    wt[creal:] = 0
    bias = torch.randn(cout, device="cuda")
    bias[creal:] = 0
    y = torch.empty((n, h, w, cout), dtype=BF16, device="cuda")
    rows = ops.conv1x1_stat_rows(n, h, w)
    stats = torch.full((rows * 2 * creal,), float("nan"), device="cuda")
    ops.conv1x1(x, wt, bias, y, stats, creal)
    want = x.float().reshape(-1, cin) @ wt.float().t() + bias
    assert rel_l2(y.float().reshape(-1, cout), want) < 4e-3
    assert float(y[..., creal:].float().abs().max()) == 0.0 if creal < cout else True
    st = stats.view(rows, 2, creal).double().sum(0)
    yf = y.float().reshape(-1, cout)[:, :creal].double()
    assert torch.allclose(st[0], yf.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(st[1], (yf * yf).sum(0), rtol=1e-4, atol=1e-2)
    # backward-data = the same GEMM with the transposed operand, no bias
    dy = _rand_bf16(n, h, w, cout)
    dy[..., creal:] = 0
    dx = torch.empty((n, h, w, cin), dtype=BF16, device="cuda")
    ops.conv1x1(dy, wt.t().contiguous(), None, dx)
    assert rel_l2(dx.float().reshape(-1, cin), dy.float().reshape(-1, cout) @ wt.float()) < 4e-3
    # weight gradient (fp32), real rows only
    dw = torch.full((creal, cin, 1, 1), float("nan"), device="cuda")
    ops.conv1x1_wgrad(x, dy, dw)
    want_dw = dy.float().reshape(-1, cout)[:, :creal].t() @ x.float().reshape(-1, cin)
    assert rel_l2(dw.view(creal, cin), want_dw) < 2e-4


@pytest.mark.parametrize("n,h,w,cin,cup,creal", [(2, 8, 24, 128, 64, 32), (1, 4, 4, 1024, 256, 256), (2, 16, 16, 256, 64, 64)])
def test_convt2x2_with_statistics_rows(n, h, w, cin, cup, creal):
    from unet_torch_b200 import ops

    torch.manual_seed(cin)
    x = _rand_bf16(n, h, w, cin)
    wt = torch.randn(cin, cup, 2, 2, device="cuda") * cin ** -0.5
    wt[:, creal:] = 0
    bias = torch.randn(cup, device="cuda")
    bias[creal:] = 0
    wf, _ = ops.prep_convt2x2_weight(wt)
    out = torch.empty((n, 2 * h, 2 * w, cup), dtype=BF16, device="cuda")
    rows = ops.convt2x2_stat_rows(n, h, w)
    stats = torch.full((rows * 2 * creal,), float("nan"), device="cuda")
    ops.convt2x2_stats(x, wf, bias, out, stats, creal)
    want = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.to(BF16).float(), bias, stride=2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), want) < 4e-3
    plain = torch.empty_like(out)
    ops.convt2x2(x, wf, bias, plain)
    assert torch.equal(plain, out)
    st = stats.view(rows, 2, creal).double().sum(0)
    of = out.float().reshape(-1, cup)[:, :creal].double()
    assert torch.allclose(st[0], of.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(st[1], (of * of).sum(0), rtol=1e-4, atol=1e-2)


def test_strided_sgemm_composition_shapes():
    """The three weight-side products of a gate (compose, dW_up, dW_q) and the two bias products, in the layouts the engine
    uses, against torch.einsum in fp64."""
    from unet_torch_b200 import ops

    torch.manual_seed(3)
    for cq, ch, chp in ((128, 32, 64), (256, 64, 64), (200, 24, 64), (1024, 256, 256)):
        w_up = torch.randn(cq, cq, 2, 2, device="cuda")
        w_q = torch.randn(ch, cq, 1, 1, device="cuda")
        b_up, b_q = torch.randn(cq, device="cuda"), torch.randn(ch, device="cuda")
        wc = torch.zeros(cq, chp, 2, 2, device="cuda")
        ops.sgemm_strided(w_up, w_q, wc, cq, ch, cq, (4 * cq, 4), (1, cq), (4 * chp, 4), batch=4, batch_strides=(1, 0, 1))
        want = torch.einsum("cdij,hd->chij", w_up.double(), w_q.double().view(ch, cq))
        assert rel_l2(wc[:, :ch], want) < 1e-5 and (chp == ch or float(wc[:, ch:].abs().max()) == 0.0)
        bc = torch.zeros(chp, device="cuda")
        ops.sgemm_strided(w_q, b_up, bc, ch, 1, cq, (cq, 1), (1, 1), (1, 1), bias_m=b_q)
        assert rel_l2(bc[:ch], w_q.double().view(ch, cq) @ b_up.double() + b_q.double()) < 1e-5
        dwc = torch.randn(cq, chp, 2, 2, device="cuda")
        dwup = torch.full((cq, cq, 2, 2), float("nan"), device="cuda")
        ops.sgemm_strided(dwc, w_q, dwup, cq, cq, ch, (4 * chp, 4), (cq, 1), (4 * cq, 4), batch=4, batch_strides=(1, 0, 1))
        assert rel_l2(dwup, torch.einsum("chij,hd->cdij", dwc[:, :ch].double(), w_q.double().view(ch, cq))) < 1e-5
        dwq = torch.full((ch, cq, 1, 1), float("nan"), device="cuda")
        for ij in range(4):
            ops.sgemm_strided(dwc, w_up, dwq, ch, cq, cq, (4, 4 * chp), (4 * cq, 4), (cq, 1), accumulate=ij > 0, offsets=(ij, ij, 0))
        want_dwq = torch.einsum("chij,cdij->hd", dwc[:, :ch].double(), w_up.double())
        assert rel_l2(dwq.view(ch, cq), want_dwq) < 1e-5
        part = torch.full((4, ch, cq), float("nan"), device="cuda")  # the engine's form: four partial products in one launch
        ops.sgemm_strided(dwc, w_up, part, ch, cq, cq, (4, 4 * chp), (4 * cq, 4), (cq, 1), batch=4, batch_strides=(1, 1, ch * cq))
        dwq2 = torch.full((ch, cq, 1, 1), float("nan"), device="cuda")
        ops.sum_batches(part, dwq2)
        assert rel_l2(dwq2.view(ch, cq), want_dwq) < 1e-5
        dbias = torch.randn(2 * ch, device="cuda")
        dbup = torch.full((cq,), float("nan"), device="cuda")
        ops.sgemm_strided(w_q, dbias, dbup, cq, 1, ch, (1, cq), (1, 1), (1, 1))
        assert rel_l2(dbup, w_q.double().view(ch, cq).t() @ dbias[:ch].double()) < 1e-5


def _bn_stats(t):  # t [P, C] fp64 -> mean, rstd (biased variance, eps 1e-5)
    m = t.mean(0)
    v = ((t - m) ** 2).mean(0)
    return m, (v + 1e-5).rsqrt()


@pytest.mark.parametrize("c,cx,n,h,w", [(32, 64, 2, 8, 24), (64, 128, 1, 16, 16), (256, 512, 2, 4, 12), (128, 256, 3, 8, 8)])
def test_gate_kernels_against_autograd(c, cx, n, h, w):
    """gate_psi_fwd / gate_apply_fwd / gate_apply_bwd / gate_bwd_reduce / gate_bwd_apply on padded maps, against torch
    autograd (fp64) through the same formulas from the same bf16 tensors, batch-statistics BatchNorms included."""
    from unet_torch_b200 import ops

    torch.manual_seed(c + cx)
    cp = max(64, c)
    P = n * h * w
    q1 = torch.zeros((n, h, w, cp), dtype=BF16, device="cuda")
    x1 = torch.zeros((n, h, w, cp), dtype=BF16, device="cuda")
    q1[..., :c] = _rand_bf16(n, h, w, c, scale=1.5) + 0.3
    x1[..., :c] = _rand_bf16(n, h, w, c, scale=0.7) - 0.2
    xs = _rand_bf16(n, h, w, cx)
    g = _rand_bf16(n, h, w, cx)
    gq, bq = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
    gx, bx = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
    wp, bp = torch.randn(c, device="cuda") * c ** -0.5, torch.randn(1, device="cuda")
    gp, betap = torch.rand(1, device="cuda") + 0.5, torch.randn(1, device="cuda") * 0.3
    # ---- fp64 autograd reference
    d = lambda t: t.double().detach().clone().requires_grad_(True)  # noqa: E731
    q1d, x1d, xsd = d(q1[..., :c].reshape(P, c)), d(x1[..., :c].reshape(P, c)), d(xs.reshape(P, cx))
    gqd, bqd, gxd, bxd, wpd, bpd, gpd, betapd = map(d, (gq, bq, gx, bx, wp, bp, gp, betap))
    mq, rq = _bn_stats(q1d)
    mx, rx = _bn_stats(x1d)
    e = torch.relu(gqd * (q1d - mq) * rq + bqd + gxd * (x1d - mx) * rx + bxd)
    s_ref = e @ wpd + bpd
    mp, rp = _bn_stats(s_ref[:, None])
    a = torch.sigmoid(gpd * (s_ref - mp) * rp + betapd)
    out_ref = xsd * a[:, None]
    (out_ref * g.double().reshape(P, cx)).sum().backward()
    # ---- kernels
    f = lambda t: t.detach().float().contiguous()  # noqa: E731
    aff_q = (f(gqd * rq), f(bqd - mq * gqd * rq), f(mq), f(rq))
    aff_x = (f(gxd * rx), f(bxd - mx * gxd * rx), f(mx), f(rx))
    aff_p = (f(gpd * rp), f(betapd - mp * gpd * rp), f(mp), f(rp))
    s = torch.empty((n, h, w), device="cuda")
    rows = ops.gate_stat_rows(P, c)
    st = torch.full((rows * 2,), float("nan"), device="cuda")
    ops.gate_psi_fwd(q1[..., :c], x1[..., :c], aff_q[0], aff_q[1], aff_x[0], aff_x[1], wp, bp, s, st)
    assert rel_l2(s.reshape(P), s_ref) < 2e-5
    st = st.view(rows, 2).double().sum(0)
    assert torch.allclose(st[0], s.double().sum(), rtol=1e-4, atol=1e-3) and torch.allclose(st[1], (s.double() ** 2).sum(), rtol=1e-4)
    cat = torch.zeros((n, h, w, 2 * cx), dtype=BF16, device="cuda")
    ops.gate_apply_fwd(xs, s, aff_p[0], aff_p[1], cat[..., :cx])
    assert rel_l2(cat[..., :cx].float().reshape(P, cx), out_ref) < 4e-3 and float(cat[..., cx:].float().abs().max()) == 0.0
    dxs = torch.empty((n, h, w, cx), dtype=BF16, device="cuda")
    dz = torch.empty((n, h, w), device="cuda")
    sums2 = torch.empty(2, dtype=torch.float64, device="cuda")
    gcat = torch.zeros((n, h, w, 2 * cx), dtype=BF16, device="cuda")
    gcat[..., :cx] = g
    ops.gate_apply_bwd(gcat[..., :cx], xs, s, *aff_p, dxs, dz, sums2)
    assert rel_l2(dxs.float().reshape(P, cx), xsd.grad) < 4e-3
    # the engine's form: no dx store here, g * A added to the W_x backward-data by gate_dx in one pass
    dz2, sums2b = torch.full_like(dz, float("nan")), torch.empty_like(sums2)
    ops.gate_apply_bwd(gcat[..., :cx], xs, s, *aff_p, None, dz2, sums2b)
    assert torch.equal(dz2, dz) and torch.equal(sums2b, sums2)
    acc = _rand_bf16(n, h, w, cx)
    want_dx = xsd.grad + acc.double().reshape(P, cx)
    ops.gate_dx(gcat[..., :cx], s, aff_p[0], aff_p[1], acc)
    assert rel_l2(acc.float().reshape(P, cx), want_dx) < 4e-3
    ds = torch.empty((n, h, w), device="cuda")
    sums = torch.empty(4 * c + 8, dtype=torch.float64, device="cuda")
    ops.gate_bwd_reduce(q1[..., :c], x1[..., :c], aff_q, aff_x, wp, s, dz, gp, aff_p[2], aff_p[3], sums2, P, ds, sums)
    outs = [torch.full_like(t, float("nan")) for t in (gq, bq, gx, bx, wp, bp, gp, betap)]
    dbias = torch.empty(2 * c, device="cuda")
    ops.gate_bwd_apply(q1[..., :c], x1[..., :c], aff_q, aff_x, gq, gx, wp, ds, sums, None, sums2, P, outs, dbias)
    torch.cuda.synchronize()
    for got, want, name in zip(outs, (gqd, bqd, gxd, bxd, wpd, bpd, gpd, betapd), ("dgq", "dbq", "dgx", "dbx", "dwp", "dbp", "dgp", "dbetap")):
        if name == "dbp":  # analytically zero (a bias in front of a BatchNorm): rounding noise of the sum on both sides
            assert abs(float(got)) <= 1e-5 * float(ds.abs().sum()) + 1e-6, float(got)
            continue
        err = float((got.double() - want.grad).norm()) / float(want.grad.norm())
        assert err < 2e-3, (name, err)
    assert rel_l2(q1[..., :c].float().reshape(P, c), q1d.grad) < 6e-3
    assert rel_l2(x1[..., :c].float().reshape(P, c), x1d.grad) < 6e-3
    assert float(q1[..., c:].float().abs().max()) == 0.0 if cp > c else True
    want_db = torch.cat([q1[..., :c].float().reshape(P, c).sum(0), x1[..., :c].float().reshape(P, c).sum(0)])
    assert torch.allclose(dbias, want_db, rtol=1e-3, atol=1e-3 * float(want_db.abs().max()) + 1e-6)


def _step(net, U, x, y, loss_type):
    out = net(x)
    pred = torch.relu(out) if loss_type.startswith("mse") else out
    loss = U.calc_loss(pred, y, loss_type=loss_type)
    net.zero_grad(set_to_none=True)
    loss.backward()
    return out.detach(), float(loss), {k: p.grad.detach().clone() for k, p in net.named_parameters()}


@pytest.mark.parametrize("case", ["att_w64_c3_k2_dicebce_64", "att_w64_c1_k3_msemc_64x96", "att_w64_c3_k2_dicebce_128"])
def test_attention_tensor_core_engine_against_reference_yardstick(golden, case):
    import unet_torch_b200 as U
    from unet_torch_b200.model import UNetEngine

    g = golden("ref_attention_bf16_yardstick.pt")[case]
    ch, ncls, width, n, h, w, seed, loss_type = g["cfg"]
    yard = g["bf16_autocast"]
    U.loss.CLASS_NUMBER = ncls
    nets = []
    for check in (False, True):
        torch.manual_seed(seed)
        net = U.UNet_attention(ch, ncls, width)
        for k, v in net.state_dict().items():
            if v.is_floating_point():
                assert abs(float(v.double().sum()) - g["sd0_checksum"][k]) <= 1e-9 * max(1.0, abs(g["sd0_checksum"][k])), k
        nets.append(net.cuda().train().set_check_mode(check))
    fast, chk = nets
    x, y = g["x"].cuda(), g["y"].cuda()
    assert isinstance(fast._engine_for(x), UNetEngine) and not isinstance(chk._engine_for(x), UNetEngine)
    o_f, l_f, g_f = _step(fast, U, x, y, loss_type)
    o_c, l_c, g_c = _step(chk, U, x, y, loss_type)
    torch.cuda.synchronize()
    # (a) against the reference's fp64 run, yardstick = the reference's own bf16-autocast error
    e_logits, e_loss = rel_l2(o_f, g["logits64"]), abs(l_f - g["loss64"]) / abs(g["loss64"])
    assert rel_l2(o_c, g["logits64"]) < 1e-4 and abs(l_c - g["loss64"]) / abs(g["loss64"]) < 1e-4  # the fp32 engine is exact
    floor = g["grad_floor"]
    err = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm()) / max(float(b.double().norm()), floor)  # noqa: E731
    small = {k: err(g_f[k], v) for k, v in g["small_grads64"].items()}
    # (b) every gradient against the fp32 check engine (which is as close to the fp64 reference as the reference's own fp32 run:
    # asserted on the small ones here - even fp32 moves some gradients of this network by percents)
    fp32_worst = max(g["fp32"]["grads"].values())
    for k, v in g["small_grads64"].items():
        assert err(g_c[k], v) <= 3 * fp32_worst + 2e-3, (k, err(g_c[k], v), fp32_worst)
    errs = {k: err(g_f[k], g_c[k]) for k in g_f}
    med, med_y = statistics.median(errs.values()), statistics.median(yard["grads"].values())
    worst = max((k for k in errs if ".psi.1." not in k), key=lambda k: errs[k] / (yard["grads"][k] + 0.02))
    print(f"{case}: logits {e_logits:.3e} (reference bf16 {yard['logits']:.3e}) loss {e_loss:.3e} ({yard['loss']:.3e}); grads median "
          f"{med:.3e} ({med_y:.3e}); worst vs yardstick {worst}: {errs[worst]:.3e} ({yard['grads'][worst]:.3e})")
    assert e_logits <= 1.5 * yard["logits"] and e_loss <= max(1.5 * yard["loss"], 2e-3)
    assert med <= 1.5 * med_y
    # the eight one-element gradients (BatchNorm2d(1) affine of every psi) are single draws with no averaging inside a tensor,
    # some of them close to zero: they are held to the yardstick as a group, by ABSOLUTE error (root mean square over the
    # eight; the yardstick's relative errors are turned back into absolute ones with the fp64 norms), every other parameter on
    # its own
    single = [k for k in errs if ".psi.1." in k]
    scale = {k: max(g["grad_norm64"][k], floor) for k in single}
    rms = lambda d: (sum(d[k] ** 2 for k in single) / len(single)) ** 0.5  # noqa: E731
    abs_y = {k: yard["grads"][k] * scale[k] for k in single}
    abs_f = {k: float((g_f[k].double() - g_c[k].double()).norm()) for k in single}
    abs_s = {k: small[k] * scale[k] for k in single}
    assert len(single) == 8 and rms(abs_f) <= 1.5 * rms(abs_y) + floor, (rms(abs_f), rms(abs_y))
    assert rms(abs_s) <= 1.5 * rms(abs_y) + floor, (rms(abs_s), rms(abs_y))
    assert all(errs[k] <= 1.5 * yard["grads"][k] + 0.02 for k in errs if k not in single), worst
    assert all(small[k] <= 1.5 * yard["grads"][k] + 0.02 for k in small if k not in single)
    # BatchNorm buffers (training statistics of the three BatchNorms of every gate included)
    sd_f, sd_c = fast.state_dict(), chk.state_dict()
    for k, v in g["buffers1"].items():
        assert rel_l2(sd_c[k], v) < 1e-4, k
        assert rel_l2(sd_f[k], v) < 0.02, (k, rel_l2(sd_f[k], v))
    assert all(int(sd_f[k]) == 1 for k in sd_f if k.endswith("num_batches_tracked"))
    # eval forward (running statistics), no autograd
    fast.eval()
    chk.eval()
    with torch.no_grad():
        e_f, e_c = fast(x), chk(x)
    assert rel_l2(e_c, e_f) <= 1.5 * yard["logits_eval"] + 2e-3, rel_l2(e_c, e_f)
    # fused inference heads (OutConv + softmax/argmax, + sigmoid threshold, + relu / divisor) run through the gated engine too
    with torch.no_grad():
        mask = fast.predict(x)
        assert mask.dtype == torch.uint8 and (mask.long() != e_f.argmax(1)).float().mean() < 2e-3
        assert (fast.predict_binary(x, 0.5).bool() != (torch.sigmoid(e_f[:, 0]) >= 0.5)).float().mean() < 2e-3
        dens, sums = fast.predict_density(x, 200.0)
        assert rel_l2(dens, torch.relu(e_f) / 200.0) < 5e-3 and rel_l2(sums, (torch.relu(e_f) / 200.0).double().sum((2, 3))) < 5e-3


def test_attention_eval_mode_backward_and_graph_replay():
    """module.eval() under autograd (frozen BatchNorm statistics) on the tensor-core engine against the fp32 check engine;
    CUDA-graph replay + FusedSGD reproduces the eager steps."""
    import unet_torch_b200 as U

    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 64, 64, device="cuda")
    y = (torch.rand(2, 64, 64, device="cuda") > 0.5).float()
    nets = []
    for check in (False, True):
        torch.manual_seed(11)
        net = U.UNet_attention(3, 2).cuda().train().set_check_mode(check)
        with torch.no_grad():
            net(x)  # one training forward: non-trivial running statistics
        nets.append(net.eval())
    (o_f, l_f, g_f), (o_c, l_c, g_c) = (_step(net, U, x, y, "dice_bce_mc") for net in nets)
    assert rel_l2(o_f, o_c) < 0.05
    errs = {k: rel_l2(g_f[k], g_c[k]) for k in g_f if float(g_c[k].norm()) > 1e-6 * max(float(v.norm()) for v in g_c.values())}
    assert statistics.median(errs.values()) < 0.25, statistics.median(errs.values())
    # behind a FROZEN BatchNorm the bias gradients of the gate convolutions are not zero: the composed-weight backward must then
    # carry the rank-one term db' (x) b_up into dW_q (tests/test_host_logic.py::test_gate_weight_composition_algebra)
    gate_w = {k: v for k, v in errs.items() if ".W_q.0." in k or ".up." in k and "attenion" in k}
    print("frozen-BN gate weight-side gradients vs the fp32 engine:", {k: round(v, 4) for k, v in gate_w.items()})
    assert len(gate_w) == 16 and max(gate_w.values()) < 0.3, gate_w
    # graphs: three FusedSGD steps eager vs replayed
    states = []
    for graphs in (False, True):
        torch.manual_seed(5)
        net = U.UNet_attention(3, 2).cuda().train()
        net.enable_cuda_graphs(graphs, share_grads=graphs)
        opt = U.FusedSGD(net, lr=0.01, momentum=0.9, weight_decay=1e-4)
        losses = []
        for _ in range(4):
            loss = U.calc_loss(net(x), y, loss_type="dice_bce_mc")
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        states.append((losses, {k: v.clone() for k, v in net.state_dict().items()}))
    (la, sa), (lb, sb) = states
    assert la[-1] < la[0]
    assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(la, lb)), (la, lb)
    for k in sa:
        if sa[k].is_floating_point():
            assert torch.allclose(sa[k], sb[k], rtol=1e-4, atol=1e-6), k


def test_fold_rows_and_folded_statistics_path(monkeypatch):
    """fold_rows (pre-reduction of one-row-per-tile statistics) against a torch sum, and a network step with the folding
    threshold lowered so that small maps take the folded path: identical BatchNorm buffers up to fp32 summation order."""
    import unet_torch_b200 as U
    from unet_torch_b200 import _lib, ops

    torch.manual_seed(1)
    for rows, ncols in ((5000, 64), (777, 2), (9001, 512), (300, 256)):
        part = torch.randn(rows, ncols, device="cuda")
        out_rows = min(296, rows)
        out = torch.full((out_rows, ncols), float("nan"), device="cuda")
        _lib.call("b200unet_fold_rows", part.data_ptr(), rows, ncols, out.data_ptr(), out_rows, torch.cuda.current_stream().cuda_stream)
        want = torch.stack([part[b::out_rows].double().sum(0) for b in range(out_rows)])
        assert torch.allclose(out.double(), want, rtol=1e-5, atol=1e-4)
    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 64, 64, device="cuda")
    y = (torch.rand(2, 64, 64, device="cuda") > 0.5).float()
    res = []
    for above in (ops.FOLD_ROWS_ABOVE, 8):
        monkeypatch.setattr(ops, "FOLD_ROWS_ABOVE", above)
        monkeypatch.setattr(ops, "FOLD_ROWS_TO", 296 if above > 8 else 7)
        torch.manual_seed(2)
        net = U.UNet_attention(3, 2).cuda().train()
        out = net(x)
        U.calc_loss(out, y, loss_type="dice_bce_mc").backward()
        res.append((out.detach(), {k: v.clone() for k, v in net.state_dict().items() if "running" in k}))
    assert rel_l2(res[1][0], res[0][0]) < 2e-2
    for k in res[0][1]:
        assert rel_l2(res[1][1][k], res[0][1][k]) < 1e-3, k


@pytest.mark.parametrize("frozen", [False, True])
def test_replay_every_gate_operator_from_the_engine_trace(frozen):
    """Operator-by-operator replay of the four attention gates of one step (tests/gate_check.py): the engine records what each
    gate read and wrote (`UNetEngine.trace`), the host recomputes every operator from the recorded inputs in the reference's
    two-module form (ConvTranspose2d -> 1x1 conv, fp64) at one-operator tolerance - forward maps, BatchNorm statistics, s, the
    gated skip; backward dz, ds, both BatchNorm backwards, every parameter gradient including the split of the composed
    ConvTranspose2d's gradient into up / W_q, and the two data gradients. frozen: module.eval() under autograd (running
    statistics), where the bias gradients of the gate convolutions do not vanish."""
    import unet_torch_b200 as U
    from gate_check import check_gate

    def nchw(t):
        return t.float().permute(0, 3, 1, 2).contiguous().cpu().double()

    torch.manual_seed(21)
    net = U.UNet_attention(3, 2).cuda().train()
    U.loss.CLASS_NUMBER = 2
    x = torch.randn(2, 3, 64, 64, device="cuda")
    y = (torch.rand(2, 64, 64, device="cuda") > 0.5).float()
    if frozen:
        with torch.no_grad():
            net(x)      # one training forward: non-trivial running statistics
        net.eval()
    eng = net._get_engine()
    eng.trace = []
    U.calc_loss(net(x), y, loss_type="dice_bce_mc").backward()
    torch.cuda.synchronize()
    trace, eng.trace = eng.trace, None
    fwd = {id(r["gate"]): r for k, r in trace if k == "gate"}
    bwd = {id(r["gate"]): r for k, r in trace if k == "gate_bwd"}
    assert len(fwd) == 4 and set(fwd) == set(bwd)
    c = lambda t: t.detach().double().cpu()  # noqa: E731
    for key, f in fwd.items():
        b, att = bwd[key], f["gate"].att
        names = dict(W_up=att.up.weight, b_up=att.up.bias, W_q=att.W_q[0].weight, b_q=att.W_q[0].bias, W_x=att.W_x[0].weight,
                     b_x=att.W_x[0].bias, gq=att.W_q[1].weight, bq=att.W_q[1].bias, gx=att.W_x[1].weight, bx=att.W_x[1].bias,
                     wpsi=att.psi[0].weight, bpsi=att.psi[0].bias, gp=att.psi[1].weight, bp=att.psi[1].bias)
        rec = dict(q=nchw(f["q"]), x=nchw(f["x"]), q1=nchw(f["q1"]), x1=nchw(f["x1"]), s=c(f["s"]), out=nchw(f["out"]),
                   aq=tuple(c(t) for t in f["aq"][:4]), ax=tuple(c(t) for t in f["ax"][:4]), ap=tuple(c(t) for t in f["ap"][:4]),
                   count=float(f["aq"][4]), frozen=frozen, params={k: c(p) for k, p in names.items()}, g=nchw(b["g"]),
                   dxs=nchw(b["dxs"]), dq=nchw(b["dq"]), dq1=nchw(b["dq1"]), dx1=nchw(b["dx1"]), dz=c(b["dz"]), ds=c(b["ds"]),
                   grads={k: c(b["grads"][p]) for k, p in names.items()})
        assert rec["count"] == rec["s"].numel() and f["training"] == (not frozen)
        bad = check_gate(rec, tol_map=6e-3, tol_sum=2e-4, tol_grad=4e-3, weights_rounded=lambda t: t.to(torch.bfloat16).double())
        assert bad == [], (f["gate"].ch, bad)
        # every recorded gradient IS the parameter's .grad
        for k, p in names.items():
            assert torch.equal(p.grad, b["grads"][p]), k
