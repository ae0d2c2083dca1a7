"""One-operator-deep host check of an attention gate (reference Attention_block, Model.py:257-296) from recorded tensors.

`check_gate(rec)` takes what one gate read and wrote in a forward + backward pass - as host fp64 tensors, maps in NCHW - and
recomputes every operator of the gate from the RECORDED inputs of that operator with torch CPU ops in the reference's
two-module form (ConvTranspose2d -> 1x1 conv, not the engine's composed weight), so each comparison is one operator deep and its
tolerance is that operator's rounding. It returns [(what, error, tolerance)] for every comparison that fails.

The records come from `UNetEngine.trace` on the GPU (tests/test_gpu_gate.py) or - to validate this checker itself without a GPU -
from an fp64 autograd run of the reference form (tests/test_host_logic.py)."""
import torch
import torch.nn.functional as F


def rel(a, b, floor=0.0):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), floor, 1e-300))


def gate_reference_run(p, q, x, g, frozen=False, stats=None, eps=1e-5):
    """fp64 autograd run of the reference's gate. p: parameter dict (W_up, b_up, W_q, b_q, W_x, b_x, gq, bq, gx, bx, wpsi, bpsi,
    gp, bp); stats: {(mean, var)} per BatchNorm ("q", "x", "p") when frozen. Returns the record `check_gate` consumes."""
    p = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    q = q.double().clone().requires_grad_(True)
    x = x.double().clone().requires_grad_(True)

    def bn(t, gamma, beta, key):
        if frozen:
            m, v = (s.double() for s in stats[key])
        else:
            m, v = t.mean((0, 2, 3)), t.var((0, 2, 3), unbiased=False)
        r = (v + eps).rsqrt()
        sh = (1, -1, 1, 1)
        return gamma.view(sh) * (t - m.view(sh)) * r.view(sh) + beta.view(sh), (gamma * r, beta - m * gamma * r, m, r)

    q1 = F.conv2d(F.conv_transpose2d(q, p["W_up"], p["b_up"], stride=2), p["W_q"], p["b_q"])
    x1 = F.conv2d(x, p["W_x"], p["b_x"])
    q1.retain_grad()
    x1.retain_grad()
    Q1, aq = bn(q1, p["gq"], p["bq"], "q")
    X1, ax = bn(x1, p["gx"], p["bx"], "x")
    e = torch.relu(Q1 + X1)
    s = F.conv2d(e, p["wpsi"].view(1, -1, 1, 1), p["bpsi"])
    s.retain_grad()
    z, ap = bn(s, p["gp"], p["bp"], "p")
    z.retain_grad()
    out = x * torch.sigmoid(z)
    (out * g.double()).sum().backward()
    d = lambda t: t.detach()  # noqa: E731
    return dict(q=d(q), x=d(x), q1=d(q1), x1=d(x1), s=d(s)[:, 0], out=d(out), aq=tuple(map(d, aq)), ax=tuple(map(d, ax)),
                ap=tuple(map(d, ap)), count=float(s.numel()), frozen=frozen, params={k: d(v) for k, v in p.items()}, g=g.double(),
                dxs=d(x.grad), dq=d(q.grad), dq1=d(q1.grad), dx1=d(x1.grad), dz=d(z.grad)[:, 0], ds=d(s.grad)[:, 0],
                grads={k: d(v.grad) for k, v in p.items()})


def check_gate(rec, tol_map=3e-3, tol_sum=1e-4, tol_grad=3e-3, weights_rounded=None):
    """rec: see gate_reference_run. tol_map: bf16-stored maps; tol_sum: fp32 sums / per-pixel fp32 values; tol_grad: parameter
    gradients. weights_rounded: optional function applied to conv weights before the host GEMMs (the engine rounds its GEMM
    operands to bf16)."""
    wr = weights_rounded or (lambda t: t)
    P = rec["params"]
    bad = []

    def cmp(what, got, want, tol, floor=0.0):
        e = rel(got, want, floor)
        if not e <= tol:
            bad.append((what, e, tol))

    sh = (1, -1, 1, 1)
    q, x, q1, x1, s, g = (rec[k].double() for k in ("q", "x", "q1", "x1", "s", "g"))
    sq, tq, mq, rq = (t.double() for t in rec["aq"])
    sx, tx, mx, rx = (t.double() for t in rec["ax"])
    sp, tp, mp, rp = (t.double() for t in rec["ap"])
    # ---- forward, operator by operator from the recorded inputs
    w_up, w_q, w_x = P["W_up"].double(), P["W_q"].double(), P["W_x"].double()
    cmp("q1 = W_q(up(q))", q1, F.conv2d(F.conv_transpose2d(q, w_up, P["b_up"].double(), stride=2), w_q, P["b_q"].double()), 2 * tol_map)
    cmp("x1 = W_x(x)", x1, F.conv2d(x, wr(w_x), P["b_x"].double()), tol_map)
    if not rec["frozen"]:
        for name, t, m, r in (("q", q1, mq, rq), ("x", x1, mx, rx), ("p", s[:, None], mp, rp)):
            cmp(f"BN_{name} mean", m, t.mean((0, 2, 3)), tol_sum, floor=0.1 * float(t.std()))
            cmp(f"BN_{name} rstd", r, (t.var((0, 2, 3), unbiased=False) + 1e-5).rsqrt(), tol_sum)
    for name, (sc, shf, m, r), gamma, beta in (("q", (sq, tq, mq, rq), P["gq"], P["bq"]), ("x", (sx, tx, mx, rx), P["gx"], P["bx"]),
                                               ("p", (sp, tp, mp, rp), P["gp"], P["bp"])):
        cmp(f"BN_{name} scale", sc, gamma.double() * r, tol_sum)
        cmp(f"BN_{name} shift", shf, beta.double() - m * gamma.double() * r, tol_sum, floor=1e-3 * float(gamma.double().norm()))
    e = torch.relu(sq.view(sh) * q1 + tq.view(sh) + sx.view(sh) * x1 + tx.view(sh))
    wpsi = P["wpsi"].double().view(-1)
    cmp("s = psi(E)", s, (e * wpsi.view(sh)).sum(1) + P["bpsi"].double(), tol_sum)
    a = torch.sigmoid(sp * s + tp)
    cmp("out = x * A", rec["out"], x * a[:, None], tol_map)
    # ---- backward
    dz_ref = (g * x).sum(1) * a * (1 - a)
    cmp("dz", rec["dz"], dz_ref, tol_sum)
    m = rec["count"]
    dz = rec["dz"].double()
    shat = (s - mp) * rp
    if rec["frozen"]:
        ds_ref = P["gp"].double() * rp * dz
    else:
        ds_ref = P["gp"].double() * rp * (dz - dz.sum() / m - shat * (dz * shat).sum() / m)
    cmp("ds = BN_p backward", rec["ds"], ds_ref, 10 * tol_sum)
    ds = rec["ds"].double()
    de = ds[:, None] * wpsi.view(sh) * (e > 0)

    def bn_bwd(dy, t, gamma, mu, r):
        that = (t - mu.view(sh)) * r.view(sh)
        k = gamma.double().view(sh) * r.view(sh)
        if rec["frozen"]:
            return k * dy
        return k * (dy - dy.sum((0, 2, 3)).view(sh) / m - that * (dy * that).sum((0, 2, 3)).view(sh) / m)

    cmp("dq1 = BN_q backward", rec["dq1"], bn_bwd(de, q1, P["gq"], mq, rq), tol_map)
    cmp("dx1 = BN_x backward", rec["dx1"], bn_bwd(de, x1, P["gx"], mx, rx), tol_map)
    G = rec["grads"]
    qhat, xhat = (q1 - mq.view(sh)) * rq.view(sh), (x1 - mx.view(sh)) * rx.view(sh)
    cmp("dgamma_q", G["gq"], (de * qhat).sum((0, 2, 3)), tol_grad)
    cmp("dbeta_q", G["bq"], de.sum((0, 2, 3)), tol_grad)
    cmp("dgamma_x", G["gx"], (de * xhat).sum((0, 2, 3)), tol_grad)
    cmp("dbeta_x", G["bx"], de.sum((0, 2, 3)), tol_grad)
    cmp("dw_psi", G["wpsi"].reshape(-1), (ds[:, None] * e).sum((0, 2, 3)), tol_grad)
    cmp("dgamma_p", G["gp"].reshape(-1), (dz * shat).sum().view(1), tol_grad, floor=1e-4 * float(dz.abs().sum()))
    cmp("dbeta_p", G["bp"].reshape(-1), dz.sum().view(1), tol_grad, floor=1e-4 * float(dz.abs().sum()))
    cmp("db_psi", G["bpsi"].reshape(-1), ds.sum().view(1), tol_grad, floor=1e-3 * float(ds.abs().sum()))
    # ---- the three linear maps, backward from the RECORDED gradients at their outputs, in the reference's two-module form
    dq1, dx1 = rec["dq1"].double(), rec["dx1"].double()
    lw = {k: P[k].double().clone().requires_grad_(True) for k in ("W_up", "b_up", "W_q", "b_q", "W_x", "b_x")}
    ql, xl = q.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = F.conv2d(F.conv_transpose2d(ql, lw["W_up"], lw["b_up"], stride=2), lw["W_q"], lw["b_q"])
    y.backward(dq1)
    yx = F.conv2d(xl, lw["W_x"], lw["b_x"])
    yx.backward(dx1)
    floor_b = 1e-3 * float(dq1.abs().sum())  # sums that cancel behind a train-mode BatchNorm: fp32 summation noise
    cmp("dq = up/W_q backward-data", rec["dq"], ql.grad, 2 * tol_map)
    cmp("dxs = g * A + W_x backward-data", rec["dxs"], g * a[:, None] + xl.grad, tol_map)
    cmp("dW_up", G["W_up"], lw["W_up"].grad, tol_grad)
    cmp("db_up", G["b_up"], lw["b_up"].grad, tol_grad, floor=floor_b)
    cmp("dW_q", G["W_q"], lw["W_q"].grad, tol_grad)
    cmp("db_q", G["b_q"], lw["b_q"].grad, tol_grad, floor=floor_b)
    cmp("dW_x", G["W_x"], lw["W_x"].grad, tol_grad)
    cmp("db_x", G["b_x"], lw["b_x"].grad, tol_grad, floor=1e-3 * float(dx1.abs().sum()))
    return bad
