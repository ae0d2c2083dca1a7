"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only); writes tests/golden/

The reference has no tests or golden vectors of its own, so these files - outputs of the reference's own
Model.UNet and loss.calc_loss on seeded inputs - are what pins the oracle and the CUDA path. /root/reference does
not exist on the GPU box; the committed vectors travel instead.
"""
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("B200UNET_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import warnings

warnings.filterwarnings("ignore")
import torch  # noqa: E402

import Model as RefModel  # noqa: E402
import loss as ref_loss  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
torch.set_num_threads(8)


def blob_labels(gen, n, h, w, ncls):
    """Blob-structured labels (smooth random field thresholded into classes)."""
    f = torch.randn(n, 1, h // 4, w // 4, generator=gen)
    f = torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
    q = torch.quantile(f.flatten(), torch.linspace(0, 1, ncls + 1)[1:-1])
    return torch.bucketize(f, q).float()


def run_case(ch, ncls, width, n, h, w, seed, loss_type, dtype):
    torch.manual_seed(seed)
    net = RefModel.UNet(ch, ncls, width)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, ch, h, w, generator=gen)
    if loss_type in ("dice_bce_mc", "CE"):
        y = blob_labels(gen, n, h, w, ncls)
    else:
        y = torch.rand(n, ncls, h, w, generator=gen) * 200 * (torch.rand(n, ncls, h, w, generator=gen) > 0.7)
    net = net.to(dtype)
    net.train()
    ref_loss.CLASS_NUMBER = ncls
    out = net(x.to(dtype))
    pred = torch.relu(out) if loss_type.startswith("mse") else out
    l = ref_loss.calc_loss(pred, y.to(dtype), loss_type=loss_type)
    net.zero_grad()
    l.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    sd1 = {k: v.clone() for k, v in net.state_dict().items()}
    net.eval()
    with torch.no_grad():
        out_eval = net(x.to(dtype))
    return dict(x=x, y=y, sd0=sd0, logits=out.detach(), loss=l.detach(), grads=grads, sd1=sd1,
                logits_eval=out_eval, loss_type=loss_type, cfg=(ch, ncls, width, n, h, w, seed))


def checksum(sd):
    out = {}
    for k, v in sd.items():
        f = v.double().flatten()
        out[k] = torch.tensor([f.sum(), f.abs().sum(), f[0], f[f.numel() // 2], f[-1]], dtype=torch.float64)
    return out


def sample(t, step=997):
    return t.flatten()[::step].clone()


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- A. narrow nets, everything stored (pins the oracle end to end)
    small = {}
    for name, args in {
        "w4_c1_k2_dicebce": (1, 2, 4, 2, 32, 32, 0, "dice_bce_mc"),
        "w4_c3_k5_dicebce": (3, 5, 4, 2, 32, 48, 35, "dice_bce_mc"),
        "w4_c3_k3_ce": (3, 3, 4, 1, 32, 32, 1063, "CE"),
        "w4_c3_k2_msemc": (3, 2, 4, 2, 32, 32, 7, "mseMC"),
    }.items():
        c32 = run_case(*args, torch.float32)
        c64 = run_case(*args, torch.float64)
        small[name] = dict(
            cfg=c32["cfg"], loss_type=c32["loss_type"], x=c32["x"], y=c32["y"], sd0=c32["sd0"],
            logits=c32["logits"], loss=c32["loss"], grads=c32["grads"], logits_eval=c32["logits_eval"],
            buffers1={k: v for k, v in c32["sd1"].items() if "running" in k or "num_batches" in k},
            logits64=c64["logits"].float(), loss64=c64["loss"],
            grads64={k: g.float() for k, g in c64["grads"].items()})
    torch.save(small, os.path.join(OUT, "ref_small_nets.pt"))

    # ---- B. full-width nets (what the CUDA path runs): weights are reproducible from the seed, store checksums
    full = {}
    for name, args in {
        "w64_c3_k2_dicebce": (3, 2, 64, 2, 32, 32, 0, "dice_bce_mc"),
        "w64_c1_k2_dicebce": (1, 2, 64, 2, 32, 32, 35, "dice_bce_mc"),
        "w64_c3_k5_ce": (3, 5, 64, 1, 32, 32, 1063, "CE"),
        "w64_c3_k2_msemc": (3, 2, 64, 2, 32, 32, 7, "mseMC"),
    }.items():
        c32 = run_case(*args, torch.float32)
        c64 = run_case(*args, torch.float64)
        small_keys = [k for k, g in c32["grads"].items() if g.numel() <= 4096]
        full[name] = dict(
            cfg=c32["cfg"], loss_type=c32["loss_type"], x=c32["x"], y=c32["y"],
            sd0_checksum=checksum(c32["sd0"]),
            logits=c32["logits"], loss=c32["loss"], logits_eval=c32["logits_eval"],
            logits64=c64["logits"].float(), loss64=c64["loss"],
            grad_norm={k: g.double().norm() for k, g in c32["grads"].items()},
            grad_norm64={k: g.double().norm() for k, g in c64["grads"].items()},
            grad_small={k: c32["grads"][k] for k in small_keys},
            grad_small64={k: c64["grads"][k].float() for k in small_keys},
            grad_sample={k: sample(g) for k, g in c32["grads"].items() if g.numel() > 4096},
            grad_sample64={k: sample(g).float() for k, g in c64["grads"].items() if g.numel() > 4096},
            buffers1_checksum=checksum({k: v for k, v in c32["sd1"].items() if "running" in k or "num_batches" in k}))
    torch.save(full, os.path.join(OUT, "ref_full_nets.pt"))

    # ---- C. operator-level vectors from the torch ops the reference calls
    gen = torch.Generator().manual_seed(123)
    ops = {}
    xp = torch.randn(2, 3, 6, 8, generator=gen)
    xp[0, 0, 0, 0:2] = 1.5                      # tie inside a window -> first wins
    xp[0, 0, 1, 0:2] = 1.5
    xp[0, 1, 2, 3] = float("nan")               # NaN propagates
    xp[1, 2, 4, 4], xp[1, 2, 4, 5], xp[1, 2, 5, 4], xp[1, 2, 5, 5] = -0.0, 0.0, -0.0, 0.0
    pv, pi = torch.nn.functional.max_pool2d(xp, 2, return_indices=True)
    ops["pool"] = dict(x=xp, out=pv, idx=pi)
    xo = torch.randn(1, 2, 5, 7, generator=gen)  # odd sizes: floor mode
    pv, pi = torch.nn.functional.max_pool2d(xo, 2, return_indices=True)
    ops["pool_odd"] = dict(x=xo, out=pv, idx=pi)
    z = torch.randn(2, 5, 16, 16, generator=gen) * 0.01
    ops["argmax_small_logits"] = dict(z=z, mask=torch.argmax(torch.softmax(z, 1), 1))
    z2 = torch.randn(2, 3, 8, 8, generator=gen)
    z2[:, 1] = z2[:, 0]                          # exact ties
    ops["argmax_ties"] = dict(z=z2, mask=torch.argmax(torch.softmax(z2, 1), 1))
    for ncls in (2, 5):
        zz = torch.randn(3, ncls, 24, 40, generator=gen, requires_grad=True)
        tt = torch.randint(0, ncls, (3, 24, 40), generator=gen).float()
        ref_loss.CLASS_NUMBER = ncls
        for lt in ("dice_bce_mc", "CE"):
            l = ref_loss.calc_loss(zz, tt, loss_type=lt)
            (g,) = torch.autograd.grad(l, zz)
            ops[f"loss_{lt}_{ncls}"] = dict(z=zz.detach(), t=tt, loss=l.detach(), grad=g)
    o = torch.randn(2, 2, 16, 16, generator=gen, requires_grad=True)
    t = torch.rand(2, 2, 16, 16, generator=gen) * 3
    l = ref_loss.calc_loss(torch.relu(o), t, loss_type="mseMC")
    (g,) = torch.autograd.grad(l, o)
    ops["loss_relu_mseMC"] = dict(o=o.detach(), t=t, loss=l.detach(), grad=g)
    o1 = torch.randn(2, 1, 16, 16, generator=gen, requires_grad=True)
    t1 = torch.rand(2, 16, 16, generator=gen)
    l = ref_loss.calc_loss(o1, t1, loss_type="mse")
    (g,) = torch.autograd.grad(l, o1)
    ops["loss_mse"] = dict(o=o1.detach(), t=t1, loss=l.detach(), grad=g)
    torch.save(ops, os.path.join(OUT, "ref_ops.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
