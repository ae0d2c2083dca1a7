"""Generate tests/golden/ref_bf16_yardstick.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden_yardstick.py        # needs /root/reference; ~1 minute

For every full-width case of ref_full_nets.pt (same seeds, inputs and labels: oracle/make_golden.py `run_case`) this runs
the reference's own Model.UNet + calc_loss three times - fp64 (truth), fp32, and fp32 parameters under
`torch.autocast("cpu", dtype=torch.bfloat16)` (what a user of the reference gets from PyTorch when asking for bf16) - and
stores the ERRORS of the reference's fp32 and bf16-autocast runs against its own fp64 run: logits rel-L2, loss relative
error, rel-L2 of every parameter's gradient. These are the yardstick the end-to-end -m gpu tests hold the B200 path to
("ours <= 1.5 x the reference's own bf16 error, per parameter and in the median"): a bf16 BatchNorm network differs from
an fp32 one by tens of percent in individual gradients, for the reference as much as for us.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import warnings

warnings.filterwarnings("ignore")
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402

RefModel, ref_loss = ref_loader.load()
OUT = os.path.join(HERE, "..", "tests", "golden")
torch.set_num_threads(8)
CASES = {  # identical to make_golden.py section B
    "w64_c3_k2_dicebce": (3, 2, 64, 2, 32, 32, 0, "dice_bce_mc"),
    "w64_c1_k2_dicebce": (1, 2, 64, 2, 32, 32, 35, "dice_bce_mc"),
    "w64_c3_k5_ce": (3, 5, 64, 1, 32, 32, 1063, "CE"),
    "w64_c3_k2_msemc": (3, 2, 64, 2, 32, 32, 7, "mseMC"),
    # larger maps (more pixels per BatchNorm channel, as in training): 4 x 128^2 like SURVEY 7.4-1
    "w64_c3_k2_dicebce_128": (3, 2, 64, 4, 128, 128, 0, "dice_bce_mc"),
}


def blob_labels(gen, n, h, w, ncls):  # make_golden.py
    f = torch.randn(n, 1, h // 4, w // 4, generator=gen)
    f = torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
    q = torch.quantile(f.flatten(), torch.linspace(0, 1, ncls + 1)[1:-1])
    return torch.bucketize(f, q).float()


def inputs(ch, ncls, n, h, w, seed, loss_type):
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, ch, h, w, generator=gen)
    if loss_type in ("dice_bce_mc", "CE"):
        y = blob_labels(gen, n, h, w, ncls)
    else:
        y = torch.rand(n, ncls, h, w, generator=gen) * 200 * (torch.rand(n, ncls, h, w, generator=gen) > 0.7)
    return x, y


def run(cfg, mode):
    ch, ncls, width, n, h, w, seed, loss_type = cfg
    torch.manual_seed(seed)
    net = RefModel.UNet(ch, ncls, width)
    x, y = inputs(ch, ncls, n, h, w, seed, loss_type)
    dtype = torch.float64 if mode == "fp64" else torch.float32
    net = net.to(dtype).train()
    ref_loss.CLASS_NUMBER = ncls
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
        out = net(x.to(dtype))
    out = out.to(dtype)
    pred = torch.relu(out) if loss_type.startswith("mse") else out
    l = ref_loss.calc_loss(pred, y.to(dtype), loss_type=loss_type)
    net.zero_grad()
    l.backward()
    return out.detach().double(), float(l), {k: p.grad.detach().double() for k, p in net.named_parameters()}


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-300))


def main():
    res = {}
    for name, cfg in CASES.items():
        o64, l64, g64 = run(cfg, "fp64")
        entry = {"cfg": cfg}
        for mode in ("fp32", "bf16_autocast"):
            o, l, g = run(cfg, mode)
            entry[mode] = dict(logits=rel(o, o64), loss=abs(l - l64) / abs(l64), grads={k: rel(g[k], g64[k]) for k in g64})
        res[name] = entry
        gr = sorted(entry["bf16_autocast"]["grads"].values())
        print(f"{name}: reference bf16-autocast vs its fp64: logits {entry['bf16_autocast']['logits']:.3e} loss "
              f"{entry['bf16_autocast']['loss']:.3e} grads median {gr[len(gr) // 2]:.3e} worst {gr[-1]:.3e} best {gr[0]:.3e}; "
              f"fp32: logits {entry['fp32']['logits']:.3e} grads worst {max(entry['fp32']['grads'].values()):.3e}")
    torch.save(res, os.path.join(OUT, "ref_bf16_yardstick.pt"))
    print(os.path.getsize(os.path.join(OUT, "ref_bf16_yardstick.pt")), "bytes")


if __name__ == "__main__":
    main()
