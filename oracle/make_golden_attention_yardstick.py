"""Generate tests/golden/ref_attention_bf16_yardstick.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden_attention_yardstick.py        # needs /root/reference; ~2 minutes

Full-width (initial_feature_map = 64) Model.UNet_attention + calc_loss, run three times like make_golden_yardstick.py: fp64
(truth), fp32, and fp32 parameters under `torch.autocast("cpu", dtype=torch.bfloat16)`. A width-64 attention network has
35 M parameters, so the fixture keeps the ERRORS of the reference's own fp32 / bf16-autocast runs against its fp64 run
(logits rel-L2, loss, rel-L2 of every parameter's gradient) - the yardstick for the tensor-core path - plus the fp64 logits,
loss and the gradients of every parameter with at most 4096 elements (BatchNorm affines, biases, psi, head) as direct
references. Weights are not stored: the package's UNet_attention consumes the RNG exactly like the reference
(tests/golden/ref_attention.pt pins that), so `torch.manual_seed(seed)` reproduces them.
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
from oracle.make_golden_yardstick import inputs, rel  # noqa: E402

RefModel, ref_loss = ref_loader.load()
OUT = os.path.join(HERE, "..", "tests", "golden")
torch.set_num_threads(8)
CASES = {
    "att_w64_c3_k2_dicebce_64": (3, 2, 64, 2, 64, 64, 0, "dice_bce_mc"),
    "att_w64_c1_k3_msemc_64x96": (1, 3, 64, 2, 64, 96, 35, "mseMC"),
    "att_w64_c3_k2_dicebce_128": (3, 2, 64, 2, 128, 128, 7, "dice_bce_mc"),
}
SMALL = 4096


def run(cfg, mode):
    ch, ncls, width, n, h, w, seed, loss_type = cfg
    torch.manual_seed(seed)
    net = RefModel.UNet_attention(ch, ncls, width)
    checksum = {k: float(v.double().sum()) for k, v in net.state_dict().items() if v.is_floating_point()}
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
    grads = {k: p.grad.detach().double() for k, p in net.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}
    net.eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
        oe = net(x.to(dtype)).double()
    return out.detach().double(), float(l), grads, bufs, oe, checksum, x, y


def main():
    res = {}
    for name, cfg in CASES.items():
        o64, l64, g64, b64, e64, cks, x, y = run(cfg, "fp64")
        floor = 1e-6 * max(float(g.norm()) for g in g64.values())
        relg = lambda a, b: float((a - b).norm() / max(float(b.norm()), floor))  # noqa: E731
        entry = {"cfg": cfg, "x": x, "y": y, "logits64": o64.float(), "loss64": l64, "logits_eval64": e64.float(),
                 "sd0_checksum": cks, "grad_floor": floor, "grad_norm64": {k: float(g.norm()) for k, g in g64.items()},
                 "small_grads64": {k: g.float() for k, g in g64.items() if g.numel() <= SMALL},
                 "buffers1": {k: v.float() for k, v in b64.items() if v.numel() <= SMALL}}
        for mode in ("fp32", "bf16_autocast"):
            o, l, g, _, oe, _, _, _ = run(cfg, mode)
            entry[mode] = dict(logits=rel(o, o64), loss=abs(l - l64) / abs(l64), logits_eval=rel(oe, e64),
                               grads={k: relg(g[k], g64[k]) for k in g64})
        res[name] = entry
        gr = sorted(entry["bf16_autocast"]["grads"].values())
        print(f"{name}: reference bf16-autocast vs its fp64: logits {entry['bf16_autocast']['logits']:.3e} (eval "
              f"{entry['bf16_autocast']['logits_eval']:.3e}) loss {entry['bf16_autocast']['loss']:.3e} grads median "
              f"{gr[len(gr) // 2]:.3e} worst {gr[-1]:.3e} best {gr[0]:.3e}; fp32: logits {entry['fp32']['logits']:.3e} grads worst "
              f"{max(entry['fp32']['grads'].values()):.3e}", flush=True)
    path = os.path.join(OUT, "ref_attention_bf16_yardstick.pt")
    torch.save(res, path)
    print(os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
