"""Generate tests/golden/ref_attention.pt from the UNMODIFIED reference's UNet_attention (Model.py:257-391); build container.

    python oracle/make_golden_attention.py

Two narrow nets (width 4: 121 k parameters, everything stored in fp32 and fp64) pin the oracle restatement (unet_oracle.unet_attention_forward)
and the CUDA path end to end: initial state, logits, loss, every parameter gradient, BatchNorm buffers after the step, and the
eval-mode logits. The reference's constructor is what consumes the RNG, so the weights are reproducible from the seed.
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
from oracle.make_golden_yardstick import blob_labels  # noqa: E402

RefModel, ref_loss = ref_loader.load()
OUT = os.path.join(HERE, "..", "tests", "golden", "ref_attention.pt")


def run_case(ch, ncls, width, n, h, w, seed, loss_type, dtype):
    torch.manual_seed(seed)
    net = RefModel.UNet_attention(ch, ncls, width)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, ch, h, w, generator=gen)
    if loss_type in ("dice_bce_mc", "CE"):
        y = blob_labels(gen, n, h, w, ncls)
    else:
        y = torch.rand(n, ncls, h, w, generator=gen) * 3 * (torch.rand(n, ncls, h, w, generator=gen) > 0.6)
    net = net.to(dtype).train()
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
    return dict(x=x, y=y, sd0=sd0, logits=out.detach(), loss=l.detach(), grads=grads, sd1=sd1, logits_eval=out_eval,
                loss_type=loss_type, cfg=(ch, ncls, width, n, h, w, seed))


def main():
    out = {}
    for name, args in {"w4_c3_k2_dicebce": (3, 2, 4, 2, 32, 32, 3, "dice_bce_mc"),
                       "w4_c1_k3_msemc": (1, 3, 4, 1, 32, 48, 17, "mseMC")}.items():
        c32, c64 = run_case(*args, torch.float32), run_case(*args, torch.float64)
        out[name] = dict(cfg=c32["cfg"], loss_type=c32["loss_type"], x=c32["x"], y=c32["y"], sd0=c32["sd0"],
                         logits=c32["logits"], loss=c32["loss"], grads=c32["grads"], logits_eval=c32["logits_eval"],
                         buffers1={k: v for k, v in c32["sd1"].items() if "running" in k or "num_batches" in k},
                         logits64=c64["logits"].float(), loss64=c64["loss"], grads64={k: g.float() for k, g in c64["grads"].items()})
        print(name, float(c32["loss"]), len(c32["sd0"]))
    torch.save(out, OUT)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
