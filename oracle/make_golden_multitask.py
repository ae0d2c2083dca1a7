"""Generate tests/golden/ref_multitask.pt from the UNMODIFIED reference's UNet_multitask (build container only).

    python oracle/make_golden_multitask.py

The training step is the one of Trainer.py:877-900: (out1, out2) = model(x); relu on both; loss = calc_loss(out1, t1,
'mse'-family) + calc_loss(out2, t2, ...); backward. A second case combines the two losses with
MultitaskUncertaintyLoss (Trainer.py:1052-1066). Narrow nets are stored in full (fp32 and fp64), the full-width net
(what the CUDA path runs) with weight checksums, logits, loss and gradient norms/samples.
"""
import os
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import make_golden as MG  # noqa: E402  (puts /root/reference on sys.path and imports Model / loss from it)

RefModel, ref_loss = MG.RefModel, MG.ref_loss
OUT = os.path.join(MG.OUT, "ref_multitask.pt")


def run_case(ch, ncls, width, n, h, w, seed, combine, dtype):
    torch.manual_seed(seed)
    net = RefModel.UNet_multitask(ch, ncls, width)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, ch, h, w, generator=gen)
    t1 = torch.rand(n, ncls, h, w, generator=gen) * 3 * (torch.rand(n, ncls, h, w, generator=gen) > 0.6)
    t2 = torch.rand(n, ncls, h, w, generator=gen) * 3 * (torch.rand(n, ncls, h, w, generator=gen) > 0.6)
    net = net.to(dtype).train()
    o1, o2 = net(x.to(dtype))
    l1 = ref_loss.calc_loss(torch.relu(o1), t1.to(dtype), loss_type="mseMC")
    l2 = ref_loss.calc_loss(torch.relu(o2), t2.to(dtype), loss_type="mseMC")
    if combine == "sum":
        loss = l1 + l2
    else:
        lv = [torch.tensor([0.3], dtype=dtype), torch.tensor([-0.2], dtype=dtype)]
        loss = ref_loss.MultitaskUncertaintyLoss()([l1, l2], lv, [True, True])
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    sd1 = {k: v.clone() for k, v in net.state_dict().items()}
    net.eval()
    with torch.no_grad():
        e1, e2 = net(x.to(dtype))
    return dict(x=x, t1=t1, t2=t2, sd0=sd0, o1=o1.detach(), o2=o2.detach(), loss=loss.detach().reshape(()), grads=grads,
                sd1=sd1, e1=e1, e2=e2, cfg=(ch, ncls, width, n, h, w, seed), combine=combine)


def main():
    out = {}
    for name, args in {"w4_sum": (3, 2, 4, 2, 32, 32, 11, "sum"), "w4_uncertainty": (1, 2, 4, 2, 32, 48, 12, "uncertainty")}.items():
        c32, c64 = run_case(*args, torch.float32), run_case(*args, torch.float64)
        out[name] = dict(kind="small", cfg=c32["cfg"], combine=c32["combine"], x=c32["x"], t1=c32["t1"], t2=c32["t2"],
                         sd0=c32["sd0"], o1=c32["o1"], o2=c32["o2"], loss=c32["loss"], grads=c32["grads"],
                         e1=c32["e1"], e2=c32["e2"],
                         buffers1={k: v for k, v in c32["sd1"].items() if "running" in k or "num_batches" in k},
                         o1_64=c64["o1"].float(), o2_64=c64["o2"].float(), loss64=c64["loss"],
                         grads64={k: g.float() for k, g in c64["grads"].items()})
    args = (3, 2, 64, 2, 32, 32, 21, "sum")
    c32, c64 = run_case(*args, torch.float32), run_case(*args, torch.float64)
    small_keys = [k for k, g in c32["grads"].items() if g.numel() <= 4096]
    out["w64_sum"] = dict(
        kind="full", cfg=c32["cfg"], combine="sum", x=c32["x"], t1=c32["t1"], t2=c32["t2"],
        sd0_checksum=MG.checksum(c32["sd0"]), o1=c32["o1"], o2=c32["o2"], loss=c32["loss"], e1=c32["e1"], e2=c32["e2"],
        o1_64=c64["o1"].float(), o2_64=c64["o2"].float(), loss64=c64["loss"],
        grad_norm64={k: g.double().norm() for k, g in c64["grads"].items()},
        grad_small64={k: c64["grads"][k].float() for k in small_keys},
        grad_sample64={k: MG.sample(g).float() for k, g in c64["grads"].items() if g.numel() > 4096},
        buffers1_checksum=MG.checksum({k: v for k, v in c32["sd1"].items() if "running" in k or "num_batches" in k}))
    torch.save(out, OUT)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
