"""Generate tests/golden/ref_loss_branches.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden_losses.py      # needs /root/reference or the staged oracle/_ref

The calc_loss branches next to the fused hot ones ('BCE', 'dice_bce', 'rmse', 'l1loss'), MultitaskUncertaintyLoss
(loss.py:309-325) and MRAccuracy (loss.py:421-440): inputs, outputs and input gradients of the reference's own functions.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402

_, ref_loss = ref_loader.load()
OUT = os.path.join(HERE, "..", "tests", "golden", "ref_loss_branches.pt")


def main():
    gen = torch.Generator().manual_seed(77)
    out = {}
    for lt, pshape, tshape in [("BCE", (3, 1, 24, 40), (3, 24, 40)), ("dice_bce", (3, 1, 24, 40), (3, 24, 40)),
                               ("rmse", (2, 2, 16, 16), (2, 2, 16, 16)), ("l1loss", (2, 2, 16, 16), (2, 2, 16, 16))]:
        p = torch.randn(pshape, generator=gen, requires_grad=True)
        t = (torch.rand(tshape, generator=gen) > 0.6).float() if lt in ("BCE", "dice_bce") else torch.rand(tshape, generator=gen) * 3
        l = ref_loss.calc_loss(p, t, loss_type=lt)
        (g,) = torch.autograd.grad(l, p)
        out[lt] = dict(pred=p.detach(), target=t, loss=l.detach(), grad=g)
    for name, flags in (("uncertainty_reg", [True, True]), ("uncertainty_mixed", [False, True])):
        l1 = torch.tensor(0.731, requires_grad=True)
        l2 = torch.tensor(2.25, requires_grad=True)
        lv = [torch.tensor([0.3], requires_grad=True), torch.tensor([-0.2], requires_grad=True)]
        tot = ref_loss.MultitaskUncertaintyLoss()([l1, l2], lv, flags)
        grads = torch.autograd.grad(tot.sum(), [l1, l2] + lv)
        out[name] = dict(losses=[l1.detach(), l2.detach()], log_vars=[v.detach() for v in lv], flags=flags,
                         total=tot.detach(), grads=[g.detach() for g in grads])
    # MRAccuracy: blobs as predictions, dots as ground truth
    pred = torch.full((3, 1, 48, 48), -4.0)
    pred[0, 0, 4:9, 4:9] = 3.0
    pred[0, 0, 20:22, 30:40] = 1.0
    pred[0, 0, 9, 9] = 0.5          # touches the first blob diagonally: 8-connectivity keeps it one component
    pred[1, 0, 10:12, 10:12] = 2.0
    tgt = torch.zeros(3, 48, 48)
    tgt[0, 5, 5] = tgt[0, 21, 33] = tgt[0, 40, 40] = 1.0
    tgt[1, 11, 11] = 1.0
    out["MRAccuracy"] = dict(pred=pred, target=tgt, value=float(ref_loss.MRAccuracy(pred, tgt)))
    pred2 = pred.clone()
    pred2[2, 0, 0:3, 0:3] = 5.0     # a component where the ground truth has none
    out["MRAccuracy_fp"] = dict(pred=pred2, target=tgt, value=float(ref_loss.MRAccuracy(pred2, tgt)))
    torch.save(out, OUT)
    print(OUT, os.path.getsize(OUT), {k: (float(v["loss"]) if "loss" in v else v.get("value", None)) for k, v in out.items()})


if __name__ == "__main__":
    main()
