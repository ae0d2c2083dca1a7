"""-m gpu: the reference's UNMODIFIED Trainer.py (staged by oracle/build_ref.py) driving this repository's drop-in `Model` /
`loss` modules - north_star: "Trainer.py ... run unchanged". Trainer.py does `from loss import calc_loss,
MultitaskUncertaintyLoss, MRAccuracy` (Trainer.py:6) and calls `self.model(inputs)`, `calc_loss(...)`, `loss.backward()`,
`self.optimizer.step()`, poly-LR writes to `param_group['lr']`, `copy.deepcopy(model.state_dict())`, `torch.save(...)`
(Trainer.py:697-760). matplotlib is absent from this image and only used for the plots after training: it is stubbed
(oracle/ref_loader.py). The checkpoints it writes must load, strictly, into the REFERENCE's own Model.UNet."""
import os

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from gpu_util import rel_l2
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref_loader.ref_dir() is None or not os.path.exists(
    os.path.join(ref_loader.ref_dir(), "Trainer.py")), reason="reference Trainer.py not staged (python oracle/build_ref.py)")]


def _loaders(n_train, n_val, batch, make_labels, h=64, w=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for phase, n in (("train", n_train), ("val", n_val)):
        x = torch.randn(n, 3, h, w, generator=g)
        out[phase] = DataLoader(TensorDataset(x, *make_labels(x, g)), batch_size=batch if phase == "train" else 1, shuffle=False)
    return out


def _trainer(tmp_path, model, model_type, loaders, optimizer, loss_function, accuracy_metric, epochs=2, lr_scheduler=True):
    import Model  # noqa: F401  (the root drop-in shims: what Trainer.py's own imports resolve to)
    import loss

    T = ref_loader.load_trainer(Model, loss)
    return T.Trainer(model, model_type, torch.cuda.FloatTensor, torch.device("cuda:0"), str(tmp_path), loaders, 4, optimizer,
                     patience=5, num_epochs=epochs, loss_function=loss_function, accuracy_metric=accuracy_metric,
                     lr_scheduler=lr_scheduler)


@pytest.mark.parametrize("fused_opt", [False, True])
def test_reference_trainer_single_train_runs_unchanged(tmp_path, fused_opt):
    """Trainer.singe_train (Trainer.py:663-829), segmentation: 'dice_bce_mc' loss and metric, SGD as train.py:344 builds it
    (or this repository's FusedSGD in its place), poly learning rate."""
    import Model
    import loss
    import unet_torch_b200 as U

    loss.CLASS_NUMBER = 2  # train.py:163
    torch.manual_seed(0)
    model = Model.UNet(3, 2).to("cuda:0")
    w0 = model.inc.double_conv[0].weight.detach().clone()
    kw = dict(lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt = U.FusedSGD(model, **kw) if fused_opt else torch.optim.SGD(model.parameters(), **kw)

    def labels(x, g):
        return ((x[:, 0] + 0.3 * torch.randn(x.shape[0], x.shape[2], x.shape[3], generator=g)) > 0).float(),

    tr = _trainer(tmp_path, model, "single", _loaders(8, 2, 4, labels), opt, "dice_bce_mc", "dice_bce_mc")
    tr.train()   # (returns None: Trainer.train drops singe_train's return value, Trainer.py:113-117)
    assert len(tr.train_loss_list) == 2 and len(tr.val_loss_list) == 2
    assert tr.train_loss_list[1] < tr.train_loss_list[0]             # it learns the synthetic task
    assert not torch.equal(model.inc.double_conv[0].weight.detach(), w0)
    lr_now = opt.param_groups[0]["lr"]
    assert 0 < lr_now < 0.01                                         # poly-LR reached the optimizer (Trainer.py:721-724)
    best = os.path.join(str(tmp_path), "models", "best.pt")
    assert os.path.exists(best) and os.path.exists(os.path.join(str(tmp_path), "models", "last_epoch.pt"))
    assert os.path.exists(os.path.join(str(tmp_path), "logs.txt"))
    # the checkpoint is a reference checkpoint: strict load into the reference's own class, same eval logits
    RefModel, _ = ref_loader.load()
    ref_net = RefModel.UNet(3, 2)
    sd = torch.load(best, map_location="cpu")
    ref_net.load_state_dict(sd, strict=True)
    ref_net.eval()
    model.load_state_dict(sd)
    model.eval()
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = ref_net(x)
        got = model(x.cuda())
    e = rel_l2(got, want)
    print(f"Trainer-produced best.pt: reference Model.UNet vs B200 UNet eval logits rel-L2 {e:.3e}")
    assert e < 3e-2
    assert int(sd["inc.double_conv.1.num_batches_tracked"]) >= 2


def test_reference_trainer_regression_runs_unchanged(tmp_path):
    """model_type 'regression': F.relu(model(x)) + 'mseMC' (Trainer.py:709-712), Adam as configseros.yml:15 / train.py:341-343
    through this repository's FusedAdam."""
    import Model
    import unet_torch_b200 as U

    torch.manual_seed(1)
    model = Model.UNet(3, 2).to("cuda:0")
    opt = U.FusedAdam(model, lr=1e-3, weight_decay=1e-4)

    def labels(x, g):
        return (torch.relu(x[:, :2]) * 2.0),

    tr = _trainer(tmp_path, model, "regression", _loaders(8, 2, 4, labels, seed=3), opt, "mseMC", "mseMC", lr_scheduler=None)
    tr.train()
    assert tr.train_loss_list[1] < tr.train_loss_list[0]
    assert os.path.exists(os.path.join(str(tmp_path), "models", "best.pt"))


def test_reference_trainer_multitask_uncertainty_runs_unchanged(tmp_path):
    """Trainer.multi_task_uc_train (Trainer.py:994-1172): UNet_multitask, two relu + 'mse' task losses combined by
    MultitaskUncertaintyLoss with CPU log-variance leaves, stepped by the Trainer's own Adam."""
    import Model

    torch.manual_seed(2)
    model = Model.UNet_multitask(3, 1).to("cuda:0")
    g = torch.Generator().manual_seed(5)

    class TwoLabels(torch.utils.data.Dataset):
        def __init__(self, n):
            self.x = torch.randn(n, 3, 32, 32, generator=g)

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, i):
            return self.x[i], (torch.relu(self.x[i, 0]), torch.relu(-self.x[i, 1]))

    loaders = {"train": DataLoader(TwoLabels(8), batch_size=4), "val": DataLoader(TwoLabels(2), batch_size=1)}
    opt = torch.optim.SGD(model.parameters(), lr=0.01)   # replaced inside multi_task_uc_train by its own Adam (Trainer.py:1009)
    tr = _trainer(tmp_path, model, "multi_task", loaders, opt, "multi_task_loss", "mse", lr_scheduler=None)
    tr.train()
    assert len(tr.train_loss_list_1) == 2 and len(tr.val_loss_list_2) == 2
    assert all(torch.isfinite(torch.tensor(v)) for v in tr.train_loss_list + tr.val_loss_list)
    sd = torch.load(os.path.join(str(tmp_path), "models", "best.pt"), map_location="cpu")
    RefModel, _ = ref_loader.load()
    RefModel.UNet_multitask(3, 1).load_state_dict(sd, strict=True)
