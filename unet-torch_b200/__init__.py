"""B200-native (sm_100a) implementation of the U-Net hot path of caki35/UNet-Torch.

Public surface mirrors the reference: `UNet(n_channels, n_classes, ...)`/`forward(x)` (Model.py:95-169) and
`calc_loss(pred, target, loss_type=...)` (loss.py:442-516). Everything on that path executes in hand-written
CUDA kernels behind the C ABI of include/b200unet.h; importing the kernels fails loudly if the library is absent.
"""
from . import _lib  # noqa: F401
from .model import UNet, UNet_multitask, UNet_attention, Attention_block, DoubleConv, Down, Up, OutConv, predict_mask, preprocess, preprocess_crop, predict_tiled  # noqa: F401
from .loss import calc_loss, DiceLoss, ce_dice_loss, relu_mse_loss, MultitaskUncertaintyLoss, MRAccuracy  # noqa: F401
from .dist import DataParallelContext, init_from_env  # noqa: F401
from .optim import FusedAdam, FusedSGD  # noqa: F401

__all__ = ["UNet", "UNet_multitask", "UNet_attention", "Attention_block", "DoubleConv", "Down", "Up", "OutConv", "calc_loss", "DiceLoss", "ce_dice_loss", "relu_mse_loss", "MultitaskUncertaintyLoss", "MRAccuracy",
           "predict_mask", "preprocess", "preprocess_crop", "predict_tiled", "DataParallelContext", "init_from_env", "FusedSGD", "FusedAdam"]
