"""Drop-in for the reference's Model.py: `from Model import UNet` now builds the B200-native network.

Put this repository first on sys.path and the reference's train.py / Trainer.py / test.py run unchanged
(train.py:6 imports UNet, UNet_multitask, UNet_attention; UNet and UNet_multitask run on the B200 engine).
"""
from unet_torch_b200 import UNet, UNet_multitask, DoubleConv, Down, Up, OutConv  # noqa: F401


def _outside_hot_path(name):
    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md section 8f)")

    _Unavailable.__name__ = name
    return _Unavailable


UNet_attention = _outside_hot_path("UNet_attention")
