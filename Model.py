"""Drop-in for the reference's Model.py: `from Model import UNet` now builds the B200-native network.

Put this repository first on sys.path and the reference's train.py / Trainer.py / test.py run unchanged
(train.py:6 imports UNet, UNet_multitask, UNet_attention: UNet and UNet_multitask run on the tensor-core engine,
UNet_attention on the library's generic fp32 CUDA engine).
"""
from unet_torch_b200 import (UNet, UNet_multitask, UNet_attention, Attention_block, DoubleConv, Down, Up,  # noqa: F401
                             OutConv)
