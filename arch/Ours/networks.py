"""Drop-in for the reference's arch/Ours/networks.py (MTD-GAN part, :15-474 and :1940-2009): with this
repository ahead of the reference on sys.path, `from arch.Ours.networks import *` (models.py:15) resolves to
the B200-native modules.  The ablation variants (:478-1936) are out of scope (SURVEY §2 #1b)."""
from mtdgan_b200.networks import (FFT_ConvBlock, ResFFT_Generator, UpsampleBlock, Multi_Task_Discriminator_Skip,
                                  MTD_GAN_Method)
from losses import NDS_Loss, EdgeLoss, CharbonnierLoss, ls_gan  # noqa: F401  (names the reference exports, :6)

__all__ = ["FFT_ConvBlock", "ResFFT_Generator", "UpsampleBlock", "Multi_Task_Discriminator_Skip", "MTD_GAN_Method",
           "NDS_Loss", "EdgeLoss", "CharbonnierLoss", "ls_gan"]
