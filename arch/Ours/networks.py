"""Drop-in for the reference's arch/Ours/networks.py (:15-2009): with this repository ahead of the reference on
sys.path, `from arch.Ours.networks import *` (models.py:15) resolves to the B200-native modules -- MTD-GAN itself and
the ablation rows (REDCNN_Generator, the five partial discriminators, the ten Ablation_* wrappers; models.py:56-75)."""
from mtdgan_b200.networks import (FFT_ConvBlock, ResFFT_Generator, UpsampleBlock, Multi_Task_Discriminator_Skip,
                                  MTD_GAN_Method, REDCNN_Generator, CLS_Discriminator, SEG_Discriminator,
                                  CLS_SEG_Discriminator, CLS_REC_Discriminator, SEG_REC_Discriminator, Ablation_CLS,
                                  Ablation_SEG, Ablation_CLS_SEG, Ablation_CLS_REC, Ablation_SEG_REC, Ablation_CLS_SEG_REC,
                                  Ablation_CLS_SEG_REC_NDS, Ablation_CLS_SEG_REC_RC, Ablation_CLS_SEG_REC_NDS_RC,
                                  Ablation_CLS_SEG_REC_NDS_RC_ResFFT)
from losses import NDS_Loss, EdgeLoss, CharbonnierLoss, ls_gan  # noqa: F401  (names the reference exports, :6)

__all__ = ["FFT_ConvBlock", "ResFFT_Generator", "UpsampleBlock", "Multi_Task_Discriminator_Skip", "MTD_GAN_Method",
           "REDCNN_Generator", "CLS_Discriminator", "SEG_Discriminator", "CLS_SEG_Discriminator", "CLS_REC_Discriminator",
           "SEG_REC_Discriminator", "Ablation_CLS", "Ablation_SEG", "Ablation_CLS_SEG", "Ablation_CLS_REC", "Ablation_SEG_REC",
           "Ablation_CLS_SEG_REC", "Ablation_CLS_SEG_REC_NDS", "Ablation_CLS_SEG_REC_RC", "Ablation_CLS_SEG_REC_NDS_RC",
           "Ablation_CLS_SEG_REC_NDS_RC_ResFFT", "NDS_Loss", "EdgeLoss", "CharbonnierLoss", "ls_gan"]
