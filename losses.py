"""Drop-in for the reference's top-level losses.py (hot part: ls_gan :10-11, NDS_Loss :13-15,
CharbonnierLoss :99-111, EdgeLoss :113-138, get_loss :186-197).  The VGG / ResNet perceptual losses
(:17-97, :140-183) are not used by MTD_GAN_Method and are out of scope (SURVEY §2 #2b)."""
from mtdgan_b200.losses import ls_gan, NDS_Loss, CharbonnierLoss, EdgeLoss, get_loss, nds_mask  # noqa: F401

__all__ = ["ls_gan", "NDS_Loss", "CharbonnierLoss", "EdgeLoss", "get_loss", "nds_mask"]
