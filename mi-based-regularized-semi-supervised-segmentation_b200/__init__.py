"""B200-native IIC mutual-information losses + UDA consistency (drop-in for the reference's loss API).

Import as ``iic_b200`` (the importable alias at the repo root); this directory carries the name the
project layout asks for, which is not a valid Python identifier.
"""
from . import _lib, augment, checks, ops  # noqa: F401
from .augment import TensorRandomFlip, draw_flip_flags, flip_stack  # noqa: F401
from .checks import check_mode, get_check_mode, raise_if_flagged, set_check_mode  # noqa: F401
from .losses.iic_loss import (IIDLoss, IIDSegmentationLoss, IIDSegmentationSmallPathLoss, compute_joint,  # noqa: F401
                              iic_losses, patch_generator)
from .losses.kl_losses import KL_div, MSELoss, dice_from_counts, sup_kl_from_logits, uda_from_logits  # noqa: F401
from .ops import set_data_parallel, data_parallel_transport, ddp_loss_scale  # noqa: F401
from .semi_seg._utils import IICLossWrapper, combine_iic_losses, iic_regularization  # noqa: F401
from .semi_seg.meters import DeferredScalarMeters  # noqa: F401

__all__ = ["IIDLoss", "IIDSegmentationLoss", "IIDSegmentationSmallPathLoss", "compute_joint", "patch_generator",
           "iic_losses", "iic_regularization", "combine_iic_losses", "DeferredScalarMeters",
           "KL_div", "MSELoss", "uda_from_logits", "sup_kl_from_logits", "dice_from_counts", "TensorRandomFlip", "draw_flip_flags", "flip_stack", "IICLossWrapper", "set_check_mode", "get_check_mode",
           "check_mode", "raise_if_flagged", "set_data_parallel", "data_parallel_transport", "ddp_loss_scale"]
