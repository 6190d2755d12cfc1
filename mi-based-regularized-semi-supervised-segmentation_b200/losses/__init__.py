from .iic_loss import (IIDLoss, IIDSegmentationLoss, IIDSegmentationSmallPathLoss, compute_joint,  # noqa: F401
                       patch_generator)
from .kl_losses import KL_div, MSELoss, dice_from_counts, sup_kl_from_logits, uda_from_logits  # noqa: F401
