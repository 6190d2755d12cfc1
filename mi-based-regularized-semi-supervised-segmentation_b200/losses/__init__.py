from .iic_loss import (IIDLoss, IIDSegmentationLoss, IIDSegmentationSmallPathLoss, compute_joint,  # noqa: F401
                       patch_generator)
from .kl_losses import KL_div, MSELoss  # noqa: F401
