"""Loss plumbing of the reference's ``semi_seg/_utils.py`` that sits on the hot path.

``IIDLoss`` (semi_seg/_utils.py:12-15) returns the loss only; ``IICLossWrapper`` (:189-224) picks the
global loss for encoder features and the small-patch local loss for decoder features.  The UNet block
names come from contrastyou/arch/unet.py:185-194 and are restated here so this module does not need
the backbone.
"""
from __future__ import annotations

from itertools import repeat
from typing import List, Union

from torch import Tensor, nn

from ..losses.iic_loss import IIDLoss as _IIDLoss
from ..losses.iic_loss import IIDSegmentationSmallPathLoss

# contrastyou/arch/unet.py:185-194
ENCODER_NAMES = ["Conv1", "Conv2", "Conv3", "Conv4", "Conv5"]
DECODER_NAMES = ["Up5", "Up_conv5", "Up4", "Up_conv4", "Up3", "Up_conv3", "Up2", "Up_conv2", "DeConv_1x1"]


class IIDLoss(_IIDLoss):
    def forward(self, x_out: Tensor, x_tf_out: Tensor):
        return super().forward(x_out, x_tf_out)[0]


def _nlist(n):
    def parse(x):
        if isinstance(x, (list, tuple)):
            assert len(x) == n, (len(x), n)
            return x
        return list(repeat(x, n))
    return parse


class IICLossWrapper(nn.Module):
    def __init__(self, feature_names: Union[str, List[str]], paddings: Union[int, List[int]],
                 patch_sizes: Union[int, List[int]]) -> None:
        super().__init__()
        if isinstance(feature_names, str):
            feature_names = [feature_names]
        self._encoder_features = [f for f in feature_names if f in ENCODER_NAMES]
        self._decoder_features = [f for f in feature_names if f in DECODER_NAMES]
        assert len(feature_names) == len(self._encoder_features) + len(self._decoder_features)
        self._LossModuleDict = nn.ModuleDict()
        for f in self._encoder_features:
            self._LossModuleDict[f] = IIDLoss()
        if len(self._decoder_features) > 0:
            paddings = _nlist(len(self._decoder_features))(paddings)
            patch_sizes = _nlist(len(self._decoder_features))(patch_sizes)
            for f, p, size in zip(self._decoder_features, paddings, patch_sizes):
                self._LossModuleDict[f] = IIDSegmentationSmallPathLoss(padding=p, patch_size=size)

    def __getitem__(self, item):
        if item in self._LossModuleDict.keys():
            return self._LossModuleDict[item]
        raise IndexError(item)

    def __iter__(self):
        for k, v in self._LossModuleDict.items():
            yield v

    def items(self):
        return self._LossModuleDict.items()

    @property
    def feature_names(self):
        return self._encoder_features + self._decoder_features
