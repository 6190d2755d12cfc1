"""Per-sample flip alignment of the two views, batched.

The reference aligns ``model(T(x))`` with ``T(model(x))`` by replaying the same random flips on images, logits and
every decoder feature map (semi_seg/epocher.py:121,148-149,160-161,264-266)::

    self._affine_transformer = TensorRandomFlip(axis=[1, 2], threshold=0.8)
    with FixRandomSeed(seed):
        x_tf = torch.stack([self._affine_transformer(x) for x in batch], dim=0)

``TensorRandomFlip.__call__`` (dc2:deepclustering2/augment/tensor_augment.py:31-41) draws one ``random.random()``
per axis, in axis order, for one (C, H, W) sample; ``FixRandomSeed`` (dc2:decorator/decorator.py:196-212) seeds the
global ``random`` module on entry and restores it on exit.  :func:`draw_flip_flags` replays exactly those draws with
a private generator, :func:`flip_stack` applies them to the whole batch in one kernel, and
``uda_from_logits(..., teacher_flips=flags)`` reads the teacher logits through the flips without a copy.
"""
from __future__ import annotations

import random
from typing import Sequence, Union

import torch
from torch import Tensor

from . import _lib
from .ops import FlipBatchFunction


def draw_flip_flags(seed: int, batch: int, axis: Union[int, Sequence[int], None] = (1, 2),
                    threshold: float = 0.8) -> Tensor:
    """The flips ``[TensorRandomFlip(axis, threshold)(x) for x in batch]`` performs under ``FixRandomSeed(seed)``,
    as one uint8 per sample (``_lib.FLIP_H`` = axis 1, ``_lib.FLIP_W`` = axis 2 of a (C, H, W) sample)."""
    if axis is None:
        return torch.zeros(batch, dtype=torch.uint8)
    axes = [axis] if isinstance(axis, int) else list(axis)
    bits = {1: _lib.FLIP_H, 2: _lib.FLIP_W, -2: _lib.FLIP_H, -1: _lib.FLIP_W}
    for a in axes:
        if a not in bits:
            raise ValueError(f"only the spatial axes 1 and 2 of a (C, H, W) sample can be flipped here, got axis {a}")
    assert 0 <= threshold <= 1
    rng = random.Random()
    rng.seed(seed)                      # the stream random.seed(seed) gives the module-level generator
    flags = []
    for _ in range(batch):
        f = 0
        for a in axes:                  # tensor_augment.py:34-36: one draw per axis, in order
            if rng.random() < threshold:
                f ^= bits[a]            # flipping the same axis twice cancels
        flags.append(f)
    return torch.tensor(flags, dtype=torch.uint8)


def flip_stack(batch: Tensor, seed_or_flags: Union[int, Tensor], axis=(1, 2), threshold: float = 0.8) -> Tensor:
    """``torch.stack([TensorRandomFlip(axis, threshold)(x) for x in batch])`` under ``FixRandomSeed(seed)`` in one
    launch; differentiable.  Pass the flags of :func:`draw_flip_flags` instead of the seed to reuse one draw for
    several maps of an iteration (images, logits, feature maps), as the reference does by re-seeding."""
    flags = seed_or_flags if isinstance(seed_or_flags, Tensor) else draw_flip_flags(seed_or_flags, len(batch), axis,
                                                                                    threshold)
    return FlipBatchFunction.apply(batch, flags.to(batch.device, non_blocking=True))


class TensorRandomFlip:
    """Batched twin of dc2's ``TensorRandomFlip``: ``T(batch, seed)`` equals the reference's seeded per-sample loop."""

    def __init__(self, axis=None, threshold=0.5) -> None:
        if isinstance(axis, int):
            self._axis = [axis]
        elif isinstance(axis, (list, tuple)):
            assert all(isinstance(a, int) for a in axis), axis
            self._axis = axis
        elif axis is None:
            self._axis = axis
        else:
            raise ValueError(str(axis))
        assert 0 <= threshold <= 1
        self._threshold = threshold

    def flags(self, seed: int, batch: int) -> Tensor:
        return draw_flip_flags(seed, batch, self._axis, self._threshold)

    def __call__(self, batch: Tensor, seed: int) -> Tensor:
        return flip_stack(batch, self.flags(seed, len(batch)))

    def __repr__(self):
        axis = "" if not self._axis else f" with axis={self._axis}."
        return f"{self.__class__.__name__}" + axis
