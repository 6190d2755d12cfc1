"""UDA consistency criteria on the B200 CUDA path.

``KL_div`` mirrors dc2:deepclustering2/loss/kl_losses.py:76-129; ``MSELoss`` is the twin of the
``torch.nn.MSELoss()`` the reference builds at semi_seg/trainer.py:137,194.  Both are called as
``criterion(softmax(tf_logits), softmax(logits_tf).detach())`` at semi_seg/epocher.py:221-224.
``uda_from_logits`` fuses those two softmaxes into the loss kernels.
"""
from __future__ import annotations

from typing import List, Optional, Union

import torch
from torch import Tensor, nn

from .. import checks
from ..ops import UDAFunction

_KIND = {"mse": 0, "kl": 1}


class KL_div(nn.Module):
    r"""KL(target || prob) = -\sum target * log((prob+eps)/(target+eps)); reduction "mean" or "sum"."""

    def __init__(self, reduction="mean", eps=1e-16, weight: Union[List[float], Tensor] = None, verbose=True):
        super().__init__()
        assert reduction in ("mean", "sum", "none"), reduction
        self._eps = eps
        self._reduction = reduction
        self._weight: Optional[Tensor] = weight
        if weight is not None:
            assert isinstance(weight, (list, Tensor)), type(weight)
            if isinstance(weight, list):
                assert all(isinstance(x, (int, float)) for x in weight)
                self._weight = torch.Tensor(weight).float()
            else:
                self._weight = weight.float()
            self._weight = self._weight / self._weight.sum() * len(self._weight)   # kl_losses.py:100
        if verbose:
            print(f"Initialized {self.__class__.__name__} \nwith weight={self._weight} and reduction={self._reduction}.")

    def forward(self, prob: Tensor, target: Tensor, **kwargs) -> Tensor:
        do_assert = not kwargs.get("disable_assert")
        if do_assert:
            assert prob.shape == target.shape
            assert not target.requires_grad
            assert prob.requires_grad
        if self._reduction == "none":
            raise NotImplementedError("KL_div(reduction='none') is not on the udaiic path; use the reference for it")
        if self._weight is not None:
            assert len(self._weight) == target.shape[1]
        loss = UDAFunction.apply(prob, target, _KIND["kl"], float(self._eps), self._weight, False,
                                 bool(do_assert and checks.want_simplex_kernels()))
        if self._reduction == "sum":
            npix = prob.numel() // prob.shape[1]
            loss = loss * float(npix)
        if do_assert:
            checks.finish(prob.device, loss, "prob / target is not a simplex over dim 1")
        return loss

    def __repr__(self):
        return f"{self.__class__.__name__}\n, weight={self._weight}"


class MSELoss(nn.Module):
    """``torch.nn.MSELoss()`` (mean over every element) with the fused CUDA forward/backward."""

    def __init__(self, reduction: str = "mean") -> None:
        super().__init__()
        assert reduction in ("mean", "sum"), reduction
        self.reduction = reduction

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        assert input.shape == target.shape, (input.shape, target.shape)
        loss = UDAFunction.apply(input, target.detach(), _KIND["mse"], 0.0, None, False, False)
        if self.reduction == "sum":
            loss = loss * float(input.numel())
        return loss


def uda_from_logits(student_logits: Tensor, teacher_logits: Tensor, kind: str = "mse", eps: float = 1e-16) -> Tensor:
    """criterion(softmax(student_logits, 1), softmax(teacher_logits, 1).detach()) in one kernel each way
    (semi_seg/epocher.py:221-224); the gradient is returned w.r.t. ``student_logits``."""
    assert student_logits.shape == teacher_logits.shape
    return UDAFunction.apply(student_logits, teacher_logits.detach(), _KIND[kind], float(eps), None, True, False)
