"""Torch-CPU port of the reference's loss path -- TEST INFRASTRUCTURE ONLY (the timed CPU baseline).

``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` time THIS file on the GPU
box's host cores: /root/reference is a Python tree that does not travel to the GPU box, so the
"reference arm" is this port (``cpu_baseline.kind == "port"``).  It deliberately issues the same
ATen operator sequence as the reference (simplex asserts with their host syncs, NCHW->CNHW
permute+contiguous, one ``F.conv2d`` whose "filter" is the whole H x W map, the min-shift, autograd
for the backward) so that its wall-clock is representative of the reference's own CPU path
(BASELINE.md section 2).  ``tests/test_oracle_golden.py`` pins it to the reference's outputs.

The product package never imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def simplex(t: torch.Tensor, axis: int = 1) -> bool:
    """dc2:deepclustering2/utils/assertion.py:56-65."""
    s = t.sum(axis).type(torch.float32)
    return bool(torch.allclose(s, torch.ones_like(s), rtol=1e-4, atol=1e-4))


def joint_global(x: torch.Tensor, y: torch.Tensor, symmetric: bool = True) -> torch.Tensor:
    """contrastyou/losses/iic_loss.py:74-94 (materialises the (N,K,K) outer product like the reference)."""
    assert simplex(x) and simplex(y)
    n, k = x.shape
    assert y.shape == (n, k)
    j = (x.unsqueeze(2) * y.unsqueeze(1)).sum(dim=0)
    if symmetric:
        j = (j + j.t()) / 2.0
    return j / j.sum()


def iid_loss(x: torch.Tensor, y: torch.Tensor, lamb: float = 1.0):
    """contrastyou/losses/iic_loss.py:43-71 -> (loss, loss_no_lamb, P)."""
    assert simplex(x) and simplex(y)
    k = x.shape[1]
    P = joint_global(x, y)
    pi = P.sum(dim=1).view(k, 1).expand(k, k)
    pj = P.sum(dim=0).view(1, k).expand(k, k)
    lp, lpi, lpj = torch.log(P + 1e-10), torch.log(pi + 1e-10), torch.log(pj + 1e-10)
    loss = (-P * (lp - lamb * lpj - lamb * lpi)).sum()
    loss_no_lamb = (-P * (lp - lpj - lpi)).sum()
    return loss, loss_no_lamb, P


def iid_segmentation_loss(x: torch.Tensor, y: torch.Tensor, padding: int, lamda: float = 1.0,
                          mask: torch.Tensor | None = None) -> torch.Tensor:
    """contrastyou/losses/iic_loss.py:107-149."""
    assert x.requires_grad and y.requires_grad
    assert simplex(x)
    assert x.shape == y.shape
    k = x.shape[1]
    if mask is not None:
        assert not mask.requires_grad
        x, y = x * mask, y * mask
    xc = x.permute(1, 0, 2, 3).contiguous()
    yc = y.permute(1, 0, 2, 3).contiguous()
    J = F.conv2d(xc, weight=yc, padding=(padding, padding))            # (Ki, Kj, T, T)
    J = J - J.min().detach() + 1e-16
    T = 2 * padding + 1
    J = J.permute(2, 3, 0, 1)
    J = J / J.sum(dim=3, keepdim=True).sum(dim=2, keepdim=True)
    J = (J + J.permute(0, 1, 3, 2)) / 2.0
    pi = J.sum(dim=2, keepdim=True).repeat(1, 1, k, 1)
    pj = J.sum(dim=3, keepdim=True).repeat(1, 1, 1, k)
    loss = (-J * (torch.log(J + 1e-16) - lamda * torch.log(pi + 1e-16)
                  - lamda * torch.log(pj + 1e-16))).sum() / (T * T)
    if torch.isnan(loss):
        raise RuntimeError(loss)
    return loss


def _windows(h, w, ph, pw, sh, sw):
    """contrastyou/losses/iic_loss.py:152-160."""
    hs = list(range(0, h - ph, sh)) + [max(h - ph, 0)]
    ws = list(range(0, w - pw, sw)) + [max(w - pw, 0)]
    return [(a, min(a + ph, h), b, min(b + pw, w)) for a in hs for b in ws]


def iid_segmentation_small_path_loss(x, y, padding: int, patch_size: int, lamda: float = 1.0, mask=None):
    """contrastyou/losses/iic_loss.py:171-186."""
    assert x.shape == y.shape
    st = patch_size // 2
    losses = []
    for (a, b, c, d) in _windows(x.shape[2], x.shape[3], patch_size, patch_size, st, st):
        m = None if mask is None else mask[:, :, a:b, c:d]
        losses.append(iid_segmentation_loss(x[:, :, a:b, c:d], y[:, :, a:b, c:d], padding, lamda, m))
    return sum(losses) / float(len(losses))


def kl_div(prob, target, eps: float = 1e-16):
    """dc2:deepclustering2/loss/kl_losses.py:107-126 (reduction='mean', weight=None)."""
    assert prob.shape == target.shape and simplex(prob) and simplex(target)
    return (-target * torch.log((prob + eps) / (target + eps))).sum(1).mean()


def mse(prob, target):
    """torch.nn.MSELoss() (semi_seg/trainer.py:137,194)."""
    return F.mse_loss(prob, target)


def local_plus_global_step(logits1, logits2, glogits1, glogits2, padding: int, patch_size: int):
    """One fwd+bwd 'step' of BASELINE config 2 on the CPU, as the reference epocher would run it
    (semi_seg/epocher.py:269-275): head softmaxes, local small-path loss + global IIDLoss, backward
    to the logits.  Returns the scalar loss value."""
    l1 = logits1.detach().requires_grad_(True)
    l2 = logits2.detach().requires_grad_(True)
    g1 = glogits1.detach().requires_grad_(True)
    g2 = glogits2.detach().requires_grad_(True)
    loc = iid_segmentation_small_path_loss(l1.softmax(1), l2.softmax(1), padding, patch_size)
    glo = iid_loss(g1.softmax(1), g2.softmax(1))[0]
    total = loc + glo
    total.backward()
    return float(total)
