"""Cluster heads with their S sub-heads batched (host half).

The reference's ``ClusterHead`` / ``LocalClusterHead`` (contrastyou/trainer/_utils.py:96-134,137-168) run every
sub-head as its own ``nn.Sequential``: S adaptive poolings + S ``Linear``s, or S 1x1 convolutions that each re-read the
feature map, and S softmaxes.  The classes below keep the constructor arguments, the module tree and therefore the
``state_dict`` keys of the reference (``_headers.<s>.<index>.weight``), so its checkpoints load with ``strict=True``,
but evaluate the "linear" head type as ONE pooling + ONE ``linear`` / ONE 1x1 ``conv2d`` with the sub-heads' weights
concatenated (stock torch / cuDNN: the backbone side stays out of scope), followed by one softmax over a
(N, S, K, ...) view.  ``forward`` returns the reference's list of S probability maps (channel-block views of one
tensor: W-stride 1, which is all the loss kernels ask for); ``logits`` returns the S logit maps for the softmax-fused
losses (``IIDSegmentationSmallPathLoss.from_logits``).  The "mlp" head type runs per sub-head as in the reference.
"""
from __future__ import annotations

from typing import List

import torch
from torch import Tensor, nn
from torch.nn import functional as F


class Flatten(nn.Module):
    def forward(self, features):
        return features.view(features.shape[0], -1)


class Identical(nn.Module):
    def forward(self, input):
        return input


class Normalize(nn.Module):
    def forward(self, input):
        return F.normalize(input, p=2, dim=1)


class SoftmaxWithT(nn.Softmax):
    """softmax(input / T) (contrastyou/trainer/_utils.py:15-23; the reference divides in place)."""

    def __init__(self, dim, T: float = 0.1) -> None:
        super().__init__(dim)
        self._T = T

    def forward(self, input: Tensor) -> Tensor:
        return super().forward(input / self._T)


def _split_heads(batched: Tensor, S: int, K: int, normalize: bool, T: float, want_probs: bool) -> List[Tensor]:
    """(N, S*K, *spatial) logits of all sub-heads -> S maps (N, K, *spatial), normalised / softmaxed per sub-head."""
    n, spatial = batched.shape[0], batched.shape[2:]
    v = batched.view(n, S, K, *spatial)
    if normalize:
        v = F.normalize(v, p=2, dim=2)
    if want_probs:
        v = torch.softmax(v / T, dim=2)
    return list(v.unbind(1))


class ClusterHead(nn.Module):
    """Encoder-side head: global average pool -> Linear -> softmax per sub-head (contrastyou/trainer/_utils.py:96-134)."""

    def __init__(self, input_dim, num_clusters=5, num_subheads=10, head_type="linear", T=1, normalize=False) -> None:
        super().__init__()
        assert head_type in ("linear", "mlp"), head_type
        self._input_dim, self._num_clusters, self._num_subheads = input_dim, num_clusters, num_subheads
        self._T, self._normalize, self._head_type = T, normalize, head_type

        def sub_header():
            layers = [nn.AdaptiveAvgPool2d((1, 1)), Flatten()]
            if head_type == "linear":
                layers += [nn.Linear(input_dim, num_clusters)]
            else:
                layers += [nn.Linear(input_dim, 128), nn.LeakyReLU(0.01, inplace=True), nn.Linear(128, num_clusters)]
            layers += [Normalize() if normalize else Identical(), SoftmaxWithT(1, T=T)]
            return nn.Sequential(*layers)

        self._headers = nn.ModuleList([sub_header() for _ in range(num_subheads)])

    def _batched(self, features: Tensor, want_probs: bool) -> List[Tensor]:
        if self._head_type != "linear":
            outs = [h[:-1](features) for h in self._headers]
            return [torch.softmax(o / self._T, dim=1) for o in outs] if want_probs else outs
        pooled = F.adaptive_avg_pool2d(features, (1, 1)).flatten(1)                 # once, not once per sub-head
        w = torch.cat([h[2].weight for h in self._headers], dim=0)
        b = torch.cat([h[2].bias for h in self._headers], dim=0)
        return _split_heads(F.linear(pooled, w, b), self._num_subheads, self._num_clusters, self._normalize, self._T,
                            want_probs)

    def logits(self, features: Tensor) -> List[Tensor]:
        return self._batched(features, False)

    def forward(self, features: Tensor) -> List[Tensor]:
        return self._batched(features, True)


class LocalClusterHead(nn.Module):
    """Decoder-side head: 1x1 conv -> softmax over the clusters per sub-head (contrastyou/trainer/_utils.py:137-168)."""

    def __init__(self, input_dim, head_type="linear", num_clusters=10, num_subheads=10, T=1, interm_dim=64,
                 normalize=False) -> None:
        super().__init__()
        assert head_type in ("linear", "mlp"), head_type
        self._num_clusters, self._num_subheads = num_clusters, num_subheads
        self._T, self._normalize, self._head_type = T, normalize, head_type

        def sub_header():
            if head_type == "linear":
                layers = [nn.Conv2d(input_dim, num_clusters, 1, 1, 0)]
            else:
                layers = [nn.Conv2d(input_dim, interm_dim, 1, 1, 0), nn.LeakyReLU(0.01, inplace=True),
                          nn.Conv2d(interm_dim, num_clusters, 1, 1, 0)]
            layers += [Normalize() if normalize else Identical(), SoftmaxWithT(1, T=T)]
            return nn.Sequential(*layers)

        self._headers = nn.ModuleList([sub_header() for _ in range(num_subheads)])

    def _batched(self, features: Tensor, want_probs: bool) -> List[Tensor]:
        if self._head_type != "linear":
            outs = [h[:-1](features) for h in self._headers]
            return [torch.softmax(o / self._T, dim=1) for o in outs] if want_probs else outs
        w = torch.cat([h[0].weight for h in self._headers], dim=0)                  # (S*K, C_f, 1, 1)
        b = torch.cat([h[0].bias for h in self._headers], dim=0)
        return _split_heads(F.conv2d(features, w, b), self._num_subheads, self._num_clusters, self._normalize,
                            self._T, want_probs)                                    # the feature map is read once

    def logits(self, features: Tensor) -> List[Tensor]:
        return self._batched(features, False)

    def forward(self, features: Tensor) -> List[Tensor]:
        return self._batched(features, True)
