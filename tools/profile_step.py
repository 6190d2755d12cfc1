"""Tiny driver for ncu: a few eager forward+backward steps of the local IIC loss at a given shape.

    python tools/profile_step.py [--B 32 --K 10 --H 224 --W 224 --pad 1 --steps 3]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402

ap = argparse.ArgumentParser()
for k, v in dict(B=32, K=10, H=224, W=224, pad=1, steps=3, patch=512).items():
    ap.add_argument(f"--{k}", type=int, default=v)
a = ap.parse_args()
dev = torch.device("cuda:0")
iic_b200.set_check_mode("deferred")   # same as bench.py: device-side simplex / NaN flags
g = torch.Generator(device=dev).manual_seed(0)
base = torch.nn.functional.interpolate(torch.randn(a.B, a.K, a.H // 8, a.W // 8, device=dev, generator=g) * 3,
                                       size=(a.H, a.W), mode="bilinear")
x = (base + 0.5 * torch.randn(a.B, a.K, a.H, a.W, device=dev, generator=g)).softmax(1).requires_grad_(True)
y = (base + 0.5 * torch.randn(a.B, a.K, a.H, a.W, device=dev, generator=g)).softmax(1).requires_grad_(True)
gx = torch.randn(a.B, a.K, device=dev, generator=g).softmax(1).requires_grad_(True)
gy = torch.randn(a.B, a.K, device=dev, generator=g).softmax(1).requires_grad_(True)
crit = iic_b200.IIDSegmentationSmallPathLoss(padding=a.pad, patch_size=a.patch)
glob = iic_b200.IIDLoss()
for _ in range(a.steps):
    loss = crit(x, y) + glob(gx, gy)[0]
    torch.autograd.grad(loss, (x, y, gx, gy))
torch.cuda.synchronize()
print("loss", loss.item())
