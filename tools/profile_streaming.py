"""Tiny driver for ncu: the streaming kernels of SURVEY 8f (UDA, UDA through flips, flip_stack, supervised KL + Dice)
at (32, 4, 224, 224), a few eager forward+backward calls each, inputs rotating over 4 sets so they come from HBM.

    python tools/profile_streaming.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402

dev = torch.device("cuda:0")
iic_b200.set_check_mode("deferred")
B, C, H, W = 32, 4, 224, 224
g = torch.Generator(device=dev).manual_seed(0)
rs = [(torch.randn(B, C, H, W, device=dev, generator=g) * 2).requires_grad_(True) for _ in range(4)]
rt = [torch.randn(B, C, H, W, device=dev, generator=g) * 2 for _ in range(4)]
rl = [torch.randint(0, C, (B, H, W), device=dev, generator=g) for _ in range(4)]
rf = iic_b200.draw_flip_flags(4321, B).to(dev)
for rep in range(2):
    for a, t, l in zip(rs, rt, rl):
        torch.autograd.grad(iic_b200.uda_from_logits(a, t, "mse"), (a,))
        torch.autograd.grad(iic_b200.uda_from_logits(a, t, "mse", teacher_flips=rf), (a,))
        iic_b200.flip_stack(t, rf)
        torch.autograd.grad(iic_b200.sup_kl_from_logits(a, l, return_dice=True)[0], (a,))
torch.cuda.synchronize()
print("ok")
