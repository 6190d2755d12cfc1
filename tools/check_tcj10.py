"""Tensor-core joint (csrc/local_fwd_tcj10.cu) against an fp64 convolution and the FFMA2 joint: error and time.

    python tools/check_tcj10.py            # on a B200
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402
from iic_b200 import _lib, ops  # noqa: E402


def joint64(x, y, pad):
    # iic_loss.py:120-123 in fp64
    x64, y64 = x.double(), y.double()
    J = F.conv2d(x64.permute(1, 0, 2, 3), y64.permute(1, 0, 2, 3), padding=pad)      # (K, K, T, T)
    return J.permute(2, 3, 0, 1).contiguous()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    shapes = [(32, 10, 224, 224, 1.0), (32, 10, 224, 224, 8.0), (4, 10, 224, 224, 1.0), (6, 9, 160, 128, 1.0),
              (5, 10, 131, 236, 2.0), (40, 4, 64, 8, 1.0), (32, 10, 56, 56, 1.0)]
    worst = 0.0
    for (B, K, H, W, sharp) in shapes:
        x = torch.softmax(sharp * torch.randn(B, K, H, W, device=dev), 1)
        y = torch.softmax(sharp * torch.randn(B, K, H, W, device=dev), 1)
        ref = joint64(x, y, 1)
        out = {}
        for name, off in (("ffma2", 1), ("tc", 0)):
            _lib.set_option("no_tcj10", off)
            J = ops._local_joint(x, y, None, 1, H, W, H, W, check_simplex=True)[0]
            torch.cuda.synchronize()
            err = ((J - ref).abs().max() / ref.sum() * 9).item()          # relative to one displacement's total mass
            rel = ((J - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()
            t = timeit(lambda: ops._local_joint(x, y, None, 1, H, W, H, W))
            out[name] = (err, rel, t)
        _lib.set_option("no_tcj10", 0)
        fl = ops.flags_tensor(dev).item()
        print(f"B={B} K={K} H={H} W={W} sharp={sharp}: " +
              "  ".join(f"{n}: err/mass {e:.2e} max-rel {r:.2e} {t:.1f} us" for n, (e, r, t) in out.items()) + f"  flags={fl}")
        worst = max(worst, out["tc"][0])
    # the assertion must still fire
    x = torch.rand(32, 10, 224, 224, device=dev)
    y = torch.softmax(torch.randn(32, 10, 224, 224, device=dev), 1)
    ops.flags_tensor(dev).zero_()
    ops._local_joint(x, y, None, 1, 224, 224, 224, 224, check_simplex=True)
    torch.cuda.synchronize()
    print("non-simplex x -> flags", ops.flags_tensor(dev).item())
    ops.flags_tensor(dev).zero_()
    print("worst tc err/mass", worst)


if __name__ == "__main__":
    main()
