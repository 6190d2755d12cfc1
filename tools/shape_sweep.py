"""Time the local IIC loss (fwd+bwd, probability inputs) over the BASELINE.json shapes on one GPU.

    python tools/shape_sweep.py [substring of the shape name]

Prints one JSON line per shape: Mpx/s, fraction of the HBM roofline (24*K bytes per pixel over the
measured copy bandwidth) and fraction of the FP32 FMA peak (6*K^2*T^2 flop per pixel).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402

SHAPES = [  # name, B, K, H, W, pad
    ("cfg2  Up_conv2 K=10 p=1", 32, 10, 224, 224, 1),
    ("cfg2' Up_conv2 K=20 p=1", 32, 20, 224, 224, 1),
    ("cfg3  Up_conv3 K=20 p=1 (8/GPU)", 8, 20, 112, 112, 1),
    ("cfg3  Up_conv2 K=20 p=3 (8/GPU)", 8, 20, 224, 224, 3),
    ("cfg4  512^2 K=20 p=3 (16/GPU)", 16, 20, 512, 512, 3),
    ("cfg5  K=128 p=1 (4/GPU sample)", 4, 128, 224, 224, 1),
    ("cfg5  K=128 p=1 (32/GPU, full)", 32, 128, 224, 224, 1),
]
dev = torch.device("cuda:0")
iic_b200.set_check_mode("deferred")
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = float(peaks["hbm_gbs"])
except Exception:  # noqa: BLE001
    hbm = 6650.0
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, B, K, H, W, pad in SHAPES:
    if only not in name:
        continue
    g = torch.Generator(device=dev).manual_seed(1)
    base = torch.nn.functional.interpolate(torch.randn(B, K, H // 8, W // 8, device=dev, generator=g) * 3, size=(H, W), mode="bilinear")
    x = (base + 0.5 * torch.randn(B, K, H, W, device=dev, generator=g)).softmax(1).requires_grad_(True)
    y = (base + 0.5 * torch.randn(B, K, H, W, device=dev, generator=g)).softmax(1).requires_grad_(True)
    crit = iic_b200.IIDSegmentationSmallPathLoss(padding=pad, patch_size=1024)

    def step():
        return torch.autograd.grad(crit(x, y), (x, y))

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    px = B * H * W
    T = 2 * pad + 1
    print(json.dumps({"shape": name, "B": B, "K": K, "H": H, "W": W, "pad": pad, "ms": round(ms, 3),
                      "mpx_s": round(px / ms / 1e3, 1),
                      "hbm_roofline_frac": round(24.0 * K * px / (ms * 1e-3) / 1e9 / hbm, 4),
                      "fp32_fma_frac": round(3.0 * K * K * T * T * px / (ms * 1e-3) / (148 * 128 * 1.965e9), 4)}))
    del x, y, base
    torch.cuda.empty_cache()
iic_b200.raise_if_flagged(dev)
