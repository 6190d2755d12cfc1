"""Timeline of one finish launch at config 2 (csrc/finish.cu built with -DIIC_FIN_TRACE into tools/_bin/libiic_fintrace.so).

    tools/build_fin_trace.sh && python tools/fin_trace.py
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402
from iic_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_bin", "libiic_fintrace.so")   # before the first load()

dev = torch.device("cuda:0")
iic_b200.set_check_mode("deferred")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(32, 10, 224, 224, device=dev, generator=g).softmax(1).requires_grad_(True)
y = torch.randn(32, 10, 224, 224, device=dev, generator=g).softmax(1).requires_grad_(True)
gx = torch.randn(32, 10, device=dev, generator=g).softmax(1).requires_grad_(True)
gy = torch.randn(32, 10, device=dev, generator=g).softmax(1).requires_grad_(True)
crit = iic_b200.IIDSegmentationSmallPathLoss(padding=1, patch_size=512)
glob = iic_b200.IIDLoss()
lib = _lib.load()
lib.iic_debug_fin_trace.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
buf = (C.c_ulonglong * 16)()
for it in range(6):
    lib.iic_debug_fin_trace(buf, 1)
    losses = iic_b200.iic_losses([(crit, x, y), (glob, gx, gy)])
    torch.cuda.synchronize()
    lib.iic_debug_fin_trace(buf, 0)
    t = [int(v) for v in buf]
    t0 = t[0]
    print(f"run {it}: phase-1 end (last CTA) {t[1] - t0} ns | last CTA: ticket {t[2] - t0}, J staged {t[3] - t0}, patch min {t[4] - t0}, "
          f"units done {t[5] - t0}, end {t[6] - t0} | CTA 0 phase 1 end {t[7] - t0}, last-index CTA {t[8] - t0}; tail: entered {t[10] - t0}, summed {t[9] - t0}")
