"""Times the pieces of the softmax-fused local term at config-2 shape: fused joint + finish, and the from-logits backward
on the tensor cores vs on the FFMA2 kernel (no_tc10).  python tools/time_from_logits.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402
from iic_b200 import ops as O  # noqa: E402

dev = torch.device("cuda:0")
iic_b200.set_check_mode("deferred")
B, K, H, W = 32, 10, 224, 224
g = torch.Generator(device=dev).manual_seed(0)
l1 = torch.randn(B, K, H, W, device=dev, generator=g) * 2
l2 = torch.randn(B, K, H, W, device=dev, generator=g) * 2
one = torch.ones((), device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


t = O.LocalTerm(l1, l2, None, 1, (H, W), (H, W), 1.0, logits=True)
(r,) = O._finish_terms([t], False)
gx, gy = torch.empty_like(l1), torch.empty_like(l2)
print("fused joint + finish (us):", round(timed(lambda: O._finish_terms([t], False)), 1))
print("from-logits backward, tensor cores (us):", round(timed(lambda: O._local_backward_logits_into(l1, l2, r["Wx"], r["Wy"], one, 1, 1.0, gx, gy)), 1))
iic_b200._lib.set_option("no_tc10", 1)
print("from-logits backward, FFMA2 (us):", round(timed(lambda: O._local_backward_logits_into(l1, l2, r["Wx"], r["Wy"], one, 1, 1.0, gx, gy)), 1))
iic_b200._lib.set_option("no_tc10", 0)
p1, p2 = l1.softmax(1), l2.softmax(1)
tp = O.LocalTerm(p1, p2, None, 1, (H, W), (H, W), 1.0)
(rp,) = O._finish_terms([tp], False)
print("probability backward, tensor cores (us):", round(timed(lambda: O._local_backward_into(p1, p2, None, rp["Wx"], rp["Wy"], one, 1, H, W, H, W, gx, gy)), 1))
print("probability joint + finish (us):", round(timed(lambda: O._finish_terms([tp], False)), 1))
