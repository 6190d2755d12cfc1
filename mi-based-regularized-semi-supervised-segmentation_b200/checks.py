"""Host-side assertion policy for the loss modules.

The reference asserts ``simplex(...)`` and ``isnan(loss)`` on the host, which costs a device
synchronisation per assert.  Here the checks run on the device and
set bits in a sticky flag word; the policy decides when the host looks at it:

  "strict"   (default) -- same observable behaviour as the reference: every loss call ends with ONE
             read of the flag word and raises AssertionError (not a simplex,
             dc2:utils/assertion.py:56-65) or RuntimeError (NaN loss, iic_loss.py:147-148,184-185).
  "deferred" -- checks still run on the device but nothing synchronises; call
             :func:`raise_if_flagged` when convenient (e.g. once per iteration, next to the
             ``.item()`` the epocher does anyway).  CUDA-graph safe.
  "off"      -- the simplex kernels are skipped too (the NaN bit is still recorded for free).
"""
from __future__ import annotations

import contextlib

import torch

from . import _lib, ops

_MODE = "strict"
_MODES = ("strict", "deferred", "off")


def set_check_mode(mode: str) -> str:
    global _MODE
    if mode not in _MODES:
        raise ValueError(f"check mode must be one of {_MODES}, got {mode!r}")
    prev, _MODE = _MODE, mode
    return prev


def get_check_mode() -> str:
    return _MODE


@contextlib.contextmanager
def check_mode(mode: str):
    prev = set_check_mode(mode)
    try:
        yield
    finally:
        set_check_mode(prev)


def want_simplex_kernels() -> bool:
    return _MODE != "off"


def device_simplex(t: torch.Tensor):
    """Enqueue the simplex assertion for `t` (axis 1) unless checks are off."""
    if _MODE != "off":
        ops.ops.simplex_check(t, 1)


def read_and_clear(device) -> int:
    """Synchronising read of the sticky flag word of the current stream; clears it."""
    f = ops.flags_tensor(device)
    v = int(f.item())
    if v:
        f.zero_()
    return v


def raise_if_flagged(device, loss=None, simplex_msg: str = "input is not a simplex over dim 1"):
    v = read_and_clear(device)
    if v & _lib.FLAG_NOT_SIMPLEX:
        raise AssertionError(simplex_msg)
    if v & _lib.FLAG_BAD_LABEL:     # class2one_hot's `assert sset(seg, list(range(C)))`
        raise AssertionError("a label lies outside [0, C)")
    if v & _lib.FLAG_NAN_LOSS:
        raise RuntimeError(loss if loss is not None else "IIC loss is NaN")


def finish(device, loss=None, simplex_msg: str = "input is not a simplex over dim 1"):
    """End-of-call hook of every loss module."""
    if _MODE == "strict":
        raise_if_flagged(device, loss, simplex_msg)
