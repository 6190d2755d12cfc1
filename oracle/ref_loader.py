"""Load the reference's own loss code from /root/reference -- TEST INFRASTRUCTURE ONLY.

Used in the build container by ``oracle/make_golden.py`` (and by a few ``not gpu`` tests that
skip when /root/reference is absent) to run the UNMODIFIED reference functions and pin the oracle
restatement to them.  /root/reference does not exist on the GPU box; nothing on the GPU path
imports this file.

Recipe: ``contrastyou/losses/iic_loss.py`` is loaded BY FILE PATH so that
``contrastyou/__init__.py``'s ``mkdir .data`` side effect never runs against the read-only mount,
with four stub modules standing in for imports that no longer exist on torch 2.11 / py3.12:
``termcolor``, ``torch._six``, ``deepclustering2.utils`` (only ``simplex`` is used, and the stub's
``simplex`` is the wheel's own function, extracted from the wheel by path) and
``contrastyou.helper`` (only ``average_iter``, taken from the reference file by path).
"""
from __future__ import annotations

import collections.abc
import importlib.util
import os
import sys
import tempfile
import types
import zipfile

_WHEEL = "deepclustering2-2.0.0-py3-none-any.whl"
_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")   # oracle/make_ref.py
_cache = {}


def _pick_root() -> str:
    """IIC_REFERENCE_ROOT if set, else /root/reference (the build container), else the verbatim copy that
    oracle/make_ref.py laid out under oracle/_ref/ (the GPU box, where /root/reference does not exist)."""
    env = os.environ.get("IIC_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _LOCAL_COPY):
        if os.path.isfile(os.path.join(cand, "contrastyou", "losses", "iic_loss.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "contrastyou", "losses", "iic_loss.py"))


def _load_by_path(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _extract_wheel_member(member: str) -> str:
    unpacked = os.path.join(REFERENCE_ROOT, member)          # oracle/_ref: the wheel is already extracted
    if os.path.isfile(unpacked):
        return unpacked
    out_dir = os.path.join(tempfile.gettempdir(), "iic_b200_dc2_extract")
    out = os.path.join(out_dir, member)
    if not os.path.isfile(out):
        with zipfile.ZipFile(os.path.join(REFERENCE_ROOT, _WHEEL)) as z:
            z.extract(member, out_dir)
    return out


def load():
    """Returns a namespace with the reference's IIDLoss, compute_joint, IIDSegmentationLoss,
    patch_generator, IIDSegmentationSmallPathLoss, KL_div, simplex, average_iter,
    weighted_average_iter -- the unmodified reference objects."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import torch

    saved = {k: sys.modules.get(k) for k in
             ("termcolor", "torch._six", "deepclustering2", "deepclustering2.utils",
              "deepclustering2.utils.general", "deepclustering2.loss",
              "deepclustering2.loss.kl_losses", "contrastyou", "contrastyou.helper")}
    try:
        tc = types.ModuleType("termcolor")
        tc.colored = lambda s, *a, **k: s
        six = types.ModuleType("torch._six")
        six.container_abcs = collections.abc
        six.string_classes = (str, bytes)
        six.int_classes = int
        six.inf = float("inf")
        sys.modules["termcolor"] = tc
        sys.modules["torch._six"] = six

        assertion = _load_by_path("_ref_dc2_assertion",
                                  _extract_wheel_member("deepclustering2/utils/assertion.py"))
        dc2 = types.ModuleType("deepclustering2")
        dc2u = types.ModuleType("deepclustering2.utils")
        dc2u.simplex = assertion.simplex
        dc2u.assert_list = assertion.assert_list
        dc2.utils = dc2u
        sys.modules["deepclustering2"] = dc2
        sys.modules["deepclustering2.utils"] = dc2u

        if not hasattr(collections, "Mapping"):  # contrastyou/helper/utils.py:12 uses collections.Mapping
            collections.Mapping = collections.abc.Mapping
            collections.MutableMapping = collections.abc.MutableMapping
            collections.Iterable = collections.abc.Iterable
        helper_utils = _load_by_path("_ref_contrastyou_helper_utils",
                                     os.path.join(REFERENCE_ROOT, "contrastyou", "helper", "utils.py"))
        cy = types.ModuleType("contrastyou")
        cyh = types.ModuleType("contrastyou.helper")
        cyh.average_iter = helper_utils.average_iter
        cy.helper = cyh
        sys.modules["contrastyou"] = cy
        sys.modules["contrastyou.helper"] = cyh

        iic = _load_by_path("_ref_iic_loss",
                            os.path.join(REFERENCE_ROOT, "contrastyou", "losses", "iic_loss.py"))
        # kl_losses.py does ``from ..utils.general import simplex, assert_list``: give it a stub
        # parent package whose utils.general exposes the wheel's own (identical) assertion helpers.
        dc2g = types.ModuleType("deepclustering2.utils.general")
        dc2g.simplex = assertion.simplex
        dc2g.assert_list = assertion.assert_list
        dc2u.general = dc2g
        dc2u.__path__ = []
        dc2.__path__ = []
        dc2l = types.ModuleType("deepclustering2.loss")
        dc2l.__path__ = []
        sys.modules["deepclustering2.utils.general"] = dc2g
        sys.modules["deepclustering2.loss"] = dc2l
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            kl = _load_by_path("deepclustering2.loss.kl_losses",
                               _extract_wheel_member("deepclustering2/loss/kl_losses.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    ns = types.SimpleNamespace(
        IIDLoss=iic.IIDLoss, compute_joint=iic.compute_joint,
        IIDSegmentationLoss=iic.IIDSegmentationLoss, patch_generator=iic.patch_generator,
        IIDSegmentationSmallPathLoss=iic.IIDSegmentationSmallPathLoss,
        KL_div=kl.KL_div, simplex=assertion.simplex, class2one_hot=assertion.class2one_hot,
        _assertion=assertion,
        average_iter=helper_utils.average_iter, weighted_average_iter=helper_utils.weighted_average_iter,
        torch=torch,
    )
    _cache["ns"] = ns
    return ns


def load_dice():
    """The reference's unmodified ``UniversalDice`` meter (dc2:meters2/individual_meters/general_dice_meter.py),
    loaded by path.  Its imports are satisfied by the wheel's own assertion helpers plus three trivial stubs for
    things the counting code never touches (the ``_Metric`` base class, ``to_float``, ``iter_average``)."""
    if "dice" in _cache:
        return _cache["dice"]
    ns = load()
    a = ns._assertion
    names = ("deepclustering2", "deepclustering2.utils", "deepclustering2.type", "deepclustering2.meters2",
             "deepclustering2.meters2.individual_meters", "deepclustering2.meters2.individual_meters._metric")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        mods = {k: types.ModuleType(k) for k in names}
        for m in mods.values():
            m.__path__ = []
        u = mods["deepclustering2.utils"]
        u.simplex, u.one_hot, u.class2one_hot, u.probs2one_hot = a.simplex, a.one_hot, a.class2one_hot, a.probs2one_hot
        u.iter_average = lambda xs: sum(xs) / len(list(xs))
        mods["deepclustering2.type"].to_float = float
        metric = mods["deepclustering2.meters2.individual_meters._metric"]
        metric._Metric = type("_Metric", (), {"__init__": lambda self: None})
        metric.MeterResultDict = dict
        sys.modules.update(mods)
        dice = _load_by_path("_ref_dc2_general_dice_meter",
                             _extract_wheel_member("deepclustering2/meters2/individual_meters/general_dice_meter.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["dice"] = dice.UniversalDice
    return dice.UniversalDice


def load_flip():
    """The reference's unmodified ``TensorRandomFlip`` (dc2:augment/tensor_augment.py:17-41) and ``FixRandomSeed``
    (dc2:decorator/decorator.py:196-212), loaded by path; returns ``(TensorRandomFlip, FixRandomSeed)``."""
    if "flip" in _cache:
        return _cache["flip"]
    ns = load()
    a = ns._assertion
    names = ("deepclustering2", "deepclustering2.utils")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        mods = {k: types.ModuleType(k) for k in names}
        for m in mods.values():
            m.__path__ = []
        mods["deepclustering2.utils"].assert_list = a.assert_list
        sys.modules.update(mods)
        aug = _load_by_path("_ref_dc2_tensor_augment", _extract_wheel_member("deepclustering2/augment/tensor_augment.py"))
        dec = _load_by_path("_ref_dc2_decorator", _extract_wheel_member("deepclustering2/decorator/decorator.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["flip"] = (aug.TensorRandomFlip, dec.FixRandomSeed)
    return _cache["flip"]
