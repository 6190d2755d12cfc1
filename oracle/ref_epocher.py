"""The reference's own training iteration (``semi_seg/epocher.py`` ``UDAIICEpocher``) on synthetic loaders -- BENCH /
TEST INFRASTRUCTURE ONLY, never imported by the product.

``import_reference()`` puts ``oracle/_ref/reference`` (the verbatim copy made by ``oracle/make_ref.py``; in the build
container ``/root/reference`` is copied to a scratch directory first, because ``contrastyou/__init__.py:5-7`` does a
``mkdir .data`` next to itself) on ``sys.path`` behind the ~30 lines of import shims the 2020 code needs on torch 2.11 /
Python 3.12: stub modules for packages that are not installed and not reached by the epocher
(``termcolor``, ``tensorboardX``, ``skimage``, ``medpy``, ``gdown``, ``matplotlib``, ``SimpleITK``, ``easydict``,
``torch_optimizer``, ``apex``), ``torch._six``, the ``collections.Mapping`` aliases and two private tqdm names.
No reference file is modified.

``build_epocher(...)`` then assembles what ``UDAIICTrainer._init`` / ``_run_epoch`` assemble (semi_seg/trainer.py:
150-166, 187-206): ``UNet(**Arch)``, ``ProjectorWrapper`` + ``IICLossWrapper`` from the yaml parameters, Adam over model
and heads, and one ``UDAIICEpocher`` over synthetic labeled / unlabeled loaders.  ``swap="b200"`` applies INTEGRATION.md
section 2 -- the three import swaps -- by monkeypatching the names the reference modules imported; nothing else differs
between the two arms.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_LOCAL_COPY = os.path.join(HERE, "_ref", "reference")
_imported = {}


class _Permissive(types.ModuleType):
    """A stub package: any attribute is another stub (callable, returns None)."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__all__ = []

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        sub = _Permissive(self.__name__ + "." + item)
        setattr(self, item, sub)
        sys.modules.setdefault(sub.__name__, sub)
        return sub

    def __call__(self, *a, **k):
        return None


class _StubFinder:
    """Meta-path finder: ``import <stub package>.<anything>`` yields another permissive stub."""

    tops = set()

    @classmethod
    def find_spec(cls, fullname, path=None, target=None):
        if fullname.split(".")[0] not in cls.tops:
            return None
        import importlib.machinery
        return importlib.machinery.ModuleSpec(fullname, cls, is_package=True)

    @staticmethod
    def create_module(spec):
        return _Permissive(spec.name)

    @staticmethod
    def exec_module(module):
        pass


def reference_root() -> str:
    env = os.environ.get("IIC_REFERENCE_ROOT")
    if env and os.path.isdir(os.path.join(env, "semi_seg")) and os.path.isdir(os.path.join(env, "deepclustering2")):
        return env
    if os.path.isdir(os.path.join(_LOCAL_COPY, "semi_seg")):
        return _LOCAL_COPY
    raise RuntimeError("oracle/_ref/reference is missing: run `python oracle/make_ref.py` where /root/reference exists")


def import_reference():
    """Returns a namespace with the reference modules ``semi_seg.epocher``, ``semi_seg._utils``, ``semi_seg.trainer``
    (not imported: it pulls the data loaders), ``contrastyou.arch.UNet`` and dc2's ``KL_div`` -- all unmodified."""
    if "ns" in _imported:
        return _imported["ns"]
    import torch
    import tqdm.utils
    root = reference_root()
    for name in ("termcolor", "tensorboardX", "skimage", "skimage.io", "medpy", "medpy.metric", "medpy.metric.binary",
                 "gdown", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "SimpleITK", "easydict", "apex",
                 "sklearn_extra"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = _Permissive(name)
                _StubFinder.tops.add(name.split(".")[0])
    if _StubFinder not in sys.meta_path:
        sys.meta_path.append(_StubFinder)
    if isinstance(sys.modules["termcolor"], _Permissive):
        sys.modules["termcolor"].colored = lambda s, *a, **k: s
    if isinstance(sys.modules["tensorboardX"], _Permissive):      # dc2:writer/SummaryWriter.py:15 subclasses it
        sys.modules["tensorboardX"].SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
    if "torch_optimizer" not in sys.modules:
        to = types.ModuleType("torch_optimizer")
        to.__all__ = []
        sys.modules["torch_optimizer"] = to
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.container_abcs = collections.abc
        six.string_classes = (str, bytes)
        six.int_classes = int
        six.inf = float("inf")
        sys.modules["torch._six"] = six
        torch._six = six
    for n in ("Mapping", "MutableMapping", "Iterable", "Callable", "Sequence"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    if not hasattr(tqdm.utils, "_OrderedDict"):
        tqdm.utils._OrderedDict = collections.OrderedDict
    if not hasattr(tqdm.utils, "_basestring"):
        tqdm.utils._basestring = str
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        epocher = importlib.import_module("semi_seg.epocher")
        utils = importlib.import_module("semi_seg._utils")
        arch = importlib.import_module("contrastyou.arch")
        dc2_loss = importlib.import_module("deepclustering2.loss")
        iic_loss = importlib.import_module("contrastyou.losses.iic_loss")
    ns = types.SimpleNamespace(root=root, epocher=epocher, utils=utils, UNet=arch.UNet, KL_div=dc2_loss.KL_div,
                               iic_loss=iic_loss, torch=torch)
    _imported["ns"] = ns
    return ns


class SyntheticLoader:
    """Endless iterator of pre-built batches in the format ``preprocess_input_with_twice_transformation`` unpacks
    (contrastyou/epocher/_utils.py:25-28): ``[[(image, target), (image_tf, target_tf)], filenames, partitions, groups]``.
    The tensors live on `device` already ("a generator of pre-built device batches")."""

    def __init__(self, torch, batch, device, labeled, n_sets=4, seed=0, hw=224, num_classes=4):
        g = torch.Generator().manual_seed(seed)
        self.sets = []
        for _ in range(n_sets):
            img = torch.rand(batch, 1, hw, hw, generator=g).to(device)
            tgt = torch.randint(0, num_classes, (batch, 1, hw, hw), generator=g).to(device)
            names = [f"patient{i:03d}_00_{i:02d}" for i in range(batch)]
            self.sets.append([[(img, tgt), (img.clone(), tgt.clone())], names, ["0"] * batch,
                              [f"patient{i:03d}" for i in range(batch)]])
        self.i = 0
        self.labeled = labeled

    def __iter__(self):
        return self

    def __next__(self):
        s = self.sets[self.i % len(self.sets)]
        self.i += 1
        return s


# config/semi.yaml:46-58 (the defaults) and the BASELINE config-1 variant
YAML_DEFAULT = dict(feature_names=["Conv5", "Up_conv3", "Up_conv2"], feature_importance=[1.0, 0.5, 0.5],
                    num_clusters=20, num_subheads=5, paddings=[1, 3], patch_sizes=1024, uda="mse",
                    uda_weight=5.0, iic_weight=0.1)
CONFIG1 = dict(feature_names=["Conv5"], feature_importance=[1.0], num_clusters=10, num_subheads=5, paddings=1,
               patch_sizes=1024, uda="mse", uda_weight=5.0, iic_weight=0.1)


def build_epocher(cfg, device, labeled_bs=4, unlabeled_bs=4, num_batches=3, swap="reference", seed=10, lr=1e-7):
    """One ``UDAIICEpocher`` exactly as ``UDAIICTrainer._run_epoch`` builds it (semi_seg/trainer.py:197-206).
    swap = "reference": the reference's own loss classes; "b200": INTEGRATION.md section 2 applied by monkeypatching
    (``semi_seg._utils`` names for the IIC losses, the trainer's criterion table for UDA)."""
    ns = import_reference()
    torch = ns.torch
    torch.manual_seed(seed)
    import random
    random.seed(seed)
    U = ns.utils
    saved = (U._IIDLoss, U.IIDSegmentationSmallPathLoss, U.IIDLoss)
    try:
        if swap == "b200":
            import iic_b200
            from iic_b200.losses import iic_loss as b200_losses
            # semi_seg/_utils.py:8  `from contrastyou.losses.iic_loss import IIDLoss as _IIDLoss, IIDSegmentationSmallPathLoss`
            U._IIDLoss = b200_losses.IIDLoss
            U.IIDSegmentationSmallPathLoss = b200_losses.IIDSegmentationSmallPathLoss
            U.IIDLoss = type("IIDLoss", (b200_losses.IIDLoss,),
                             {"forward": lambda self, a, b: b200_losses.IIDLoss.forward(self, a, b)[0]})  # :12-15
            reg_criterion = {"mse": iic_b200.MSELoss(), "kl": iic_b200.KL_div(verbose=False)}[cfg["uda"]]  # trainer.py:194
        else:
            reg_criterion = {"mse": torch.nn.MSELoss(), "kl": ns.KL_div(verbose=False)}[cfg["uda"]]
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model = ns.UNet(input_dim=1, num_classes=4).to(device)
            proj = U.ProjectorWrapper()
            head = dict(num_clusters=cfg["num_clusters"], num_subheads=cfg["num_subheads"], head_types="linear",
                        normalize=False)
            proj.init_encoder(feature_names=cfg["feature_names"], **head)
            proj.init_decoder(feature_names=cfg["feature_names"], **head)
            proj = proj.to(device)
            wrapper = U.IICLossWrapper(feature_names=cfg["feature_names"], paddings=cfg["paddings"],
                                       patch_sizes=cfg["patch_sizes"])
            sup = ns.KL_div(verbose=False)
    finally:
        U._IIDLoss, U.IIDSegmentationSmallPathLoss, U.IIDLoss = saved
    from itertools import chain
    optim = torch.optim.Adam(chain(model.parameters(), proj.parameters()), lr=lr, weight_decay=1e-5)
    fi = [float(x) for x in cfg["feature_importance"]]
    fi = [x / sum(fi) for x in fi]                                   # semi_seg/trainer.py:46-49
    lab = SyntheticLoader(torch, labeled_bs, device, True, seed=seed + 1)
    unl = SyntheticLoader(torch, unlabeled_bs, device, False, seed=seed + 2)
    ep = ns.epocher.UDAIICEpocher(model, proj, optim, lab, unl, sup, reg_criterion, wrapper, num_batches=num_batches,
                                  cur_epoch=0, device=device, feature_position=list(cfg["feature_names"]),
                                  feature_importance=fi, cons_weight=cfg["uda_weight"], iic_weight=cfg["iic_weight"])
    return ep, model, proj


def run_epoch(ep):
    """``_Epocher.run()`` (dc2:epoch/_epocher.py:82-87) with its tqdm bar silenced; returns the meter dict."""
    import contextlib
    import io
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        return ep.run()


if __name__ == "__main__":
    import time
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ep, *_ = build_epocher(CONFIG1, dev, num_batches=2)
    t0 = time.perf_counter()
    out = run_epoch(ep)
    print(f"{dev}: 2 iterations of the reference UDAIICEpocher (config 1) in {time.perf_counter() - t0:.2f} s")
    print({k: v for k, v in out.items()} if hasattr(out, "items") else out)
