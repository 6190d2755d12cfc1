"""Importable alias of the package directory ``mi-based-regularized-semi-supervised-segmentation_b200``.

The project layout names the package directory after the reference repository; that name contains
hyphens and cannot be imported directly, so this tiny package points its ``__path__`` at it.
``import iic_b200`` / ``from iic_b200.losses.iic_loss import IIDLoss`` then resolve into that tree.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "mi-based-regularized-semi-supervised-segmentation_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _os, _f
