from ._utils import IICLossWrapper, IIDLoss  # noqa: F401
from .meters import DeferredScalarMeters  # noqa: F401
