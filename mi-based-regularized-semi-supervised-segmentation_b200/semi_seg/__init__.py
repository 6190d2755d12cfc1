from ._utils import IICLossWrapper, IIDLoss  # noqa: F401
