from ._utils import ClusterHead, LocalClusterHead  # noqa: F401
