"""Deferred scalar meters for the epocher.

The reference records every scalar of an iteration with ``meter.add(tensor.item())`` -- ``sup_loss``, ``reg_loss``,
``mi``, one ``individual_mis`` entry per feature layer, ``uda`` (semi_seg/epocher.py:181-187,225,278-282): five or
more device synchronisations per iteration, each of which stalls the host behind the whole iteration's kernels.
:class:`DeferredScalarMeters` keeps the values on the device: ``record(sup_loss=..., reg_loss=..., ...)`` stacks the
iteration's 0-d tensors into one row of a preallocated (capacity, M) buffer (one ``torch.stack`` + one row copy, no
synchronisation, CUDA-graph friendly), and ``summary()`` does ONE device-to-host transfer and returns, per name, the
same ``{"mean": ...}`` dict dc2's ``AverageValueMeter.summary()`` gives
(dc2:deepclustering2/meters2/individual_meters/averagemeter.py:8-52).  ``tracking_status()`` mirrors
``MeterInterface.tracking_status`` for the progress bar and can be called every N iterations instead of every one.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import Tensor


class DeferredScalarMeters:
    def __init__(self, names: List[str], capacity: int = 1024, device=None) -> None:
        assert len(names) == len(set(names)) and capacity > 0
        self.names = list(names)
        self._index = {n: i for i, n in enumerate(self.names)}
        self._capacity = int(capacity)
        self._device = device
        self._buf: Optional[Tensor] = None
        self._n = 0
        self._host_sum = torch.zeros(len(names), dtype=torch.float64)      # rows already folded to the host
        self._host_rows = torch.zeros(len(names), dtype=torch.float64)     # per name: how many non-NaN values
        self._host_n = 0

    def reset(self) -> None:
        self._n = 0
        self._host_sum.zero_()
        self._host_rows = torch.zeros(len(self.names), dtype=torch.float64)
        self._host_n = 0

    def record(self, **values) -> None:
        """One iteration's scalars (0-d tensors or Python numbers); names that are missing record NaN and do not count."""
        row = []
        dev = self._device
        for n in self.names:
            v = values.get(n)
            if isinstance(v, Tensor):
                dev = dev or v.device
        dev = dev or torch.device("cpu")
        for n in self.names:
            v = values.get(n, float("nan"))
            row.append(v.detach().to(torch.float32).reshape(()) if isinstance(v, Tensor)
                       else torch.tensor(float(v), dtype=torch.float32, device=dev))
        unknown = set(values) - set(self.names)
        assert not unknown, f"unknown meter names {sorted(unknown)}"
        if self._buf is None:
            self._buf = torch.empty((self._capacity, len(self.names)), dtype=torch.float32, device=dev)
            self._device = dev
        if self._n == self._capacity:
            self._fold()
        self._buf[self._n].copy_(torch.stack(row), non_blocking=True)
        self._n += 1

    def _fold(self) -> None:
        """Move the recorded rows to the host (one transfer) and free the buffer rows."""
        if self._n:
            rows = self._buf[:self._n].to("cpu", torch.float64)
            self._host_sum += torch.nan_to_num(rows, nan=0.0).sum(0)
            self._host_rows += (~torch.isnan(rows)).sum(0).to(torch.float64)
            self._host_n += self._n
            self._n = 0

    def summary(self) -> Dict[str, Dict[str, float]]:
        """{name: {"mean": value}} over everything recorded since the last reset; ONE device synchronisation."""
        self._fold()
        counts = self._host_rows
        out = {}
        for i, n in enumerate(self.names):
            out[n] = {"mean": float(self._host_sum[i] / counts[i]) if counts[i] > 0 else float("nan")}
        return out

    def tracking_status(self) -> Dict[str, Dict[str, float]]:
        return self.summary()
