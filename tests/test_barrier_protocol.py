"""The mbarrier protocols of the row-block tensor-core backward kernels under a randomised scheduler
(tools/barrier_sim.py): the configurations that ship must neither deadlock nor let an issuer pass a wait before the
transform it waits for has happened; the two configurations that were found broken on the way must be caught."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import barrier_sim as S  # noqa: E402


@pytest.mark.parametrize("name", [n for n in S.CONFIGS if "(bug)" not in n and "(reverted)" not in n])
def test_shipped_protocols_hold(name):
    assert S.check(name, S.CONFIGS[name], runs=25) is None


@pytest.mark.parametrize("name", [n for n in S.CONFIGS if "(bug)" in n or "(reverted)" in n])
def test_broken_protocols_are_caught(name):
    bad = S.check(name, S.CONFIGS[name], runs=25)
    assert bad is not None and ("Race" in bad or "Deadlock" in bad)


@pytest.mark.parametrize("name", list(S.FORWARD_CONFIGS))
def test_joint_kernel_protocols_hold(name):
    assert S.check(name, S.FORWARD_CONFIGS[name], runs=15) is None
