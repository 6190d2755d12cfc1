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


@pytest.mark.parametrize("name", [n for n in S.JOINT10_CONFIGS if "(bug)" not in n])
def test_config2_joint_protocol_holds(name):
    """csrc/local_fwd_tcj10.cu: x ring released by the issuers' x_free commits, y ring by y_done, accumulator sets by the drain."""
    assert S.check(name, S.JOINT10_CONFIGS[name], runs=15) is None


def test_config2_joint_ring_release_through_y_done_is_caught():
    """The first ring-4 version of the joint (it hung on the GPU): a waiter two phases behind on a parity barrier."""
    name = "tcj10 ring 4 released through y_done (bug)"
    bad = S.check(name, S.JOINT10_CONFIGS[name], runs=25)
    assert bad is not None and ("Race" in bad or "Deadlock" in bad)
