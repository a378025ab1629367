"""Synchronisation protocol of the opt-in forward kernel (csrc/attn_fwd2.cuh) under random schedules: no deadlock, no
mbarrier phase aliasing, no buffer overwritten before its consumer is done — and the model does notice when one of the
waits is removed.  See tools/attn_v2_protocol_sim.py for what is modelled."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import attn_v2_protocol_sim as sim  # noqa: E402


@pytest.mark.parametrize("n_tiles", [1, 2, 3, 4, 6])          # S = 327 -> 3 key tiles, 707 -> 6
def test_protocol_has_no_deadlock_or_hazard(n_tiles):
    for seed in range(400):
        assert sim.run(n_tiles, seed) > 0


@pytest.mark.parametrize("skip", ["sfree", "o_before_k", "o_before_p", "s_before_v"])
def test_model_detects_a_missing_wait(skip):
    caught = 0
    for seed in range(300):
        try:
            sim.run(4, seed, skip=(skip,))
        except (AssertionError, sim.Deadlock):
            caught += 1
    assert caught > 0, f"removing the '{skip}' wait went unnoticed in 300 schedules"
