"""Synchronisation protocol of the opt-in forward kernel (csrc/attn_fwd2.cuh) under random schedules: no deadlock, no
mbarrier phase aliasing, no buffer overwritten before its consumer is done — and the model does notice when one of the
waits is removed.  See tools/attn_v2_protocol_sim.py for what is modelled."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import attn_v2_protocol_sim as sim  # noqa: E402


@pytest.mark.parametrize("n_warps", [4, 8])                   # attn_fwd2 / attn_fwd3
@pytest.mark.parametrize("n_tiles", [1, 2, 3, 4, 6])          # S = 327 -> 3 key tiles, 707 -> 6
def test_protocol_has_no_deadlock_or_hazard(n_tiles, n_warps):
    for seed in range(300):
        assert sim.run(n_tiles, seed, n_warps=n_warps) > 0


@pytest.mark.parametrize("skip,n_warps", [("sfree", 4), ("o_before_k", 4), ("o_before_p", 4), ("s_before_v", 4),
                                          ("sfree", 8), ("o_before_p", 8), ("pair", 8), ("s_before_k", 8),
                                          ("o_before_v", 8)])
def test_model_detects_a_missing_wait(skip, n_warps):
    caught = 0
    for seed in range(300):
        try:
            sim.run(4, seed, skip=(skip,), n_warps=n_warps)
        except (AssertionError, sim.Deadlock):
            caught += 1
    assert caught > 0, f"removing the '{skip}' wait went unnoticed in 300 schedules"
