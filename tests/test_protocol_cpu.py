"""The mbarrier protocol of the persistent attention kernel (csrc/attention_pipe.cu), executed as a discrete-event model
under random interleavings (tools/sim_attn_pipe.py): no deadlock, no phase aliasing, no buffer hazard — and the model does
notice protocols that are deliberately broken."""
import importlib.util
import os
import random

import pytest

_spec = importlib.util.spec_from_file_location(
    "sim_attn_pipe", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sim_attn_pipe.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


@pytest.mark.parametrize("prefix", [0, 1])
def test_protocol_is_clean_under_random_schedules(prefix):
    for seed in range(40):
        for njobs in (1, 2, 3, 5):
            sim.Sim(njobs, prefix, random.Random(1000 * seed + njobs)).run()


@pytest.mark.parametrize("broken", ["no_q_empty_wait", "no_kv_empty_wait", "no_s_empty_wait", "no_pv0_wait"])
def test_model_catches_broken_protocols(broken):
    caught = 0
    for seed in range(30):
        try:
            sim.Sim(4, 1, random.Random(seed), broken=broken).run()
        except AssertionError:
            caught += 1
    assert caught > 0
