"""Cross-GPU flag protocol of the reduce-scatter / all-reduce kernels (csrc/comm.cu sync_begin / sync_end) replayed on
the CPU under randomised schedules (tools/flag_protocol_model.py): liveness, every read sees the version written for
its call, no gradient buffer is overwritten while a peer still reads it; injected faults must be caught."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import flag_protocol_model as fm  # noqa: E402


@pytest.mark.parametrize("world,calls,ctas", [(2, 8, 4), (4, 6, 3), (8, 4, 2), (3, 5, 1)])
def test_protocol_holds(world, calls, ctas):
    for seed in range(25):
        fm.simulate(world, calls, ctas, seed)


@pytest.mark.parametrize("bug", ["no_ready_wait", "no_done_wait"])
def test_injected_faults_are_detected(bug):
    caught = 0
    for seed in range(40):
        try:
            fm.simulate(4, 6, 3, seed, bug=bug)
        except fm.FlagProtocolError:
            caught += 1
    assert caught > 0, f"fault {bug} was never detected"


def test_early_sequence_store_is_harmless():
    for seed in range(25):
        fm.simulate(4, 6, 3, seed, bug="early_seq_store")
