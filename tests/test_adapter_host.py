"""CPU half of the C++ drop-in boundary test: all seven games behind open_spiel::LoadGame, run through the reference's
own basic_tests harness (RandomSimTest, RandomSimTestWithUndo, CheckChanceOutcomes, RandomSimTestCustomObserver) and in
lock-step against the stock C++ games — see open_spiel_b200/adapter/adapter_host_test.cc.  The binary is built by
__graft_entry__.build() where the reference headers exist and travels to the GPU box."""
import os
import subprocess

import pytest

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "open_spiel_b200", "adapter", "_build", "adapter_host_test")


@pytest.mark.skipif(not os.path.exists(BIN), reason="adapter_host_test not built (needs the reference headers)")
def test_reference_harness_and_lockstep_on_all_seven_dropins():
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=900)
    tail = out.stdout[-2000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert "adapter_host_test ok" in out.stdout, tail
    for g in ["tic_tac_toe", "connect_four", "breakthrough", "hex", "go(board_size=9,komi=7.5)", "kuhn_poker", "leduc_poker", "mnk", "othello", "y(board_size=9)", "havannah(board_size=4)"]:
        assert "ok " + g + "\n" in out.stdout, g
