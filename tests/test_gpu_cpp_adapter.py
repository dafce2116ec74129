"""GPU: the C++ open_spiel::Game/State adapter (open_spiel_b200/adapter) — LoadGame("connect_four") returns the B200
implementation, which then passes the reference's own tests/basic_tests.cc RandomSimTest harness and a lock-step
comparison against the stock C++ game.  The binary is built where the reference headers exist
(make -C open_spiel_b200/adapter) and shipped; skipped when absent."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "open_spiel_b200", "adapter", "_build", "adapter_test")


@pytest.mark.skipif(not os.path.exists(BIN), reason="adapter_test not built (needs the reference headers)")
def test_cpp_adapter_passes_reference_harness():
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "adapter_test ok" in out.stdout
