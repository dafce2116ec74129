"""CPU checks on the C-ABI library: it loads, exports every symbol include/b2s.h declares, and its
host-only entry points behave (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2s.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from open_spiel_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libb2s.so does not export " + n
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"


def test_game_info_matches_reference_constants():
    import open_spiel_b200 as b2
    g = b2.load_game("connect_four")
    assert (g.num_distinct_actions(), g.max_game_length(), g.observation_tensor_shape()) == (7, 42, [3, 6, 7])
    g = b2.load_game("tic_tac_toe")
    assert (g.num_distinct_actions(), g.max_game_length(), g.observation_tensor_size()) == (9, 9, 27)
    g = b2.load_game("breakthrough")
    assert (g.num_distinct_actions(), g.max_game_length(), g.observation_tensor_size()) == (768, 209, 192)
    with pytest.raises(b2.SpielError):
        b2.load_game("chess")
    with pytest.raises(b2.SpielError):
        b2.load_game("connect_four(rows=9,columns=9)")      # does not fit 64 bits: rejected, no fallback


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import open_spiel_b200 as b2
    with pytest.raises(b2.SpielError):
        b2.load_game("connect_four").new_batch(16)
    from open_spiel_b200 import _lib
    h = C.c_void_p()
    rc = _lib.lib().b2s_batch_create(1, None, 16, 0, C.byref(h))
    assert rc != 0 and b"no CUDA device" in _lib.lib().b2s_last_error()
