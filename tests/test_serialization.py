"""CPU: the text formats of open_spiel_b200/serialization.py against the UNMODIFIED reference (oracle/_ref):
information-state strings rebuilt from tensors, hex-float doubles, CFRSolverBase::Serialize / DeserializeCFRSolver in both
directions (the reference loads what we write; we parse what it writes), State::Serialize."""
import ctypes
import random
import struct

import numpy as np
import pytest

from open_spiel_b200 import serialization as ser
from oracle_lib import OracleGame, infostate_tensors
import ref_lib

needs_ref = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


def test_hex_double_is_printf_percent_a():
    libc = ctypes.CDLL(None)
    libc.snprintf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_double]

    def c_a(x):
        b = ctypes.create_string_buffer(64)
        libc.snprintf(b, 64, b"%a", ctypes.c_double(x))
        return b.value.decode()

    rnd = random.Random(7)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 1e-6, 1 / 3, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 0.1]
    vals += [struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0] for _ in range(20000)]
    for v in vals:
        if v == v:
            assert ser.hex_double(v) == c_a(v)
            assert ser.parse_double(ser.hex_double(v)) == v or v in (float("inf"), float("-inf"))


@pytest.mark.parametrize("name", ["kuhn_poker", "leduc_poker"])
def test_information_state_strings_from_tensors(name):
    # every information state of the game: the oracle's string (pinned to the reference playthroughs) vs ours from its tensor
    tensors = infostate_tensors(OracleGame(name))
    assert len(tensors) == {"kuhn_poker": 12, "leduc_poker": 936}[name]
    f = ser.INFORMATION_STATE_STRING[name]
    for key, blob in tensors.items():
        assert f(np.frombuffer(blob, dtype=np.float32)) == key


def _layout_from(ref_table, name):
    """A CFRSolver.table()-shaped dict (flat arrays + offsets + tensor keys) holding the reference's table."""
    tensors = infostate_tensors(OracleGame(name))
    keys = sorted(ref_table)
    offsets, legal = [0], []
    cols = {f: [] for f in ("regrets", "cum_policy", "cur_policy")}
    for k in keys:
        v = ref_table[k]
        legal += v["legal"]
        for f in cols:
            cols[f] += v[f]
        offsets.append(len(legal))
    t = {"offsets": np.array(offsets, dtype=np.int32), "legal_actions": np.array(legal, dtype=np.int32),
         "keys": np.stack([np.frombuffer(tensors[k], dtype=np.float32) for k in keys])}
    t.update({f: np.array(v) for f, v in cols.items()})
    return keys, t


@needs_ref
@pytest.mark.parametrize("name,iters", [("kuhn_poker", 37), ("leduc_poker", 6)])
def test_cfr_solver_text_format_both_directions(name, iters):
    rg = ref_lib.RefGame(name)
    ref = ref_lib.RefCFR(rg)
    ref.iterate(iters)
    text = ref_lib.cfr_serialize(ref)
    # (1) we parse what the reference writes
    parsed = ser.deserialize_cfr_solver(text)
    assert parsed["game"] == ref_lib.game_to_string(rg) and parsed["solver_type"] == "CFRSolver" and parsed["iteration"] == iters
    assert parsed["table"] == ref.table()
    # (2) we write what the reference writes: same header, same set of table entries byte for byte (the reference emits
    # its unordered_map in hash order, so only the order of entries may differ)
    keys, layout = _layout_from(ref.table(), name)
    assert ser.table_keys(name, layout) == keys
    mine = ser.serialize_cfr_solver(ref_lib.game_to_string(rg), "CFRSolver", iters, keys, layout)
    head_r, _, vals_r = text.partition("[SolverValuesTable]\n")
    head_m, _, vals_m = mine.partition("[SolverValuesTable]\n")
    assert head_m == head_r
    pairs = lambda v: sorted(zip(v.split(ser.DELIMITER)[0::2], v.split(ser.DELIMITER)[1::2]))   # noqa: E731
    assert pairs(vals_m) == pairs(vals_r)
    # (3) the reference loads what we write, and continues training from it exactly like the original
    loaded = ref_lib.cfr_deserialize(rg, mine)
    assert loaded.table() == ref.table()
    loaded.iterate(3)
    ref.iterate(3)
    assert loaded.table() == ref.table()
    # (4) and back into flat arrays in a given row order (CFRSolver.load_table's arguments)
    r, c, p = ser.table_arrays_from(parsed["table"], keys, layout)
    assert np.array_equal(r, layout["regrets"]) and np.array_equal(c, layout["cum_policy"]) and np.array_equal(p, layout["cur_policy"])


@needs_ref
@pytest.mark.parametrize("gs", ["connect_four", "go(board_size=5)", "kuhn_poker", "leduc_poker", "breakthrough"])
def test_state_serialize_format(gs):
    rng = random.Random(3)
    rg = ref_lib.RefGame(gs)
    st = rg.new_initial_state()
    hist = []
    for _ in range(12):
        if st.is_terminal():
            break
        a = rng.choice(st.legal_actions())
        st.apply_action(a)
        hist.append(a)
    text = ser.serialize_state(hist)
    assert text == ref_lib.state_serialize(st)
    back = ref_lib.deserialize_state(rg, text)
    assert back.history() == hist and back.to_string() == st.to_string()
