"""Pin the CPU oracle against the reference's playthrough traces (tests/golden/playthroughs/*.json,
extracted by tests/golden/make_golden.py from open_spiel/integration_tests/playthroughs/*.txt)."""
import glob
import json
import math
import os

import numpy as np
import pytest

from oracle_lib import OracleGame

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "playthroughs", "*.json")))


def check_state(st, g, game):
    """Compare one oracle state with one golden record."""
    if not g["detailed"]:       # non-detailed states carry only the applied action
        return
    assert [l.rstrip() for l in st.to_string().strip("\n").split("\n")] == \
        [l.rstrip() for l in g["to_string"].strip("\n").split("\n")]
    assert st.is_terminal() == g["is_terminal"]
    assert st.current_player() == g["current_player"]
    assert st.history() == g["history"]
    if "legal_actions" in g:
        assert st.legal_actions() == g["legal_actions"]
    elif g["is_terminal"]:
        assert st.legal_actions() == []
    rets = st.returns()
    if "returns" in g:          # chance nodes print no Returns()
        assert rets == g["returns"]
        for r, txt in zip(rets, g["returns_text"]):   # sign of zero (hex prints "-0")
            assert (math.copysign(1.0, r) < 0) == txt.startswith("-")
    if "chance_outcomes" in g:
        co = st.chance_outcomes()
        assert [a for a, _ in co] == [a for a, _ in g["chance_outcomes"]]
        for (_, p), (_, q) in zip(co, g["chance_outcomes"]):
            assert abs(p - q) < 1e-6
    for name, vals in g["tensors"].items():
        player = int(name[name.index("(") + 1:name.index(")")])
        if name.startswith("ObservationTensor"):
            t = st.observation_tensor(player)
        else:
            t = st.information_state_tensor(player)
        np.testing.assert_array_equal(t, np.array(vals, dtype=np.float32), err_msg=name)
    for name, val in g["strings"].items():
        if name.startswith("InformationStateString("):
            p = int(name[name.index("(") + 1:name.index(")")])
            assert st.information_state_string(p) == val, name
        elif name.startswith("ObservationString("):
            p = int(name[name.index("(") + 1:name.index(")")])
            assert st.observation_string(p) == val, name


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_oracle_replays_reference_playthrough(path):
    gold = json.load(open(path, encoding="utf-8"))
    try:
        game = OracleGame(gold["game"])
    except ValueError:
        pytest.skip("oracle does not implement " + gold["game"])
    hdr = gold["header"]
    assert game.num_distinct_actions == int(hdr["NumDistinctActions"])
    assert game.max_game_length == int(hdr["MaxGameLength"])
    assert game.num_players == int(hdr["NumPlayers"])
    if "ObservationTensorSize" in hdr:
        assert game.observation_tensor_size == int(hdr["ObservationTensorSize"])
    if "InformationStateTensorSize" in hdr:
        assert game.information_state_tensor_size == int(hdr["InformationStateTensorSize"])
    st = game.new_initial_state()
    states = gold["states"]
    for k, g in enumerate(states):
        check_state(st, g, game)
        if k < len(gold["actions"]):
            a = gold["actions"][k]
            assert a in st.legal_actions()
            st.apply_action(a)
    assert st.is_terminal()
