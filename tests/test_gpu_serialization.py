"""GPU: training state crosses the boundary in the reference's own text format, in both directions, without losing a bit:
device CFR tables -> CFRSolverBase::Serialize text -> the UNMODIFIED reference's DeserializeCFRSolver -> both continue
training -> tables still identical; and reference text -> device solver.  Plus State::Serialize round trips."""
import numpy as np
import pytest

import open_spiel_b200 as b2
import ref_lib
from oracle_lib import OracleGame, infostate_tensors

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not shipped")]


def same(dev_table, ref_table, tensors):
    by_key = {dev_table["keys"][k].tobytes(): k for k in range(len(dev_table["players"]))}
    assert len(by_key) == len(ref_table)
    for key, v in ref_table.items():
        k = by_key[tensors[key]]
        lo, hi = dev_table["offsets"][k], dev_table["offsets"][k + 1]
        for f in ("regrets", "cum_policy", "cur_policy"):
            assert np.array_equal(dev_table[f][lo:hi], np.array(v[f])), (key, f)


@pytest.mark.parametrize("name,iters", [("kuhn_poker", 25), ("leduc_poker", 7)])
def test_device_tables_to_reference_and_back(name, iters):
    game, rg = b2.load_game(name), ref_lib.RefGame(name)
    tensors = infostate_tensors(OracleGame(name))
    dev = b2.CFRSolver(game)
    dev.evaluate_and_update_policy(iters)
    text = dev.serialize()
    ref = ref_lib.cfr_deserialize(rg, text)                 # stock DeserializeCFRSolver
    same(dev.table(), ref.table(), tensors)
    ref.iterate(4)
    dev.evaluate_and_update_policy(4)
    same(dev.table(), ref.table(), tensors)                 # the reference continued from our checkpoint exactly
    # and the other way: a fresh device solver resumes from the reference's own text
    dev2 = b2.CFRSolver(game)
    parsed = dev2.load_serialized(ref_lib.cfr_serialize(ref))
    assert parsed["iteration"] == iters + 4 and dev2.info().iteration == iters + 4
    ref.iterate(3)
    dev2.evaluate_and_update_policy(3)
    same(dev2.table(), ref.table(), tensors)


def test_state_serialize_round_trip_through_the_reference():
    rng = np.random.RandomState(5)
    for gs in ("connect_four", "go(board_size=5)", "leduc_poker", "othello", "havannah(board_size=4,swap=True)", "y(board_size=5)", "mnk(m=4,n=4,k=3)"):
        game, rg = b2.load_game(gs), ref_lib.RefGame(gs)
        st = game.new_initial_state()
        for _ in range(9):
            if st.is_terminal():
                break
            la = st.legal_actions()
            st.apply_action(int(la[rng.randint(len(la))]))
        text = st.serialize()
        rs = ref_lib.deserialize_state(rg, text)            # the reference loads our state
        assert rs.history() == st.history() and rs.legal_actions() == st.legal_actions()
        back = game.deserialize_state(ref_lib.state_serialize(rs))   # and we load the reference's
        assert back.history() == st.history() and back.legal_actions() == st.legal_actions()
        assert np.array_equal(np.asarray(back.observation_tensor(0)), np.asarray(st.observation_tensor(0)))


def test_information_state_strings_and_tabular_policy_match_the_reference():
    rng = np.random.RandomState(11)
    for name in ("kuhn_poker", "leduc_poker"):
        game, rg = b2.load_game(name), ref_lib.RefGame(name)
        for _ in range(6):
            st, rs = game.new_initial_state(), rg.new_initial_state()
            while not rs.is_terminal():
                if rs.current_player() >= 0:
                    p = rs.current_player()
                    assert st.information_state_string(p) == rs.information_state_string(p)
                else:
                    assert [a for a, _ in st.chance_outcomes()] == [a for a, _ in rs.chance_outcomes()]
                    assert [pr for _, pr in st.chance_outcomes()] == [pr for _, pr in rs.chance_outcomes()]
                la = rs.legal_actions()
                a = int(la[rng.randint(len(la))])
                st.apply_action(a)
                rs.apply_action(a)
        dev = b2.CFRSolver(game)
        ref = ref_lib.RefCFR(rg)
        dev.evaluate_and_update_policy(9)
        ref.iterate(9)
        pol = dev.tabular_average_policy()
        table = ref.table()
        assert set(pol) == set(table)
        for key, v in table.items():
            total = 0.0
            for c in v["cum_policy"]:          # sequential sum as CFRAveragePolicy does (Python 3.12's sum() is compensated)
                total += c
            want = [c / total if total > 0 else 1.0 / len(v["legal"]) for c in v["cum_policy"]]
            assert [a for a, _ in pol[key]] == v["legal"] and [p for _, p in pol[key]] == want
