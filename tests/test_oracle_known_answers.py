"""Known-answer cases of the reference's own per-game tests, run against the oracle (CPU) — and, marked gpu,
against the device path through the scalar State adapter."""
import pytest

from oracle_lib import OracleGame


def play(game_string, actions, loader):
    st = loader(game_string)
    for a in actions:
        st.apply_action(a)
    return st


def oracle_loader(gs):
    return OracleGame(gs).new_initial_state()


def device_loader(gs):
    import open_spiel_b200 as b2
    return b2.load_game(gs).new_initial_state()


LOADERS = [pytest.param(oracle_loader, id="oracle"), pytest.param(device_loader, id="device", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("loader", LOADERS)
def test_connect_four_fast_loss(loader):
    # connect_four_test.cc:38-58
    st = play("connect_four", [3, 3, 4, 4, 2, 2], loader)
    assert not st.is_terminal()
    st.apply_action(1)
    assert st.is_terminal() and st.returns() == [1.0, -1.0]
    if loader is oracle_loader:
        assert st.to_string() == ".......\n.......\n.......\n.......\n..ooo..\n.xxxx..\n"


@pytest.mark.parametrize("loader", LOADERS)
def test_connect_four_full_board_draw(loader):
    # connect_four_test.cc:67-85: the full board ooxxxoo/xxoooxx/... is a draw; reach it column by column
    rows_top_down = ["ooxxxoo", "xxoooxx", "ooxxxoo", "xxoooxx", "ooxxxoo", "xxoooxx"]
    cols = [[rows_top_down[5 - r][c] for r in range(6)] for c in range(7)]     # bottom-up per column
    # x moves first; find an order of drops that alternates x / o and never completes a line early
    import itertools
    heights = [0] * 7
    st = loader("connect_four")
    order = []
    for ply in range(42):
        want = "x" if ply % 2 == 0 else "o"
        for c in itertools.chain(range(7)):
            if heights[c] < 6 and cols[c][heights[c]] == want:
                trial = play("connect_four", order + [c], loader)
                if ply == 41 or not trial.is_terminal():
                    order.append(c)
                    heights[c] += 1
                    break
        else:
            pytest.skip("no alternating drop order found for the draw position")
    st = play("connect_four", order, loader)
    assert st.is_terminal() and st.returns() == [0.0, 0.0] and st.legal_actions() == []


@pytest.mark.parametrize("loader", LOADERS)
def test_connect_four_arbitrary_sizes(loader):
    # connect_four_test.cc:319-395: 4x5 board with x_in_row=3, MaxGameLength
    st = play("connect_four(rows=4,columns=5,x_in_row=3)", [0, 1, 0, 1], loader)
    assert not st.is_terminal()
    st.apply_action(0)                       # x has three in column 0
    assert st.is_terminal() and st.returns() == [1.0, -1.0]
    st = play("connect_four(rows=7,columns=8,x_in_row=5)", [0, 7, 1, 7, 2, 7, 3, 7], loader)
    assert not st.is_terminal()
    st.apply_action(4)                       # x: five in the bottom row
    assert st.is_terminal() and st.returns() == [1.0, -1.0]


@pytest.mark.parametrize("loader", LOADERS)
def test_hex_board_orientation_and_swap(loader):
    # hex_test.cc:31-47: 3 columns x 4 rows, black connects north-south
    st = play("hex(num_cols=3,num_rows=4)", [1, 2, 4, 5, 7, 8, 10], loader)
    assert st.is_terminal() and st.returns() == [1.0, -1.0]
    # hex_test.cc:49-68: swap rule
    st = play("hex(board_size=3,swap=True)", [1, 9], loader)
    la = st.legal_actions()
    assert 1 in la and 3 not in la and st.current_player() == 0


def test_go_13x13_all_actions_legal_at_start():
    # go_test.cc:54-67 (oracle only: the device path is limited to 9x9)
    g = OracleGame("go(board_size=13)")
    st = g.new_initial_state()
    assert g.num_distinct_actions == 170 and len(st.legal_actions()) == 170


def test_leduc_three_players_starting_player():
    # leduc_poker_test.cc:66-93 (oracle only: the device path is 2-player)
    st = play("leduc_poker(players=3,starting_player=1)", [0, 2, 4], oracle_loader)
    assert st.current_player() == 1
    st.apply_action(0)
    assert st.current_player() == 2
    st.apply_action(2)
    assert st.current_player() == 0
    st.apply_action(1)
    assert st.is_chance_node()
    st.apply_action(3)
    assert st.current_player() == 2


def test_kuhn_has_54_non_chance_states():
    # kuhn_poker_test.cc:43-50
    seen = set()

    def walk(st):
        if not st.is_chance_node():
            seen.add(tuple(st.history()))
        if st.is_terminal():
            return
        for a in st.legal_actions():
            c = st.clone()
            c.apply_action(a)
            walk(c)

    walk(OracleGame("kuhn_poker").new_initial_state())
    assert len(seen) == 54


def test_tic_tac_toe_reachable_state_count():
    # tic_tac_toe.h:48 kNumberStates = 5478
    seen = set()

    def walk(st):
        key = st.to_string()
        if key in seen:
            return
        seen.add(key)
        if st.is_terminal():
            return
        for a in st.legal_actions():
            c = st.clone()
            c.apply_action(a)
            walk(c)

    walk(OracleGame("tic_tac_toe").new_initial_state())
    assert len(seen) == 5478
