"""GPU parity: b2s_record_trajectories (batched RecordBatchedTrajectory, algorithms/trajectories.cc:98-200, uniform
policies) vs oracle/algorithms/trajectories.cc on the same Philox stream: every field of every episode, padding
included, must be identical.  The oracle recorder itself is pinned to the unmodified reference's
RecordBatchedTrajectory by tests/test_trajectories_oracle.py."""
import numpy as np
import pytest
import torch

import open_spiel_b200 as b2
from oracle_lib import OracleGame, oracle_record_trajectory

pytestmark = pytest.mark.gpu

CASES = [
    ("tic_tac_toe", 300, 0),
    ("connect_four", 300, 0),
    ("connect_four(rows=4,columns=5,x_in_row=3)", 200, 0),
    ("breakthrough(rows=6,columns=6)", 64, 0),
    ("breakthrough", 40, 0),
    ("hex(board_size=5)", 128, 0),
    ("hex", 40, 0),
    ("go(board_size=5)", 64, 0),
    ("go(board_size=9)", 24, 0),
    ("kuhn_poker", 400, 0),
    ("leduc_poker", 400, 0),
    ("kuhn_poker(players=3)", 200, 0),
    ("leduc_poker(players=3)", 200, 0),
    ("othello", 64, 0),
    ("mnk(m=5,n=5,k=4)", 96, 0),
    ("y(board_size=7)", 96, 0),
    ("havannah(board_size=4,swap=True)", 96, 0),
    ("connect_four", 100, 20),         # explicit max_unroll_length shorter than max_game_length is an error if exceeded
]


@pytest.mark.parametrize("gs,n,prefix", CASES[:-1], ids=[c[0] for c in CASES[:-1]])
def test_device_recorder_equals_oracle_recorder(gs, n, prefix):
    game, og = b2.load_game(gs), OracleGame(gs)
    batch = game.new_batch(n)
    seed, off = 0xABCDEF, 1000
    T = game.max_game_length()
    tr = batch.record_trajectories(seed, lane_offset=off)
    assert batch.error_count()[0] == 0
    A = game.num_distinct_actions()
    legal = tr.legal_actions().cpu().numpy()
    obs = tr.observations.cpu().numpy()
    actions, players = tr.actions.cpu().numpy(), tr.player_ids.cpu().numpy()
    valid, nit = tr.valid.cpu().numpy(), tr.next_is_terminal.cpu().numpy()
    rewards, lengths = tr.rewards.cpu().numpy(), tr.lengths.cpu().numpy()
    pol = tr.player_policies().cpu().numpy()
    assert legal.shape == (n, T, A) and obs.shape[:2] == (n, T)
    init = og.new_initial_state()
    for i in range(n):
        o = oracle_record_trajectory(init, seed, off + i, T)
        assert lengths[i] == o["length"], (gs, i)
        assert np.array_equal(actions[i], o["actions"]), (gs, i)
        assert np.array_equal(players[i], o["player_ids"]), (gs, i)
        assert np.array_equal(valid[i], o["valid"]) and np.array_equal(nit[i], o["next_is_terminal"]), (gs, i)
        assert np.array_equal(legal[i], o["legal_actions"]), (gs, i)
        assert np.array_equal(obs[i], o["observations"]), (gs, i)
        assert np.array_equal(rewards[i].astype(np.float64), o["rewards"]), (gs, i)
        la = o["legal_actions"].astype(np.float64)
        assert np.array_equal(pol[i], np.where(o["valid"][:, None] == 1, la / la.sum(-1, keepdims=True), 1.0))
    # the batch is left at the terminal states
    assert bool(batch.status()[1].all())


def test_recorder_from_mid_game_states_and_short_unroll():
    game, og = b2.load_game("connect_four"), OracleGame("connect_four")
    n = 128
    batch = game.new_batch(n)
    rng = np.random.RandomState(3)
    states = [og.new_initial_state() for _ in range(n)]
    for _ in range(10):                                   # 10 random plies first (nobody can have won after <7; some may)
        acts = np.full(n, -1, dtype=np.int32)
        for i, st in enumerate(states):
            if not st.is_terminal():
                la = st.legal_actions()
                acts[i] = la[rng.randint(len(la))]
                st.apply_action(int(acts[i]))
        batch.apply_actions(torch.from_numpy(acts).to(batch._dev))
    T = 32                                                # 42 - 10: always enough
    tr = batch.record_trajectories(77, max_unroll_length=T)
    assert batch.error_count()[0] == 0
    actions, lengths = tr.actions.cpu().numpy(), tr.lengths.cpu().numpy()
    for i, st in enumerate(states):
        o = oracle_record_trajectory(st, 77, i, T)
        assert lengths[i] == o["length"] and np.array_equal(actions[i], o["actions"])
    # an unroll length that cannot hold the longest episode is reported (the reference CHECK-fails, trajectories.cc:64-68)
    batch2 = game.new_batch(n)
    batch2.record_trajectories(77, max_unroll_length=8)
    assert batch2.error_count()[0] > 0


def test_time_major_buffers_are_the_transposed_views():
    game = b2.load_game("tic_tac_toe")
    batch = game.new_batch(64)
    tr = batch.record_trajectories(1)
    assert tr.time_major["actions"].shape == (9, 64) and tr.actions.shape == (64, 9)
    assert tr.actions.data_ptr() == tr.time_major["actions"].data_ptr()
    assert int(tr.valid.sum()) == int(tr.lengths.sum())
