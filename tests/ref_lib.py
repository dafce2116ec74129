"""ctypes binding of oracle/_ref/libspiel_ref_c.so — the UNMODIFIED reference compiled against the abseil
shim (oracle/ref_build.mk).  Same Python surface as oracle_lib.OracleGame/OracleState.  Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libspiel_ref_c.so")
_LIB = None


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        vp, i64p, dp, fp, cp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_char_p
        L.ref_last_error.restype = cp
        L.ref_load_game.restype = vp
        L.ref_load_game.argtypes = [cp]
        L.ref_free_game.argtypes = [vp]
        for f in ("ref_num_distinct_actions", "ref_num_players", "ref_max_game_length", "ref_max_chance_outcomes",
                  "ref_observation_tensor_size", "ref_information_state_tensor_size"):
            getattr(L, f).argtypes = [vp]
        for f in ("ref_min_utility", "ref_max_utility"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [vp]
        L.ref_new_initial_state.restype = vp
        L.ref_new_initial_state.argtypes = [vp]
        L.ref_clone.restype = vp
        L.ref_clone.argtypes = [vp]
        L.ref_free_state.argtypes = [vp]
        L.ref_current_player.argtypes = [vp]
        L.ref_is_terminal.argtypes = [vp]
        L.ref_legal_actions.argtypes = [vp, i64p, C.c_int]
        L.ref_apply_action.argtypes = [vp, C.c_int64]
        L.ref_returns.argtypes = [vp, dp]
        L.ref_observation_tensor.argtypes = [vp, C.c_int, fp, C.c_int]
        L.ref_information_state_tensor.argtypes = [vp, C.c_int, fp, C.c_int]
        L.ref_to_string.argtypes = [vp, cp, C.c_int]
        L.ref_information_state_string.argtypes = [vp, C.c_int, cp, C.c_int]
        L.ref_observation_string.argtypes = [vp, C.c_int, cp, C.c_int]
        L.ref_chance_outcomes.argtypes = [vp, i64p, dp, C.c_int]
        L.ref_history.argtypes = [vp, i64p, C.c_int]
        L.ref_cfr_new.restype = vp
        L.ref_cfr_new.argtypes = [vp]
        L.ref_cfr_free.argtypes = [vp]
        L.ref_cfr_iterate.argtypes = [vp, C.c_int]
        L.ref_cfr_num_infostates.argtypes = [vp]
        L.ref_cfr_get.argtypes = [vp, cp, i64p, dp, dp, dp, C.c_int]
        L.ref_cfr_keys.argtypes = [vp, cp, C.c_int]
        L.ref_cfr_exploitability.restype = C.c_double
        L.ref_cfr_exploitability.argtypes = [vp, vp]
        L.ref_cfr_nash_conv.restype = C.c_double
        L.ref_cfr_nash_conv.argtypes = [vp, vp]
        L.ref_replay_batch.restype = C.c_long
        L.ref_replay_batch.argtypes = [vp, C.c_long, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
        L.ref_mcts_search.argtypes = [vp, vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, i64p,
                                      C.POINTER(C.c_int), dp, C.c_int, i64p, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


class RefGame:
    def __init__(self, game_string):
        L = lib()
        self._g = L.ref_load_game(game_string.encode())
        if not self._g:
            raise ValueError("reference: " + L.ref_last_error().decode())
        self.name = game_string.split("(")[0]
        self.num_distinct_actions = L.ref_num_distinct_actions(self._g)
        self.num_players = L.ref_num_players(self._g)
        self.max_game_length = L.ref_max_game_length(self._g)
        self.max_chance_outcomes = L.ref_max_chance_outcomes(self._g)
        self.observation_tensor_size = L.ref_observation_tensor_size(self._g)
        self.information_state_tensor_size = L.ref_information_state_tensor_size(self._g)

    def new_initial_state(self):
        return RefState(self, lib().ref_new_initial_state(self._g))


class RefState:
    def __init__(self, game, ptr):
        self.game, self._s = game, ptr

    def __del__(self):
        try:
            lib().ref_free_state(self._s)
        except Exception:
            pass

    def clone(self):
        return RefState(self.game, lib().ref_clone(self._s))

    def current_player(self):
        return lib().ref_current_player(self._s)

    def is_terminal(self):
        return bool(lib().ref_is_terminal(self._s))

    def is_chance_node(self):
        return self.current_player() == -1

    def legal_actions(self):
        cap = max(self.game.num_distinct_actions, self.game.max_chance_outcomes, 1) + 8
        buf = (C.c_int64 * cap)()
        n = lib().ref_legal_actions(self._s, buf, cap)
        return list(buf[:n])

    def apply_action(self, a):
        if lib().ref_apply_action(self._s, int(a)):
            raise RuntimeError("reference: " + lib().ref_last_error().decode())

    def returns(self):
        buf = (C.c_double * self.game.num_players)()
        lib().ref_returns(self._s, buf)
        return list(buf)

    def observation_tensor(self, player=0):
        out = np.zeros(self.game.observation_tensor_size, dtype=np.float32)
        lib().ref_observation_tensor(self._s, player, out.ctypes.data_as(C.POINTER(C.c_float)), out.size)
        return out

    def information_state_tensor(self, player=0):
        out = np.zeros(self.game.information_state_tensor_size, dtype=np.float32)
        lib().ref_information_state_tensor(self._s, player, out.ctypes.data_as(C.POINTER(C.c_float)), out.size)
        return out

    def _str(self, fn, *args):
        buf = C.create_string_buffer(8192)
        fn(self._s, *args, buf, 8192)
        return buf.value.decode()

    def to_string(self):
        return self._str(lib().ref_to_string)

    def information_state_string(self, player=0):
        return self._str(lib().ref_information_state_string, player)

    def observation_string(self, player=0):
        return self._str(lib().ref_observation_string, player)

    def chance_outcomes(self):
        cap = self.game.max_chance_outcomes + 8
        a = (C.c_int64 * cap)()
        p = (C.c_double * cap)()
        n = lib().ref_chance_outcomes(self._s, a, p, cap)
        return [(a[i], p[i]) for i in range(n)]

    def history(self):
        buf = (C.c_int64 * 2048)()
        n = lib().ref_history(self._s, buf, 2048)
        return list(buf[:n])


class RefCFR:
    """The unmodified reference's algorithms::CFRSolver (through oracle/ref_glue/ref_c_api.cc)."""

    def __init__(self, game):
        self.game = game
        self._c = lib().ref_cfr_new(game._g)

    def __del__(self):
        try:
            lib().ref_cfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        assert lib().ref_cfr_iterate(self._c, iters) == 0

    def table(self):
        L = lib()
        buf = C.create_string_buffer(1 << 20)
        L.ref_cfr_keys(self._c, buf, 1 << 20)
        out = {}
        for key in buf.value.decode().split("\n"):
            if key == "" and not out:
                pass
            legal = (C.c_int64 * 16)()
            r, cu, cp = (C.c_double * 16)(), (C.c_double * 16)(), (C.c_double * 16)()
            n = L.ref_cfr_get(self._c, key.encode(), legal, r, cu, cp, 16)
            if n < 0:
                continue
            out[key] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n]),
                        "cur_policy": list(cp[:n])}
        return out

    def exploitability(self):
        return lib().ref_cfr_exploitability(self.game._g, self._c)

    def nash_conv(self):
        return lib().ref_cfr_nash_conv(self.game._g, self._c)


def ref_mcts(game, state, uct_c, max_simulations, n_rollouts=1, solve=True, seed=0, max_memory_mb=1000):
    """The unmodified reference's MCTSBot::MCTSearch (RandomRolloutEvaluator); returns root children stats."""
    L = lib()
    L.ref_mcts_search_mb.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_int]
    cap = game.num_distinct_actions + 4
    acts, vis, rew = (C.c_int64 * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
    best, rv = C.c_int64(), C.c_int()
    n = L.ref_mcts_search_mb(game._g, state._s, uct_c, max_simulations, n_rollouts, int(solve), seed, acts, vis, rew, cap,
                             C.byref(best), C.byref(rv), int(max_memory_mb))
    assert n >= 0, L.ref_last_error()
    return {"children": [(acts[i], vis[i], rew[i]) for i in range(n)], "best_action": best.value, "root_visits": rv.value}


def sizeof_search_node():
    return lib().ref_sizeof_search_node()


def ref_record_batched_trajectory(game, batch_size, seed, T):
    """The unmodified reference's RecordBatchedTrajectory (algorithms/trajectories.cc:98-118) with GetUniformPolicy for
    every player and include_full_observations; dict of [B, T, ...] numpy arrays."""
    import numpy as np
    L = lib()
    A, P, F = game.num_distinct_actions, game.num_players, game.information_state_tensor_size
    obs = np.zeros((batch_size, T, F), dtype=np.float32)
    legal = np.zeros((batch_size, T, A), dtype=np.int32)
    pol = np.zeros((batch_size, T, A), dtype=np.float64)
    actions = np.zeros((batch_size, T), dtype=np.int64)
    players, valid, nit = (np.zeros((batch_size, T), dtype=np.int32) for _ in range(3))
    rewards = np.zeros((batch_size, P), dtype=np.float64)
    L.ref_record_batched_trajectory.restype = C.c_int
    L.ref_record_batched_trajectory.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8
    n = L.ref_record_batched_trajectory(game._g, batch_size, seed, T, obs.ctypes.data, legal.ctypes.data, pol.ctypes.data,
                                        actions.ctypes.data, players.ctypes.data, valid.ctypes.data, nit.ctypes.data,
                                        rewards.ctypes.data)
    assert n == T, L.ref_last_error()
    return {"observations": obs, "legal_actions": legal, "player_policies": pol, "actions": actions, "player_ids": players,
            "valid": valid, "next_is_terminal": nit, "rewards": rewards}


class RefMCCFR:
    """The unmodified reference's algorithms::ExternalSamplingMCCFRSolver(game, seed) (AverageType::kSimple)."""

    def __init__(self, game, seed=0, full_average=False):
        L = lib()
        L.ref_mccfr_new.restype = C.c_void_p
        L.ref_mccfr_new.argtypes = [C.c_void_p, C.c_int]
        L.ref_mccfr_new_full.restype = C.c_void_p
        L.ref_mccfr_new_full.argtypes = [C.c_void_p, C.c_int]
        L.ref_mccfr_free.argtypes = [C.c_void_p]
        L.ref_mccfr_iterate.argtypes = [C.c_void_p, C.c_int]
        L.ref_mccfr_keys.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_mccfr_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_int]
        L.ref_mccfr_nash_conv.restype = C.c_double
        L.ref_mccfr_nash_conv.argtypes = [C.c_void_p, C.c_void_p]
        self.game = game
        self._c = (L.ref_mccfr_new_full if full_average else L.ref_mccfr_new)(game._g, seed)

    def __del__(self):
        try:
            lib().ref_mccfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        assert lib().ref_mccfr_iterate(self._c, iters) == 0, lib().ref_last_error()

    def table(self):
        L = lib()
        buf = C.create_string_buffer(1 << 20)
        L.ref_mccfr_keys(self._c, buf, 1 << 20)
        out = {}
        for key in buf.value.decode().split("\n"):
            legal = (C.c_int64 * 16)()
            r, cu = (C.c_double * 16)(), (C.c_double * 16)()
            n = L.ref_mccfr_get(self._c, key.encode(), legal, r, cu, 16)
            if n < 0:
                continue
            out[key] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n])}
        return out

    def nash_conv(self):
        return lib().ref_mccfr_nash_conv(self.game._g, self._c)


def _text(fn, *args, cap=1 << 22):
    buf = C.create_string_buffer(cap)
    n = fn(*args, buf, cap)
    assert 0 <= n < cap, lib().ref_last_error()
    return buf.value.decode()


def game_to_string(game):
    L = lib()
    L.ref_game_to_string.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    return _text(L.ref_game_to_string, game._g)


def cfr_serialize(ref_cfr):
    L = lib()
    L.ref_cfr_serialize.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    return _text(L.ref_cfr_serialize, ref_cfr._c)


def cfr_deserialize(game, text):
    """DeserializeCFRSolver(text) of the unmodified reference, wrapped as a RefCFR."""
    L = lib()
    L.ref_cfr_deserialize.restype = C.c_void_p
    L.ref_cfr_deserialize.argtypes = [C.c_char_p]
    ptr = L.ref_cfr_deserialize(text.encode())
    assert ptr, L.ref_last_error()
    r = RefCFR.__new__(RefCFR)
    r.game, r._c = game, ptr
    return r


def state_serialize(state):
    L = lib()
    L.ref_state_serialize.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    return _text(L.ref_state_serialize, state._s, cap=1 << 16)


def deserialize_state(game, text):
    L = lib()
    L.ref_deserialize_state.restype = C.c_void_p
    L.ref_deserialize_state.argtypes = [C.c_void_p, C.c_char_p]
    ptr = L.ref_deserialize_state(game._g, text.encode())
    assert ptr, L.ref_last_error()
    return RefState(game, ptr)


def replay_batch(game_string, hist, final_actions, mask_words, threads=0):
    """Replays n recorded lanes on the unmodified reference (ref_replay_batch): hist int32 [n, L] (-1 = no move) then
    final_actions int32 [n].  Returns dict of numpy arrays: mask_before / mask_after uint32 [n, W], terminal uint8 [n],
    cur_player int8 [n], returns float32 [n, P], obs_bits uint32 [n, ceil(obs/32)] and the number of failed lanes."""
    L = lib()
    g = RefGame(game_string)
    hist = np.ascontiguousarray(hist, dtype=np.int32)
    fin = np.ascontiguousarray(final_actions, dtype=np.int32)
    n, plies = hist.shape
    P = L.ref_num_players(g._g)
    OW = (L.ref_observation_tensor_size(g._g) + 31) // 32
    out = {"mask_before": np.zeros((n, mask_words), np.uint32), "terminal": np.zeros(n, np.uint8), "cur_player": np.zeros(n, np.int8),
           "returns": np.zeros((n, P), np.float32), "mask_after": np.zeros((n, mask_words), np.uint32),
           "obs_bits": np.zeros((n, OW), np.uint32)}
    threads = threads or min(len(os.sched_getaffinity(0)), 64)
    bad = L.ref_replay_batch(g._g, n, plies, hist.ctypes.data, fin.ctypes.data, mask_words, OW, threads, out["mask_before"].ctypes.data,
                             out["terminal"].ctypes.data, out["cur_player"].ctypes.data, out["returns"].ctypes.data,
                             out["mask_after"].ctypes.data, out["obs_bits"].ctypes.data)
    out["failed_lanes"] = int(bad)
    return out


class RefOSMCCFR:
    """The unmodified reference's algorithms::OutcomeSamplingMCCFRSolver(game, epsilon, seed)."""

    def __init__(self, game, seed=0, epsilon=0.6):
        L = lib()
        L.ref_osmccfr_new.restype = C.c_void_p
        L.ref_osmccfr_new.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_osmccfr_free.argtypes = [C.c_void_p]
        L.ref_osmccfr_iterate.argtypes = [C.c_void_p, C.c_int]
        L.ref_osmccfr_keys.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_osmccfr_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.ref_osmccfr_nash_conv.restype = C.c_double
        L.ref_osmccfr_nash_conv.argtypes = [C.c_void_p, C.c_void_p]
        self.game = game
        self._c = L.ref_osmccfr_new(game._g, epsilon, seed)

    def __del__(self):
        try:
            lib().ref_osmccfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        assert lib().ref_osmccfr_iterate(self._c, iters) == 0, lib().ref_last_error()

    def table(self):
        L = lib()
        buf = C.create_string_buffer(1 << 20)
        L.ref_osmccfr_keys(self._c, buf, 1 << 20)
        out = {}
        for key in buf.value.decode().split("\n"):
            legal = (C.c_int64 * 16)()
            r, cu = (C.c_double * 16)(), (C.c_double * 16)()
            n = L.ref_osmccfr_get(self._c, key.encode(), legal, r, cu, 16)
            if n < 0:
                continue
            out[key] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n])}
        return out

    def nash_conv(self):
        return lib().ref_osmccfr_nash_conv(self.game._g, self._c)
