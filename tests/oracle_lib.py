"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    srcs = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "oracle"))
            if "_ref" not in dp and "absl_shim" not in dp for f in fs if f.endswith((".cc", ".h"))]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    L = C.CDLL(so)
    vp, i64p, dp, fp, cp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_char_p
    L.orc_load_game.restype = vp
    L.orc_load_game.argtypes = [cp, C.c_int, C.POINTER(cp), dp]
    for f in ("orc_num_distinct_actions", "orc_num_players", "orc_max_game_length",
              "orc_observation_tensor_size", "orc_information_state_tensor_size", "orc_max_chance_outcomes"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [vp]
    for f in ("orc_min_utility", "orc_max_utility"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [vp]
    L.orc_free_game.argtypes = [vp]
    L.orc_new_initial_state.restype = vp
    L.orc_new_initial_state.argtypes = [vp]
    L.orc_clone.restype = vp
    L.orc_clone.argtypes = [vp]
    L.orc_free_state.argtypes = [vp]
    L.orc_current_player.argtypes = [vp]
    L.orc_is_terminal.argtypes = [vp]
    L.orc_legal_actions.argtypes = [vp, i64p, C.c_int]
    L.orc_apply_action.argtypes = [vp, C.c_int64]
    L.orc_returns.argtypes = [vp, dp]
    L.orc_observation_tensor.argtypes = [vp, C.c_int, fp]
    L.orc_information_state_tensor.argtypes = [vp, C.c_int, fp]
    L.orc_to_string.argtypes = [vp, cp, C.c_int]
    L.orc_information_state_string.argtypes = [vp, C.c_int, cp, C.c_int]
    L.orc_observation_string.argtypes = [vp, C.c_int, cp, C.c_int]
    L.orc_chance_outcomes.argtypes = [vp, i64p, dp, C.c_int]
    L.orc_history.argtypes = [vp, i64p, C.c_int]
    L.orc_error.argtypes = [vp, cp, C.c_int]
    _LIB = L
    return L


def parse_game_string(s):
    """'go(board_size=7,komi=4.5)' -> ('go', {'board_size': 7.0, 'komi': 4.5})."""
    s = s.strip()
    if "(" not in s:
        return s, {}
    name, rest = s.split("(", 1)
    rest = rest.rstrip(")")
    params = {}
    for kv in rest.split(","):
        if not kv.strip():
            continue
        k, v = kv.split("=")
        v = v.strip()
        params[k.strip()] = {"True": 1.0, "true": 1.0, "False": 0.0, "false": 0.0}.get(v, None)
        if params[k.strip()] is None:
            params[k.strip()] = float(v)
    return name, params


class OracleGame:
    def __init__(self, game_string, **kw):
        name, params = parse_game_string(game_string)
        params.update({k: float(v) for k, v in kw.items()})
        self.name, self.params = name, params
        L = lib()
        keys = (C.c_char_p * len(params))(*[k.encode() for k in params])
        vals = (C.c_double * len(params))(*[params[k] for k in params])
        self._g = L.orc_load_game(name.encode(), len(params), keys, vals)
        if not self._g:
            raise ValueError("oracle: unknown game " + game_string)
        self.num_distinct_actions = L.orc_num_distinct_actions(self._g)
        self.num_players = L.orc_num_players(self._g)
        self.max_game_length = L.orc_max_game_length(self._g)
        self.observation_tensor_size = L.orc_observation_tensor_size(self._g)
        self.information_state_tensor_size = L.orc_information_state_tensor_size(self._g)
        self.max_chance_outcomes = L.orc_max_chance_outcomes(self._g)

    def new_initial_state(self):
        return OracleState(self, lib().orc_new_initial_state(self._g))

    def __del__(self):
        try:
            lib().orc_free_game(self._g)
        except Exception:
            pass


class OracleState:
    def __init__(self, game, ptr):
        self.game, self._s = game, ptr

    def __del__(self):
        try:
            lib().orc_free_state(self._s)
        except Exception:
            pass

    def clone(self):
        return OracleState(self.game, lib().orc_clone(self._s))

    def current_player(self):
        return lib().orc_current_player(self._s)

    def is_terminal(self):
        return bool(lib().orc_is_terminal(self._s))

    def is_chance_node(self):
        return self.current_player() == -1

    def legal_actions(self):
        cap = max(self.game.num_distinct_actions, self.game.max_chance_outcomes, 1) + 8
        buf = (C.c_int64 * cap)()
        n = lib().orc_legal_actions(self._s, buf, cap)
        return list(buf[:n])

    def rollout_candidates(self):
        cap = max(self.game.num_distinct_actions, self.game.max_chance_outcomes, 1) + 8
        buf = (C.c_int64 * cap)()
        lib().orc_rollout_candidates.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        n = lib().orc_rollout_candidates(self._s, buf, cap)
        return list(buf[:n])

    def apply_action(self, a):
        if lib().orc_apply_action(self._s, int(a)):
            buf = C.create_string_buffer(256)
            lib().orc_error(self._s, buf, 256)
            raise RuntimeError("oracle: " + buf.value.decode())

    def returns(self):
        buf = (C.c_double * self.game.num_players)()
        lib().orc_returns(self._s, buf)
        return list(buf)

    def observation_tensor(self, player=0):
        import numpy as np
        out = np.zeros(self.game.observation_tensor_size, dtype=np.float32)
        lib().orc_observation_tensor(self._s, player, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def information_state_tensor(self, player=0):
        import numpy as np
        out = np.zeros(self.game.information_state_tensor_size, dtype=np.float32)
        lib().orc_information_state_tensor(self._s, player, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def _str(self, fn, *args):
        buf = C.create_string_buffer(4096)
        fn(self._s, *args, buf, 4096)
        return buf.value.decode()

    def to_string(self):
        return self._str(lib().orc_to_string)

    __str__ = to_string

    def information_state_string(self, player=0):
        return self._str(lib().orc_information_state_string, player)

    def observation_string(self, player=0):
        return self._str(lib().orc_observation_string, player)

    def chance_outcomes(self):
        cap = self.game.max_chance_outcomes + 8
        a = (C.c_int64 * cap)()
        p = (C.c_double * cap)()
        n = lib().orc_chance_outcomes(self._s, a, p, cap)
        return [(a[i], p[i]) for i in range(n)]

    def history(self):
        cap = 1024
        buf = (C.c_int64 * cap)()
        n = lib().orc_history(self._s, buf, cap)
        return list(buf[:n])


def oracle_mcts(state, uct_c, max_simulations, n_rollouts=1, solve=True, seed=0, tree_index=0, puct=False,
                reference_rng=False, max_nodes=1):
    """oracle/algorithms/mcts.cc: one search; returns dict(children=[(action, visits, reward, outcome_p0)], ...)."""
    import math
    L = lib()
    L.orc_mcts_search_gc.restype = C.c_int
    L.orc_mcts_search_gc.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                     C.POINTER(C.c_long), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    cap = state.game.num_distinct_actions + 4
    acts, vis = (C.c_int64 * cap)(), (C.c_int * cap)()
    rew, outc = (C.c_double * cap)(), (C.c_double * cap)()
    best, rv, ro, nodes, ran, gcs = C.c_int64(), C.c_int(), C.c_double(), C.c_long(), C.c_int(), C.c_int()
    n = L.orc_mcts_search_gc(state.game._g, state._s, uct_c, max_simulations, n_rollouts, int(solve), seed, tree_index,
                             acts, vis, rew, outc, cap, C.byref(best), C.byref(rv), C.byref(ro), C.byref(nodes), C.byref(ran), int(puct),
                             int(reference_rng), int(max_nodes), C.byref(gcs))
    return {"children": [(acts[i], vis[i], rew[i], outc[i]) for i in range(n)], "best_action": best.value,
            "root_visits": rv.value, "root_outcome_p0": ro.value, "nodes": nodes.value, "sims_run": ran.value, "gc_runs": gcs.value}


def oracle_record_trajectory(state, seed, lane, T, forced=None):
    """oracle/algorithms/trajectories.cc: one episode from `state`; dict of [T]-padded numpy rows + rewards + length."""
    import numpy as np
    L = lib()
    g = state.game
    use_info = g.information_state_tensor_size > 0
    A, F, P = g.num_distinct_actions, (g.information_state_tensor_size if use_info else g.observation_tensor_size), g.num_players
    legal = np.zeros((T, A), dtype=np.int32)
    obs = np.zeros((T, F), dtype=np.float32)
    actions = np.zeros(T, dtype=np.int64)
    players, valid, nit = (np.zeros(T, dtype=np.int32) for _ in range(3))
    rewards = np.zeros(P, dtype=np.float64)
    L.orc_record_trajectory.restype = C.c_int
    L.orc_record_trajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int] + [C.c_void_p] * 7 + \
                                       [C.c_void_p, C.c_int]
    fa = np.asarray(forced, dtype=np.int64) if forced is not None else None
    n = L.orc_record_trajectory(g._g, state._s, seed, lane, T, int(use_info), legal.ctypes.data, obs.ctypes.data,
                                actions.ctypes.data, players.ctypes.data, valid.ctypes.data, nit.ctypes.data,
                                rewards.ctypes.data, fa.ctypes.data if fa is not None else None,
                                len(fa) if fa is not None else 0)
    return {"length": n, "legal_actions": legal, "observations": obs, "actions": actions, "player_ids": players,
            "valid": valid, "next_is_terminal": nit, "rewards": rewards}


class OracleCFR:
    """oracle/algorithms/cfr.cc: restatement of algorithms::CFRSolver / CFRPlusSolver."""

    def __init__(self, game, linear_averaging=False, regret_matching_plus=False):
        L = lib()
        L.orc_cfr_new.restype = C.c_void_p
        L.orc_cfr_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_cfr_free.argtypes = [C.c_void_p]
        L.orc_cfr_iterate.argtypes = [C.c_void_p, C.c_int]
        L.orc_cfr_num_infosets.argtypes = [C.c_void_p]
        L.orc_cfr_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
        self.game = game
        self._c = L.orc_cfr_new(game._g, int(linear_averaging), int(regret_matching_plus))

    def __del__(self):
        try:
            lib().orc_cfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        lib().orc_cfr_iterate(self._c, iters)

    def table(self):
        """{info_state_string: dict(legal, regrets, cum_policy, cur_policy, player)}"""
        L = lib()
        out = {}
        for k in range(L.orc_cfr_num_infosets(self._c)):
            key = C.create_string_buffer(512)
            legal = (C.c_int64 * 16)()
            r, cu, cp = (C.c_double * 16)(), (C.c_double * 16)(), (C.c_double * 16)()
            pl = C.c_int()
            n = L.orc_cfr_get(self._c, k, key, 512, legal, r, cu, cp, 16, C.byref(pl))
            out[key.value.decode()] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n]),
                                       "cur_policy": list(cp[:n]), "player": pl.value}
        return out


class OracleMCCFR:
    """oracle/algorithms/mccfr.cc: restatement of algorithms::ExternalSamplingMCCFRSolver (kSimple averaging).
    rng_mode 0 = the reference's std::mt19937 stream, 1 = the device solver's position-keyed Philox stream."""

    def __init__(self, game, seed=0, rng_mode=1, traversals_per_update=1, full_average=False):
        L = lib()
        L.orc_mccfr_new.restype = C.c_void_p
        L.orc_mccfr_new.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.orc_mccfr_free.argtypes = [C.c_void_p]
        L.orc_mccfr_iterate.argtypes = [C.c_void_p, C.c_int]
        L.orc_mccfr_num_infosets.argtypes = [C.c_void_p]
        L.orc_mccfr_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
        self.game = game
        self._c = L.orc_mccfr_new(game._g, seed, rng_mode, traversals_per_update)
        L.orc_mccfr_set_full_average.argtypes = [C.c_void_p, C.c_int]
        L.orc_mccfr_set_full_average(self._c, int(full_average))

    def __del__(self):
        try:
            lib().orc_mccfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        assert lib().orc_mccfr_iterate(self._c, iters) == 0, "sampling failed (sum of probabilities < z)"

    def table(self):
        L = lib()
        out = {}
        for k in range(L.orc_mccfr_num_infosets(self._c)):
            key = C.create_string_buffer(512)
            legal = (C.c_int64 * 16)()
            r, cu = (C.c_double * 16)(), (C.c_double * 16)()
            pl = C.c_int()
            n = L.orc_mccfr_get(self._c, k, key, 512, legal, r, cu, 16, C.byref(pl))
            out[key.value.decode()] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n]),
                                       "player": pl.value}
        return out


class OracleOSMCCFR:
    """oracle/algorithms/os_mccfr.cc: restatement of algorithms::OutcomeSamplingMCCFRSolver(game, epsilon, seed).
    rng_mode 0 = the reference's mt19937 stream through the shim's distributions, 1 = the device's Philox stream."""

    def __init__(self, game, seed=0, rng_mode=1, trajectories_per_update=1, epsilon=0.6):
        L = lib()
        L.orc_osmccfr_new.restype = C.c_void_p
        L.orc_osmccfr_new.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double]
        L.orc_osmccfr_free.argtypes = [C.c_void_p]
        L.orc_osmccfr_iterate.argtypes = [C.c_void_p, C.c_int]
        L.orc_osmccfr_num_infosets.argtypes = [C.c_void_p]
        L.orc_osmccfr_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
        self.game = game
        self._c = L.orc_osmccfr_new(game._g, seed, rng_mode, trajectories_per_update, epsilon)

    def __del__(self):
        try:
            lib().orc_osmccfr_free(self._c)
        except Exception:
            pass

    def iterate(self, iters=1):
        assert lib().orc_osmccfr_iterate(self._c, iters) == 0, "sampling failed"

    def table(self):
        L = lib()
        out = {}
        for k in range(L.orc_osmccfr_num_infosets(self._c)):
            key = C.create_string_buffer(512)
            legal = (C.c_int64 * 16)()
            r, cu = (C.c_double * 16)(), (C.c_double * 16)()
            pl = C.c_int()
            n = L.orc_osmccfr_get(self._c, k, key, 512, legal, r, cu, 16, C.byref(pl))
            out[key.value.decode()] = {"legal": list(legal[:n]), "regrets": list(r[:n]), "cum_policy": list(cu[:n]),
                                       "player": pl.value}
        return out


def infostate_tensors(game):
    """{info_state_string: information-state tensor bytes} for every decision node of the oracle's game tree."""
    out = {}

    def walk(st):
        if st.is_terminal():
            return
        if not st.is_chance_node():
            p = st.current_player()
            key = st.information_state_string(p)
            if key not in out:
                out[key] = st.information_state_tensor(p).tobytes()
        for a in st.legal_actions():
            c = st.clone()
            c.apply_action(a)
            walk(c)

    walk(game.new_initial_state())
    return out
