"""Playthrough writer: dumps a game trajectory in the reference's playthrough text format.

Replaces open_spiel/python/algorithms/generate_playthrough.py:211-521 for sequential games, on top of
a pyspiel-compatible module (the in-tree `pyspiel` built from open_spiel_b200/adapter/pyspiel_module.cc, or the stock
one): same header block, same per-state lines, same tensor pictures, same "interesting state" rule, so a file written
here for a game/action sequence the reference also recorded (open_spiel/integration_tests/playthroughs/*.txt) is
byte-identical to the reference's.  `replay(text, pyspiel)` re-runs a recorded file: this is the reference's own
playthrough regression test (python/tests/playthrough_test.py) pointed at the drop-in module.
"""
import collections
import re

import numpy as np

_FLOAT_DIGITS = 6            # generate_playthrough.py:44


def _escape(text):
    return text.replace("\\", r"\\").replace("\n", r"\n")


def _bits(vec):
    """0/1 vector as a row of circles; long vectors collapse to zeros(n) / binvec(n, 0x...) when that is shorter."""
    full = "".join("◯" if v == 0 else "◉" for v in vec)
    short = None
    if len(vec) > 250:
        if all(v == 0 for v in vec):
            short = "zeros(%d)" % len(vec)
        elif all(v in (0, 1) for v in vec):
            width = (len(vec) + 15) // 16
            as_int = int("".join("1" if b else "0" for b in vec), 2)
            short = "binvec(%d, 0x%s)" % (len(vec), format(as_int, "x").rjust(width, "0"))
    return short if short and len(short) < len(full) else full


def _tensor_lines(tensor, label, max_cols=120):
    """generate_playthrough.py:100-137."""
    binary = bool(np.logical_or(tensor == 0, tensor == 1).all())
    if (not tensor.shape) or tensor.shape == (0,) or tensor.ndim > 3 or not binary:
        return ["%s = [%s]" % (label, ", ".join(str(round(v, 5)) for v in tensor.ravel()))]
    if tensor.ndim == 1:
        return ["%s: %s" % (label, _bits(tensor))]
    if tensor.ndim == 2:
        if len(label) + tensor.shape[1] + 2 < max_cols:
            lines, pad = ["%s: %s" % (label, _bits(tensor[0]))], " " * (len(label) + 2)
        else:
            lines, pad = ["%s:" % label, _bits(tensor[0])], ""
        lines.extend(pad + _bits(row) for row in tensor[1:])
        return lines
    lines = ["%s:" % label]
    bands = []                      # planes side by side, a new band when the line would exceed max_cols
    for plane in tensor:
        rows = [_bits(r) for r in plane]
        if not bands or len(bands[-1][0]) + len(rows[0]) + 2 > max_cols:
            bands.append(rows)
        else:
            bands[-1] = [a + "  " + b for a, b in zip(bands[-1], rows)]
    for i, band in enumerate(bands):
        if i:
            lines.append("")
        lines.extend(band)
    return lines


def _fmt_float(x):
    return ("{:.%dg}" % _FLOAT_DIGITS).format(x)


def _fmt_params(d, as_game=False):
    def fmt(v):
        return _fmt_params(v, True) if isinstance(v, dict) else _escape(str(v))
    if as_game:
        return d["name"] + "(" + ",".join("%s=%s" % (k, fmt(v)) for k, v in sorted(d.items()) if k != "name") + ")"
    return "{" + ",".join("%s=%s" % (k, fmt(v)) for k, v in sorted(d.items())) + "}"


class Observation:
    """python/observation.py `_Observation`: flat tensor + named views + string form of one observer."""

    def __init__(self, pyspiel, game, observer):
        self._obs = pyspiel._Observation(game, observer)
        self._info = self._obs.tensors_info() if self._obs.has_tensor() else []
        self._has_string = self._obs.has_string()
        self.tensor = None
        self.dict = {}
        if self._obs.has_tensor():
            self._refresh()

    def _refresh(self):
        self.tensor = np.asarray(self._obs.tensor(), dtype=np.float32)
        self.dict = {}
        off = 0
        for name, shape in self._info:
            size = int(np.prod(shape, dtype=np.int64))
            self.dict[name] = self.tensor[off:off + size].reshape(shape)
            off += size

    def set_from(self, state, player):
        self._obs.set_from(state, player)
        self._refresh()

    def string_from(self, state, player):
        return self._obs.string_from(state, player) if self._has_string else None


def make_observation(pyspiel, game, iig_type=None, params=None):
    observer = game.make_observer(iig_type, params or {})
    return None if observer is None else Observation(pyspiel, game, observer)


def _shapes(d):
    if len(d) == 1:
        return str(list(d[min(d)].shape))
    return ", ".join("%s: %s" % (k, list(v.shape)) for k, v in d.items())


class _Interesting:
    """First state of every player, first two chance nodes, first three decisions of a player, then every tenth
    (ShouldDisplayStateTracker, generate_playthrough.py:189-208)."""

    def __init__(self):
        self.seen = collections.defaultdict(int)

    def __call__(self, state):
        p = state.current_player()
        n = self.seen[p]
        self.seen[p] += 1
        if n == 0:
            return True
        if p == -1:
            return n < 2
        return n < 3 or n % 10 == 0


def playthrough_lines(pyspiel, game_string, action_sequence=None, seed=None, observation_params_string=None):
    """The playthrough of `game_string` as a list of lines; actions beyond `action_sequence` are drawn uniformly."""
    actions_in = list(action_sequence or [])
    lines = []
    show = [True]

    def out(text, force=False):
        if force or show[0]:
            lines.append(text)

    game = pyspiel.load_game(game_string)
    out("game: %s" % game_string)
    if observation_params_string:
        out("observation_params: %s" % observation_params_string)
    if seed is None:
        seed = np.random.randint(2**32 - 1)
    gt = game.get_type()
    obs_params = pyspiel.game_parameters_from_string(observation_params_string) if observation_params_string else None
    default_obs = make_observation(pyspiel, game, None, obs_params)
    info_obs = make_observation(pyspiel, game, pyspiel.IIGObservationType(perfect_recall=True))
    public_obs = private_obs = None
    if gt.information == gt.Information.IMPERFECT_INFORMATION:
        public_obs = make_observation(pyspiel, game, pyspiel.IIGObservationType(
            public_info=True, perfect_recall=False, private_info=pyspiel.PrivateInfoType.NONE))
        private_obs = make_observation(pyspiel, game, pyspiel.IIGObservationType(
            public_info=False, perfect_recall=False, private_info=pyspiel.PrivateInfoType.SINGLE_PLAYER))

    out("")
    out("GameType.chance_mode = %s" % gt.chance_mode)
    out("GameType.dynamics = %s" % gt.dynamics)
    out("GameType.information = %s" % gt.information)
    out('GameType.long_name = "%s"' % gt.long_name)
    out("GameType.max_num_players = %s" % gt.max_num_players)
    out("GameType.min_num_players = %s" % gt.min_num_players)
    out("GameType.parameter_specification = [%s]" % ", ".join('"%s"' % p for p in sorted(gt.parameter_specification)))
    out("GameType.provides_information_state_string = %s" % gt.provides_information_state_string)
    out("GameType.provides_information_state_tensor = %s" % gt.provides_information_state_tensor)
    out("GameType.provides_observation_string = %s" % gt.provides_observation_string)
    out("GameType.provides_observation_tensor = %s" % gt.provides_observation_tensor)
    out("GameType.provides_factored_observation_string = %s" % gt.provides_factored_observation_string)
    out("GameType.reward_model = %s" % gt.reward_model)
    out('GameType.short_name = "%s"' % gt.short_name)
    out("GameType.utility = %s" % gt.utility)
    out("")
    out("NumDistinctActions() = %s" % game.num_distinct_actions())
    out("PolicyTensorShape() = %s" % game.policy_tensor_shape())
    out("MaxChanceOutcomes() = %s" % game.max_chance_outcomes())
    out("GetParameters() = %s" % _fmt_params(game.get_parameters()))
    out("NumPlayers() = %s" % game.num_players())
    out("MinUtility() = {:.5}".format(game.min_utility()))
    out("MaxUtility() = {:.5}".format(game.max_utility()))
    out("UtilitySum() = %s" % game.utility_sum())
    if info_obs and info_obs.tensor is not None:
        out("InformationStateTensorShape() = %s" % _shapes(info_obs.dict))
        out("InformationStateTensorLayout() = %s" % game.information_state_tensor_layout())
        out("InformationStateTensorSize() = %s" % len(info_obs.tensor))
    if default_obs and default_obs.tensor is not None:
        out("ObservationTensorShape() = %s" % _shapes(default_obs.dict))
        out("ObservationTensorLayout() = %s" % game.observation_tensor_layout())
        out("ObservationTensorSize() = %s" % len(default_obs.tensor))
    out("MaxGameLength() = %s" % game.max_game_length())
    out('ToString() = "%s"' % str(game))

    players = list(range(game.num_players()))
    state = game.new_initial_states()[-1]
    rng = np.random.RandomState(seed)
    interesting = _Interesting()
    idx = 0
    while True:
        show[0] = interesting(state)
        out("", force=True)
        out("# State %d" % idx, force=True)
        for ln in str(state).splitlines():
            out(("# %s" % ln).rstrip())
        out("IsTerminal() = %s" % state.is_terminal())
        out("History() = %s" % [int(a) for a in state.history()])
        out('HistoryString() = "%s"' % state.history_str())
        out("IsChanceNode() = %s" % state.is_chance_node())
        out("IsSimultaneousNode() = %s" % state.is_simultaneous_node())
        out("CurrentPlayer() = %s" % state.current_player())
        if info_obs:
            for p in players:
                s = info_obs.string_from(state, p)
                if s is not None:
                    out('InformationStateString(%d) = "%s"' % (p, _escape(s)))
        if info_obs and info_obs.tensor is not None:
            for p in players:
                info_obs.set_from(state, p)
                for name, t in info_obs.dict.items():
                    label = "InformationStateTensor(%d)" % p + ("" if name == "info_state" else "." + name)
                    for ln in _tensor_lines(t, label):
                        out(ln)
        if default_obs:
            for p in players:
                s = default_obs.string_from(state, p)
                if s is not None:
                    out('ObservationString(%d) = "%s"' % (p, _escape(s)))
        if public_obs:
            s = public_obs.string_from(state, 0)
            if s is not None:
                out('PublicObservationString() = "%s"' % _escape(s))
            for p in players:
                s = private_obs.string_from(state, p)
                if s is not None:
                    out('PrivateObservationString(%d) = "%s"' % (p, _escape(s)))
        if default_obs and default_obs.tensor is not None:
            for p in players:
                default_obs.set_from(state, p)
                for name, t in default_obs.dict.items():
                    label = "ObservationTensor(%d)" % p + ("" if name == "observation" else "." + name)
                    for ln in _tensor_lines(t, label):
                        out(ln)
        if gt.chance_mode == gt.ChanceMode.SAMPLED_STOCHASTIC:
            out('SerializeState() = "%s"' % _escape(state.serialize()))
        if not state.is_chance_node():
            out("Rewards() = [%s]" % ", ".join(_fmt_float(x) for x in state.rewards()))
            out("Returns() = [%s]" % ", ".join(_fmt_float(x) for x in state.returns()))
        if state.is_terminal():
            break
        if state.is_simultaneous_node() or state.is_mean_field_node():
            raise ValueError("playthrough writer: sequential games only")
        if state.is_chance_node():
            out("ChanceOutcomes() = [%s]" % ", ".join("(%s,%s)" % (o, _fmt_float(p)) for o, p in state.chance_outcomes()))
        legal = state.legal_actions()
        out("LegalActions() = [%s]" % ", ".join(str(a) for a in legal))
        out("StringLegalActions() = [%s]" % ", ".join('"%s"' % state.action_to_string(state.current_player(), a) for a in legal))
        if idx < len(actions_in):
            action = actions_in[idx]
            if isinstance(action, str):
                action = state.string_to_action(state.current_player(), action)
        else:
            action = int(rng.choice(legal))
        out("")
        out('# Apply action "%s"' % state.action_to_string(state.current_player(), action), force=True)
        out("action: %s" % action, force=True)
        state.apply_action(action)
        idx += 1
    return lines


def playthrough(pyspiel, game_string, action_sequence=None, seed=None, observation_params_string=None):
    return "\n".join(playthrough_lines(pyspiel, game_string, action_sequence, seed, observation_params_string)) + "\n"


def recorded_params(text):
    """(game string, action ids, observation params) of a recorded playthrough (generate_playthrough.py:465-512, action-id mode)."""
    game, actions, obs_params = None, [], None
    for line in text.splitlines():
        m = re.fullmatch(r"game: (.*)", line)
        if m:
            game = m.group(1)
            continue
        m = re.fullmatch(r"observation_params: (.*)", line)
        if m:
            obs_params = m.group(1)
            continue
        m = re.fullmatch(r"action: (.*)", line)
        if m:
            actions.append(int(m.group(1)))
    if game is None:
        raise ValueError("no 'game:' line in the playthrough")
    return game, actions, obs_params


def replay(text, pyspiel):
    """Re-runs a recorded playthrough on `pyspiel`; returns the regenerated text (equal to `text` iff nothing changed)."""
    game, actions, obs_params = recorded_params(text)
    return playthrough(pyspiel, game, actions, observation_params_string=obs_params)
