"""Wire formats either side of the device path, bit-compatible with the reference's text formats, so that what the device
produces can be handed to stock OpenSpiel and vice versa:

  * State::Serialize / Game::DeserializeState (spiel.cc:411-430, 757-791): the action history, one action per line;
  * CFRSolverBase::Serialize / DeserializeCFRSolver (algorithms/cfr.cc:284-307, 509-532, 639-661, 704-777): sections
    [Meta] [Game] [SolverType] [SolverSpecificState] [SolverValuesTable], table entries
    "<info state string><~>legal;regrets;cum_policy;cur_policy" with doubles as C "%a" hex floats (lossless);
  * the information-state STRINGS the reference keys its tables by (kuhn_poker.cc:109-166, leduc_poker.cc:198-239),
    rebuilt from the information-state TENSORS the device solver keys by.

Pure host code (no device work): the tables come from CFRSolver.table() / go back through CFRSolver.load_table().
"""
import numpy as np

DELIMITER = "<~>"
SERIALIZATION_VERSION = 1


# ---- doubles as printf("%a") ----------------------------------------------------------------------------------------
def hex_double(x):
    """C printf("%a", x) (glibc): shortest hex mantissa, e.g. 0x1p+0, 0x1.8p-3, 0x0p+0, -0x1.999999999999ap-4."""
    x = float(x)
    if x != x:
        return "-nan" if np.signbit(x) else "nan"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    h = x.hex()                                    # [-]0x1.xxxxxxxxxxxxxp+e  (always 13 mantissa digits)
    sign = "-" if h[0] == "-" else ""
    body = h[len(sign):]
    mant, exp = body.split("p")
    lead, _, frac = mant.partition(".")
    frac = frac.rstrip("0")
    return sign + lead + ("." + frac if frac else "") + "p" + exp


def parse_double(s):
    s = s.strip()
    return float.fromhex(s) if "0x" in s.lower() else float(s)


# ---- State::Serialize ------------------------------------------------------------------------------------------------
def serialize_state(history_actions):
    """State::Serialize (spiel.cc:411-430) for games without stochastic-sampled chance: one action id per line."""
    return "".join("%d\n" % a for a in history_actions)


def deserialize_state(game, text):
    """Game::DeserializeState (spiel.cc:757-791): replay the listed actions from the initial state."""
    state = game.new_initial_state()
    for line in text.split("\n"):
        line = line.strip()
        if line:
            state.apply_action(int(line))
    return state


# ---- information-state strings from tensors -------------------------------------------------------------------------
def kuhn_information_state_string(tensor):
    """kuhn_poker.cc:109-166 (2 players): "<card>" + one 'p'/'b' per betting action; tensor = player(2).card(3).betting(3x2)."""
    t = np.asarray(tensor).reshape(-1)
    card = int(np.argmax(t[2:5]))
    out = str(card)
    for i in range(3):
        pair = t[5 + 2 * i: 7 + 2 * i]
        if pair[0] == 0 and pair[1] == 0:
            break
        out += "p" if pair[0] == 1 else "b"
    return out


def leduc_information_state_string(tensor):
    """leduc_poker.cc:198-239 (2 players, default parameters): the perfect-recall observer string of the player the
    tensor was made for (folds are not visible in the tensor, so states after a fold are out of its domain: use the
    pyspiel module's State.information_state_string for those).
    tensor = player(2) . private card(6) . public card(6) . betting[2 rounds][4 actions][2] (call = 10, raise = 01)."""
    t = np.asarray(tensor).reshape(-1)
    player = int(np.argmax(t[0:2]))
    private = int(np.argmax(t[2:8]))
    public = int(np.argmax(t[8:14])) if t[8:14].any() else None
    rounds = [[], []]
    for r in range(2):
        for i in range(4):
            pair = t[14 + r * 8 + 2 * i: 16 + r * 8 + 2 * i]
            if pair[0] == 1:
                rounds[r].append(1)        # call
            elif pair[1] == 1:
                rounds[r].append(2)        # raise
            else:
                break
    # replay the betting to recover pot and money (leduc_poker.cc:298-414: ante 1, raises of 2 then 4, 100 starting chips)
    # and whose turn it is (cur_player_, printed as "[Player: ...]" whoever observes, leduc_poker.cc:218)
    money, ante, pot, stakes = [99.0, 99.0], [1, 1], 2, 1
    cur = 0
    for r, seq in enumerate(rounds):
        actor, raises, calls = 0, 0, 0
        for a in seq:
            if a == 2:
                stakes += 2 if r == 0 else 4
                raises, calls = raises + 1, 0
            else:
                calls += 1
            pay = stakes - ante[actor]
            ante[actor] += pay
            money[actor] -= pay
            pot += pay
            actor ^= 1
        complete = (raises == 0 and calls == 2) or (raises > 0 and calls == 1)      # ReadyForNextRound, :680-683
        if r == 0 and not complete:
            cur = actor
            break
        if r == 0 and public is None:
            cur = -1                                   # waiting for the public card: kChancePlayerId
            break
        if r == 1:
            cur = actor if not complete else actor ^ 1   # after the last call the mover stays (the state is terminal)
    rnd = 2 if public is not None else 1
    fmt = lambda v: ("%d" % v) if float(v).is_integer() else repr(float(v))   # noqa: E731
    s = "[Observer: %d][Private: %d][Round %d][Player: %d][Pot: %d][Money: %s %s]" % (
        player, private, rnd, cur, pot, fmt(money[0]), fmt(money[1]))
    if public is not None:
        s += "[Public: %d]" % public
    s += "[Round1: %s][Round2: %s]" % (" ".join(map(str, rounds[0])), " ".join(map(str, rounds[1])))
    return s


INFORMATION_STATE_STRING = {"kuhn_poker": kuhn_information_state_string, "leduc_poker": leduc_information_state_string}


def table_keys(game_name, table):
    """Information-state strings of every row of a CFRSolver.table()."""
    f = INFORMATION_STATE_STRING[game_name]
    return [f(k) for k in table["keys"]]


# ---- CFR solver text format ----------------------------------------------------------------------------------------
def serialize_values_table(keys, table, delimiter=DELIMITER):
    """SerializeCFRInfoStateValuesTable (cfr.cc:639-661), double_precision = -1 (hex floats)."""
    parts = []
    for k, key in enumerate(keys):
        lo, hi = int(table["offsets"][k]), int(table["offsets"][k + 1])
        vals = ";".join([",".join(str(int(a)) for a in table["legal_actions"][lo:hi])] +
                        [",".join(hex_double(v) for v in table[f][lo:hi]) for f in ("regrets", "cum_policy", "cur_policy")])
        parts += [key, vals]
    return delimiter.join(parts)


def serialize_cfr_solver(game_string, solver_type, iteration, keys, table, delimiter=DELIMITER):
    """CFRSolverBase::Serialize (cfr.cc:284-307); solver_type "CFRSolver" / "CFRPlusSolver"."""
    return ("# Automatically generated by OpenSpiel CFRSolverBase::Serialize\n[Meta]\nVersion: %d\n\n[Game]\n%s\n[SolverType]\n%s\n"
            "[SolverSpecificState]\n%d\n[SolverValuesTable]\n" % (SERIALIZATION_VERSION, game_string, solver_type, iteration)
            ) + serialize_values_table(keys, table, delimiter)


def deserialize_cfr_solver(text, delimiter=DELIMITER):
    """PartiallyDeserializeCFRSolver + DeserializeCFRInfoStateValuesTable (cfr.cc:663-777).  Returns
    dict(game, solver_type, iteration, table={key: dict(legal, regrets, cum_policy, cur_policy)})."""
    head, _, values = text.partition("[SolverValuesTable]\n")
    sections, cur = {}, None
    for line in head.split("\n"):
        if not line or line[0] == "#":
            continue
        if line in ("[Meta]", "[Game]", "[SolverType]", "[SolverSpecificState]"):
            cur = line
            sections[cur] = []
        elif cur is not None:
            sections[cur].append(line)
    table = {}
    if values:
        splits = values.split(delimiter)
        for i in range(0, len(splits) - 1, 2):
            legal, regrets, cum, cur_p = (f.split(",") for f in splits[i + 1].split(";"))
            table[splits[i]] = {"legal": [int(a) for a in legal], "regrets": [parse_double(v) for v in regrets],
                                "cum_policy": [parse_double(v) for v in cum], "cur_policy": [parse_double(v) for v in cur_p]}
    return {"game": "".join(sections.get("[Game]", [])), "solver_type": "".join(sections.get("[SolverType]", [])),
            "iteration": int(sections.get("[SolverSpecificState]", ["0"])[0]), "table": table}


def table_arrays_from(parsed_table, keys, table_layout):
    """Flat (regrets, cum_policy, cur_policy) arrays in the row order of `table_layout` (a CFRSolver.table()) from a parsed
    {key: values} table — the arguments of CFRSolver.load_table()."""
    E = len(table_layout["legal_actions"])
    out = {f: np.zeros(E) for f in ("regrets", "cum_policy", "cur_policy")}
    for k, key in enumerate(keys):
        lo, hi = int(table_layout["offsets"][k]), int(table_layout["offsets"][k + 1])
        v = parsed_table[key]
        assert v["legal"] == [int(a) for a in table_layout["legal_actions"][lo:hi]], key
        for f in out:
            out[f][lo:hi] = v[f]
    return out["regrets"], out["cum_policy"], out["cur_policy"]
