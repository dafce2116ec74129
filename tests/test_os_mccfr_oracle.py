"""CPU: pins the oracle's outcome-sampling MCCFR (oracle/algorithms/os_mccfr.cc) and the kFull averaging of its
external-sampling MCCFR to the UNMODIFIED reference solvers (algorithms/outcome_sampling_mccfr.cc,
external_sampling_mccfr.cc:188-230, built by oracle/ref_build.mk).  Fed the reference's own random stream — std::mt19937
through the uniform_real / discrete distributions of the abseil shim the reference is built against — with one episode per
update, the restatements must reproduce the reference's tables BIT FOR BIT."""
import pytest

from oracle_lib import OracleGame, OracleMCCFR, OracleOSMCCFR
import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


def same_tables(rt, mt):
    assert set(rt) == set(mt)
    for key, v in rt.items():
        assert v["legal"] == mt[key]["legal"]
        assert v["regrets"] == mt[key]["regrets"], (key, v["regrets"], mt[key]["regrets"])
        assert v["cum_policy"] == mt[key]["cum_policy"], (key, v["cum_policy"], mt[key]["cum_policy"])


@pytest.mark.parametrize("name,seed,eps,steps", [("kuhn_poker", 0, 0.6, [1, 9, 90, 900]), ("kuhn_poker", 1234, 0.3, [50, 2000]),
                                                 ("leduc_poker", 0, 0.6, [1, 20, 400, 3000]), ("leduc_poker", 7, 0.9, [1500])])
def test_oracle_outcome_sampling_equals_reference_bitwise(name, seed, eps, steps):
    ref = ref_lib.RefOSMCCFR(ref_lib.RefGame(name), seed, eps)
    mine = OracleOSMCCFR(OracleGame(name), seed=seed, rng_mode=0, trajectories_per_update=1, epsilon=eps)
    for k in steps:
        ref.iterate(k)
        mine.iterate(k)
        same_tables(ref.table(), mine.table())


@pytest.mark.parametrize("name,seed,steps", [("kuhn_poker", 3, [1, 30, 300]), ("leduc_poker", 5, [1, 10, 60])])
def test_oracle_full_average_equals_reference_bitwise(name, seed, steps):
    ref = ref_lib.RefMCCFR(ref_lib.RefGame(name), seed, full_average=True)
    mine = OracleMCCFR(OracleGame(name), seed=seed, rng_mode=0, traversals_per_update=1, full_average=True)
    for k in steps:
        ref.iterate(k)
        mine.iterate(k)
        same_tables(ref.table(), mine.table())


def test_reference_known_answer_outcome_sampling_kuhn():
    """outcome_sampling_mccfr_test.cc: 10000 iterations on kuhn_poker give NashConv < 0.17 with its seed."""
    ref = ref_lib.RefOSMCCFR(ref_lib.RefGame("kuhn_poker"), 39823987)
    ref.iterate(10000)
    assert ref.nash_conv() < 0.17
