"""CPU: pins the oracle's CFR restatement against the UNMODIFIED reference CFRSolver (oracle/_ref): identical
tables, bit for bit, after every one of several iterations on kuhn_poker and leduc_poker — the reference's own
"two implementations agree" criterion (python/algorithms/cfr_test.py:240-272), tightened from 1e-10 to exact —
plus the known answers of algorithms/cfr_test.cc (Kuhn exploitability, Leduc NashConv)."""
import pytest

import ref_lib
from oracle_lib import OracleCFR, OracleGame

needs_ref = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("gs,iters", [("kuhn_poker", [1, 1, 3, 5, 40]), ("leduc_poker", [1, 1, 3])])
def test_oracle_cfr_tables_equal_reference_bitwise(gs, iters):
    og, rg = OracleGame(gs), ref_lib.RefGame(gs)
    o, r = OracleCFR(og), ref_lib.RefCFR(rg)
    for k in iters:
        o.iterate(k)
        r.iterate(k)
        to, tr = o.table(), r.table()
        assert set(to) == set(tr) and len(to) == {"kuhn_poker": 12, "leduc_poker": 936}[gs]
        for key in to:
            for f in ("legal", "regrets", "cum_policy", "cur_policy"):
                assert to[key][f] == tr[key][f], (gs, key, f)


@needs_ref
def test_reference_known_answers():
    # cfr_test.cc:36-62 — Kuhn: exploitability <= 0.05 after 300 iterations; cfr_test.cc:299-301 Leduc NashConv <= 2 after 10
    r = ref_lib.RefCFR(ref_lib.RefGame("kuhn_poker"))
    r.iterate(300)
    assert r.exploitability() <= 0.05
    r = ref_lib.RefCFR(ref_lib.RefGame("leduc_poker"))
    r.iterate(10)
    assert r.nash_conv() <= 2.0


def test_oracle_cfr_converges_on_kuhn():
    # average policy at "0" (player 0 holding the jack): never... sanity on known Kuhn structure:
    o = OracleCFR(OracleGame("kuhn_poker"))
    o.iterate(300)
    t = o.table()
    assert len(t) == 12
    # with the king facing a bet ("2pb"), calling is dominant: average policy puts ~all mass on bet/call
    cp = t["2pb"]["cum_policy"]
    assert cp[1] / (cp[0] + cp[1]) > 0.99
    # with the jack facing a bet ("0pb"), folding is dominant
    cp = t["0pb"]["cum_policy"]
    assert cp[0] / (cp[0] + cp[1]) > 0.99
