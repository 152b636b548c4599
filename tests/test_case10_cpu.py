"""Golden vectors from the reference's own build: examples/Case10, Case06 and Case07 `dns.out.ref`
(tests/golden/case*_dns.out.ref).  The oracle, started from the restated initial condition of each case
(tests/tlab_cases.py), must reproduce the ten logged iterations -- time, dt, CFL number, diffusion number, min/max
dilatation -- to the printed digits.  This pins grid generation, the compact schemes on the stretched grids,
OPR_Burgers, the Poisson solver, the RK4-5 advance, TIME_COURANT and DNS_BOUNDS_CONTROL of the oracle against the
reference at once.

Case10: every digit (6 significant digits of a dilatation of 4e-4).  Case06 / Case07: the dilatation is 1e-8, i.e. what
the discrete operators leave of an O(1) cancellation; the sixth printed digit (1e-14 absolute) is round-off and may
differ by one unit, five digits are required."""
import pytest

import tlab_cases as tc


@pytest.mark.parametrize("name,dil_digits", [("case10", 6), ("case07", 6), ("case06", 5)])
def test_oracle_reproduces_reference_log(name, dil_digits):
    from oracle import fdm, dns as OD
    c = tc.CASES[name]
    x, y, z = tc.grids(c)
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    o = OD.Dns(go, **tc.dns_kwargs(c, OD, y))
    o.s[0][...] = tc.initial_scalar(c, x, y)
    rows = tc.run(o, 10)
    assert len(rows) == 11 == len(tc.reference_log(name))
    assert tc.compare_with_reference_log(name, rows, dil_digits) == []
