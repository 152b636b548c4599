"""Golden vector from the reference's own build: examples/Case10/dns.out.ref (tests/golden/case10_dns.out.ref).
The oracle, started from the restated initial condition of the case (tests/case10.py), must reproduce the ten logged
iterations -- time, dt, CFL number, diffusion number, min/max dilatation -- to every printed digit.  This pins grid
generation, the compact schemes on the stretched grid, OPR_Burgers, the Poisson solver, the RK4-5 advance, TIME_COURANT
and DNS_BOUNDS_CONTROL of the oracle against the reference at once."""
import case10


def test_oracle_reproduces_case10_log():
    from oracle import fdm, dns as OD
    x, y, z = case10.grids()
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    o = OD.Dns(go, **case10.dns_kwargs(OD, y))
    o.s[0][...] = case10.initial_scalar(x, y)
    rows = case10.run(o, 10)
    assert len(rows) == 11 == len(case10.reference_log())
    assert case10.compare_with_reference_log(rows) == []
    assert abs(rows[0]["dt"] - 0.457764e-2) < 1e-8          # the log's time column after one step
