"""Oracle parity at the benchmark LINE LENGTHS and eigenvalue range (thin slabs of the C3 and C4 grids).

The oracle cannot run 1024x512x1024 in a test, but every operator here acts on lines: a slab that keeps one or two
directions at full length and the benchmark's grid spacing in the thin one has the same lines, the same LU factors, the
same chunk windows and -- for OPR_Poisson -- the same range of eigenvalues lambda = kx'^2 + kz'^2 (0 ... both Nyquist
wavenumbers, singular modes included) as the full grid (opr_elliptic.f90:199-203).  C3 = 1024 x 512 x 1024 and
C4 = 2048 x 1024 x 2048 of BASELINE.json; spacing 2 pi / 1024 and 2 pi / 2048 in x and z, tanh-stretched y.
Tolerances are the north_star's: 1e-12 per operator call, 1e-10 on fields after RK steps.
"""
import numpy as np
import pytest

from common import grid_tanh, smooth_field, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _grids(nx, ny, nz, nfull):
    """x and z keep the spacing of the nfull-point benchmark direction, whatever their own length"""
    h = 2.0 * np.pi / nfull
    return np.arange(nx) * h, grid_tanh(ny), np.arange(nz) * h


def _plans(nx, ny, nz, nfull):
    from oracle import fdm
    from tlab_b200 import opr
    x, y, z = _grids(nx, ny, nz, nfull)
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    return (x, y, z), go, gg


def _field(shape, grids, seed, rough=0.05):
    """smooth part + white noise: the noise excites every wavenumber up to Nyquist (all eigenvalues of the Poisson stage)"""
    rng = np.random.default_rng(seed)
    return smooth_field(shape, grids, seed=seed) + rough * rng.standard_normal(shape)


# (nx, ny, nz, nfull): full-length x and y lines / full-length y and z lines of C3; the three C4 line lengths
SLABS = [(1024, 512, 16, 1024), (16, 512, 1024, 1024), (2048, 64, 16, 2048), (16, 1024, 16, 2048), (16, 64, 2048, 2048)]


@pytest.mark.parametrize("slab", SLABS)
def test_operators_at_benchmark_line_lengths(cuda, slab):
    """OPR_Partial (P1, P2_P1) and OPR_Burgers along every direction against the oracle."""
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz, nfull = slab
    grids, go, gg = _plans(nx, ny, nz, nfull)
    shape = (nz, ny, nx)
    a = _field(shape, grids, 5)
    v = _field(shape, grids, 6)
    u, vel = torch.from_numpy(a).to(cuda), torch.from_numpy(v).to(cuda)
    visc, schmidt = 1.0 / 5000.0, [1.0]
    B = O.Burgers(go, visc, schmidt)
    opr.OPR_Burgers_Initialize(gg, visc, schmidt)
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    Bg = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    bcs = [[0, 0], [0, 0]]
    bad = []
    for idir in range(3):
        res = torch.full_like(u, float("nan"))
        tmp = torch.full_like(u, float("nan"))
        P[idir](O.OPR_P1, nx, ny, nz, bcs, gg[idir], u, res)
        e = rel_l2(res.cpu().numpy(), O.opr_partial(idir, O.OPR_P1, bcs, go[idir], a))
        if not e <= TOL:
            bad.append(("P1", idir, e))
        P[idir](O.OPR_P2_P1, nx, ny, nz, bcs, gg[idir], u, res, tmp)
        r2, r1 = O.opr_partial(idir, O.OPR_P2_P1, bcs, go[idir], a)
        e2, e1 = rel_l2(res.cpu().numpy(), r2), rel_l2(tmp.cpu().numpy(), r1)
        if not (e2 <= TOL and e1 <= TOL):
            bad.append(("P2_P1", idir, e2, e1))
        for is_ in (0, 1):
            Bg[idir](opr.OPR_B_U_IN, is_, nx, ny, nz, bcs, u, vel, res)
            e = rel_l2(res.cpu().numpy(), B.apply(idir, is_, bcs, a, v))
            if not e <= TOL:
                bad.append(("burgers", idir, is_, e))
    assert not bad, bad


@pytest.mark.parametrize("slab", SLABS)
def test_poisson_over_the_benchmark_eigenvalue_range(cuda, slab):
    """OPR_Poisson with white-noise forcing (every mode excited, both singular pairs present) against the oracle; the
    eigenvalues cover 0 ... mwn_x(Nyquist)^2 + mwn_z(Nyquist)^2 of the full grid."""
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz, nfull = slab
    grids, go, gg = _plans(nx, ny, nz, nfull)
    rng = np.random.default_rng(3)
    f = rng.standard_normal((nz, ny, nx))
    hb = 0.3 * rng.standard_normal((nz, nx))
    ht = 0.3 * rng.standard_normal((nz, nx))
    ell = O.Elliptic(go)
    p_ref, dpdy_ref = O.opr_poisson(ell, f, hb, ht)
    opr.OPR_Elliptic_Initialize(gg)
    p = torch.from_numpy(f).to(cuda)
    t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=cuda)
    t2 = torch.zeros_like(t1)
    dpdy = torch.zeros_like(p)
    opr.OPR_Poisson(nx, ny, nz, opr.BCS_NN, p, t1, t2, torch.from_numpy(hb).to(cuda), torch.from_numpy(ht).to(cuda), dpdy)
    e1, e2 = rel_l2(p.cpu().numpy(), p_ref), rel_l2(dpdy.cpu().numpy(), dpdy_ref)
    assert e1 <= TOL and e2 <= TOL, (e1, e2)
    # per-mode check in spectral space: no single eigenvalue may hide behind the L2 norm of the field
    c = np.fft.fft(np.fft.rfft(p.cpu().numpy(), axis=2), axis=0)
    cr = np.fft.fft(np.fft.rfft(p_ref, axis=2), axis=0)
    num = np.linalg.norm(c - cr, axis=1)
    den = np.linalg.norm(cr, axis=1)
    worst = float((num / np.maximum(den, 1e-300)).max())
    assert worst <= 1e-10, worst


@pytest.mark.parametrize("slab", [(1024, 512, 16, 1024), (16, 512, 1024, 1024), (16, 1024, 16, 2048)])
def test_substep_at_benchmark_line_lengths(cuda, slab):
    """One RK substep (sources + RHS_GLOBAL_INCOMPRESSIBLE_1 incl. Poisson + update) against the oracle: hq, hs <= 1e-12
    relative to the oracle (the bench configuration: free-slip top, linear buoyancy, one scalar)."""
    from oracle import fdm, dns as OD
    from tlab_b200 import opr, dns as GD
    nx, ny, nz, nfull = slab
    (x, y, z), go, gg = _plans(nx, ny, nz, nfull)
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[1.0], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
              buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(D,), bcs_scal_jmax=(N,))
    o, g = OD.Dns(go, **kw), GD.Dns(gg, **kw)
    shape = (nz, ny, nx)
    wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
    for i in range(3):
        f = 0.05 * _field(shape, (x, y, z), 31 + i, rough=0.01) * wall
        o.q[i][...] = f
        g.set("q%d" % (i + 1), f)
    sc = 0.5 + 0.02 * _field(shape, (x, y, z), 40, rough=0.01) * wall
    o.s[0][...] = sc
    g.set("s1", sc)
    dte = 1e-4
    o.dte = dte
    o.sources_flow()
    o.rhs_global_incompressible_1()
    g.substep(dte, 0.0, False)
    for i in range(3):
        e = rel_l2(g.get("hq%d" % (i + 1)), o.hq[i])
        assert e <= 1e-12, ("hq", i, e)
    assert rel_l2(g.get("hs1"), o.hs[0]) <= 1e-12
    assert rel_l2(g.get("p"), o.last_pressure) <= 1e-11
    g.close()
