"""Pins the CPU oracle (numpy restatement of the reference) against
  * golden vectors produced by the reference's own compact_lib.py (tests/golden/make_golden.py),
  * the uniform-grid coefficient limits quoted in the reference sources,
  * the reference's self-consistency recipes (src/valid/: vpartial, vburgers, vintegral, vpoisson),
  * analytic convergence orders.
"""
import os

import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched, smooth_field, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "compact_lib_c1n6.npz"))


def test_der1_matches_compact_lib_periodic(golden):
    from oracle import fdm
    x, u, du = golden["per_x"], golden["per_u"], golden["per_du"]
    g = fdm.Plan(x, True, True)
    got = fdm.der1_solve(0, g.der1, g.der1.lu, u.copy())
    assert rel_l2(got, du) <= 1e-13


def test_der1_matches_compact_lib_nonuniform(golden):
    from oracle import fdm
    x, u, du = golden["nonuni_x"], golden["nonuni_u"], golden["nonuni_du"]
    g = fdm.Plan(x, False, False)
    got = fdm.der1_solve(0, g.der1, g.der1.lu, u.copy())
    assert rel_l2(got, du) <= 1e-13


def test_der1_matches_compact_lib_uniform_biased(golden):
    from oracle import fdm
    x, u, du = golden["uni_x"], golden["uni_u"], golden["uni_du"]
    g = fdm.Plan(x, False, True)
    got = fdm.der1_solve(0, g.der1, g.der1.lu, u.copy())
    # compact_lib uses central differences for the uniform Jacobian, tlab the scheme itself: equal to round-off
    assert rel_l2(got, du) <= 1e-12


def test_uniform_coefficient_limits():
    """(2/11, 1, 2/11 | 3/44, 12/11, -51/22, 12/11, 3/44)/h^2 and the 1/3, 14/9, 1/9 first-derivative stencil
    (fdm_comx_direct.f90:45,104; fdm_com1_jacobian.f90:108-109; fdm_com2_jacobian.f90:102-103)."""
    from oracle import fdm
    n, h = 32, 0.125
    x = np.arange(n) * h
    g = fdm.Plan(x, True, True, mode2=fdm.FDM_COM6_JACOBIAN)
    b1 = 7.0 / 9.0
    assert np.allclose(g.der1.lhs[5, 1:4] * b1 / h, [1 / 3, 1, 1 / 3], rtol=1e-13)
    assert np.allclose(g.der1.rhs[5, 1:6] * b1, [-1 / 36, -7 / 9, 0, 7 / 9, 1 / 36], rtol=1e-13, atol=1e-16)
    c1 = 12.0 / 11.0
    assert np.allclose(g.der2.lhs[5, 1:4] * c1 / h ** 2, [2 / 11, 1, 2 / 11], rtol=1e-12)
    assert np.allclose(g.der2.rhs[5, 1:6] * c1, [3 / 44, 12 / 11, -51 / 22, 12 / 11, 3 / 44], rtol=1e-13)


@pytest.mark.parametrize("periodic", [True, False])
def test_sixth_order_convergence(periodic):
    """vpartial.f90:100-184: Gaussian/sine test functions, 6th-order slopes in the interior."""
    from oracle import fdm
    errs = []
    for n in (32, 64, 128):
        if periodic:
            x = grid_periodic(n)
            u, du, d2u = np.sin(3 * x), 3 * np.cos(3 * x), -9 * np.sin(3 * x)
            g = fdm.Plan(x, True, True)
        else:
            x = grid_stretched(n + 1, 1.0, 0.1)
            u = np.exp(-((x - 0.5) / 0.15) ** 2)
            du = -2 * (x - 0.5) / 0.15 ** 2 * u
            d2u = (-2 / 0.15 ** 2 + 4 * (x - 0.5) ** 2 / 0.15 ** 4) * u
            g = fdm.Plan(x, False, False)
        r1 = fdm.der1_solve(0, g.der1, g.der1.lu, u[:, None])
        r2 = fdm.der2_solve(g.der2, g.der2.lu, u[:, None], r1)
        errs.append((np.abs(r1[:, 0] - du).max(), np.abs(r2[:, 0] - d2u).max()))
    # periodic: 6th order everywhere; biased: the global "3-5-6-5-3" scheme is limited by its closures
    # (doc/numerical.tex:30-186), at least 3rd order in the maximum norm
    for k in range(2):
        slope = np.log2(errs[0][k] / errs[1][k]), np.log2(errs[1][k] / errs[2][k])
        assert min(slope) > (5.5 if periodic else 2.5), (k, errs, slope)


def test_neumann_variants_zero_the_wall_derivative():
    from oracle import fdm
    y = grid_tanh(65)
    g = fdm.Plan(y, False, False)
    u = np.cos(np.pi * y)[:, None] + 0.2 * y[:, None] ** 2
    base = fdm.der1_solve(0, g.der1, g.der1.lu, u.copy())
    for ibc, ends in ((1, (0,)), (2, (64,)), (3, (0, 64))):
        d = fdm.der1_solve(ibc, g.der1, g.der1.lu, u.copy())
        for e in ends:
            assert d[e, 0] == 0.0
        assert np.abs(d[8:-8] - base[8:-8]).max() < 1e-3      # interior unaffected beyond the closure's reach


def test_thomas_solvers_against_dense():
    from oracle import fdm
    rng = np.random.default_rng(1)
    n, m = 40, 5
    a, b, c = rng.uniform(0.1, 0.4, n), rng.uniform(1.0, 2.0, n), rng.uniform(0.1, 0.4, n)
    f = rng.standard_normal((n, m))
    A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
    la, lb, lc = a.copy(), b.copy(), c.copy()
    fdm.tridfs(la, lb, lc)
    x = f.copy()
    fdm.tridss(la, lb, lc, x)
    assert rel_l2(x, np.linalg.solve(A, f)) < 1e-13
    # circulant
    Ap = A.copy()
    Ap[0, n - 1], Ap[n - 1, 0] = a[0], c[n - 1]
    pa, pb, pc, pd, pe = a.copy(), b.copy(), c.copy(), np.zeros(n), np.zeros(n)
    fdm.tridpfs(pa, pb, pc, pd, pe)
    x = f.copy()
    fdm.tridpss(pa, pb, pc, pd, pe, x)
    assert rel_l2(x, np.linalg.solve(Ap, f)) < 1e-13
    # pentadiagonal (forward LU and reverse LE)
    d5 = [rng.uniform(0.05, 0.2, n), rng.uniform(0.1, 0.4, n), rng.uniform(1.5, 2.0, n), rng.uniform(0.1, 0.4, n),
          rng.uniform(0.05, 0.2, n)]
    P5 = np.diag(d5[2]) + np.diag(d5[1][1:], -1) + np.diag(d5[0][2:], -2) + np.diag(d5[3][:-1], 1) + np.diag(d5[4][:-2], 2)
    w = [v.copy() for v in d5]
    fdm.pentadfs(*w)
    x = f.copy()
    fdm.pentadss(*w, x)
    assert rel_l2(x, np.linalg.solve(P5, f)) < 1e-13
    w = [v.copy() for v in d5]
    fdm.pentadfs2(*w)
    x = f.copy()
    fdm.pentadss2(*w, x)
    assert rel_l2(x, np.linalg.solve(P5, f)) < 1e-13
    # circulant pentadiagonal: the Woodbury closure of PENTADPFS assumes constant diagonals (uniform periodic grid)
    cst = [np.full(n, 0.09), np.full(n, 0.56), np.full(n, 1.0), np.full(n, 0.56), np.full(n, 0.09)]
    Pp = np.zeros((n, n))
    for i in range(n):
        for off, dg in zip((-2, -1, 0, 1, 2), cst):
            Pp[i, (i + off) % n] = dg[i]
    w = [v.copy() for v in cst] + [np.zeros(n), np.zeros(n)]
    fdm.pentadpfs(*w)
    x = f.copy()
    fdm.pentadpss(*w, x)
    assert rel_l2(x, np.linalg.solve(Pp, f)) < 1e-12


def test_burgers_identity():
    """vburgers.f90:76-153: OPR_Burgers(SELF, a) == visc * d2 a - a * d a."""
    from oracle import fdm, operators as O
    nx, ny, nz = 24, 25, 16
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    g = [fdm.Plan(x, True, True), fdm.Plan(y, False, False), fdm.Plan(z, True, True)]
    a = smooth_field((nz, ny, nx), (x, y, z), seed=3)
    B = O.Burgers(g, 0.01, [0.7])
    bcs = [[0, 0], [0, 0]]
    for idir in range(3):
        r = B.apply(idir, 0, bcs, a, a)
        d2, d1 = O.opr_partial(idir, O.OPR_P2_P1, bcs, g[idir], a)
        assert rel_l2(r, 0.01 * d2 - a * d1) < 1e-14


def test_integral_operator_and_ode_solvers():
    """vintegral.f90:136-340: FDM_Int1 and OPR_ODE2_Factorize_NN / _NN_Sing against analytic solutions."""
    from oracle import fdm, integral as I
    n = 129
    y = grid_stretched(n, 1.0, 0.2)
    g = fdm.Plan(y, False, False)
    uex = np.cos(3 * y) + y ** 2
    vex = -3 * np.sin(3 * y) + 2 * y
    for lam2 in (0.0, 4.0, 400.0):
        lam = np.sqrt(lam2)
        fi = {1: I.int1_initialize(g.der1, lam, 1), 2: I.int1_initialize(g.der1, -lam, 2)}
        fex = -9 * np.cos(3 * y) + 2 - lam2 * uex
        f = np.zeros((n, 2, 1))
        f[:, 0, 0], f[:, 1, 0] = fex, -2 * fex
        bcs = np.zeros((2, 2, 1))
        bcs[0, :, 0] = [vex[0], -2 * vex[0]]
        bcs[1, :, 0] = [vex[-1], -2 * vex[-1]]
        if lam2 == 0.0:
            u, v = I.ode2_factorize_nn_sing(fi, f, bcs)
            assert np.abs(u[:, 0, 0] - (uex - uex[0])).max() < 1e-6
        else:
            u, v = I.ode2_factorize_nn(fi, fi[1].rhs, fi[2].rhs, f, bcs)
            assert np.abs(u[:, 0, 0] - uex).max() < 1e-6
            assert np.abs(u[:, 1, 0] + 2 * uex).max() < 2e-6      # linearity across lines
        assert np.abs(v[:, 0, 0] - vex).max() < 1e-5


def test_poisson_round_trip():
    """vpoisson.f90:162-248 with delta-delta operators: OPR_Poisson inverts div(grad) to round-off."""
    from oracle import fdm, operators as O
    nx, ny, nz = 32, 33, 32
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    g = [fdm.Plan(x, True, True), fdm.Plan(y, False, False), fdm.Plan(z, True, True)]
    a = smooth_field((nz, ny, nx), (x, y, z), seed=21)       # |k| <= 8 < Nyquist (16): no 2-Delta content
    bcs = [[0, 0], [0, 0]]

    def d(i, f):
        return O.opr_partial(i, O.OPR_P1, bcs, g[i], f)
    f = d(0, d(0, a)) + d(1, d(1, a)) + d(2, d(2, a))
    ay = d(1, a)
    p, dpdy = O.opr_poisson(O.Elliptic(g), f, ay[:, 0, :], ay[:, -1, :])
    assert rel_l2(p, a - a[:, 0, :].mean()) < 1e-12
    assert rel_l2(dpdy, ay) < 1e-12


def test_rk_step_is_divergence_free_and_matches_tables():
    from oracle import fdm, operators as O, dns as OD
    kdt, ktime, kco = OD.rk_coefficients(OD.RKM_EXP4)
    assert abs(sum(kdt[i] * np.prod([1.0]) for i in range(5))) > 0     # tables present
    assert abs(ktime[1] - kdt[0]) == 0.0 and len(kco) == 4
    nx, ny, nz = 16, 17, 16
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    g = [fdm.Plan(x, True, True), fdm.Plan(y, False, False), fdm.Plan(z, True, True)]
    o = OD.Dns(g, visc=1e-3, schmidt=[1.0])
    wall = np.sin(np.pi * y / y[-1])[None, :, None]
    for i in range(3):
        o.q[i][...] = 0.3 * smooth_field((nz, ny, nx), (x, y, z), seed=i, nmodes=3) * wall
    o.s[0][...] = 0.5
    o.runge_kutta(1e-3)
    bcs = [[0, 0], [0, 0]]
    div = sum(O.opr_partial(i, O.OPR_P1, bcs, g[i], o.q[i]) for i in range(3))
    assert np.abs(div[:, 1:-1, :]).max() < 1e-12       # "zero divergence down to round-off" (opr_elliptic.f90:109)


def test_ode_solvers_with_dirichlet_data():
    """OPR_ODE2_Factorize_DD / _DD_Sing (opr_odes.f90:188-260, 391-478; vintegral.f90's recipe): analytic solutions, linearity across
    lines, and -- to round-off -- the first-order systems they are built from: v = delta u + ..., i.e. the pair (u, v) returned
    satisfies u' = v and v' - lambda^2 u = f in the discrete sense of FDM_Int1 away from the walls (checked through the NN solver:
    feeding the wall derivatives of the DD solution to OPR_ODE2_Factorize_NN returns the same u).  ORACLE ONLY."""
    from oracle import fdm, integral as I
    n = 129
    y = grid_stretched(n, 1.0, 0.2)
    g = fdm.Plan(y, False, False)
    uex = np.cos(3 * y) + y ** 2
    vex = -3 * np.sin(3 * y) + 2 * y
    for lam2 in (0.0, 4.0, 400.0):
        lam = np.sqrt(lam2)
        fi = {1: I.int1_initialize(g.der1, lam, 1), 2: I.int1_initialize(g.der1, -lam, 2)}
        fex = -9 * np.cos(3 * y) + 2 - lam2 * uex
        f = np.zeros((n, 2, 1))
        f[:, 0, 0], f[:, 1, 0] = fex, -2 * fex
        bcs = np.zeros((2, 2, 1))
        bcs[0, :, 0] = [uex[0], -2 * uex[0]]
        bcs[1, :, 0] = [uex[-1], -2 * uex[-1]]
        if lam2 == 0.0:
            u, v = I.ode2_factorize_dd_sing(fi, f.copy(), bcs)
        else:
            u, v = I.ode2_factorize_dd(fi, fi[1].rhs, fi[2].rhs, f.copy(), bcs)
        assert np.abs(u[:, 0, 0] - uex).max() < 1e-6
        assert np.abs(u[:, 1, 0] + 2 * uex).max() < 2e-6          # linearity across lines
        assert np.abs(v[:, 0, 0] - vex).max() < 1e-5
        assert u[0, 0, 0] == uex[0] and abs(u[-1, 0, 0] - uex[-1]) <= 1e-13 * abs(uex[-1]) + 1e-15
        if lam2 > 0.0:
            # the same discrete solution through the Neumann solver, fed with the wall derivatives the DD solver returned
            b2 = np.zeros((2, 2, 1))
            b2[0], b2[1] = v[0], v[-1]
            u2, v2 = I.ode2_factorize_nn(fi, fi[1].rhs, fi[2].rhs, f.copy(), b2)
            assert np.abs(u2 - u).max() <= 1e-9 * np.abs(u).max()
            assert np.abs(v2 - v).max() <= 1e-9 * np.abs(v).max()


def test_poisson_round_trip_with_dirichlet_data():
    """vpoisson.f90:212-241, the DD branch: OPR_Poisson with the wall values of a given field inverts div(grad) built from delta-delta
    operators to round-off.  ORACLE ONLY (the CUDA path takes BCS_NN)."""
    from oracle import fdm, operators as O
    nx, ny, nz = 16, 33, 16
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    g = [fdm.Plan(x, True, True), fdm.Plan(y, False, False), fdm.Plan(z, True, True)]
    a = smooth_field((nz, ny, nx), (x, y, z), seed=22)
    bcs = [[0, 0], [0, 0]]

    def d(i, f):
        return O.opr_partial(i, O.OPR_P1, bcs, g[i], f)
    f = d(0, d(0, a)) + d(1, d(1, a)) + d(2, d(2, a))
    p, dpdy = O.opr_poisson(O.Elliptic(g), f, a[:, 0, :], a[:, -1, :], ibc=fdm.BCS_DD)
    assert rel_l2(p, a) < 1e-11
    assert rel_l2(dpdy, d(1, a)) < 1e-10
