"""Oracle of the direct Poisson solver (oracle/integral2.py: FDM_Int2_*, OPR_Poisson_FourierXZ_Direct; SURVEY 8 row f4, ORACLE ONLY --
the CUDA path of this variant is not built).  The reference stores no output for this path, so the restatement is pinned on what
its result must satisfy, computed independently of it with dense matrices of the direct second derivative:
  * rows 3 .. n-2:  (B2 - lambda A2) p = A2 f  (the discrete equation p'' - lambda p = f with p'' = A2^-1 B2 p);
  * rows 2 and n-1: the same with p''_1, p''_n eliminated through the wall rows of the scheme;
  * Dirichlet ends keep the given value, Neumann ends satisfy the reference's biased fourth-order wall formula
    p'_1 = b1 p1 + b2 p2 + b3 p3 + b4 p4 + a2 p''_2 (fdm_integral.f90:561-618) with p''_2 = f_2 + lambda p_2;
  * the whole solver: the discrete Laplacian (Fourier multipliers of the x / z second derivatives + direct scheme in y) of the
    returned p equals the forcing in the interior rows, and p reproduces a smooth field at truncation level."""
import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched, dense_direct2 as _dense


@pytest.fixture(scope="module")
def setup():
    from oracle import fdm
    out = {}
    for kind, n in (("tanh", 65), ("stretched", 48)):
        y = grid_tanh(n) if kind == "tanh" else grid_stretched(n)
        plan = fdm.Plan(y, False, False, mode2=fdm.FDM_COM6_DIRECT)
        A, B = _dense(plan.der2, n)
        # the dense matrices are the scheme: sixth-order second derivative of a smooth function
        u = np.sin(2.0 * y) + 0.3 * y ** 2
        d2 = np.linalg.solve(A, B @ u)
        assert np.abs(d2 - (-4.0 * np.sin(2.0 * y) + 0.6))[3:-3].max() < 2e-5
        from oracle import operators as O
        ref = O.opr_partial(1, O.OPR_P2, [[0, 0], [0, 0]], plan, (u[None, :, None] * np.ones((1, n, 2))))[0, :, 0]
        assert np.abs(d2 - ref).max() <= 1e-10 * np.abs(ref).max()
        out[kind] = (y, plan, A, B)
    return out


@pytest.mark.parametrize("kind", ["tanh", "stretched"])
@pytest.mark.parametrize("ibc", ["DD", "ND", "DN", "NN"])
def test_int2_solution_satisfies_the_discrete_equation(setup, kind, ibc):
    from oracle import fdm, integral2 as I2
    y, plan, A, B = setup[kind]
    n = y.size
    code = {"DD": fdm.BCS_DD, "ND": fdm.BCS_ND, "DN": fdm.BCS_DN, "NN": fdm.BCS_NN}[ibc]
    lam = np.array([0.7, 12.5, 431.0, 0.0 if ibc != "NN" else 3.0e4])
    M = lam.size
    rng = np.random.default_rng(2)
    f = rng.standard_normal((n, 2, M))
    bc_b, bc_t = rng.standard_normal((2, M)), rng.standard_normal((2, M))
    fdmi = I2.int2_initialize(y, plan.der2, lam, code)
    u = np.zeros((n, 2, M))
    u[0], u[n - 1] = bc_b, bc_t
    I2.int2_solve(fdmi, fdmi.rhs, f.copy(), u)
    xp = I2._pad(y)
    cb = I2.coef_c1n4_biased(xp, 1)
    ct = I2.coef_c1n4_biased(xp, n, backwards=True)
    for m in range(M):
        for l in range(2):
            p, ff = u[:, l, m], f[:, l, m]
            scale = np.abs(B @ p).max() + lam[m] * np.abs(A @ p).max() + np.abs(A @ ff).max()
            # second derivative implied by the equation in rows 2 .. n-1, by the wall rows of the scheme in rows 1 and n
            pp = ff + lam[m] * p
            pp[0] = (B[0] @ p - A[0, 1:] @ pp[1:]) / A[0, 0]
            pp[n - 1] = (B[n - 1] @ p - A[n - 1, :n - 1] @ pp[:n - 1]) / A[n - 1, n - 1]
            res = A @ pp - B @ p
            assert np.abs(res[1:n - 1]).max() <= 2e-11 * scale, (ibc, m, l, np.abs(res[1:n - 1]).max() / scale)
            # boundary data
            if ibc in ("DD", "DN"):
                assert p[0] == bc_b[l, m]
            else:
                d1 = cb[1] * p[0] + cb[2] * p[1] + cb[3] * p[2] + cb[4] * p[3] + cb[5] * pp[1]
                assert abs(d1 - bc_b[l, m]) <= 1e-10 * (np.abs(cb[1:5]).max() * np.abs(p[:4]).max() + abs(cb[5] * pp[1]))
            if ibc in ("DD", "ND"):
                assert p[n - 1] == bc_t[l, m]
            else:
                d1 = ct[1] * p[n - 1] + ct[2] * p[n - 2] + ct[3] * p[n - 3] + ct[4] * p[n - 4] + ct[5] * pp[n - 2]
                assert abs(d1 - bc_t[l, m]) <= 1e-10 * (np.abs(ct[1:5]).max() * np.abs(p[-4:]).max() + abs(ct[5] * pp[n - 2]))


def test_wall_formula_is_fourth_order():
    """p'_1 = b1 p1 + b2 p2 + b3 p3 + b4 p4 + a2 p''_2; uniform grid: (-29/6, 54/6, -27/6, 2/6)/h and 3h (the reference's comment)."""
    from oracle import integral2 as I2
    h = 0.05
    x = I2._pad(np.arange(8) * h)
    c = I2.coef_c1n4_biased(x, 1)
    assert np.allclose(c[1:5] * h, [-29 / 6, 54 / 6, -27 / 6, 2 / 6], rtol=1e-12)
    assert np.isclose(c[5] / h, 3.0, rtol=1e-12)
    cb = I2.coef_c1n4_biased(x, 8, backwards=True)
    assert np.allclose(cb[1:5] * h, [29 / 6, -54 / 6, 27 / 6, -2 / 6], rtol=1e-12) and np.isclose(cb[5] / h, -3.0, rtol=1e-12)


def test_direct_poisson_solver():
    from oracle import fdm, integral2 as I2, operators as O
    nx, ny, nz = 16, 49, 8
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    g = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    ell = I2.EllipticDirect(g, y)
    Lx, Lz = x[-1] + x[1] - x[0], z[-1] + z[1] - z[0]
    X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]
    a = (np.cos(2 * np.pi * X / Lx) * np.sin(2 * np.pi * Z / Lz) + 0.5 * np.cos(4 * np.pi * X / Lx)) * np.cos(1.3 * Y) + 0.2 * np.sin(0.7 * Y)
    day = -(np.cos(2 * np.pi * X / Lx) * np.sin(2 * np.pi * Z / Lz) + 0.5 * np.cos(4 * np.pi * X / Lx)) * 1.3 * np.sin(1.3 * Y) \
        + 0.14 * np.cos(0.7 * Y) + 0.0 * Z
    bcs = [[0, 0], [0, 0]]
    gd = fdm.Plan(y, False, False, name="yd", mode2=fdm.FDM_COM6_DIRECT)

    def laplacian(q):
        return (O.opr_partial(0, O.OPR_P2, bcs, g[0], q) + O.opr_partial(2, O.OPR_P2, bcs, g[2], q)
                + O.opr_partial(1, O.OPR_P2, bcs, gd, q))
    f = laplacian(a)
    p, dpdy = I2.opr_poisson_direct(ell, f, day[:, 0, :], day[:, ny - 1, :])
    # the discrete equation, interior rows (the wall rows carry boundary data)
    res = laplacian(p) - f
    assert np.abs(res[:, 2:ny - 2, :]).max() <= 1e-10 * np.abs(f).max()
    # the solution: p = a up to the constant fixed by p(bottom) = 0 in the mean mode, at truncation level (Neumann wall formula)
    shift = (p - a).mean()
    assert np.abs(p - a - shift).max() <= 2e-4 * np.abs(a).max()
    assert abs(p[:, 0, :].mean()) <= 1e-12
    assert np.abs(dpdy - day).max() <= 2e-3 * np.abs(day).max()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_direct_second_derivative_is_exact_for_polynomials_on_a_random_grid(seed):
    """An independent pin of oracle/fdm_direct.py (FDM_C2N6_Direct, fdm_comx_direct.f90:305-412; the reference holds no golden
    for it): on an arbitrary non-uniform grid the compact relation A2 p'' = B2 p must hold EXACTLY for polynomials up to degree 6
    in the interior rows and up to degree 4 in the two rows at either wall (coef_c2n3_biased / coef_c2n4) -- and must fail beyond
    (so the test is not vacuous).  Any wrong coefficient breaks the exactness."""
    from oracle import fdm
    rng = np.random.default_rng(seed)
    n = 24
    x = np.cumsum(0.5 + rng.random(n))
    x = (x - x[0]) / (x[-1] - x[0])
    plan = fdm.Plan(x, False, False, mode2=fdm.FDM_COM6_DIRECT)
    A, B = _dense(plan.der2, n)
    xc = x - 0.5
    for k in range(9):
        p = xc ** k
        pp = k * (k - 1) * xc ** (k - 2) if k >= 2 else 0.0 * xc
        r = np.abs(A @ pp - B @ p) / (np.abs(A) @ np.abs(pp) + np.abs(B) @ np.abs(p))
        wall = np.r_[r[:2], r[-2:]]
        if k <= 4:
            assert wall.max() <= 1e-14, (k, wall)
        if k <= 6:
            assert r[2:-2].max() <= 1e-14, (k, r[2:-2].max())
        if k == 5:
            assert wall.min() > 1e-9          # third / fourth-order wall closures
        if k == 8:
            assert r[3:-3].max() > 1e-4       # sixth-order interior scheme
