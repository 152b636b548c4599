"""What the circulant form and the unscaled constant chunks of the fast kernels rest on (csrc/plan.cu build_sys2, lines2.cu,
march.cu), pinned on the CPU against the oracle's own factors and solvers:

  * the reference's Jacobian-form matrices are A0 diag(s) -- constant-coefficient scheme matrix times the Jacobian factor of the
    COLUMN (fdm_com1_jacobian.f90, fdm_com2_jacobian.f90:263-274) -- so the LU factors that TRIDFS / TRIDPFS leave have clean
    forward multipliers, d_j = d0_j / s_j and g_j = g0_j s_{j+1} / s_j;
  * a periodic line is solved exactly by two CYCLIC first-order recurrences with the converged constants, carried out chunk by
    chunk (zero-inflow sweeps + windows of a few chunk ends wrapping around the line), then scaled by rho_j;
  * a non-periodic stretched line is solved by constant chunks in the variable w = x / rho everywhere but next to the walls.
These are numpy restatements of the algorithms (not of the CUDA code); the GPU parity tests cover the kernels."""
import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched

C, LB = 16, 6


def _factors(plan, second, periodic):
    """(a, d, g) of  y_j = f_j + a_j y_{j-1},  x_j = d_j y_j + g_j x_{j+1}  as plan.cu derives them, and the column factor s."""
    der = plan.der2 if second else plan.der1
    lu = der.lu[1:, 1:]
    n = lu.shape[0]
    s = der.lhs[1:, 2].copy()
    if periodic:
        alpha, beta, gamma = lu[:, 0], lu[:, 1], lu[:, 2]
        a = np.zeros(n)
        a[1:] = alpha[1:] * beta[:-1] / beta[1:]
        return a, beta.copy(), gamma.copy(), s
    return lu[:, 0].copy(), lu[:, 1].copy(), lu[:, 2] * lu[:, 1], s


@pytest.mark.parametrize("second", [False, True])
@pytest.mark.parametrize("case", [("periodic", 256), ("periodic", 2048), ("tanh", 512), ("stretched", 256)])
def test_reference_lu_is_the_lu_of_a_column_scaled_constant_matrix(case, second):
    from oracle import fdm
    kind, n = case
    per = kind == "periodic"
    x = grid_periodic(n) if per else (grid_tanh(n) if kind == "tanh" else grid_stretched(n))
    plan = fdm.Plan(x, per, per)
    a, d, g, s = _factors(plan, second, per)
    # the column factor is the Jacobian (squared for the second derivative), noise and all
    jac = plan.jac[1:, 1] ** (2 if second else 1)
    assert np.abs(s / jac / (s[n // 2] / jac[n // 2]) - 1).max() < 1e-15
    m = n // 2
    lo, hi = 4 * C, n - 2 * C                      # away from the start-up of the LU and from the last rows
    ratio = s[lo:hi] / s[lo + 1:hi + 1]
    assert np.abs(a[lo:hi] / a[m] - 1).max() < 4e-15
    assert np.abs(d[lo:hi] * s[lo:hi] / (d[m] * s[m]) - 1).max() < 4e-15
    assert np.abs(g[lo:hi] * ratio / (g[m] * s[m] / s[m + 1]) - 1).max() < 4e-15
    if per and n >= 2048:
        # ... while the raw factors of a long uniform line carry Jacobian noise hundreds of times the round-off of the clean ones
        assert np.abs(d[lo:hi] / d[m] - 1).max() > 1e-13


def _const_chunk_tables(ca, cd, cg):
    P = ca ** np.arange(1, C + 1)
    R = cg ** np.arange(C, 0, -1)
    Q = np.zeros(C)
    q = 0.0
    for j in range(C - 1, -1, -1):
        q = cd * P[j] + cg * q
        Q[j] = q
    return Q, R


def _local(f, ca, cd, cg):
    """zero-inflow sweeps of every chunk: (x^, y_end)."""
    T = f.size // C
    xh, ye = np.zeros_like(f), np.zeros(T)
    for t in range(T):
        e = 0.0
        y = np.zeros(C)
        for j in range(C):
            e = f[t * C + j] + ca * e
            y[j] = e
        ye[t] = e
        xb = 0.0
        for j in range(C - 1, -1, -1):
            xb = cd * y[j] + cg * xb
            xh[t * C + j] = xb
    return xh, ye


@pytest.mark.parametrize("second", [False, True])
@pytest.mark.parametrize("n", [128, 1024, 2048])
def test_circulant_form_reproduces_tridpss(n, second):
    from oracle import fdm
    plan = fdm.Plan(grid_periodic(n), True, True)
    der = plan.der2 if second else plan.der1
    a, d, g, s = _factors(plan, second, True)
    mid = slice(n // 4, n - n // 4)
    ca, cd, cg = a[mid].mean(), d[mid].mean(), g[mid].mean()
    rho = (1.0 / s) / np.mean(1.0 / s[mid])
    rng = np.random.default_rng(3)
    f = rng.standard_normal(n)
    ref = f.copy()
    fdm.tridpss(*[der.lu[1:, k].copy() for k in range(1, 6)], ref)
    T = n // C
    Q, R = _const_chunk_tables(ca, cd, cg)
    xh, ye = _local(f, ca, cd, cg)
    wf, wb = (ca ** C) ** np.arange(LB), (cg ** C) ** np.arange(LB)
    A = np.array([sum(wf[k] * ye[(t - 1 - k) % T] for k in range(LB)) for t in range(T)])
    z = xh[::C] + Q[0] * A
    B = np.array([sum(wb[k] * z[(t + 1 + k) % T] for k in range(LB)) for t in range(T)])
    x = (xh.reshape(T, C) + np.outer(A, Q) + np.outer(B, R)).reshape(n) * rho
    assert np.abs(x - ref).max() <= 2e-14 * np.abs(ref).max()
    # without the column scaling the constant-coefficient solution misses the reference by its Jacobian noise
    if n >= 2048:
        assert np.abs(x / rho - ref).max() > 10 * np.abs(x - ref).max()


@pytest.mark.parametrize("second", [False, True])
@pytest.mark.parametrize("kind,n", [("tanh", 512), ("stretched", 256), ("tanh", 1024)])
def test_unscaled_constant_chunks_reproduce_tridss(kind, n, second):
    from oracle import fdm
    x = grid_tanh(n) if kind == "tanh" else grid_stretched(n)
    plan = fdm.Plan(x, False, False)
    der = plan.der2 if second else plan.der1
    a, d, g, s = _factors(plan, second, False)
    m = n // 2
    ca, cd, cg = a[m], d[m], g[m] * s[m] / s[m + 1]
    rho = s[m] / s
    T = n // C
    # constant chunks: flat to 2^-48 with the scaling divided out
    def const(t):
        i = np.arange(t * C, (t + 1) * C)
        if i[-1] + 1 >= n:
            return False
        tol = 2.0 ** -48
        return (np.abs(a[i] / ca - 1).max() <= tol and np.abs(d[i] / rho[i] / cd - 1).max() <= tol and
                np.abs(g[i] * s[i] / s[i + 1] / cg - 1).max() <= tol)
    isc = np.array([const(t) for t in range(T)])
    assert isc[3:T - 1].all() and not isc[0] and not isc[T - 1]      # tables only next to the walls
    rng = np.random.default_rng(4)
    f = rng.standard_normal(n)
    ref = f.copy()
    fdm.tridss(der.lu[1:, 1].copy(), der.lu[1:, 2].copy(), der.lu[1:, 3].copy(), ref)
    # chunk by chunk: table chunks with their own factors, constant chunks in w = x / rho; everything crossing chunks in x
    Qc, Rc = _const_chunk_tables(ca, cd, cg)
    xh, ye = np.zeros(n), np.zeros(T)
    Q, R, Af = np.zeros(n), np.zeros(n), np.zeros(T)
    for t in range(T):
        i0 = t * C
        if isc[t]:
            xl, yl = _local(f[i0:i0 + C], ca, cd, cg)
            xh[i0:i0 + C], ye[t] = xl, yl[0]
            Af[t] = ca ** C
        else:
            e, y, P = 0.0, np.zeros(C), np.zeros(C)
            w = 1.0
            for j in range(C):
                e = f[i0 + j] + a[i0 + j] * e
                y[j] = e
                w *= a[i0 + j]
                P[j] = w
            ye[t], Af[t] = e, w
            xb, q, w = 0.0, 0.0, 1.0
            for j in range(C - 1, -1, -1):
                xb = d[i0 + j] * y[j] + g[i0 + j] * xb
                xh[i0 + j] = xb
                w *= g[i0 + j]
                R[i0 + j] = w
                q = d[i0 + j] * P[j] + g[i0 + j] * q
                Q[i0 + j] = q
    A = np.zeros(T)
    for t in range(1, T):
        A[t] = ye[t - 1] + Af[t - 1] * A[t - 1]
    # chunk starts in x:  constant chunk  x_0 = rho_0 (w^_0 + Qc_0 A + Rc_0 B_w),  B_w = B_x / rho(next start)
    xsol = np.zeros(n)
    Bx = 0.0
    for t in range(T - 1, -1, -1):
        i0 = t * C
        if isc[t]:
            Bw = Bx / rho[i0 + C]
            xsol[i0:i0 + C] = rho[i0:i0 + C] * (xh[i0:i0 + C] + Qc * A[t] + Rc * Bw)
        else:
            xsol[i0:i0 + C] = xh[i0:i0 + C] + Q[i0:i0 + C] * A[t] + R[i0:i0 + C] * Bx
        Bx = xsol[i0]
    assert np.abs(xsol - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("second", [False, True])
@pytest.mark.parametrize("n,P", [(256, 2), (1024, 4), (1024, 8), (2048, 8)])
def test_circulant_form_split_over_ranks(n, P, second):
    """What splitz_ends_kernel / splitz_march_kernel<CIRC> exchange (csrc/splitz.cu): a rank that owns Tl consecutive chunks of a
    periodic line needs, beyond its own data, the forward ends y of the previous rank's last 3 chunks and (y, x^_0) of the next
    rank's first 3 chunks -- rank P-1 and rank 0 are neighbours like any other pair -- and reproduces TRIDPSS on its chunks."""
    from oracle import fdm
    LBM = 3
    plan = fdm.Plan(grid_periodic(n), True, True)
    der = plan.der2 if second else plan.der1
    a, d, g, s = _factors(plan, second, True)
    mid = slice(n // 4, n - n // 4)
    ca, cd, cg = a[mid].mean(), d[mid].mean(), g[mid].mean()
    rho = (1.0 / s) / np.mean(1.0 / s[mid])
    f = np.random.default_rng(11).standard_normal(n)
    ref = f.copy()
    fdm.tridpss(*[der.lu[1:, k].copy() for k in range(1, 6)], ref)
    T, Tl = n // C, n // C // P
    Q, R = _const_chunk_tables(ca, cd, cg)
    wf, wb = (ca ** C) ** np.arange(LBM), (cg ** C) ** np.arange(LBM)
    assert abs((ca ** C) ** LBM) < 2.0 ** -56 and abs((cg ** C) ** LBM) < 2.0 ** -56        # the dropped fourth term
    # phase 1 on every rank: zero-inflow sweeps of its first and last LBM chunks, ends published to the neighbours
    xh_all, ye_all = _local(f, ca, cd, cg)           # (a rank only ever reads its own chunks and the published ends of these)
    out = np.zeros(n)
    for r in range(P):
        t0 = r * Tl
        prev_y = [ye_all[(t0 - 1 - k) % T] for k in range(LBM)]                  # from rank r-1: its last LBM chunks
        next_y = [ye_all[(t0 + Tl + k) % T] for k in range(LBM)]                  # from rank r+1: its first LBM chunks
        next_x0 = [xh_all[((t0 + Tl + k) % T) * C] for k in range(LBM)]
        ye = np.concatenate([prev_y[::-1], ye_all[t0:t0 + Tl], next_y])          # index LBM + local chunk
        A = np.array([sum(wf[k] * ye[LBM + t - 1 - k] for k in range(LBM)) for t in range(Tl + LBM)])
        z = np.concatenate([xh_all[t0 * C:(t0 + Tl) * C:C], next_x0]) + Q[0] * A  # chunk starts, own and beyond the slab
        B = np.array([sum(wb[k] * z[t + 1 + k] for k in range(LBM)) for t in range(Tl)])
        x = xh_all[t0 * C:(t0 + Tl) * C].reshape(Tl, C) + np.outer(A[:Tl], Q) + np.outer(B, R)
        out[t0 * C:(t0 + Tl) * C] = x.reshape(-1) * rho[t0 * C:(t0 + Tl) * C]
    assert np.abs(out - ref).max() <= 2e-14 * np.abs(ref).max()
