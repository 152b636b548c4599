"""Oracle: first-order integral operators and second-order ODE solvers in y.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows /root/reference/src:
  fdm/fdm_integral.f90     FDM_Int1_CreateSystem (:91-214), FDM_Int1_Initialize (:58-87),
                           FDM_Int1_Solve (:219-314)
  operators/opr_odes.f90   OPR_ODE2_Factorize_DN_Sing (:37-96), _NN_Sing (:165-183),
                           _NN (:265-386), _DD (:391-478), _DD_Sing (:188-260); the DD pair is ORACLE ONLY (the CUDA path solves BCS_NN)

Storage: fdmi.lhs(n+1, 5+1), fdmi.rhs(n+1, 3+1) 1-based padded;
rhs_b(1:5, 0:7) as [6][8]; rhs_t(0:4, 8) as [5][9].
Line data: (n, nlines, M) arrays (Fortran f(nlines, n) transposed, M modes at once).
"""
import numpy as np

from .fdm import (BCS_MIN, BCS_MAX, BCS_BOTH, fdm_bcs_reduce, tridfs, tridss, pentadfs, pentadss,
                  matmul_3d, matmul_5d)


class Integral:
    """fdm_integral_dt (fdm_integral.f90:18-26)."""

    def __init__(self):
        self.mode_fdm = 0
        self.lam = 0.0
        self.bc = 0
        self.rhs_b = np.zeros((6, 8))
        self.rhs_t = np.zeros((5, 9))
        self.lhs = None
        self.rhs = None


def int1_create_system(g, lam, ibc):
    """fdm_integral.f90:91-214; g is a fdm.Derivative (first derivative plan).

    lam may be a scalar or a vector of M eigenvalues; every lambda-dependent array
    carries a trailing axis of length M (1 for a scalar) so that the statements
    below are the reference's, broadcast over modes."""
    fdmi = Integral()
    lam = np.atleast_1d(np.asarray(lam, dtype=np.float64))
    M = lam.shape[0]
    ndl, ndr = g.nb_diag
    idl = ndl // 2 + 1
    idr = ndr // 2 + 1
    nx = g.size
    assert abs(idl - idr) <= 1
    fdmi.mode_fdm = g.mode_fdm
    fdmi.lam = lam
    fdmi.bc = ibc
    glhs = g.lhs[:, :ndl + 1, None]
    grhs = g.rhs[:, :ndr + 1, None]
    lhs = np.zeros((nx + 1, ndr + 1, M))
    rhs = np.zeros((nx + 1, ndl + 1, 1))
    fdmi.rhs_b = np.zeros((6, 8, M))
    fdmi.rhs_t = np.zeros((5, 9, M))

    # new rhs diagonals (array A), independent of lambda
    rhs[:, 1:] = glhs[:, 1:ndl + 1]
    rhsr_b = np.zeros((6, 8, 1))
    rhsr_t = np.zeros((5, 9, 1))
    fdm_bcs_reduce(ibc, rhs, grhs, rhsr_b, rhsr_t)

    rhs_b = fdmi.rhs_b
    rhs_t = fdmi.rhs_t
    if ibc == BCS_MIN:
        rhs_b[1:idl + 2, 1:ndl + 1] = rhs[1:idl + 2, 1:ndl + 1]
        for ir in range(1, idr):
            rhs_b[1 + ir, idl - ir] = -rhsr_b[1 + ir, idr - ir]
    elif ibc == BCS_MAX:
        rhs_t[0:idl + 1, 1:ndl + 1] = rhs[nx - idl:nx + 1, 1:ndl + 1]
        for ir in range(1, idr):
            rhs_t[idl - ir, idl + ir] = -rhsr_t[idr - ir, idr + ir]

    # new lhs diagonals (array C = B + h lambda A), dependent on lambda
    lhs[:, 1:] = grhs[:, 1:ndr + 1]
    lhs[1:, idr] = lhs[1:, idr] + lam * glhs[1:, idl]
    for i in range(1, idl):
        lhs[1 + i:nx + 1, idr - i] = lhs[1 + i:nx + 1, idr - i] + lam * glhs[1 + i:nx + 1, idl - i]
        lhs[1:nx - i + 1, idr + i] = lhs[1:nx - i + 1, idr + i] + lam * glhs[1:nx - i + 1, idl + i]

    if ibc == BCS_MIN:
        lhs[1:idr + 1, 1:ndr + 1] = rhsr_b[1:idr + 1, 1:ndr + 1]
        lhs[1, idr + 1:idr + idl] = lhs[1, idr + 1:idr + idl] - lam * rhs_b[1, idl + 1:ndl + 1]
        for ir in range(1, idr):
            lhs[1 + ir, idr - idl + 1:idr + idl] = lhs[1 + ir, idr - idl + 1:idr + idl] + lam * rhs_b[1 + ir, 1:ndl + 1]
    elif ibc == BCS_MAX:
        lhs[nx - idr + 1:nx + 1, 1:ndr + 1] = rhsr_t[1:idr + 1, 1:ndr + 1]
        lhs[nx, idr - idl + 1:idr] = lhs[nx, idr - idl + 1:idr] - lam * rhs_t[idl, 1:idl]
        for ir in range(1, idr):
            lhs[nx - ir, idr - idl + 1:idr + idl] = lhs[nx - ir, idr - idl + 1:idr + idl] + lam * rhs_t[idl - ir, 1:ndl + 1]

    # normalization such that new central diagonal in rhs is 1
    m = max(idr, idl + 1)
    for ir in range(1, m + 1):
        dummy = 1.0 / rhs[ir, idl]
        rhs_b[ir, 0:ndl + 1] = rhs_b[ir, 0:ndl + 1] * dummy
        dummy = 1.0 / rhs[nx - ir + 1, idl]
        rhs_t[idl - ir + 1, 1:ndl + 2] = rhs_t[idl - ir + 1, 1:ndl + 2] * dummy
        dummy = 1.0 / rhs[ir, idl]
        rhs[ir, 1:ndl + 1] = rhs[ir, 1:ndl + 1] * dummy
        lhs[ir, 1:ndr + 1] = lhs[ir, 1:ndr + 1] * dummy
        dummy = 1.0 / rhs[nx - ir + 1, idl]
        rhs[nx - ir + 1, 1:ndl + 1] = rhs[nx - ir + 1, 1:ndl + 1] * dummy
        lhs[nx - ir + 1, 1:ndr + 1] = lhs[nx - ir + 1, 1:ndr + 1] * dummy

    # interior points: normalization such that 1. upper-diagonal is 1
    for ir in range(m + 1, nx - m + 1):
        dummy = 1.0 / rhs[ir, idl + 1]
        rhs[ir, 1:ndl + 1] = rhs[ir, 1:ndl + 1] * dummy
        lhs[ir, 1:ndr + 1] = lhs[ir, 1:ndr + 1] * dummy

    # reducing system in the opposite end to account for the case of extended stencils
    if ibc == BCS_MIN:
        fdm_bcs_reduce(BCS_MAX, lhs, rhs, rhs_t=rhs_t)
    elif ibc == BCS_MAX:
        fdm_bcs_reduce(BCS_MIN, lhs, rhs, rhs_b=rhs_b)

    fdmi.lhs = lhs
    fdmi.rhs = rhs
    return fdmi


def int1_initialize(g, lam, ibc):
    """fdm_integral.f90:58-87."""
    fdmi = int1_create_system(g, lam, ibc)
    nx = fdmi.lhs.shape[0] - 1
    nd = fdmi.lhs.shape[1] - 1
    cols = [fdmi.lhs[2:nx, k] for k in range(1, nd + 1)]   # rows 2..nx-1
    if nd == 3:
        tridfs(*cols)
    elif nd == 5:
        pentadfs(*cols)
    else:
        raise NotImplementedError
    return fdmi


def int1_solve(fdmi, rhsi, f, result, want_du=False):
    """fdm_integral.f90:219-314.  f, result: (n, nlines); result holds the bc on entry.
    Returns du_boundary (nlines,) when want_du."""
    nx = fdmi.lhs.shape[0] - 1
    ndl = fdmi.lhs.shape[1] - 1
    idl = ndl // 2 + 1
    ndr = rhsi.shape[1] - 1
    idr = ndr // 2 + 1

    if fdmi.bc == BCS_MIN:
        result[nx - 1] = f[nx - 1]
    elif fdmi.bc == BCS_MAX:
        result[0] = f[0]

    if ndr == 3:
        bcs_b, bcs_t = matmul_3d(rhsi, f, result, BCS_BOTH, fdmi.rhs_b, fdmi.rhs_t, want_bcs=True)
    else:
        bcs_b, bcs_t = matmul_5d(rhsi, f, result, BCS_BOTH, fdmi.rhs_b, fdmi.rhs_t, want_bcs=True)

    cols = [fdmi.lhs[2:nx, k] for k in range(1, ndl + 1)]
    if ndl == 3:
        tridss(*cols, result[1:nx - 1])
    else:
        pentadss(*cols, result[1:nx - 1])

    L = fdmi.lhs
    du = None

    def R(n):
        return result[n - 1]

    def F(n):
        return f[n - 1]

    if fdmi.bc == BCS_MAX:
        r1 = bcs_b
        for ic in range(1, idl):
            r1 = r1 + L[1, idl + ic] * R(1 + ic)
        ic = idl
        r1 = r1 + L[1, 1] * R(1 + ic)
        result[0] = r1
        if want_du:
            du = L[nx, idl] * R(nx)
            for ic in range(1, idl):
                du = du + L[nx, idl - ic] * R(nx - ic)
            ic = idl
            du = du + L[nx, ndl] * R(nx - ic)
            for ic in range(1, idr):
                du = du + rhsi[nx, idr - ic] * F(nx - ic)

    if fdmi.bc == BCS_MIN:
        rn = bcs_t
        for ic in range(1, idl):
            rn = rn + L[nx, idl - ic] * R(nx - ic)
        ic = idl
        rn = rn + L[nx, ndl] * R(nx - ic)
        result[nx - 1] = rn
        if want_du:
            du = L[1, idl] * R(1)
            for ic in range(1, idl):
                du = du + L[1, idl + ic] * R(1 + ic)
            ic = idl
            du = du + L[1, 1] * R(1 + ic)
            for ic in range(1, idr):
                du = du + rhsi[1, idr + ic] * F(1 + ic)
    return du


# ###########################################################################
# opr_odes.f90.  fdmi = {BCS_MIN: Integral, BCS_MAX: Integral}; line data are
# (n, nlines, M) with M the number of modes solved at once (see int1_create_system).
def ode2_factorize_dn_sing(fdmi, f, bcs):
    """opr_odes.f90:37-96.  f(n, nlines, M) (modified), bcs(2, nlines, M) -> u, v."""
    nx = fdmi[BCS_MIN].lhs.shape[0] - 1
    M = f.shape[2]
    u = np.zeros_like(f)
    v = np.zeros_like(f)
    f[0] = 0.0
    v[nx - 1] = bcs[1]
    int1_solve(fdmi[BCS_MAX], fdmi[BCS_MAX].rhs, f, v)
    f1 = np.zeros((nx, 1, M))
    f1[0] = 1.0
    v1 = np.zeros((nx, 1, M))
    int1_solve(fdmi[BCS_MAX], fdmi[BCS_MAX].rhs, f1, v1)
    u[0] = bcs[0]
    du0_n = int1_solve(fdmi[BCS_MIN], fdmi[BCS_MIN].rhs, v, u, want_du=True)
    u1 = np.zeros((nx, 1, M))
    du1_n = int1_solve(fdmi[BCS_MIN], fdmi[BCS_MIN].rhs, v1, u1, want_du=True)
    ff = 1.0 / (du1_n[0] - v1[0, 0])
    du0_n = (v[0] - du0_n) * ff
    for i in range(nx):
        u[i] = u[i] + du0_n * u1[i, 0]
        v[i] = v[i] + du0_n * v1[i, 0]
    return u, v


def ode2_factorize_nn_sing(fdmi, f, bcs):
    """opr_odes.f90:165-183."""
    bcs = bcs.copy()
    bcs[0] = 0.0
    return ode2_factorize_dn_sing(fdmi, f, bcs)


def ode2_factorize_nn(fdmi, rhsi_b, rhsi_t, f, bcs):
    """opr_odes.f90:265-386.  f(n, nlines, M) (modified), bcs(2, nlines, M) -> u, v."""
    lam = fdmi[BCS_MIN].lam
    nx = fdmi[BCS_MIN].lhs.shape[0] - 1
    M = f.shape[2]
    u = np.zeros_like(f)
    v = np.zeros_like(f)

    f[nx - 1] = 0.0
    v[0] = 0.0
    int1_solve(fdmi[BCS_MIN], rhsi_b, f, v)

    # v^(1), e^(-), dd: wrk1d(1:3, :, 2); f1 = wrk1d(1:3, :, 1)
    w1 = np.zeros((nx, 3, M))
    w2 = np.zeros((nx, 3, M))
    w1[nx - 1, 0] = 1.0     # f1(1, nx)
    w2[0, 0] = 0.0          # v1(1)
    w2[0, 1] = 1.0          # em(1)
    w2[0, 2] = 0.0          # dd(1)
    int1_solve(fdmi[BCS_MIN], rhsi_b, w1, w2)
    v1 = w2[:, 0]
    em = w2[:, 1]

    u[nx - 1] = 0.0
    du0_n = int1_solve(fdmi[BCS_MAX], rhsi_t, v, u, want_du=True)

    # u^(1), s^(+), e^(+) overwrite wrk1d(1:3, :, 1); forcing is (v1, em, dd = 0)
    w1[nx - 1, 0] = 0.0     # u1(nx)
    w1[nx - 1, 1] = 0.0     # sp(nx)
    w2[:, 2] = 0.0          # dd(:)
    w1[nx - 1, 2] = 1.0     # ep(nx)
    der_bcs = int1_solve(fdmi[BCS_MAX], rhsi_t, w2, w1, want_du=True)
    u1 = w1[:, 0]
    sp = w1[:, 1]
    ep = w1[:, 2]
    du1_n, dsp_n, dep_n = der_bcs[0], der_bcs[1], der_bcs[2]

    a = np.zeros((4, 4, M))
    a[1, 1] = 1.0 + lam * sp[0]
    a[2, 1] = em[nx - 1]
    a[3, 1] = dsp_n
    a[1, 2] = lam * ep[0]
    a[2, 2] = lam
    a[3, 2] = dep_n
    a[1, 3] = lam * u1[0]
    a[2, 3] = v1[nx - 1]
    a[3, 3] = du1_n
    # LU decomposition
    a[1, 2] = a[1, 2] / a[1, 1]
    a[2, 2] = a[2, 2] - a[2, 1] * a[1, 2]
    a[3, 2] = a[3, 2] - a[3, 1] * a[1, 2]
    a[1, 3] = a[1, 3] / a[1, 1]
    a[2, 3] = (a[2, 3] - a[2, 1] * a[1, 3]) / a[2, 2]
    a[3, 3] = a[3, 3] - a[3, 1] * a[1, 3] - a[3, 2] * a[2, 3]
    # Solution
    v[0] = (bcs[0] - lam * u[0]) / a[1, 1]
    u[nx - 1] = (bcs[1] - v[nx - 1] - a[2, 1] * v[0]) / a[2, 2]
    fn = (bcs[1] - du0_n - a[3, 1] * v[0] - a[3, 2] * u[nx - 1]) / a[3, 3]
    u[nx - 1] = u[nx - 1] - a[2, 3] * fn
    v[0] = v[0] - a[1, 2] * u[nx - 1] - a[1, 3] * fn
    # Result
    i = nx - 1
    v[i] = v[i] + fn * v1[i] + v[0] * em[i] + lam * u[i]
    for i in range(nx - 2, 0, -1):
        u[i] = u[i] + fn * u1[i] + v[0] * sp[i] + u[nx - 1] * ep[i]
        v[i] = v[i] + fn * v1[i] + v[0] * em[i] + lam * u[i]
    i = 0
    u[i] = u[i] + fn * u1[i] + v[0] * sp[i] + u[nx - 1] * ep[i]
    v[i] = v[i] + lam * u[i]
    return u, v


def ode2_factorize_dd_sing(fdmi, f, bcs):
    """opr_odes.f90:188-260 (Dirichlet / Dirichlet, lambda = 0).  f(n, nlines, M) (modified), bcs(2, nlines, M) -> u, v."""
    nx = fdmi[BCS_MIN].lhs.shape[0] - 1
    M = f.shape[2]
    u = np.zeros_like(f)
    v = np.zeros_like(f)
    # v^(0) in v' = f, v_1 given (0 for now, to be found later on)
    f[nx - 1] = 0.0
    v[0] = 0.0
    int1_solve(fdmi[BCS_MIN], fdmi[BCS_MIN].rhs, f, v)
    # v^(1)
    f1 = np.zeros((nx, 1, M))
    f1[nx - 1] = 1.0
    v1 = np.zeros((nx, 1, M))
    int1_solve(fdmi[BCS_MIN], fdmi[BCS_MIN].rhs, f1, v1)
    # u^(0) in u' = v, u_n given
    u[nx - 1] = bcs[1]
    du0_n = int1_solve(fdmi[BCS_MAX], fdmi[BCS_MAX].rhs, v, u, want_du=True)
    # u^(1)
    u1 = np.zeros((nx, 1, M))
    du1_n = int1_solve(fdmi[BCS_MAX], fdmi[BCS_MAX].rhs, v1, u1, want_du=True)
    # s^(+): the node displacements x(:) - x_n
    f1[:] = 1.0
    sp = np.zeros((nx, 1, M))
    int1_solve(fdmi[BCS_MAX], fdmi[BCS_MAX].rhs, f1, sp)
    # constraint
    fn = 1.0 / (du1_n[0] - v1[nx - 1, 0])
    du0_n = (v[nx - 1] - du0_n) * fn
    # contribution from v_1 to satisfy the bc at the bottom
    dummy = 1.0 / sp[0, 0]
    v[0] = (bcs[0] - (u[0] + du0_n * u1[0, 0])) * dummy
    u[0] = bcs[0]
    for i in range(1, nx):
        u[i] = u[i] + du0_n * u1[i, 0] + v[0] * sp[i, 0]
        v[i] = v[i] + du0_n * v1[i, 0] + v[0]
    return u, v


def ode2_factorize_dd(fdmi, rhsi_b, rhsi_t, f, bcs):
    """opr_odes.f90:391-478 (Dirichlet / Dirichlet).  f(n, nlines, M) (modified), bcs(2, nlines, M) -> u, v."""
    lam = fdmi[BCS_MIN].lam
    nx = fdmi[BCS_MIN].lhs.shape[0] - 1
    M = f.shape[2]
    u = np.zeros_like(f)
    v = np.zeros_like(f)
    # v^(0) in v' + lambda v = f, v_1 given (0 for now, to be found later on)
    f[nx - 1] = 0.0
    v[0] = 0.0
    int1_solve(fdmi[BCS_MIN], rhsi_b, f, v)
    # v^(1) and e^(-)
    w1 = np.zeros((nx, 2, M))
    w2 = np.zeros((nx, 2, M))
    w1[nx - 1, 0] = 1.0      # f1(1, nx)
    w2[0, 0] = 0.0           # v1(1)
    w2[0, 1] = 1.0           # em(1)
    int1_solve(fdmi[BCS_MIN], rhsi_b, w1, w2)
    v1, em = w2[:, 0].copy(), w2[:, 1].copy()
    # u^(0) in u' - lambda u = v, u_n given
    u[nx - 1] = bcs[1]
    du0_n = int1_solve(fdmi[BCS_MAX], rhsi_t, v, u, want_du=True)
    # u^(1) and s^(+): forcing (v1, em), both zero at the top
    w1[nx - 1, 0] = 0.0      # u1(nx)
    w1[nx - 1, 1] = 0.0      # sp(nx)
    der_bcs = int1_solve(fdmi[BCS_MAX], rhsi_t, w2, w1, want_du=True)
    u1, sp = w1[:, 0], w1[:, 1]
    du1_n, dsp_n = der_bcs[0], der_bcs[1]
    # constraint and bottom boundary condition
    aa = du1_n - v1[nx - 1]
    bb = dsp_n - em[nx - 1]
    dummy = 1.0 / (aa * sp[0] - bb * u1[0])
    t = lam * bcs[1] - du0_n + v[nx - 1]
    v0 = (aa * (bcs[0] - u[0]) - u1[0] * t) * dummy          # q1
    fn = (sp[0] * t - bb * (bcs[0] - u[0])) * dummy
    v[0] = v0
    for i in range(nx - 1, 0, -1):
        u[i] = u[i] + fn * u1[i] + v[0] * sp[i]
        v[i] = v[i] + fn * v1[i] + v[0] * em[i] + lam * u[i]
    u[0] = bcs[0]
    v[0] = v[0] + lam * u[0]
    return u, v
