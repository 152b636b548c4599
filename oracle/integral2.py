"""Oracle: second-order integral operator in y and the direct Poisson solver built on it (second slice).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  ORACLE ONLY: the CUDA path of this solver variant is not built yet;
this restatement and its tests are the first gate for it.  Parity pinning: the reference holds no
golden vector for this path (valid/elliptic/vpoisson.f90 prints round-trip norms, no stored output), so the restatement is
pinned on the identities it must satisfy -- the discrete equation (B2 - lambda A2) p = A2 f of the direct second derivative in
the interior rows, the biased wall formula of the Neumann rows and the prescribed boundary data (tests/test_integral2_cpu.py).

Follows /root/reference/src:
  fdm/fdm_integral.f90        FDM_Int2_Initialize (:334-364), FDM_Int2_CreateSystem (:369-548, coef_c1n4_biased :561-618),
                              FDM_Int2_Solve (:627-673)
  operators/opr_elliptic.f90  OPR_Elliptic_Initialize, TYPE_DIRECT branch (:150-160, :221-240),
                              OPR_Poisson_FourierXZ_Direct (:368-455)

Storage as in oracle/integral.py: lhs(n+1, ndr+1, M), rhs(n+1, ndl+1, 1) 1-based padded, rhs_b(1:5, 0:7) as [6][8][M],
rhs_t(0:4, 1:8) as [5][9][M]; M eigenvalues at once on the trailing axis.  Line data (n, nlines[, M]).
"""
import numpy as np

from .fdm import (BCS_DD, BCS_DN, BCS_ND, BCS_NN, BCS_BOTH, fdm_bcs_reduce, tridfs, tridss, pentadfs, pentadss,
                  matmul_3d, matmul_5d)
from .fdm_direct import Pi, Pi_p, Pi_pp_3, Lag, Lag_p, _pad
from .integral import Integral


def coef_c1n4_biased(x, i, backwards=False):
    """fdm_integral.f90:561-618: p'_1 = b_1 p_1 + b_2 p_2 + b_3 p_3 + b_4 p_4 + a_2 p''_2 (x 1-based padded)."""
    i1 = i
    if backwards:
        i2, i3, i4 = i - 1, i - 2, i - 3
    else:
        i2, i3, i4 = i + 1, i + 2, i + 3
    dx1 = x[i2] - x[i1]
    dx3 = x[i2] - x[i3]
    dx4 = x[i2] - x[i4]
    set_m = [i1, i3, i4]
    a2 = 0.5 * (Pi(x, i1, set_m) - dx1 * Pi_p(x, i1, set_m)) / Pi_p(x, i2, set_m)
    b2 = (Pi_p(x, i1, set_m) * (2.0 * Pi_p(x, i2, set_m) + dx1 * Pi_pp_3(x, i2, set_m))
          - Pi(x, i1, set_m) * Pi_pp_3(x, i2, set_m))
    b2 = 0.5 * b2 / Pi(x, i2, set_m) / Pi_p(x, i2, set_m)
    D = Lag(x, i2, i1, set_m) + dx1 * Lag_p(x, i2, i1, set_m)
    b1 = (Lag(x, i1, i1, set_m) * (Lag(x, i2, i1, set_m) + 2 * dx1 * Lag_p(x, i2, i1, set_m))
          - dx1 * Lag_p(x, i1, i1, set_m) * (Lag(x, i2, i1, set_m) + dx1 * Lag_p(x, i2, i1, set_m)))
    b1 = -b1 / dx1 / D
    D = Lag(x, i2, i3, set_m) + dx3 * Lag_p(x, i2, i3, set_m)
    b3 = (Lag(x, i1, i3, set_m) * (Lag(x, i2, i3, set_m) + 2 * dx1 * Lag_p(x, i2, i3, set_m))
          - dx1 * Lag_p(x, i1, i3, set_m) * (Lag(x, i2, i3, set_m) + dx1 * Lag_p(x, i2, i3, set_m)))
    b3 = -b3 / dx3 / D
    D = Lag(x, i2, i4, set_m) + dx4 * Lag_p(x, i2, i4, set_m)
    b4 = (Lag(x, i1, i4, set_m) * (Lag(x, i2, i4, set_m) + 2 * dx1 * Lag_p(x, i2, i4, set_m))
          - dx1 * Lag_p(x, i1, i4, set_m) * (Lag(x, i2, i4, set_m) + dx1 * Lag_p(x, i2, i4, set_m)))
    b4 = -b4 / dx4 / D
    return np.array([0.0, b1, b2, b3, b4, a2])          # coef(1:5) at [1:6]


def int2_create_system(nodes, g, lam2, ibc):
    """fdm_integral.f90:369-548; g is the fdm.Derivative of the second derivative to be inverted."""
    fdmi = Integral()
    lam2 = np.atleast_1d(np.asarray(lam2, dtype=np.float64))
    M = lam2.shape[0]
    x = _pad(np.asarray(nodes, dtype=np.float64))
    ndl, ndr = g.nb_diag
    idl = ndl // 2 + 1
    idr = ndr // 2 + 1
    nx = g.size
    assert abs(idl - idr) <= 1
    fdmi.mode_fdm = g.mode_fdm
    fdmi.lam = lam2
    fdmi.bc = ibc
    glhs = g.lhs[:, :ndl + 1, None]
    grhs = g.rhs[:, :ndr + 1, None]
    lhs = np.zeros((nx + 1, ndr + 1, M))
    rhs = np.zeros((nx + 1, ndl + 1, 1))
    rhs_b = fdmi.rhs_b = np.zeros((6, 8, M))
    rhs_t = fdmi.rhs_t = np.zeros((5, 9, M))

    # new rhs diagonals (array A22R), independent of lambda
    rhs[:, 1:] = glhs[:, 1:ndl + 1]
    rhsr_b = np.zeros((6, 8, 1))
    rhsr_t = np.zeros((5, 9, 1))
    fdm_bcs_reduce(BCS_BOTH, rhs, grhs, rhsr_b, rhsr_t)

    rhs_b[1:idl + 2, 1:ndl + 1] = rhs[1:idl + 2, 1:ndl + 1]
    for ir in range(1, idr):                    # change sign in b^R_{21} for nonzero bc
        rhs_b[1 + ir, idl - ir] = -rhsr_b[1 + ir, idr - ir]
    rhs_t[0:idl + 1, 1:ndl + 1] = rhs[nx - idl:nx + 1, 1:ndl + 1]
    for ir in range(1, idr):                    # change sign in b^R_{2n} for nonzero bc
        rhs_t[idl - ir, idl + ir] = -rhsr_t[idr - ir, idr + ir]

    # new lhs diagonals (array C22R); the rhs centre diagonal is not saved because it was 1
    lhs[:, 1:] = grhs[:, 1:ndr + 1]
    lhs[1:, idr] = lhs[1:, idr] - lam2 * glhs[1:, idl]
    for i in range(1, idl):
        lhs[1 + i:nx + 1, idr - i] = lhs[1 + i:nx + 1, idr - i] - lam2 * glhs[1 + i:nx + 1, idl - i]
        lhs[1:nx - i + 1, idr + i] = lhs[1:nx - i + 1, idr + i] - lam2 * glhs[1:nx - i + 1, idl + i]

    lhs[2:idr + 1, 1:ndr + 1] = rhsr_b[2:idr + 1, 1:ndr + 1]
    for ir in range(1, idr):
        lhs[1 + ir, idr - idl + 1:idr + idl] = lhs[1 + ir, idr - idl + 1:idr + idl] - lam2 * rhs_b[1 + ir, 1:ndl + 1]
    lhs[nx - idr + 1:nx, 1:ndr + 1] = rhsr_t[1:idr, 1:ndr + 1]
    for ir in range(1, idr):
        lhs[nx - ir, idr - idl + 1:idr + idl] = lhs[nx - ir, idr - idl + 1:idr + idl] - lam2 * rhs_t[idl - ir, 1:ndl + 1]

    # corrections to BCS_DD to account for Neumann, fourth-order formula for the derivative at the boundary
    if ibc in (BCS_ND, BCS_NN):
        coef = coef_c1n4_biased(x, 1)
        lhs[1, :] = 0.0
        lhs[1, 1:4] = (-coef[2:5] / coef[1])[:, None]          # vector d_2
        rhs_b[1, :] = 0.0
        rhs_b[1, idl] = 1.0 / coef[1]                           # coefficient d_1
        rhs_b[1, idl + 1] = -coef[5] / coef[1]                  # vector e_2, only 1 component
        lhs[1, 1] = lhs[1, 1] + lam2 * rhs_b[1, idl + 1]        # d + lambda^2 h^2 e
        for ir in range(1, idr):
            c0 = idr - ir + 1
            lhs[1 + ir, c0:c0 + 3] = lhs[1 + ir, c0:c0 + 3] - rhs_b[1 + ir, idl - ir] * lhs[1, 1:4]        # reduced C matrix
            rhs_b[1 + ir, idl - ir + 1] = rhs_b[1 + ir, idl - ir + 1] + rhs_b[1 + ir, idl - ir] * rhs_b[1, idl + 1]   # reduced A
            rhs_b[1 + ir, idl - ir] = rhs_b[1 + ir, idl - ir] * rhs_b[1, idl]                               # d_1 b^R_{21}

    if ibc in (BCS_DN, BCS_NN):
        coef = coef_c1n4_biased(x, nx, backwards=True)
        lhs[nx, :] = 0.0
        lhs[nx, ndr - 2:ndr + 1] = (-coef[[4, 3, 2]] / coef[1])[:, None]     # vector d_n-1
        rhs_t[idl, :] = 0.0
        rhs_t[idl, idl] = 1.0 / coef[1]                         # coefficient d_n
        rhs_t[idl, idl - 1] = -coef[5] / coef[1]                # vector e_n-1, only 1 component
        lhs[nx, ndr] = lhs[nx, ndr] + lam2 * rhs_t[idl, idl - 1]
        for ir in range(1, idr):
            c0 = ir
            lhs[nx - ir, c0:c0 + 3] = lhs[nx - ir, c0:c0 + 3] - rhs_t[idl - ir, idl + ir] * lhs[nx, ndr - 2:ndr + 1]   # reduced C
            rhs_t[idl - ir, idl + ir - 1] = rhs_t[idl - ir, idl + ir - 1] + rhs_t[idl - ir, idl + ir] * rhs_t[idl, idl - 1]
            rhs_t[idl - ir, idl + ir] = rhs_t[idl - ir, idl + ir] * rhs_t[idl, idl]                          # d_n b^R_{2n}

    # normalization such that the new central diagonal in rhs is 1
    m = max(idr, idl + 1)
    for ir in range(2, m + 1):
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

    # interior points: normalization such that the first upper diagonal is 1
    for ir in range(m + 1, nx - m + 1):
        dummy = 1.0 / rhs[ir, idl + 1]
        rhs[ir, 1:ndl + 1] = rhs[ir, 1:ndl + 1] * dummy
        lhs[ir, 1:ndr + 1] = lhs[ir, 1:ndr + 1] * dummy

    fdmi.lhs = lhs
    fdmi.rhs = rhs
    return fdmi


def int2_initialize(nodes, g, lam2, ibc):
    """fdm_integral.f90:334-364: LU of rows 2 .. nx-1."""
    fdmi = int2_create_system(nodes, g, lam2, ibc)
    nx = fdmi.lhs.shape[0] - 1
    nd = fdmi.lhs.shape[1] - 1
    cols = [fdmi.lhs[2:nx, k] for k in range(1, nd + 1)]
    if nd == 3:
        tridfs(*cols)
    elif nd == 5:
        pentadfs(*cols)
    else:
        raise NotImplementedError
    return fdmi


def int2_solve(fdmi, rhsi, f, result):
    """fdm_integral.f90:627-673.  f, result: (n, nlines, M) as in integral.int1_solve; result holds the boundary data in rows
    1 and n on entry (the values for Dirichlet ends, the derivatives for Neumann ends)."""
    nx = fdmi.lhs.shape[0] - 1
    ndl = fdmi.lhs.shape[1] - 1
    ndr = rhsi.shape[1] - 1
    if ndr == 3:
        bcs_b, bcs_t = matmul_3d(rhsi, f, result, BCS_BOTH, fdmi.rhs_b, fdmi.rhs_t, want_bcs=True)
    elif ndr == 5:
        bcs_b, bcs_t = matmul_5d(rhsi, f, result, BCS_BOTH, fdmi.rhs_b, fdmi.rhs_t, want_bcs=True)
    else:
        raise NotImplementedError
    L = fdmi.lhs
    cols = [L[2:nx, k] for k in range(1, ndl + 1)]
    if ndl == 3:
        tridss(*cols, result[1:nx - 1])
    elif ndl == 5:
        pentadss(*cols, result[1:nx - 1])
    else:
        raise NotImplementedError
    # corrections to BCS_DD to account for Neumann
    if fdmi.bc in (BCS_ND, BCS_NN):
        result[0] = bcs_b + L[1, 1] * result[1] + L[1, 2] * result[2] + L[1, 3] * result[3]
    if fdmi.bc in (BCS_DN, BCS_NN):
        result[nx - 1] = bcs_t + L[nx, ndl] * result[nx - 2] + L[nx, ndl - 1] * result[nx - 3] + L[nx, ndl - 2] * result[nx - 4]
    return result


# ###########################################################################
class EllipticDirect:
    """OPR_Elliptic_Initialize, TYPE_DIRECT (EllipticOrder = CompactDirect6; opr_elliptic.f90:111-125,150-160,221-240), serial:
    lambda(k, i) = mwn2_x(i) + mwn2_z(k) from the SECOND-derivative plans, one FDM_Int2 system per mode with BCS_NN (BCS_DN for
    the single singular mode (1, 1): second-order FDMs are non-zero at Nyquist), the y plan built with the direct scheme."""

    def __init__(self, g, y_nodes):
        from . import fdm
        gx, gy, gz = g
        self.g = g
        self.nx, self.ny, self.nz = gx.size, gy.size, gz.size
        self.isize_line = self.nx // 2 + 1
        self.norm = 1.0 / float(gx.size * gz.size)
        self.y = np.asarray(y_nodes, dtype=np.float64)
        self.fdm_loc = fdm.Plan(self.y, False, False, name="y_elliptic", mode2=fdm.FDM_COM6_DIRECT)
        lam = np.zeros((self.nz, self.isize_line))
        for i in range(self.isize_line):
            for k in range(self.nz):
                lam[k, i] = gx.der2.mwn[i] + (gz.der2.mwn[k] if gz.size > 1 else 0.0)
        self.lam = lam

    def is_sing(self):
        m = np.zeros((self.nz, self.isize_line), dtype=bool)
        m[0, 0] = True
        return m


def opr_poisson_direct(ell, p, bcs_hb, bcs_ht, want_dpdy=True):
    """OPR_Poisson_FourierXZ_Direct (opr_elliptic.f90:368-455), BCS_NN.  p(nz, ny, nx) forcing; bcs_hb, bcs_ht (nz, nx) the
    wall-normal derivatives.  Returns (p, dpdy) with dpdy = OPR_Partial_Y(OPR_P1, p) on the flow's own y plan."""
    from . import operators as O
    nz, ny, nx = p.shape
    p = p.copy()
    p[:, 0, :] = bcs_hb
    p[:, ny - 1, :] = bcs_ht
    c = np.fft.rfft(p, axis=2)
    if nz > 1:
        c = np.fft.fft(c, axis=0)
    c = c * ell.norm
    der2 = ell.fdm_loc.der2
    sing = ell.is_sing()
    out = np.zeros_like(c)
    for mask, singular in ((~sing, False), (sing, True)):
        kk, ii = np.nonzero(mask)
        if len(kk) == 0:
            continue
        lam = ell.lam[kk, ii]
        M = len(kk)
        fdmi = int2_initialize(ell.y, der2, lam, BCS_DN if singular else BCS_NN)
        cm = c[kk, :, ii]
        f = np.zeros((ny, 2, M))
        f[:, 0, :] = cm.real.T
        f[:, 1, :] = cm.imag.T
        u = np.zeros_like(f)
        u[0] = f[0]                                 # bottom boundary conditions
        u[ny - 1] = f[ny - 1]                       # top boundary conditions
        if singular:
            u[0] = 0.0                              # compatibility constraint: the reference value of p at the bottom is zero
        int2_solve(fdmi, fdmi.rhs, f, u)
        out[kk, :, ii] = (u[:, 0, :] + 1j * u[:, 1, :]).T
    if nz > 1:
        out = np.fft.ifft(out, axis=0, norm="forward")
    sol = np.fft.irfft(out, n=nx, axis=2, norm="forward")
    if not want_dpdy:
        return sol, None
    bcs_p = [[0, 0], [0, 0]]
    return sol, O.opr_partial(1, O.OPR_P1, bcs_p, ell.g[1], sol)
