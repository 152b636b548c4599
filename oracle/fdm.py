"""Oracle: compact finite-difference plans and line solves (numpy restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows, routine by routine, /root/reference/src:
  fdm/fdm_com1_jacobian.f90   Create_System_1der, FDM_C1N4/C1N6[_Penta]_Jacobian
  fdm/fdm_com2_jacobian.f90   Create_System_2der, FDM_C2N4/C2N6/C2N6_Hyper_Jacobian
  fdm/fdm_base.f90            FDM_Bcs_Neumann (:194-300), FDM_Bcs_Reduce (:304-391)
  utils/linear3.f90           TRIDFS/TRIDSS (:29-150), TRIDPFS/TRIDPSS (:269-442)
  utils/linear5.f90           PENTADFS/PENTADSS (:30-131), PENTADFS2/SS2 (:156-244),
                              PENTADPFS/PENTADPSS (:273-411)
  fdm/fdm_matmul.f90          MatMul_* (:70-642)
  fdm/fdm_derivative.f90      FDM_Der1/Der2_{Initialize,CreateSystem,Solve}
  fdm/fdm.f90                 FDM_CreatePlan (:143-252)

Index convention: coefficient tables are stored "Fortran-shaped" with a dummy
leading row/column so that the 1-based (or 0-based, for rhs_b(:,0:) /
rhs_t(0:,:)) indices of the reference can be used verbatim.  Line data are
numpy arrays of shape (n, nlines): first index is the position along the line
(0-based), second the line -- i.e. the Fortran u(nlines, n) transposed, so that
every statement below is a whole-vector operation over lines, like the
reference's inner loops.
"""
import numpy as np

pi_wp = 3.14159265358979323846

BCS_PERIODIC = -1
BCS_DD, BCS_ND, BCS_DN, BCS_NN = 0, 1, 2, 3
BCS_NONE, BCS_MIN, BCS_MAX, BCS_BOTH = 0, 1, 2, 3

FDM_COM4_JACOBIAN = 4
FDM_COM6_JACOBIAN_PENTA = 5
FDM_COM6_JACOBIAN = 6
FDM_COM6_JACOBIAN_HYPER = 7
FDM_COM6_DIRECT = 16          # fdm_derivative.f90:57


def _cshift(a1, s):
    """Fortran cshift on a 1-based padded vector (index 0 unused)."""
    out = np.zeros_like(a1)
    out[1:] = np.roll(a1[1:], -s)
    return out


def _c(a, like):
    """Append singleton axes so a per-row coefficient broadcasts over line data."""
    return a.reshape(a.shape + (1,) * (like.ndim - a.ndim))


def _pad1(v):
    out = np.zeros(len(v) + 1)
    out[1:] = v
    return out


# ###########################################################################
# fdm_com1_jacobian.f90:195-291
def create_system_1der(dx, ndl, ndr, coef_int, bc1=None, bc2=None, bc3=None):
    """dx: 1-based padded Jacobian.  Returns lhs(n+1, ndl+1), rhs(n+1, ndr+1)."""
    nx = len(dx) - 1
    lhs = np.zeros((nx + 1, ndl + 1))
    rhs = np.zeros((nx + 1, ndr + 1))
    c = _pad1(coef_int)
    idl = ndl // 2 + 1
    idr = ndr // 2 + 1

    lhs[1:, idl] = 1.0
    for ic in range(1, idl):
        lhs[1:, idl - ic] = c[ic]
        lhs[1:, idl + ic] = c[ic]

    rhs[1:, idr] = 0.0
    for ic in range(1, idr):
        rhs[1:, idr - ic] = -c[ic + 2]
        rhs[1:, idr + ic] = c[ic + 2]

    if bc1 is not None:
        b = _pad1(bc1)
        n = 1
        lhs[n, :] = 0.0
        lhs[n, idl] = 1.0
        if idl > 1:
            icmax = min(idl - 1, 2)
            lhs[n, idl + 1:idl + icmax + 1] = b[1:icmax + 1]
        rhs[n, :] = 0.0
        icmax = min(idr, 4)
        rhs[n, idr:idr + icmax] = b[3:3 + icmax]
        # extended stencil; the reference reads coef_bc1(3+icmax), which is out of
        # bounds (undefined) when idr = 4.  The oracle takes the intended 0.
        rhs[n, 1] = b[3 + icmax] if 3 + icmax < len(b) else 0.0
        n = nx
        lhs[n, 1:] = lhs[1, :0:-1]
        rhs[n, 1:] = -rhs[1, :0:-1]

    if bc2 is not None:
        b = _pad1(bc2)
        n = 2
        if ndl == 3:
            lhs[n, 1:] = [b[1], 1.0, b[2]]
        elif ndl == 5:
            lhs[n, 1:] = [0.0, b[1], 1.0, b[2], 0.0]
        rhs[n, :] = 0.0
        icmax = min(idr + 1, 4)
        rhs[n, idr - 1:idr + icmax - 1] = b[3:3 + icmax]
        n = nx - 1
        lhs[n, 1:] = lhs[2, :0:-1]
        rhs[n, 1:] = -rhs[2, :0:-1]

    if bc3 is not None:
        b = _pad1(bc3)
        n = 3
        if ndl == 5:
            lhs[n, 1:] = [0.0, b[1], 1.0, b[2], 0.0]
        rhs[n, :] = 0.0
        icmax = min(idr + 2, 6)
        rhs[n, idr - 2:idr + icmax - 2] = b[3:3 + icmax]
        n = nx - 2
        lhs[n, 1:] = lhs[3, :0:-1]
        rhs[n, 1:] = -rhs[3, :0:-1]

    # multiply by the Jacobian
    lhs[:, idl] = lhs[:, idl] * dx
    for ic in range(1, idl):
        lhs[:, idl - ic] = lhs[:, idl - ic] * _cshift(dx, -ic)
        lhs[:, idl + ic] = lhs[:, idl + ic] * _cshift(dx, +ic)

    lhs[1:, 1:] = lhs[1:, 1:] / c[3]
    rhs[1:, 1:] = rhs[1:, 1:] / c[3]
    return lhs, rhs


# fdm_com1_jacobian.f90:38-192
def c1n4_jacobian(dx, periodic):
    coef = [0.25, 0.0, 0.75, 0.0, 0.0]
    if periodic:
        lhs, rhs = create_system_1der(dx, 3, 3, coef)
    else:
        bc1 = [2.0, 0.0, -2.5, 2.0, 0.5, 0.0]
        lhs, rhs = create_system_1der(dx, 3, 3, coef, bc1)
    return lhs, rhs, (3, 3), coef


def c1n6_jacobian(dx, periodic):
    coef = [1.0 / 3.0, 0.0, 7.0 / 9.0, 1.0 / 36.0, 0.0]
    if periodic:
        lhs, rhs = create_system_1der(dx, 3, 5, coef)
    else:
        bc1 = [2.0, 0.0, -2.5, 2.0, 0.5, 0.0]
        bc2 = [1.0 / 6.0, 0.5, -5.0 / 9.0, -0.5, 1.0, 1.0 / 18.0]
        lhs, rhs = create_system_1der(dx, 3, 5, coef, bc1, bc2)
    return lhs, rhs, (3, 5), coef


def c1n6_jacobian_penta(dx, periodic):
    c1 = 0.56
    c2 = 0.4 * (-1.0 / 3.0 + c1)
    coef = [c1, c2,
            0.5 * (1.0 / 6.0) * (9.0 + c1 - 20.0 * c2),
            0.25 * (1.0 / 15.0) * (-9.0 + 32.0 * c1 + 62.0 * c2),
            (1.0 / 6.0) * (1.0 / 10.0) * (1.0 - 3.0 * c1 + 12.0 * c2)]
    if periodic:
        lhs, rhs = create_system_1der(dx, 5, 7, coef)
    else:
        bc1 = [2.0, 0.0, -2.5, 2.0, 0.5, 0.0]
        bc2 = [1.0 / 6.0, 0.5, -5.0 / 9.0, -0.5, 1.0, 1.0 / 18.0]
        bc3 = [1.0 / 3.0, 1.0 / 3.0, -1.0 / 36.0, -7.0 / 9.0, 0.0, 7.0 / 9.0, 1.0 / 36.0, 0.0]
        lhs, rhs = create_system_1der(dx, 5, 7, coef, bc1, bc2, bc3)
    return lhs, rhs, (5, 7), coef


# ###########################################################################
# fdm_com2_jacobian.f90:179-282
def create_system_2der(dx1, dx2, ndl, ndr, coef_int, bc1=None, bc2=None, bc3=None):
    """dx1, dx2: 1-based padded 1st/2nd-order Jacobians.
    Returns lhs(n+1, ndl+1), rhs(n+1, ndr+ndl+1) (rhs_d1 in columns ndr+1..ndr+ndl)."""
    nx = len(dx1) - 1
    lhs = np.zeros((nx + 1, ndl + 1))
    rhs = np.zeros((nx + 1, ndr + 1))
    rhs_d1 = np.zeros((nx + 1, ndl + 1))
    c = _pad1(coef_int)
    idl = ndl // 2 + 1
    idr = ndr // 2 + 1

    lhs[1:, idl] = 1.0
    for ic in range(1, idl):
        lhs[1:, idl - ic] = c[ic]
        lhs[1:, idl + ic] = c[ic]

    rhs[1:, idr] = 0.0
    for ic in range(1, idr):
        rhs[1:, idr] = rhs[1:, idr] - 2.0 * c[ic + 2]
        rhs[1:, idr - ic] = c[ic + 2]
        rhs[1:, idr + ic] = c[ic + 2]

    if bc1 is not None:
        b = _pad1(bc1)
        n = 1
        lhs[n, :] = 0.0
        lhs[n, idl] = 1.0
        if idl > 1:
            icmax = min(idl - 1, 2)
            lhs[n, idl + 1:idl + icmax + 1] = b[1:icmax + 1]
        rhs[n, :] = 0.0
        icmax = min(idr, 4)
        rhs[n, idr:idr + icmax] = b[3:3 + icmax]
        # see the note in create_system_1der: out-of-bounds read in the reference
        # for 7-diagonal rhs (Hyper); intended value is 0.
        rhs[n, 1] = b[3 + icmax] if 3 + icmax < len(b) else 0.0
        n = nx
        lhs[n, 1:] = lhs[1, :0:-1]
        rhs[n, 1:] = rhs[1, :0:-1]

    if bc2 is not None:
        b = _pad1(bc2)
        n = 2
        if ndl == 3:
            lhs[n, 1:] = [b[1], 1.0, b[2]]
        rhs[n, :] = 0.0
        icmax = min(idr + 1, 4)
        rhs[n, idr - 1:idr + icmax - 1] = b[3:3 + icmax]
        n = nx - 1
        lhs[n, 1:] = lhs[2, :0:-1]
        rhs[n, 1:] = rhs[2, :0:-1]

    if bc3 is not None:
        b = _pad1(bc3)
        n = 3
        if ndl == 3:
            lhs[n, 1:] = [b[1], 1.0, b[2]]
        rhs[n, :] = 0.0
        icmax = min(idr + 2, 6)
        rhs[n, idr - 2:idr + icmax - 2] = b[3:3 + icmax]
        n = nx - 2
        lhs[n, 1:] = lhs[3, :0:-1]
        rhs[n, 1:] = rhs[3, :0:-1]

    # multiply by the Jacobians
    rhs_d1[:, idl] = -lhs[:, idl] * dx2
    for ic in range(1, idl):
        rhs_d1[:, idl - ic] = -lhs[:, idl - ic] * _cshift(dx2, -ic)
        rhs_d1[:, idl + ic] = -lhs[:, idl + ic] * _cshift(dx2, +ic)

    lhs[:, idl] = lhs[:, idl] * dx1 * dx1
    for ic in range(1, idl):
        lhs[:, idl - ic] = lhs[:, idl - ic] * _cshift(dx1, -ic) * _cshift(dx1, -ic)
        lhs[:, idl + ic] = lhs[:, idl + ic] * _cshift(dx1, +ic) * _cshift(dx1, +ic)

    lhs[1:, 1:] = lhs[1:, 1:] / c[3]
    rhs[1:, 1:] = rhs[1:, 1:] / c[3]
    rhs_d1[1:, 1:] = rhs_d1[1:, 1:] / c[3]

    full = np.zeros((nx + 1, ndr + ndl + 1))
    full[:, :ndr + 1] = rhs
    full[:, ndr + 1:] = rhs_d1[:, 1:]
    return lhs, full


# fdm_com2_jacobian.f90:39-176
def c2n4_jacobian(dx1, dx2, periodic):
    coef = [0.1, 0.0, 1.2, 0.0, 0.0]
    if periodic:
        lhs, rhs = create_system_2der(dx1, dx2, 3, 5, coef)
    else:
        bc1 = [11.0, 0.0, 13.0, -27.0, 15.0, -1.0]
        lhs, rhs = create_system_2der(dx1, dx2, 3, 5, coef, bc1)
    return lhs, rhs, (3, 5), coef


def c2n6_jacobian(dx1, dx2, periodic):
    coef = [2.0 / 11.0, 0.0, 12.0 / 11.0, 3.0 / 44.0, 0.0]
    if periodic:
        lhs, rhs = create_system_2der(dx1, dx2, 3, 5, coef)
    else:
        bc1 = [11.0, 0.0, 13.0, -27.0, 15.0, -1.0]
        bc2 = [0.1, 0.1, 1.2, -2.4, 1.2, 0.0]
        lhs, rhs = create_system_2der(dx1, dx2, 3, 5, coef, bc1, bc2)
    return lhs, rhs, (3, 5), coef


def c2n6_hyper_jacobian(dx1, dx2, periodic):
    kc = pi_wp ** 2.0
    coef = [(272.0 - 45.0 * kc) / (416.0 - 90.0 * kc),
            0.0,
            (48.0 - 135.0 * kc) / (1664.0 - 360.0 * kc),
            (528.0 - 81.0 * kc) / (208.0 - 45.0 * kc) / 4.0,
            -(432.0 - 63.0 * kc) / (1664.0 - 360.0 * kc) / 9.0]
    if periodic:
        lhs, rhs = create_system_2der(dx1, dx2, 3, 7, coef)
    else:
        bc1 = [11.0, 0.0, 13.0, -27.0, 15.0, -1.0]
        bc2 = [0.1, 0.1, 1.2, -2.4, 1.2, 0.0]
        bc3 = [2.0 / 11.0, 2.0 / 11.0, 3.0 / 44.0, 12.0 / 11.0, -51.0 / 22.0, 12.0 / 11.0, 3.0 / 44.0, 0.0]
        lhs, rhs = create_system_2der(dx1, dx2, 3, 7, coef, bc1, bc2, bc3)
    return lhs, rhs, (3, 7), coef


# ###########################################################################
# fdm_base.f90:194-300.  lhs(n+1, ndl+1) modified in place; rhs(n+1, ndr+1);
# rhs_b stored as [row 0..4][col 0..7] with rows 1..4 used (rhs_b(4,0:7));
# rhs_t stored as [row 0..4][col 0..7] with cols 1..7 used (rhs_t(0:4,7)).
def fdm_bcs_neumann(ibc, lhs, rhs, rhs_b, rhs_t):
    ndl = lhs.shape[1] - 1
    idl = ndl // 2 + 1
    ndr = rhs.shape[1] - 1
    idr = ndr // 2 + 1
    nx = lhs.shape[0] - 1
    assert idl >= idr - 1 and idr >= idl

    if ibc in (BCS_ND, BCS_NN):
        rhs_b[1:idr + 1, 1:ndr + 1] = rhs[1:idr + 1, 1:ndr + 1]
        dummy = 1.0 / rhs[1, idr]
        rhs_b[1, 1:ndr + 1] = -rhs_b[1, 1:ndr + 1] * dummy
        for ir in range(1, idr):
            for ic in range(idr + 1, ndr + 1):
                rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] + rhs_b[1 + ir, idr - ir] * rhs_b[1, ic]
            ic = ndr + 1
            rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] + rhs_b[1 + ir, idr - ir] * rhs_b[1, 1]
        lhs[1, 1:ndl + 1] = lhs[1, 1:ndl + 1] * dummy
        for ir in range(1, idr):
            for ic in range(idl + 1, ndl + 1):
                lhs[1 + ir, ic - ir] = lhs[1 + ir, ic - ir] - rhs_b[1 + ir, idr - ir] * lhs[1, ic]
            ic = idr
            rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] * lhs[1, idl]
        for ir in range(1, idl):
            ic = idr
            rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] - lhs[1 + ir, idl - ir]
        rhs_b[1, idr] = lhs[1, idl]

    if ibc in (BCS_DN, BCS_NN):
        rhs_t[1:idr + 1, 1:ndr + 1] = rhs[nx - idr + 1:nx + 1, 1:ndr + 1]
        dummy = 1.0 / rhs[nx, idr]
        rhs_t[idr, 1:ndr + 1] = -rhs_t[idr, 1:ndr + 1] * dummy
        for ir in range(1, idr):
            for ic in range(1, idr):
                rhs_t[idr - ir, ic + ir] = rhs[nx - ir, ic + ir] + rhs[nx - ir, idr + ir] * rhs_t[idr, ic]
            ic = 0
            rhs_t[idr - ir, ic + ir] = rhs_t[idr - ir, ic + ir] + rhs[nx - ir, idr + ir] * rhs_t[idr, ndr]
        lhs[nx, 1:ndl + 1] = lhs[nx, 1:ndl + 1] * dummy
        for ir in range(1, idr):
            for ic in range(1, idl):
                lhs[nx - ir, ic + ir] = lhs[nx - ir, ic + ir] - rhs[nx - ir, idr + ir] * lhs[nx, ic]
            ic = idr
            rhs_t[idr - ir, ic + ir] = rhs_t[idr - ir, ic + ir] * lhs[nx, idl]
        for ir in range(1, idl):
            ic = idr
            rhs_t[idr - ir, ic + ir] = rhs_t[idr - ir, ic + ir] - lhs[nx - ir, idl + ir]
        rhs_t[idr, idr] = lhs[nx, idl]


# fdm_base.f90:304-391.  rhs_b: rows 1.., cols 0..; rhs_t: rows 0.., cols 1..
def fdm_bcs_reduce(ibc, lhs, rhs=None, rhs_b=None, rhs_t=None):
    ndl = lhs.shape[1] - 1
    idl = ndl // 2 + 1
    ndr = rhs.shape[1] - 1
    idr = ndr // 2 + 1
    nx = lhs.shape[0] - 1
    nx_t = idr
    m = max(idl, idr + 1)

    if ibc in (BCS_MIN, BCS_BOTH):
        dummy = 1.0 / lhs[1, idl]
        lhs[1, 1:ndl + 1] = -lhs[1, 1:ndl + 1] * dummy
        lhs[1, idl] = 1.0
        for ir in range(1, idl):
            for ic in range(idl + 1, ndl + 1):
                lhs[1 + ir, ic - ir] = lhs[1 + ir, ic - ir] + lhs[1 + ir, idl - ir] * lhs[1, ic]
            ic = ndl + 1
            lhs[1 + ir, ic - ir] = lhs[1 + ir, ic - ir] + lhs[1 + ir, idl - ir] * lhs[1, 1]
        if rhs_b is not None:
            rhs_b[1:m + 1, 1:ndr + 1] = rhs[1:m + 1, 1:ndr + 1]
            rhs_b[1, 1:ndr + 1] = rhs_b[1, 1:ndr + 1] * dummy
            for ir in range(1, idl):
                for ic in range(idr, ndr + 1):
                    rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] - lhs[1 + ir, idl - ir] * rhs_b[1, ic]
                ic = ndr + 1
                rhs_b[1 + ir, ic - ir] = rhs_b[1 + ir, ic - ir] - lhs[1 + ir, idl - ir] * rhs_b[1, 1]

    if ibc in (BCS_MAX, BCS_BOTH):
        dummy = 1.0 / lhs[nx, idl]
        lhs[nx, 1:ndl + 1] = -lhs[nx, 1:ndl + 1] * dummy
        lhs[nx, idl] = 1.0
        for ir in range(1, idl):
            ic = 0
            lhs[nx - ir, ic + ir] = lhs[nx - ir, ic + ir] + lhs[nx - ir, idl + ir] * lhs[nx, ndl]
            for ic in range(1, idl):
                lhs[nx - ir, ic + ir] = lhs[nx - ir, ic + ir] + lhs[nx - ir, idl + ir] * lhs[nx, ic]
        if rhs_t is not None:
            rhs_t[nx_t - m + 1:nx_t + 1, 1:ndr + 1] = rhs[nx - m + 1:nx + 1, 1:ndr + 1]
            rhs_t[nx_t, 1:ndr + 1] = rhs_t[nx_t, 1:ndr + 1] * dummy
            for ir in range(1, idl):
                ic = 0
                rhs_t[nx_t - ir, ic + ir] = rhs_t[nx_t - ir, ic + ir] - lhs[nx - ir, idl + ir] * rhs_t[nx_t, ndr]
                for ic in range(1, idr + 1):
                    rhs_t[nx_t - ir, ic + ir] = rhs_t[nx_t - ir, ic + ir] - lhs[nx - ir, idl + ir] * rhs_t[nx_t, ic]


# ###########################################################################
# utils/linear3.f90.  Diagonals are 0-based vectors of length nmax here.
def tridfs(a, b, c):
    """linear3.f90:29-51; in place."""
    nmax = len(a)
    for n in range(1, nmax):
        a[n] = a[n] / b[n - 1]
        b[n] = b[n] - a[n] * c[n - 1]
    a[:] = -a
    b[:] = 1.0 / b
    c[:] = -c


def tridss(a, b, c, f):
    """linear3.f90:56-150; f(nmax, nlines) in place."""
    nmax = len(a)
    for n in range(1, nmax):
        f[n] = f[n] + a[n] * f[n - 1]
    f[nmax - 1] = f[nmax - 1] * b[nmax - 1]
    for n in range(nmax - 2, -1, -1):
        f[n] = (f[n] + c[n] * f[n + 1]) * b[n]


def tridpfs(a, b, c, d, e):
    """linear3.f90:269-316; in place (Fortran index n -> n-1)."""
    nmax = len(a)
    c[0] = c[0] / b[0]
    e[0] = a[0] / b[0]
    d[0] = c[nmax - 1]
    for n in range(1, nmax - 2):
        b[n] = b[n] - a[n] * c[n - 1]
        c[n] = c[n] / b[n]
        e[n] = -a[n] * e[n - 1] / b[n]
        d[n] = -d[n - 1] * c[n - 1]
    m = nmax - 2  # Fortran nmax-1
    b[m] = b[m] - a[m] * c[m - 1]
    e[m] = (c[m] - a[m] * e[m - 1]) / b[m]
    d[m] = a[nmax - 1] - d[m - 1] * c[m - 1]
    s = 0.0
    for n in range(0, nmax - 1):
        s = s + d[n] * e[n]
    b[nmax - 1] = b[nmax - 1] - s
    for n in range(nmax):
        b[n] = 1.0 / b[n]
        a[n] = -a[n] * b[n]
        c[n] = -c[n]
        e[n] = -e[n]


def tridpss(a, b, c, d, e, f):
    """linear3.f90:321-442; f(nmax, nlines) in place."""
    nmax = len(a)
    f[0] = f[0] * b[0]
    for n in range(1, nmax - 1):
        f[n] = f[n] * b[n] + a[n] * f[n - 1]
    wrk = np.zeros_like(f[0])
    for n in range(0, nmax - 1):
        wrk = wrk + d[n] * f[n]
    f[nmax - 1] = (f[nmax - 1] - wrk) * b[nmax - 1]
    f[nmax - 2] = e[nmax - 2] * f[nmax - 1] + f[nmax - 2]
    for n in range(nmax - 3, -1, -1):
        f[n] = f[n] + c[n] * f[n + 1] + e[n] * f[nmax - 1]


# utils/linear5.f90
def pentadfs(a, b, c, d, e):
    """linear5.f90:30-71; in place, 0-based (Fortran n -> n-1)."""
    nmax = len(a)
    b[1] = b[1] / c[0]
    c[1] = c[1] - b[1] * d[0]
    d[1] = d[1] - b[1] * e[0]
    for n in range(2, nmax - 1):
        a[n] = a[n] / c[n - 2]
        b[n] = (b[n] - a[n] * d[n - 2]) / c[n - 1]
        c[n] = c[n] - b[n] * d[n - 1] - a[n] * e[n - 2]
        d[n] = d[n] - b[n] * e[n - 1]
    n = nmax - 1
    a[n] = a[n] / c[n - 2]
    b[n] = (b[n] - a[n] * d[n - 2]) / c[n - 1]
    c[n] = c[n] - b[n] * d[n - 1] - a[n] * e[n - 2]
    a[2:] = -a[2:]
    b[1:] = -b[1:]
    c[:] = 1.0 / c
    d[:nmax - 1] = -d[:nmax - 1]
    e[:nmax - 2] = -e[:nmax - 2]


def pentadss(a, b, c, d, e, f):
    """linear5.f90:76-131; f(nmax, nlines) in place."""
    nmax = len(a)
    f[1] = f[1] + f[0] * b[1]
    for n in range(2, nmax):
        f[n] = f[n] + f[n - 1] * b[n] + f[n - 2] * a[n]
    n = nmax - 1
    f[n] = f[n] * c[n]
    n = nmax - 2
    f[n] = (f[n] + f[n + 1] * d[n]) * c[n]
    for n in range(nmax - 3, -1, -1):
        f[n] = (f[n] + f[n + 1] * d[n] + f[n + 2] * e[n]) * c[n]


def pentadfs2(a, b, c, d, e):
    """linear5.f90:156-204; reverse-order LE decomposition, in place."""
    nmax = len(a)
    n = nmax - 1
    e[n] = 1.0
    d[n] = 1.0
    n = nmax - 2
    e[n] = 1.0
    d[n] = d[n] / c[n + 1]
    c[n] = c[n] - d[n] * b[n + 1]
    b[n] = b[n] - d[n] * a[n + 1]
    for n in range(nmax - 3, 1, -1):
        e[n] = e[n] / c[n + 2]
        d[n] = (d[n] - e[n] * b[n + 2]) / c[n + 1]
        c[n] = c[n] - d[n] * b[n + 1] - e[n] * a[n + 2]
        b[n] = b[n] - d[n] * a[n + 1]
    n = 1
    e[n] = e[n] / c[n + 2]
    d[n] = (d[n] - e[n] * b[n + 2]) / c[n + 1]
    c[n] = c[n] - d[n] * b[n + 1] - e[n] * a[n + 2]
    b[n] = b[n] - d[n] * a[n + 1]
    a[n] = 1.0
    n = 0
    e[n] = e[n] / c[n + 2]
    d[n] = (d[n] - e[n] * b[n + 2]) / c[n + 1]
    c[n] = c[n] - d[n] * b[n + 1] - e[n] * a[n + 2]
    b[n] = 1.0
    a[n] = 1.0


def pentadss2(a, b, c, d, e, f):
    """linear5.f90:209-244; f(nmax, nlines) in place."""
    nmax = len(a)
    n = nmax - 2
    f[n] = f[n] - f[n + 1] * d[n]
    for n in range(nmax - 3, -1, -1):
        f[n] = f[n] - f[n + 1] * d[n] - f[n + 2] * e[n]
    f[0] = f[0] / c[0]
    f[1] = (f[1] - f[0] * b[1]) / c[1]
    for n in range(2, nmax):
        f[n] = (f[n] - f[n - 1] * b[n] - f[n - 2] * a[n]) / c[n]


def pentadpfs(a, b, c, d, e, f, g):
    """linear5.f90:273-345; circulant pentadiagonal, in place."""
    nmax = len(a)
    a0, b0, en, dn = a[0], b[0], e[nmax - 1], d[nmax - 1]
    b[1] = b[1] - d[nmax - 1]
    c[0] = c[0] - e[nmax - 1]
    c[1] = c[1] - e[nmax - 1]
    c[nmax - 2] = c[nmax - 2] - a[0]
    c[nmax - 1] = c[nmax - 1] - a[0]
    d[nmax - 2] = d[nmax - 2] - b[0]
    a[0] = 0.0
    a[1] = 0.0
    b[0] = 0.0
    d[nmax - 1] = 0.0
    e[nmax - 1] = 0.0
    e[nmax - 2] = 0.0
    pentadfs2(a, b, c, d, e)
    a[0] = a0
    b[0] = b0
    e[nmax - 1] = en
    d[nmax - 1] = dn
    f[:] = 0.0
    f[0] = 1.0
    f[nmax - 2] = 1.0
    g[:] = 0.0
    g[1] = 1.0
    g[nmax - 1] = 1.0
    pentadss2(a, b, c, d, e, f)
    pentadss2(a, b, c, d, e, g)


def pentadpss(a, b, c, d, e, f, g, frc):
    """linear5.f90:350-411; frc(nmax, nlines) in place."""
    nmax = len(a)
    pentadss2(a, b, c, d, e, frc)
    N = nmax - 1
    m1 = e[N] * f[0] + a[0] * f[N - 1] + b[0] * f[N] + 1.0
    m2 = e[N] * g[0] + a[0] * g[N - 1] + b[0] * g[N]
    m3 = d[N] * f[0] + e[N] * f[1] + a[0] * f[N]
    m4 = d[N] * g[0] + e[N] * g[1] + a[0] * g[N] + 1.0
    di = 1 / (m1 * m4 - m2 * m3)
    d11 = di * (m4 * e[N] - m2 * d[N])
    d12 = di * (m4 * b[0] - m2 * a[0])
    d13 = di * m4 * a[0]
    d14 = di * m2 * e[N]
    d21 = di * (m1 * d[N] - m3 * e[N])
    d22 = di * (m1 * a[0] - m3 * b[0])
    d23 = di * m3 * a[0]
    d24 = di * m1 * e[N]
    dummy1 = d11 * frc[0] + d12 * frc[N] + d13 * frc[N - 1] - d14 * frc[1]
    dummy2 = d21 * frc[0] + d22 * frc[N] - d23 * frc[N - 1] + d24 * frc[1]
    # the reference updates the interior rows 3..nmax-3 first (they do not enter
    # dummy1/2), then the boundary rows with the still-unmodified values.
    for n in list(range(2, nmax - 3)) + [0, 1, nmax - 3, nmax - 2, nmax - 1]:
        frc[n] = frc[n] - dummy1 * f[n] - dummy2 * g[n]


# ###########################################################################
# fdm_matmul.f90.  rhs is 1-based padded (n+1, nd+1); u, f are (n, nlines).
# U(n)/F(n) below are the Fortran u(:,n)/f(:,n).
def _acc(n):
    return n - 1


def matmul_3d(rhs, u, f, ibc, rhs_b=None, rhs_t=None, want_bcs=False):
    """fdm_matmul.f90:70-121. Returns (bcs_b, bcs_t) when want_bcs."""
    nx = rhs.shape[0] - 1
    r = rhs
    bcs_b = bcs_t = None
    if ibc in (BCS_MIN, BCS_BOTH):
        if want_bcs:
            bcs_b = f[0] * rhs_b[1, 2] + u[1] * rhs_b[1, 3] + u[2] * rhs_b[1, 1]
        f[1] = f[0] * rhs_b[2, 1] + u[1] * rhs_b[2, 2] + u[2] * rhs_b[2, 3]
        f[2] = f[0] * rhs_b[3, 0] + u[1] * rhs_b[3, 1] + u[2] * rhs_b[3, 2] + u[3] * rhs_b[3, 3]
    else:
        f[0] = u[0] * r[1, 2] + u[1] * r[1, 3] + u[2] * r[1, 1]
        f[1] = u[0] * r[2, 1] + u[1] * r[2, 2] + u[2] * r[2, 3]
        f[2] = u[1] * r[3, 1] + u[2] * r[3, 2] + u[3] * r[3, 3]
    # interior n = 4 .. nx-3
    lo, hi = 4, nx - 3
    if hi >= lo:
        s = slice(lo - 1, hi)
        f[s] = u[lo - 2:hi - 1] * _c(r[lo:hi + 1, 1], u) + u[s] * _c(r[lo:hi + 1, 2], u) + u[lo:hi + 1]
    if ibc in (BCS_MAX, BCS_BOTH):
        f[nx - 3] = u[nx - 4] * rhs_t[0, 1] + u[nx - 3] * rhs_t[0, 2] + u[nx - 2] * rhs_t[0, 3] + f[nx - 1] * rhs_t[0, 4]
        f[nx - 2] = u[nx - 3] * rhs_t[1, 1] + u[nx - 2] * rhs_t[1, 2] + f[nx - 1] * rhs_t[1, 3]
        if want_bcs:
            bcs_t = u[nx - 3] * rhs_t[2, 3] + u[nx - 2] * rhs_t[2, 1] + f[nx - 1] * rhs_t[2, 2]
    else:
        f[nx - 3] = u[nx - 4] * r[nx - 2, 1] + u[nx - 3] * r[nx - 2, 2] + u[nx - 2] * r[nx - 2, 3]
        f[nx - 2] = u[nx - 3] * r[nx - 1, 1] + u[nx - 2] * r[nx - 1, 2] + u[nx - 1] * r[nx - 1, 3]
        f[nx - 1] = u[nx - 3] * r[nx, 3] + u[nx - 2] * r[nx, 1] + u[nx - 1] * r[nx, 2]
    return bcs_b, bcs_t


def matmul_3d_add(rhs, u, f):
    """fdm_matmul.f90:126-153.  rhs: 1-based padded (n+1, 4)."""
    nx = rhs.shape[0] - 1
    r = rhs
    f[0] = f[0] + u[0] * r[1, 2] + u[1] * r[1, 3] + u[2] * r[1, 1]
    f[1:nx - 1] = f[1:nx - 1] + u[0:nx - 2] * _c(r[2:nx, 1], u) + u[1:nx - 1] * _c(r[2:nx, 2], u) + u[2:nx] * _c(r[2:nx, 3], u)
    f[nx - 1] = f[nx - 1] + u[nx - 3] * r[nx, 3] + u[nx - 2] * r[nx, 1] + u[nx - 1] * r[nx, 2]


def matmul_3d_antisym(rhs, u, f, ibc, rhs_b=None, rhs_t=None, want_bcs=False):
    """fdm_matmul.f90:157-212."""
    nx = rhs.shape[0] - 1
    r = rhs
    bcs_b = bcs_t = None
    if ibc == BCS_PERIODIC:
        f[0] = u[1] - u[nx - 1]
        f[1] = u[2] - u[0]
    elif ibc in (BCS_ND, BCS_NN):  # same codes as BCS_MIN, BCS_BOTH
        if want_bcs:
            bcs_b = f[0] * rhs_b[1, 2] + u[1] * rhs_b[1, 3] + u[2] * rhs_b[1, 1]
        f[1] = f[0] * rhs_b[2, 1] + u[1] * rhs_b[2, 2] + u[2] * rhs_b[2, 3]
    else:
        f[0] = u[0] * r[1, 2] + u[1] * r[1, 3] + u[2] * r[1, 1]
        f[1] = u[0] * r[2, 1] + u[1] * r[2, 2] + u[2] * r[2, 3]
    f[2:nx - 2] = u[3:nx - 1] - u[1:nx - 3]
    if ibc == BCS_PERIODIC:
        f[nx - 2] = u[nx - 1] - u[nx - 3]
        f[nx - 1] = u[0] - u[nx - 2]
    elif ibc in (BCS_DN, BCS_NN):
        f[nx - 2] = u[nx - 3] * rhs_t[1, 1] + u[nx - 2] * rhs_t[1, 2] + f[nx - 1] * rhs_t[1, 3]
        if want_bcs:
            bcs_t = u[nx - 3] * rhs_t[2, 3] + u[nx - 2] * rhs_t[2, 1] + f[nx - 1] * rhs_t[2, 2]
    else:
        f[nx - 2] = u[nx - 3] * r[nx - 1, 1] + u[nx - 2] * r[nx - 1, 2] + u[nx - 1] * r[nx - 1, 3]
        f[nx - 1] = u[nx - 3] * r[nx, 3] + u[nx - 2] * r[nx, 1] + u[nx - 1] * r[nx, 2]
    return bcs_b, bcs_t


def matmul_5d(rhs, u, f, ibc, rhs_b=None, rhs_t=None, want_bcs=False):
    """fdm_matmul.f90:266-320."""
    nx = rhs.shape[0] - 1
    r = rhs
    bcs_b = bcs_t = None
    if ibc in (BCS_MIN, BCS_BOTH):
        if want_bcs:
            bcs_b = f[0] * rhs_b[1, 3] + u[1] * rhs_b[1, 4] + u[2] * rhs_b[1, 5] + u[3] * rhs_b[1, 1]
        f[1] = f[0] * rhs_b[2, 2] + u[1] * rhs_b[2, 3] + u[2] * rhs_b[2, 4] + u[3] * rhs_b[2, 5]
        f[2] = f[0] * rhs_b[3, 1] + u[1] * rhs_b[3, 2] + u[2] * rhs_b[3, 3] + u[3] * rhs_b[3, 4] + u[4] * rhs_b[3, 5]
        f[3] = (f[0] * rhs_b[4, 0] + u[1] * rhs_b[4, 1] + u[2] * rhs_b[4, 2] + u[3] * rhs_b[4, 3]
                + u[4] * rhs_b[4, 4] + u[5] * rhs_b[4, 5])
    else:
        f[0] = u[0] * r[1, 3] + u[1] * r[1, 4] + u[2] * r[1, 5] + u[3] * r[1, 1]
        f[1] = u[0] * r[2, 2] + u[1] * r[2, 3] + u[2] * r[2, 4] + u[3] * r[2, 5]
        f[2] = u[0] * r[3, 1] + u[1] * r[3, 2] + u[2] * r[3, 3] + u[3] * r[3, 4] + u[4] * r[3, 5]
        f[3] = u[1] * r[4, 1] + u[2] * r[4, 2] + u[3] * r[4, 3] + u[4] * r[4, 4] + u[5] * r[4, 5]
    lo, hi = 5, nx - 4
    if hi >= lo:
        s = slice(lo - 1, hi)
        f[s] = (u[lo - 3:hi - 2] * _c(r[lo:hi + 1, 1], u) + u[lo - 2:hi - 1] * _c(r[lo:hi + 1, 2], u)
                + u[s] * _c(r[lo:hi + 1, 3], u) + u[lo:hi + 1] + u[lo + 1:hi + 2] * _c(r[lo:hi + 1, 5], u))
    if ibc in (BCS_MAX, BCS_BOTH):
        f[nx - 4] = (u[nx - 6] * rhs_t[0, 1] + u[nx - 5] * rhs_t[0, 2] + u[nx - 4] * rhs_t[0, 3]
                     + u[nx - 3] * rhs_t[0, 4] + u[nx - 2] * rhs_t[0, 5] + f[nx - 1] * rhs_t[0, 6])
        f[nx - 3] = (u[nx - 5] * rhs_t[1, 1] + u[nx - 4] * rhs_t[1, 2] + u[nx - 3] * rhs_t[1, 3]
                     + u[nx - 2] * rhs_t[1, 4] + f[nx - 1] * rhs_t[1, 5])
        f[nx - 2] = u[nx - 4] * rhs_t[2, 1] + u[nx - 3] * rhs_t[2, 2] + u[nx - 2] * rhs_t[2, 3] + f[nx - 1] * rhs_t[2, 4]
        if want_bcs:
            bcs_t = u[nx - 4] * rhs_t[3, 5] + u[nx - 3] * rhs_t[3, 1] + u[nx - 2] * rhs_t[3, 2] + f[nx - 1] * rhs_t[3, 3]
    else:
        f[nx - 4] = (u[nx - 6] * r[nx - 3, 1] + u[nx - 5] * r[nx - 3, 2] + u[nx - 4] * r[nx - 3, 3]
                     + u[nx - 3] * r[nx - 3, 4] + u[nx - 2] * r[nx - 3, 5])
        f[nx - 3] = (u[nx - 5] * r[nx - 2, 1] + u[nx - 4] * r[nx - 2, 2] + u[nx - 3] * r[nx - 2, 3]
                     + u[nx - 2] * r[nx - 2, 4] + u[nx - 1] * r[nx - 2, 5])
        f[nx - 2] = u[nx - 4] * r[nx - 1, 1] + u[nx - 3] * r[nx - 1, 2] + u[nx - 2] * r[nx - 1, 3] + u[nx - 1] * r[nx - 1, 4]
        f[nx - 1] = u[nx - 4] * r[nx, 5] + u[nx - 3] * r[nx, 1] + u[nx - 2] * r[nx, 2] + u[nx - 1] * r[nx, 3]
    return bcs_b, bcs_t


def matmul_5d_antisym(rhs, u, f, ibc, rhs_b=None, rhs_t=None, want_bcs=False):
    """fdm_matmul.f90:359-419."""
    nx = rhs.shape[0] - 1
    r = rhs
    r5 = r[4, 5]
    bcs_b = bcs_t = None
    if ibc == BCS_PERIODIC:
        f[0] = u[1] - u[nx - 1] + r5 * (u[2] - u[nx - 2])
        f[1] = u[2] - u[0] + r5 * (u[3] - u[nx - 1])
        f[2] = u[3] - u[1] + r5 * (u[4] - u[0])
    elif ibc in (BCS_ND, BCS_NN):
        if want_bcs:
            bcs_b = f[0] * rhs_b[1, 3] + u[1] * rhs_b[1, 4] + u[2] * rhs_b[1, 5] + u[3] * rhs_b[1, 1]
        f[1] = f[0] * rhs_b[2, 2] + u[1] * rhs_b[2, 3] + u[2] * rhs_b[2, 4] + u[3] * rhs_b[2, 5]
        f[2] = f[0] * rhs_b[3, 1] + u[1] * rhs_b[3, 2] + u[2] * rhs_b[3, 3] + u[3] * rhs_b[3, 4] + u[4] * rhs_b[3, 5]
    else:
        f[0] = u[0] * r[1, 3] + u[1] * r[1, 4] + u[2] * r[1, 5] + u[3] * r[1, 1]
        f[1] = u[0] * r[2, 2] + u[1] * r[2, 3] + u[2] * r[2, 4] + u[3] * r[2, 5]
        f[2] = u[0] * r[3, 1] + u[1] * r[3, 2] + u[2] * r[3, 3] + u[3] * r[3, 4] + u[4] * r[3, 5]
    # interior n = 4 .. nx-3
    f[3:nx - 3] = u[4:nx - 2] - u[2:nx - 4] + r5 * (u[5:nx - 1] - u[1:nx - 5])
    if ibc == BCS_PERIODIC:
        f[nx - 3] = u[nx - 2] - u[nx - 4] + r5 * (u[nx - 1] - u[nx - 5])
        f[nx - 2] = u[nx - 1] - u[nx - 3] + r5 * (u[0] - u[nx - 4])
        f[nx - 1] = u[0] - u[nx - 2] + r5 * (u[1] - u[nx - 3])
    elif ibc in (BCS_DN, BCS_NN):
        f[nx - 3] = (u[nx - 5] * rhs_t[1, 1] + u[nx - 4] * rhs_t[1, 2] + u[nx - 3] * rhs_t[1, 3]
                     + u[nx - 2] * rhs_t[1, 4] + f[nx - 1] * rhs_t[1, 5])
        f[nx - 2] = u[nx - 4] * rhs_t[2, 1] + u[nx - 3] * rhs_t[2, 2] + u[nx - 2] * rhs_t[2, 3] + f[nx - 1] * rhs_t[2, 4]
        if want_bcs:
            bcs_t = u[nx - 4] * rhs_t[3, 5] + u[nx - 3] * rhs_t[3, 1] + u[nx - 2] * rhs_t[3, 2] + f[nx - 1] * rhs_t[3, 3]
    else:
        f[nx - 3] = (u[nx - 5] * r[nx - 2, 1] + u[nx - 4] * r[nx - 2, 2] + u[nx - 3] * r[nx - 2, 3]
                     + u[nx - 2] * r[nx - 2, 4] + u[nx - 1] * r[nx - 2, 5])
        f[nx - 2] = u[nx - 4] * r[nx - 1, 1] + u[nx - 3] * r[nx - 1, 2] + u[nx - 2] * r[nx - 1, 3] + u[nx - 1] * r[nx - 1, 4]
        f[nx - 1] = u[nx - 4] * r[nx, 5] + u[nx - 3] * r[nx, 1] + u[nx - 2] * r[nx, 2] + u[nx - 1] * r[nx, 3]
    return bcs_b, bcs_t


def matmul_5d_sym(rhs, u, f, ibc):
    """fdm_matmul.f90:423-485."""
    nx = rhs.shape[0] - 1
    r = rhs
    r5 = r[3, 5]
    r3 = r[3, 3]
    if ibc == BCS_PERIODIC:
        f[0] = r3 * u[0] + u[1] + u[nx - 1] + r5 * (u[2] + u[nx - 2])
        f[1] = r3 * u[1] + u[2] + u[0] + r5 * (u[3] + u[nx - 1])
    else:
        f[0] = u[0] * r[1, 3] + u[1] * r[1, 4] + u[2] * r[1, 5] + u[3] * r[1, 1]
        f[1] = u[0] * r[2, 2] + u[1] * r[2, 3] + u[2] * r[2, 4] + u[3] * r[2, 5]
        if ibc in (BCS_ND, BCS_NN):
            f[0] = 0.0
    f[2:nx - 2] = r3 * u[2:nx - 2] + u[3:nx - 1] + u[1:nx - 3] + r5 * (u[4:nx] + u[0:nx - 4])
    if ibc == BCS_PERIODIC:
        f[nx - 2] = r3 * u[nx - 2] + u[nx - 1] + u[nx - 3] + r5 * (u[0] + u[nx - 4])
        f[nx - 1] = r3 * u[nx - 1] + u[0] + u[nx - 2] + r5 * (u[1] + u[nx - 3])
    else:
        f[nx - 2] = u[nx - 4] * r[nx - 1, 1] + u[nx - 3] * r[nx - 1, 2] + u[nx - 2] * r[nx - 1, 3] + u[nx - 1] * r[nx - 1, 4]
        f[nx - 1] = u[nx - 4] * r[nx, 5] + u[nx - 3] * r[nx, 1] + u[nx - 2] * r[nx, 2] + u[nx - 1] * r[nx, 3]
        if ibc in (BCS_DN, BCS_NN):
            f[nx - 1] = 0.0


def matmul_7d_antisym(rhs, u, f, ibc, rhs_b=None, rhs_t=None, want_bcs=False):
    """fdm_matmul.f90:491-558."""
    nx = rhs.shape[0] - 1
    r = rhs
    r6 = r[5, 6]
    r7 = r[5, 7]
    bcs_b = bcs_t = None
    N = nx

    def U(n):
        return u[n - 1]

    if ibc == BCS_PERIODIC:
        f[0] = U(2) - U(N) + r6 * (U(3) - U(N - 1)) + r7 * (U(4) - U(N - 2))
        f[1] = U(3) - U(1) + r6 * (U(4) - U(N)) + r7 * (U(5) - U(N - 1))
        f[2] = U(4) - U(2) + r6 * (U(5) - U(1)) + r7 * (U(6) - U(N))
        f[3] = U(5) - U(3) + r6 * (U(6) - U(2)) + r7 * (U(7) - U(1))
    elif ibc in (BCS_ND, BCS_NN):
        F1 = f[0]
        if want_bcs:
            bcs_b = F1 * rhs_b[1, 4] + U(2) * rhs_b[1, 5] + U(3) * rhs_b[1, 6] + U(4) * rhs_b[1, 7] + U(5) * rhs_b[1, 1]
        f[1] = F1 * rhs_b[2, 3] + U(2) * rhs_b[2, 4] + U(3) * rhs_b[2, 5] + U(4) * rhs_b[2, 6] + U(5) * rhs_b[2, 7]
        f[2] = (F1 * rhs_b[3, 2] + U(2) * rhs_b[3, 3] + U(3) * rhs_b[3, 4] + U(4) * rhs_b[3, 5]
                + U(5) * rhs_b[3, 6] + U(6) * rhs_b[3, 7])
        f[3] = (F1 * rhs_b[4, 1] + U(2) * rhs_b[4, 2] + U(3) * rhs_b[4, 3] + U(4) * rhs_b[4, 4]
                + U(5) * rhs_b[4, 5] + U(6) * rhs_b[4, 6] + U(7) * rhs_b[4, 7])
    else:
        f[0] = U(1) * r[1, 4] + U(2) * r[1, 5] + U(3) * r[1, 6] + U(4) * r[1, 7] + U(5) * r[1, 1]
        f[1] = U(1) * r[2, 3] + U(2) * r[2, 4] + U(3) * r[2, 5] + U(4) * r[2, 6] + U(5) * r[2, 7]
        f[2] = U(1) * r[3, 2] + U(2) * r[3, 3] + U(3) * r[3, 4] + U(4) * r[3, 5] + U(5) * r[3, 6] + U(6) * r[3, 7]
        f[3] = (U(1) * r[4, 1] + U(2) * r[4, 2] + U(3) * r[4, 3] + U(4) * r[4, 4] + U(5) * r[4, 5]
                + U(6) * r[4, 6] + U(7) * r[4, 7])
    # interior n = 5 .. nx-4
    f[4:N - 4] = (u[5:N - 3] - u[3:N - 5] + r6 * (u[6:N - 2] - u[2:N - 6]) + r7 * (u[7:N - 1] - u[1:N - 7]))
    if ibc == BCS_PERIODIC:
        f[N - 4] = U(N - 2) - U(N - 4) + r6 * (U(N - 1) - U(N - 5)) + r7 * (U(N) - U(N - 6))
        f[N - 3] = U(N - 1) - U(N - 3) + r6 * (U(N) - U(N - 4)) + r7 * (U(1) - U(N - 5))
        f[N - 2] = U(N) - U(N - 2) + r6 * (U(1) - U(N - 3)) + r7 * (U(2) - U(N - 4))
        f[N - 1] = U(1) - U(N - 1) + r6 * (U(2) - U(N - 2)) + r7 * (U(3) - U(N - 3))
    elif ibc in (BCS_DN, BCS_NN):
        FN = f[N - 1]
        f[N - 4] = (U(N - 6) * rhs_t[1, 1] + U(N - 5) * rhs_t[1, 2] + U(N - 4) * rhs_t[1, 3] + U(N - 3) * rhs_t[1, 4]
                    + U(N - 2) * rhs_t[1, 5] + U(N - 1) * rhs_t[1, 6] + FN * rhs_t[1, 7])
        f[N - 3] = (U(N - 5) * rhs_t[2, 1] + U(N - 4) * rhs_t[2, 2] + U(N - 3) * rhs_t[2, 3] + U(N - 2) * rhs_t[2, 4]
                    + U(N - 1) * rhs_t[2, 5] + FN * rhs_t[2, 6])
        f[N - 2] = (U(N - 4) * rhs_t[3, 1] + U(N - 3) * rhs_t[3, 2] + U(N - 2) * rhs_t[3, 3] + U(N - 1) * rhs_t[3, 4]
                    + FN * rhs_t[3, 5])
        if want_bcs:
            bcs_t = (U(N - 4) * rhs_t[4, 7] + U(N - 3) * rhs_t[4, 1] + U(N - 2) * rhs_t[4, 2] + U(N - 1) * rhs_t[4, 3]
                     + FN * rhs_t[4, 4])
    else:
        f[N - 4] = (U(N - 6) * r[N - 3, 1] + U(N - 5) * r[N - 3, 2] + U(N - 4) * r[N - 3, 3] + U(N - 3) * r[N - 3, 4]
                    + U(N - 2) * r[N - 3, 5] + U(N - 1) * r[N - 3, 6] + U(N) * r[N - 3, 7])
        f[N - 3] = (U(N - 5) * r[N - 2, 1] + U(N - 4) * r[N - 2, 2] + U(N - 3) * r[N - 2, 3] + U(N - 2) * r[N - 2, 4]
                    + U(N - 1) * r[N - 2, 5] + U(N) * r[N - 2, 6])
        f[N - 2] = (U(N - 4) * r[N - 1, 1] + U(N - 3) * r[N - 1, 2] + U(N - 2) * r[N - 1, 3] + U(N - 1) * r[N - 1, 4]
                    + U(N) * r[N - 1, 5])
        f[N - 1] = U(N - 4) * r[N, 7] + U(N - 3) * r[N, 1] + U(N - 2) * r[N, 2] + U(N - 1) * r[N, 3] + U(N) * r[N, 4]
    return bcs_b, bcs_t


def matmul_7d_sym(rhs, u, f, ibc):
    """fdm_matmul.f90:562-642."""
    nx = rhs.shape[0] - 1
    r = rhs
    r7 = r[4, 7]
    r6 = r[4, 6]
    r4 = r[4, 4]
    N = nx

    def U(n):
        return u[n - 1]

    if ibc == BCS_PERIODIC:
        f[0] = r4 * U(1) + U(2) + U(N) + r6 * (U(3) + U(N - 1)) + r7 * (U(4) + U(N - 2))
        f[1] = r4 * U(2) + U(3) + U(1) + r6 * (U(4) + U(N)) + r7 * (U(5) + U(N - 1))
        f[2] = r4 * U(3) + U(4) + U(2) + r6 * (U(5) + U(1)) + r7 * (U(6) + U(N))
    else:
        f[0] = U(1) * r[1, 4] + U(2) * r[1, 5] + U(3) * r[1, 6] + U(4) * r[1, 7] + U(5) * r[1, 1]
        f[1] = U(1) * r[2, 3] + U(2) * r[2, 4] + U(3) * r[2, 5] + U(4) * r[2, 6] + U(5) * r[2, 7]
        f[2] = U(1) * r[3, 2] + U(2) * r[3, 3] + U(3) * r[3, 4] + U(4) * r[3, 5] + U(5) * r[3, 6] + U(6) * r[3, 7]
        if ibc in (BCS_ND, BCS_NN):
            f[0] = 0.0
    # interior n = 4 .. nx-3
    f[3:N - 3] = (r4 * u[3:N - 3] + u[4:N - 2] + u[2:N - 4] + r6 * (u[5:N - 1] + u[1:N - 5])
                  + r7 * (u[6:N] + u[0:N - 6]))
    if ibc == BCS_PERIODIC:
        f[N - 3] = r4 * U(N - 2) + U(N - 1) + U(N - 3) + r6 * (U(N) + U(N - 4)) + r7 * (U(1) + U(N - 5))
        f[N - 2] = r4 * U(N - 1) + U(N) + U(N - 2) + r6 * (U(1) + U(N - 3)) + r7 * (U(2) + U(N - 4))
        f[N - 1] = r4 * U(N) + U(1) + U(N - 1) + r6 * (U(2) + U(N - 2)) + r7 * (U(3) + U(N - 3))
    else:
        f[N - 3] = (U(N - 5) * r[N - 2, 1] + U(N - 4) * r[N - 2, 2] + U(N - 3) * r[N - 2, 3] + U(N - 2) * r[N - 2, 4]
                    + U(N - 1) * r[N - 2, 5] + U(N) * r[N - 2, 6])
        f[N - 2] = (U(N - 4) * r[N - 1, 1] + U(N - 3) * r[N - 1, 2] + U(N - 2) * r[N - 1, 3] + U(N - 1) * r[N - 1, 4]
                    + U(N) * r[N - 1, 5])
        f[N - 1] = U(N - 4) * r[N, 7] + U(N - 3) * r[N, 1] + U(N - 2) * r[N, 2] + U(N - 1) * r[N, 3] + U(N) * r[N, 4]
        if ibc in (BCS_DN, BCS_NN):
            f[N - 1] = 0.0


# ###########################################################################
# fdm_derivative.f90
class Derivative:
    """fdm_derivative_dt (fdm_derivative.f90:16-29)."""

    def __init__(self, mode_fdm):
        self.mode_fdm = mode_fdm
        self.size = 0
        self.periodic = False
        self.need_1der = False
        self.nb_diag = (0, 0)
        self.rhs_b = np.zeros((5, 8))   # rhs_b(4, 0:7): rows 1..4, cols 0..7
        self.rhs_t = np.zeros((5, 8))   # rhs_t(0:4, 7): rows 0..4, cols 1..7
        self.lhs = None                 # (n+1, ndl+1) 1-based padded
        self.rhs = None                 # (n+1, ndr[+ndl]+1)
        self.mwn = None                 # (n,)
        self.lu = None                  # (n+1, ncol+1) 1-based padded
        self.coef = None


def _wavenumbers(nx):
    wn = np.zeros(nx)
    for i in range(1, nx + 1):
        if i <= nx // 2 + 1:
            wn[i - 1] = 2.0 * pi_wp * float(i - 1) / float(nx)
        else:
            wn[i - 1] = 2.0 * pi_wp * float(i - 1 - nx) / float(nx)
    return wn


def der1_create_system(x, dx, g, periodic):
    """fdm_derivative.f90:146-214.  dx is 1-based padded."""
    g.size = len(x)
    g.periodic = periodic
    if g.mode_fdm == FDM_COM4_JACOBIAN:
        g.lhs, g.rhs, g.nb_diag, g.coef = c1n4_jacobian(dx, periodic)
    elif g.mode_fdm in (FDM_COM6_JACOBIAN, FDM_COM6_JACOBIAN_HYPER):
        g.lhs, g.rhs, g.nb_diag, g.coef = c1n6_jacobian(dx, periodic)
    elif g.mode_fdm == FDM_COM6_JACOBIAN_PENTA:
        g.lhs, g.rhs, g.nb_diag, g.coef = c1n6_jacobian_penta(dx, periodic)
    else:
        raise NotImplementedError("direct schemes are outside the oracle's scope")
    if periodic:
        wn = _wavenumbers(g.size)
        c = _pad1(g.coef)
        # note: cos(wn) (not cos(2 wn)) multiplies coef(2), as in the reference (:207)
        g.mwn = (2.0 * (c[3] * np.sin(wn) + c[4] * np.sin(2.0 * wn) + c[5] * np.sin(3.0 * wn))
                 / (1.0 + 2.0 * c[1] * np.cos(wn) + 2.0 * c[2] * np.cos(wn)))
    else:
        g.mwn = np.zeros(g.size)


def der1_initialize(x, dx, g, periodic, bcs_cases):
    """fdm_derivative.f90:63-142."""
    der1_create_system(x, dx, g, periodic)
    n = g.size
    ndl, ndr = g.nb_diag
    if g.periodic:
        g.lu = np.zeros((n + 1, ndl + 2 + 1))
        g.lu[:, 1:ndl + 1] = g.lhs[:, 1:ndl + 1]
        cols = [g.lu[1:, k] for k in range(1, ndl + 3)]
        if ndl == 3:
            tridpfs(*cols)
        else:
            pentadpfs(*cols)
    else:
        g.lu = np.zeros((n + 1, 5 * 4 + 1))
        for ib, case in enumerate(bcs_cases, start=1):
            ip = (ib - 1) * 5
            blk = g.lu[:, ip:ip + ndl + 1]          # 1-based padded view (col 0 is scratch)
            save0 = blk[:, 0].copy()
            blk[:, 1:] = g.lhs[:, 1:ndl + 1]
            fdm_bcs_neumann(case, blk, g.rhs[:, :ndr + 1], g.rhs_b, g.rhs_t)
            blk[:, 0] = save0
            nmin, nmax = 1, n
            if case in (BCS_ND, BCS_NN):
                nmin += 1
            if case in (BCS_DN, BCS_NN):
                nmax -= 1
            cols = [g.lu[nmin:nmax + 1, ip + k] for k in range(1, ndl + 1)]
            if ndl == 3:
                tridfs(*cols)
            else:
                pentadfs2(*cols)


def der1_solve(ibc, g, lu1, u):
    """fdm_derivative.f90:218-278.  u(n, nlines) -> result(n, nlines)."""
    n = g.size
    ndl, ndr = g.nb_diag
    result = np.zeros_like(u)
    ibc_loc = ibc
    ip = ibc_loc * 5
    if g.periodic:
        ibc_loc = BCS_PERIODIC
    nmin, nmax = 1, n
    if ibc_loc in (BCS_ND, BCS_NN):
        result[0] = 0.0
        nmin += 1
    if ibc_loc in (BCS_DN, BCS_NN):
        result[n - 1] = 0.0
        nmax -= 1
    mm = {3: matmul_3d_antisym, 5: matmul_5d_antisym, 7: matmul_7d_antisym}[ndr]
    mm(g.rhs, u, result, ibc_loc, g.rhs_b, g.rhs_t)
    if g.periodic:
        cols = [lu1[1:, k] for k in range(1, ndl + 3)]
        if ndl == 3:
            tridpss(*cols, result)
        else:
            pentadpss(*cols, result)
    else:
        cols = [lu1[nmin:nmax + 1, ip + k] for k in range(1, ndl + 1)]
        if ndl == 3:
            tridss(*cols, result[nmin - 1:nmax])
        else:
            pentadss2(*cols, result[nmin - 1:nmax])
    return result


def der2_create_system(x, dx1, dx2, g, periodic, uniform):
    """fdm_derivative.f90:337-409."""
    g.size = len(x)
    g.periodic = periodic
    if g.mode_fdm == FDM_COM4_JACOBIAN:
        g.lhs, g.rhs, g.nb_diag, g.coef = c2n4_jacobian(dx1, dx2, periodic)
    elif g.mode_fdm in (FDM_COM6_JACOBIAN, FDM_COM6_JACOBIAN_PENTA):
        g.lhs, g.rhs, g.nb_diag, g.coef = c2n6_jacobian(dx1, dx2, periodic)
    elif g.mode_fdm == FDM_COM6_JACOBIAN_HYPER:
        g.lhs, g.rhs, g.nb_diag, g.coef = c2n6_hyper_jacobian(dx1, dx2, periodic)
    elif g.mode_fdm == FDM_COM6_DIRECT:
        # FDM_C2N6_Direct(g%size, x, g%lhs, g%rhs, g%nb_diag); g%need_1der = .false.  (fdm_derivative.f90:381-383)
        from .fdm_direct import c2n6_direct
        g.lhs, g.rhs, g.nb_diag = c2n6_direct(x)
        g.coef = np.zeros(5)
    else:
        raise NotImplementedError("CompactDirect4 is outside the oracle's scope")
    if g.mode_fdm == FDM_COM6_DIRECT:
        g.need_1der = False
    elif not uniform:
        g.need_1der = True
    if periodic:
        wn = _wavenumbers(g.size)
        c = _pad1(g.coef)
        g.mwn = (2.0 * (c[3] * (1.0 - np.cos(wn)) + c[4] * (1.0 - np.cos(2.0 * wn)) + c[5] * (1.0 - np.cos(3.0 * wn)))
                 / (1.0 + 2.0 * c[1] * np.cos(wn) + 2.0 * c[2] * np.cos(2.0 * wn)))
    else:
        g.mwn = np.zeros(g.size)


def der2_initialize(x, dx1, dx2, g, periodic, uniform):
    """fdm_derivative.f90:282-333."""
    der2_create_system(x, dx1, dx2, g, periodic, uniform)
    n = g.size
    ndl, ndr = g.nb_diag
    assert ndl == 3
    if g.periodic:
        g.lu = np.zeros((n + 1, ndl + 2 + 1))
        g.lu[:, 1:ndl + 1] = g.lhs[:, 1:ndl + 1]
        tridpfs(*[g.lu[1:, k] for k in range(1, 6)])
    else:
        g.lu = np.zeros((n + 1, ndl + 1))
        g.lu[:, 1:ndl + 1] = g.lhs[:, 1:ndl + 1]
        tridfs(*[g.lu[1:, k] for k in range(1, 4)])


def der2_solve(g, lu, u, du):
    """fdm_derivative.f90:413-459."""
    ndl, ndr = g.nb_diag
    result = np.zeros_like(u)
    ibc = BCS_PERIODIC if g.periodic else BCS_DD
    if g.mode_fdm == FDM_COM6_DIRECT:
        matmul_5d(g.rhs[:, :ndr + 1], u, result, ibc)           # g%matmul => MatMul_5d (fdm_derivative.f90:318-323)
    else:
        mm = {5: matmul_5d_sym, 7: matmul_7d_sym}[ndr]
        mm(g.rhs[:, :ndr + 1], u, result, ibc)
    if g.need_1der:
        ip = ndr
        sub = np.zeros((g.size + 1, 4))
        sub[:, 1:] = g.rhs[:, ip + 1:ip + 4]
        matmul_3d_add(sub, du, result)
    if g.periodic:
        tridpss(*[lu[1:, k] for k in range(1, 6)], result)
    else:
        tridss(*[lu[1:, k] for k in range(1, 4)], result)
    return result


# ###########################################################################
# fdm.f90:14-29, 143-252
class Plan:
    """fdm_dt: plan for one direction."""

    def __init__(self, nodes, periodic, uniform=None, mode1=FDM_COM6_JACOBIAN, mode2=FDM_COM6_JACOBIAN_HYPER,
                 name='x'):
        self.name = name
        self.x_nodes = np.asarray(nodes, dtype=np.float64).copy()
        self.size = len(self.x_nodes)
        self.periodic = bool(periodic)
        self.uniform = bool(periodic) if uniform is None else bool(uniform)
        self.der1 = Derivative(mode1)
        self.der2 = Derivative(mode2)
        self.scale = 1.0
        self.nodes = None
        self.jac = None
        create_plan(self)


def create_plan(g):
    """FDM_CreatePlan, fdm.f90:143-252."""
    x = g.x_nodes
    nx = g.size
    if g.periodic and g.der2.mode_fdm == FDM_COM6_DIRECT:
        g.der2.mode_fdm = FDM_COM6_JACOBIAN_HYPER                  # they are the same for uniform grids (fdm.f90:158)
    if nx > 1:
        g.scale = x[nx - 1] - x[0]
        if g.periodic:
            g.scale = g.scale * (1.0 + 1.0 / float(nx - 1))
    else:
        g.scale = 1.0
    g.jac = np.zeros((nx + 1, 4))   # jac(1:nx, 1:3), 1-based padded
    if nx == 1:
        g.jac[:, :] = 1.0
        g.nodes = x.copy()
        return

    # first-order derivative: Jacobian from the scheme on a unit grid
    g.nodes = np.array([float(i - 1) for i in range(1, nx + 1)])
    g.jac[1:, 1] = 1.0
    der1_initialize(g.nodes, g.jac[:, 1].copy(), g.der1, periodic=False, bcs_cases=[BCS_DD])
    g.der1.periodic = False
    g.jac[1:, 1] = der1_solve(BCS_NONE, g.der1, g.der1.lu, x.reshape(nx, 1))[:, 0]

    g.nodes = x.copy()
    der1_initialize(g.nodes, g.jac[:, 1].copy(), g.der1, periodic=g.periodic,
                    bcs_cases=[BCS_DD, BCS_ND, BCS_DN, BCS_NN])
    if g.periodic:
        g.der1.mwn = g.der1.mwn / g.jac[1, 1]

    # second-order derivative
    g.nodes = np.array([float(i - 1) for i in range(1, nx + 1)])
    g.jac[1:, 2] = 1.0
    g.jac[1:, 3] = 0.0
    der2_initialize(g.nodes, g.jac[:, 2].copy(), g.jac[:, 3].copy(), g.der2, periodic=False, uniform=True)
    g.der2.periodic = False
    g.jac[1:, 3] = der2_solve(g.der2, g.der2.lu, x.reshape(nx, 1), g.jac[1:, 2].reshape(nx, 1).copy())[:, 0]

    g.nodes = x.copy()
    g.jac[:, 2] = g.jac[:, 1]
    der2_initialize(g.nodes, g.jac[:, 2].copy(), g.jac[:, 3].copy(), g.der2, g.periodic, g.uniform)
    if g.der2.periodic:
        g.der2.mwn = g.der2.mwn / (g.jac[1, 1] ** 2)
