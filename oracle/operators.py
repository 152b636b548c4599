"""Oracle: OPR_Partial, OPR_Burgers, OPR_Poisson, BOUNDARY_BCS_NEUMANN_Y (numpy).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows /root/reference/src:
  operators/opr_partial.f90   OPR_Partial_X (:31-150), _Z (:154-262), _Y (:266-377)
  physics/opr_burgers.f90     OPR_Burgers_Initialize (:52-115), _X/_Y/_Z (:190-431), _1D (:439-521)
  operators/opr_elliptic.f90  OPR_Elliptic_Initialize (:86-250), OPR_Poisson_FourierXZ_Factorize (:263-364)
  operators/opr_fourier.f90   OPR_Fourier_X/Z_Forward/Backward (:219-433) -- layouts and sign only;
                              FFTW3 (external, un-pinned) is replaced by numpy's pocketfft.
  tools/dns/boundary_bcs.f90  BOUNDARY_BCS_NEUMANN_Y (:368-473)

Fields are numpy arrays of shape (nz, ny, nx), C order, i.e. exactly the memory
layout of the Fortran a(nx, ny, nz).  The reference's local transposes
(TLab_Transpose) only re-order memory; the oracle reshapes instead, the values
entering every line solve are the same.
"""
import numpy as np

from . import fdm
from . import integral as I
from .fdm import BCS_DD, BCS_NN, BCS_ND, BCS_DN, BCS_MIN, BCS_MAX

OPR_P1, OPR_P2, OPR_P2_P1 = 1, 2, 3
OPR_B_SELF, OPR_B_U_IN = 0, 1


def to_lines(a, idir):
    """(nz, ny, nx) -> (n, nlines) with the derivative direction first."""
    nz, ny, nx = a.shape
    if idir == 0:
        return np.ascontiguousarray(a.reshape(nz * ny, nx).T)
    if idir == 1:
        return np.ascontiguousarray(a.transpose(1, 0, 2).reshape(ny, nz * nx))
    return np.ascontiguousarray(a.reshape(nz, ny * nx))


def from_lines(l, idir, shape):
    nz, ny, nx = shape
    if idir == 0:
        return np.ascontiguousarray(l.T.reshape(nz, ny, nx))
    if idir == 1:
        return np.ascontiguousarray(l.reshape(ny, nz, nx).transpose(1, 0, 2))
    return np.ascontiguousarray(l.reshape(nz, ny, nx))


def opr_partial(idir, type_, bcs, g, u):
    """OPR_Partial_{X,Y,Z}.  bcs is the Fortran bcs(2,2) given as [[b11, b12], [b21, b22]].
    Returns result (and tmp1 = first derivative for OPR_P2_P1)."""
    if g.size == 1:
        z = np.zeros_like(u)
        return (z, np.zeros_like(u)) if type_ == OPR_P2_P1 else z
    ibc = bcs[0][0] + bcs[1][0] * 2
    ul = to_lines(u, idir)
    if type_ == OPR_P1:
        return from_lines(fdm.der1_solve(ibc, g.der1, g.der1.lu, ul), idir, u.shape)
    if type_ == OPR_P2:
        du = fdm.der1_solve(ibc, g.der1, g.der1.lu, ul) if g.der2.need_1der else ul
        return from_lines(fdm.der2_solve(g.der2, g.der2.lu, ul, du), idir, u.shape)
    if type_ == OPR_P2_P1:
        du = fdm.der1_solve(ibc, g.der1, g.der1.lu, ul)
        d2 = fdm.der2_solve(g.der2, g.der2.lu, ul, du)
        return from_lines(d2, idir, u.shape), from_lines(du, idir, u.shape)
    raise ValueError(type_)


# ###########################################################################
class Burgers:
    """fdmDiffusion(ig)%lu(:, :, 0:inb_scal), opr_burgers.f90:77-115."""

    def __init__(self, g, visc, schmidt):
        self.g = g
        self.lu = []
        for ig in range(3):
            per_is = []
            if g[ig].size > 1:
                assert g[ig].der2.nb_diag[0] == 3
                for is_ in range(0, len(schmidt) + 1):
                    dummy = visc if is_ == 0 else visc / schmidt[is_ - 1]
                    lu = g[ig].der2.lu.copy()
                    if g[ig].periodic:
                        lu[:, 2] = g[ig].der2.lu[:, 2] * dummy
                        lu[:, 4] = g[ig].der2.lu[:, 4] / dummy
                    else:
                        lu[:, 2] = g[ig].der2.lu[:, 2] * dummy
                        lu[:, 3] = g[ig].der2.lu[:, 3] / dummy
                    per_is.append(lu)
            self.lu.append(per_is)

    def apply(self, idir, is_, bcs, s, u):
        """OPR_Burgers_{X,Y,Z}: result = (visc/Sc) d2s - u ds along idir."""
        g = self.g[idir]
        if g.size == 1:
            return np.zeros_like(s)
        assert bcs[0][1] + bcs[1][1] == 0
        ibc = bcs[0][0] + bcs[1][0] * 2
        sl = to_lines(s, idir)
        vl = to_lines(u, idir)
        dsdx = fdm.der1_solve(ibc, g.der1, g.der1.lu, sl)
        res = fdm.der2_solve(g.der2, self.lu[idir][is_], sl, dsdx)
        res = res - vl * dsdx
        return from_lines(res, idir, s.shape)


# ###########################################################################
class Elliptic:
    """OPR_Elliptic_Initialize, opr_elliptic.f90:86-250 (TYPE_FACTORIZE, serial)."""

    def __init__(self, g):
        self.g = g
        gx, gy, gz = g
        self.nx, self.ny, self.nz = gx.size, gy.size, gz.size
        self.isize_line = self.nx // 2 + 1
        self.norm = 1.0 / float(gx.size * gz.size)
        self.i_sing = [1, gx.size // 2 + 1]
        self.k_sing = [1, gz.size // 2 + 1]
        lam = np.zeros((self.nz, self.isize_line))
        for i in range(1, self.isize_line + 1):
            for k in range(1, self.nz + 1):
                if gz.size > 1:
                    lam[k - 1, i - 1] = gx.der1.mwn[i - 1] ** 2.0 + gz.der1.mwn[k - 1] ** 2.0
                else:
                    lam[k - 1, i - 1] = gx.der1.mwn[i - 1] ** 2.0
        self.lam = lam
        self.fdm_loc = gy       # same schemes as g(2): identical plan (opr_elliptic.f90:109-110,125)

    def is_sing(self):
        m = np.zeros((self.nz, self.isize_line), dtype=bool)
        for i in self.i_sing:
            for k in self.k_sing:
                m[k - 1, i - 1] = True
        return m


def opr_poisson(ell, p, bcs_hb, bcs_ht, ibc=BCS_NN):
    """OPR_Poisson_FourierXZ_Factorize (opr_elliptic.f90:263-364), BCS_NN (wall-normal derivatives given) or BCS_DD (values
    given).  p(nz, ny, nx) forcing; bcs_hb, bcs_ht (nz, nx).  Returns (p, dpdy)."""
    assert ibc in (BCS_NN, BCS_DD)
    nz, ny, nx = p.shape
    p = p.copy()
    p[:, 0, :] = bcs_hb
    p[:, ny - 1, :] = bcs_ht
    c = np.fft.rfft(p, axis=2)                     # OPR_Fourier_X_Forward (r2c, unnormalised)
    if nz > 1:
        c = np.fft.fft(c, axis=0)                  # OPR_Fourier_Z_Forward (c2c, sign -1)
    c = c * ell.norm
    der1 = ell.fdm_loc.der1
    sing = ell.is_sing()
    out_u = np.zeros_like(c)
    out_v = np.zeros_like(c)

    for mask, singular in ((~sing, False), (sing, True)):
        kk, ii = np.nonzero(mask)
        if len(kk) == 0:
            continue
        lam = ell.lam[kk, ii]
        M = len(kk)
        fi = {BCS_MIN: I.int1_initialize(der1, np.sqrt(lam), BCS_MIN),
              BCS_MAX: I.int1_initialize(der1, -np.sqrt(lam), BCS_MAX)}
        # the reference stores lambda(k,i) in fdm_int1(BCS_MIN)%lambda = sqrt(lambda): opr_elliptic.f90:205-206
        cm = c[kk, :, ii]                          # (M, ny)
        f = np.zeros((ny, 2, M))
        f[:, 0, :] = cm.real.T
        f[:, 1, :] = cm.imag.T
        bcs = np.zeros((2, 2, M))
        bcs[0] = f[0]
        bcs[1] = f[ny - 1]
        if ibc == BCS_DD:
            if singular:
                u, v = I.ode2_factorize_dd_sing(fi, f, bcs)
            else:
                u, v = I.ode2_factorize_dd(fi, fi[BCS_MIN].rhs, fi[BCS_MAX].rhs, f, bcs)
        elif singular:
            u, v = I.ode2_factorize_nn_sing(fi, f, bcs)
        else:
            u, v = I.ode2_factorize_nn(fi, fi[BCS_MIN].rhs, fi[BCS_MAX].rhs, f, bcs)
        out_u[kk, :, ii] = (u[:, 0, :] + 1j * u[:, 1, :]).T
        out_v[kk, :, ii] = (v[:, 0, :] + 1j * v[:, 1, :]).T

    def backward(cc):
        if nz > 1:
            cc = np.fft.ifft(cc, axis=0, norm="forward")       # unnormalised, sign +1
        return np.fft.irfft(cc, n=nx, axis=2, norm="forward")  # c2r, unnormalised

    return backward(out_u), backward(out_v)


# ###########################################################################
def boundary_bcs_neumann_y(ibc, g, u):
    """BOUNDARY_BCS_NEUMANN_Y (boundary_bcs.f90:368-473).  u(nz, ny, nx) -> bcs_hb, bcs_ht (nz, nx)."""
    nz, ny, nx = u.shape
    bcs_hb = np.zeros((nz, nx))
    bcs_ht = np.zeros((nz, nx))
    if g.size == 1:
        return bcs_hb, bcs_ht
    d = g.der1
    ndl, ndr = d.nb_diag
    ul = to_lines(u, 1)
    dst = np.zeros_like(ul)
    ip = ibc * 5
    nmin, nmax = 1, g.size
    if ibc in (BCS_ND, BCS_NN):
        dst[0] = 0.0
        nmin += 1
    if ibc in (BCS_DN, BCS_NN):
        dst[ny - 1] = 0.0
        nmax -= 1
    mm = {3: fdm.matmul_3d_antisym, 5: fdm.matmul_5d_antisym, 7: fdm.matmul_7d_antisym}[ndr]
    hb, ht = mm(d.rhs, ul, dst, ibc, d.rhs_b, d.rhs_t, want_bcs=True)
    cols = [d.lu[nmin:nmax + 1, ip + k] for k in range(1, ndl + 1)]
    if ndl == 3:
        fdm.tridss(*cols, dst[nmin - 1:nmax])
    else:
        fdm.pentadss2(*cols, dst[nmin - 1:nmax])
    idl = ndl // 2 + 1
    if ibc in (BCS_ND, BCS_NN):
        for ic in range(1, idl):
            hb = hb + d.lu[1, ip + idl + ic] * dst[ic]
        bcs_hb = hb.reshape(nz, nx)
    if ibc in (BCS_DN, BCS_NN):
        for ic in range(1, idl):
            ht = ht + d.lu[ny, ip + idl - ic] * dst[ny - 1 - ic]
        bcs_ht = ht.reshape(nz, nx)
    return bcs_hb, bcs_ht
