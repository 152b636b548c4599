"""CPU baseline of the RK substep: driver of oracle/cpp/tlab_cpu.cpp (C++17 + OpenMP).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Nothing under tlab_b200/ imports this module; bench.py uses it for the
`cpu_baseline` leg and for `--impl reference` (the image has no Fortran compiler and no FFTW, so the reference itself cannot be
built: this is a restatement, kind "port").

Division of labour: the numpy oracle builds every table exactly as the reference does (scheme coefficients, TRIDFS / TRIDPFS
factors, the diffusivity-scaled factors of OPR_Burgers_Initialize, FDM_Int1_Initialize for every eigenvalue of the Poisson
problem) -- that is initialisation, untimed in the reference's own profiling as well -- and the C++ library runs the sweeps of a
substep on all host cores.  The x / z transforms use scipy's pocketfft with `workers` threads (the reference uses FFTW3).
The class mirrors oracle.dns.Dns statement by statement; tests/test_cpu_baseline.py checks it against that oracle.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

from . import fdm
from . import integral as I
from .fdm import BCS_MIN, BCS_MAX, BCS_PERIODIC, BCS_DD, BCS_ND, BCS_DN, BCS_NN
from .dns import rk_coefficients, RKM_EXP4, DNS_BCS_DIRICHLET, DNS_BCS_NEUMANN
from .operators import Burgers, Elliptic

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "tlab_cpu.cpp")
BW = 8


class Band(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("periodic", ctypes.c_int), ("sym", ctypes.c_int), ("nb", ctypes.c_int),
                ("rc", ctypes.c_double), ("r2", ctypes.c_double), ("r3", ctypes.c_double),
                ("bot", (ctypes.c_double * BW) * 4), ("top", (ctypes.c_double * BW) * 4)]


class Tri(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("periodic", ctypes.c_int), ("nmin", ctypes.c_int), ("nmax", ctypes.c_int),
                ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("c", ctypes.c_void_p), ("d", ctypes.c_void_p),
                ("e", ctypes.c_void_p)]


def _cpu_tag():
    try:
        flags = [l for l in open("/proc/cpuinfo") if l.startswith("flags")][0]
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def build_lib(verbose=False):
    """g++ -O3 -march=native -fopenmp, compiled on the machine that runs it (the tag is a hash of the CPU flags, so a library
    built in another container is not reused on a host with a different instruction set)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libtlab_cpu_%s.so" % _cpu_tag())
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(SRC):
        cmd = ["g++", "-std=c++17", "-O3", "-march=native", "-fopenmp", "-fno-math-errno", "-shared", "-fPIC", SRC, "-o", lib]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout)
        if verbose:
            print(" ".join(cmd))
    return lib


_LIB = None


def load():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_lib())
        L.cpu_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ---------------------------------------------------------------------------------------------------------------------
def _probe_band(apply, n, periodic):
    """Band of a banded product `apply(u(n, 1)) -> f(n, 1)` of the oracle: interior stencil from a unit vector in the middle of the
    line, dense rows at the ends from unit vectors next to the walls (exact coefficients, whatever special cases the product has)."""
    b = Band()
    b.n, b.periodic = n, int(periodic)
    k = n // 2
    e = np.zeros((n, 1))
    e[k, 0] = 1.0
    col = apply(e)[:, 0]
    c1 = col[k - 1]                      # coefficient of u(i+1) in row i
    assert abs(c1 - 1.0) < 1e-14, c1     # the reference normalises the first off-diagonal to 1
    b.sym = int(abs(col[k + 1] - col[k - 1]) < 1e-14)
    b.rc = float(col[k]) if b.sym else 0.0
    b.r2 = float(col[k - 2])
    b.r3 = float(col[k - 3])
    assert abs(col[k - 4]) == 0.0
    b.nb = 0 if periodic else 4
    if not periodic:
        for kk in range(BW):
            e = np.zeros((n, 1))
            e[kk, 0] = 1.0
            col = apply(e)[:, 0]
            for i in range(4):
                b.bot[i][kk] = float(col[i])
            e = np.zeros((n, 1))
            e[n - 1 - kk, 0] = 1.0
            col = apply(e)[:, 0]
            for q in range(4):
                b.top[q][kk] = float(col[n - 1 - q])
    return b


class _Keep:
    """owns the numpy arrays a Tri points to"""

    def __init__(self):
        self.arrays = []

    def tri(self, n, periodic, cols, nmin=1, nmax=None):
        t = Tri()
        t.n, t.periodic, t.nmin, t.nmax = n, int(periodic), nmin, n if nmax is None else nmax
        ptrs = []
        for c in cols:
            a = np.ascontiguousarray(c, dtype=np.float64)
            self.arrays.append(a)
            ptrs.append(a.ctypes.data)
        while len(ptrs) < 5:
            ptrs.append(None)
        t.a, t.b, t.c, t.d, t.e = ptrs
        return t


class CpuDns:
    """oracle.dns.Dns with the sweeps in C++/OpenMP (same constructor arguments, same fields q, s, hq, hs)."""

    def __init__(self, g, visc, schmidt, rkm_mode=RKM_EXP4, buoyancy_type='none', buoyancy_params=(0.0, 0.0),
                 buoyancy_vector=(0.0, 0.0, 0.0), bbackground=None, bcs_flow_jmin=(DNS_BCS_DIRICHLET,) * 3,
                 bcs_flow_jmax=(DNS_BCS_DIRICHLET,) * 3, bcs_scal_jmin=None, bcs_scal_jmax=None, scal_limit=True, scal_min=0.0,
                 scal_max=1.0, mode_chunk=8192):
        self.L = load()
        self.g = g
        self.nx, self.ny, self.nz = g[0].size, g[1].size, g[2].size
        self.N = self.nx * self.ny * self.nz
        self.visc, self.schmidt = visc, list(schmidt)
        self.inb_scal = len(self.schmidt)
        self.kdt, self.ktime, self.kco = rk_coefficients(rkm_mode)
        self.rkm_endstep = len(self.kdt)
        self.buoyancy_type, self.buoyancy_params, self.buoyancy_vector = buoyancy_type, buoyancy_params, buoyancy_vector
        self.bbackground = np.zeros(self.ny) if bbackground is None else np.asarray(bbackground, float)
        self.bcs_flow_jmin, self.bcs_flow_jmax = tuple(bcs_flow_jmin), tuple(bcs_flow_jmax)
        self.bcs_scal_jmin = tuple(bcs_scal_jmin) if bcs_scal_jmin else (DNS_BCS_DIRICHLET,) * self.inb_scal
        self.bcs_scal_jmax = tuple(bcs_scal_jmax) if bcs_scal_jmax else (DNS_BCS_DIRICHLET,) * self.inb_scal
        self.scal_limit, self.scal_min, self.scal_max = scal_limit, scal_min, scal_max
        self.threads = self.L.cpu_threads()
        shape = (self.nz, self.ny, self.nx)
        self.q = [np.zeros(shape) for _ in range(3)]
        self.s = [np.zeros(shape) for _ in range(self.inb_scal)]
        self.hq = [np.zeros(shape) for _ in range(3)]
        self.hs = [np.zeros(shape) for _ in range(self.inb_scal)]
        self.tmp1 = np.zeros(shape)
        self.tmp3 = np.zeros(shape)
        self.dte = 0.0
        self.keep = _Keep()
        self._line_tables()
        self._poisson_tables(mode_chunk)

    # ---- tables ----------------------------------------------------------------------------------------------------
    def _line_tables(self):
        g = self.g
        bur = Burgers(g, self.visc, self.schmidt)
        self.b1, self.t1, self.b2, self.t2, self.rd1 = [], [], [], [], []
        for d in range(3):
            gd = g[d]
            n = gd.size
            if n == 1:
                self.b1.append(None); self.t1.append(None); self.b2.append(None); self.t2.append(None); self.rd1.append(None)
                continue
            der1, der2 = gd.der1, gd.der2
            ndl, ndr = der1.nb_diag
            mm1 = {3: fdm.matmul_3d_antisym, 5: fdm.matmul_5d_antisym, 7: fdm.matmul_7d_antisym}[ndr]
            ibc = BCS_PERIODIC if gd.periodic else BCS_DD

            def apply1(u, mm1=mm1, der1=der1, ibc=ibc):
                f = np.zeros_like(u)
                mm1(der1.rhs, u, f, ibc, der1.rhs_b, der1.rhs_t)
                return f
            self.b1.append(_probe_band(apply1, n, gd.periodic))
            lu = der1.lu
            if gd.periodic:
                self.t1.append(self.keep.tri(n, True, [lu[1:, k] for k in range(1, 6)]))
            else:
                self.t1.append(self.keep.tri(n, False, [lu[1:, k] for k in range(1, 4)]))
            ndr2 = der2.nb_diag[1]
            mm2 = {5: fdm.matmul_5d_sym, 7: fdm.matmul_7d_sym}[ndr2]

            def apply2(u, mm2=mm2, der2=der2, ibc=ibc, ndr2=ndr2):
                f = np.zeros_like(u)
                mm2(der2.rhs[:, :ndr2 + 1], u, f, ibc)
                return f
            self.b2.append(_probe_band(apply2, n, gd.periodic))
            per_is = []
            for lu2 in bur.lu[d]:
                if gd.periodic:
                    per_is.append(self.keep.tri(n, True, [lu2[1:, k] for k in range(1, 6)]))
                else:
                    per_is.append(self.keep.tri(n, False, [lu2[1:, k] for k in range(1, 4)]))
            self.t2.append(per_is)
            if der2.need_1der:
                r = np.ascontiguousarray(der2.rhs[1:, ndr2 + 1:ndr2 + 4], dtype=np.float64)
                self.keep.arrays.append(r)
                self.rd1.append(r)
            else:
                self.rd1.append(None)
        # BOUNDARY_BCS_NEUMANN_Y: Neumann variants of the first derivative along y
        self.neu = {}
        gy = g[1]
        if gy.size > 1 and not gy.periodic:
            d = gy.der1
            ndl, ndr = d.nb_diag
            assert ndl == 3
            idl = ndl // 2 + 1
            mm1 = {3: fdm.matmul_3d_antisym, 5: fdm.matmul_5d_antisym, 7: fdm.matmul_7d_antisym}[ndr]
            n = gy.size
            for ibc in (BCS_ND, BCS_DN, BCS_NN):
                fb, ft = np.zeros(BW), np.zeros(BW)

                def applyn(u, ibc=ibc):
                    f = np.zeros_like(u)
                    mm1(d.rhs, u, f, ibc, d.rhs_b, d.rhs_t, want_bcs=True)
                    return f
                band = _probe_band(applyn, n, False)
                for kk in range(BW):
                    for side, arr in ((0, fb), (1, ft)):
                        e = np.zeros((n, 1))
                        e[kk if side == 0 else n - 1 - kk, 0] = 1.0
                        f = np.zeros_like(e)
                        hb, ht = mm1(d.rhs, e, f, ibc, d.rhs_b, d.rhs_t, want_bcs=True)
                        val = hb if side == 0 else ht
                        arr[kk] = float(val[0]) if val is not None else 0.0
                ip = ibc * 5
                nmin, nmax = 1, n
                if ibc in (BCS_ND, BCS_NN):
                    nmin += 1
                if ibc in (BCS_DN, BCS_NN):
                    nmax -= 1
                tri = self.keep.tri(n, False, [d.lu[nmin:nmax + 1, ip + k] for k in range(1, 4)], nmin, nmax)
                self.neu[ibc] = (band, tri, fb, ft, float(d.lu[1, ip + idl + 1]), float(d.lu[n, ip + idl - 1]))

    def _poisson_tables(self, chunk):
        ell = Elliptic(self.g)
        self.ell = ell
        nz, nxh, n = self.nz, ell.isize_line, self.ny
        M = nz * nxh
        lam = np.ascontiguousarray(ell.lam.reshape(M))
        self.lam = lam
        self.sing = np.ascontiguousarray(ell.is_sing().reshape(M).astype(np.uint8))
        der1 = ell.fdm_loc.der1
        self.Lmin = np.empty((M, n, 5))
        self.Lmax = np.empty((M, n, 5))
        self.rbmin, self.rtmin = np.empty((M, 10)), np.empty((M, 10))
        self.rbmax, self.rtmax = np.empty((M, 10)), np.empty((M, 10))
        for m0 in range(0, M, chunk):
            m1 = min(M, m0 + chunk)
            sq = np.sqrt(lam[m0:m1])
            for side, lam_s, Ls, rbs, rts in ((BCS_MIN, sq, self.Lmin, self.rbmin, self.rtmin),
                                              (BCS_MAX, -sq, self.Lmax, self.rbmax, self.rtmax)):
                fi = I.int1_initialize(der1, lam_s, side)
                Ls[m0:m1] = fi.lhs[1:, 1:6, :].transpose(2, 0, 1)
                rb, rt = fi.rhs_b, fi.rhs_t
                rbs[m0:m1, 0:3] = rb[1, 1:4, :].T
                rbs[m0:m1, 3:6] = rb[2, 1:4, :].T
                rbs[m0:m1, 6:10] = rb[3, 0:4, :].T
                rts[m0:m1, 0:4] = rt[0, 1:5, :].T
                rts[m0:m1, 4:7] = rt[1, 1:4, :].T
                rts[m0:m1, 7:10] = rt[2, 1:4, :].T
                rhs = np.ascontiguousarray(fi.rhs[1:, 1:4, 0])
                if side == BCS_MIN:
                    self.rhsmin = rhs
                else:
                    self.rhsmax = rhs

    # ---- operators ---------------------------------------------------------------------------------------------------
    def _burgers(self, d, is_, s, vel, out):
        if self.g[d].size == 1:
            return
        self.L.cpu_line_op(4, d, self.nx, self.ny, self.nz, ctypes.byref(self.b1[d]), ctypes.byref(self.t1[d]),
                           ctypes.byref(self.b2[d]), ctypes.byref(self.t2[d][is_]), _p(self.rd1[d]), _p(s), None,
                           ctypes.c_double(0.0), _p(vel), _p(out), 1)

    def _partial(self, d, u, u2, scale, out, accumulate):
        if self.g[d].size == 1:
            if accumulate == 0:
                out[...] = 0.0
            return
        self.L.cpu_line_op(1, d, self.nx, self.ny, self.nz, ctypes.byref(self.b1[d]), ctypes.byref(self.t1[d]), None, None,
                           None, _p(u), _p(u2), ctypes.c_double(scale), None, _p(out), accumulate)

    def _poisson(self, p, hb, ht, dpdy):
        """OPR_Poisson_FourierXZ_Factorize (opr_elliptic.f90:263-364): p forcing -> solution, dpdy"""
        import scipy.fft as sfft
        nz, ny, nx = p.shape
        self.L.cpu_set_planes(nx, ny, nz, _p(p), _p(hb), _p(ht))
        w = self.threads
        c = sfft.rfft(p, axis=2, workers=w)
        if nz > 1:
            c = sfft.fft(c, axis=0, workers=w, overwrite_x=True)
        c = np.ascontiguousarray(c)
        self.L.cpu_scale(ctypes.c_longlong(2 * c.size), ctypes.c_double(self.ell.norm), _p(c.view(np.float64)))
        cv = np.empty_like(c)
        self.L.cpu_poisson_modes(self.ell.isize_line, ny, nz, _p(c.view(np.float64)), _p(cv.view(np.float64)), _p(self.lam),
                                 _p(self.sing), _p(self.Lmin), _p(self.Lmax), _p(self.rbmin), _p(self.rtmin), _p(self.rbmax),
                                 _p(self.rtmax), _p(self.rhsmin), _p(self.rhsmax))
        for src, dst in ((c, p), (cv, dpdy)):
            if nz > 1:
                src = sfft.ifft(src, axis=0, norm="forward", workers=w, overwrite_x=True)
            dst[...] = sfft.irfft(src, n=nx, axis=2, norm="forward", workers=w)

    # ---- the substep (oracle.dns.Dns, statement by statement) -------------------------------------------------------------
    def sources_flow(self):
        if self.buoyancy_type == 'none':
            return
        assert self.buoyancy_type == 'linear'
        c1, c0 = self.buoyancy_params
        for iq in range(3):
            if abs(self.buoyancy_vector[iq]) > 0.0:
                self.L.cpu_buoyancy_linear(self.nx, self.ny, self.nz, ctypes.c_double(self.buoyancy_vector[iq]), ctypes.c_double(c1),
                                           ctypes.c_double(c0), _p(self.bbackground), _p(self.s[0]), _p(self.hq[iq]))

    def rhs_global_incompressible_1(self):
        u, v, w = self.q
        hq, hs, s = self.hq, self.hs, self.s
        B = self._burgers
        B(0, 0, u, u, hq[0]); B(1, 0, u, v, hq[0]); B(2, 0, u, w, hq[0])
        B(1, 0, v, v, hq[1]); B(0, 0, v, u, hq[1]); B(2, 0, v, w, hq[1])
        B(2, 0, w, w, hq[2]); B(0, 0, w, u, hq[2]); B(1, 0, w, v, hq[2])
        for is_ in range(self.inb_scal):
            B(0, is_ + 1, s[is_], u, hs[is_]); B(1, is_ + 1, s[is_], v, hs[is_]); B(2, is_ + 1, s[is_], w, hs[is_])
        dummy = 1.0 / self.dte
        tmp1, tmp3 = self.tmp1, self.tmp3
        self._partial(1, hq[1], v, dummy, tmp1, 0)
        self._partial(0, hq[0], u, dummy, tmp1, +1)
        self._partial(2, hq[2], w, dummy, tmp1, +1)
        nx, ny, nz = self.nx, self.ny, self.nz
        hb, ht = np.empty((nz, nx)), np.empty((nz, nx))
        self.L.cpu_get_planes(nx, ny, nz, _p(hq[1]), _p(hb), _p(ht))
        self._poisson(tmp1, hb, ht, tmp3)
        self._partial(0, tmp1, None, 0.0, hq[0], -1)
        self.L.cpu_sub(ctypes.c_longlong(self.N), _p(tmp3), _p(hq[1]))
        self._partial(2, tmp1, None, 0.0, hq[2], -1)
        for arr, tmin, tmax in ([(hq[i], self.bcs_flow_jmin[i], self.bcs_flow_jmax[i]) for i in range(3)] +
                                [(hs[i], self.bcs_scal_jmin[i], self.bcs_scal_jmax[i]) for i in range(self.inb_scal)]):
            ibc = (1 if tmin == DNS_BCS_NEUMANN else 0) + (2 if tmax == DNS_BCS_NEUMANN else 0)
            hb[...] = 0.0
            ht[...] = 0.0
            if ibc > 0:
                band, tri, fb, ft, lub, lut = self.neu[ibc]
                self.L.cpu_neumann_y(ibc, nx, ny, nz, ctypes.byref(band), ctypes.byref(tri), _p(fb), _p(ft), ctypes.c_double(lub),
                                     ctypes.c_double(lut), _p(arr), _p(hb), _p(ht))
            self.L.cpu_set_planes(nx, ny, nz, _p(arr), _p(hb), _p(ht))

    def substep(self, dte, kco=None):
        """TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT + DNS_BOUNDS_LIMIT + hq = hq*kco (time.f90:261-298, 559-670)"""
        self.dte = dte
        n = ctypes.c_longlong(self.N)
        self.sources_flow()
        self.rhs_global_incompressible_1()
        for iq in range(3):
            self.L.cpu_axpy(n, ctypes.c_double(dte), _p(self.hq[iq]), _p(self.q[iq]))
        for is_ in range(self.inb_scal):
            self.L.cpu_axpy(n, ctypes.c_double(dte), _p(self.hs[is_]), _p(self.s[is_]))
            if self.scal_limit:
                self.L.cpu_clip(n, ctypes.c_double(self.scal_min), ctypes.c_double(self.scal_max), _p(self.s[is_]))
        if kco is not None:
            for a in self.hq + self.hs:
                self.L.cpu_scale(n, ctypes.c_double(kco), _p(a))

    def runge_kutta(self, dtime):
        for a in self.hq + self.hs:
            self.L.cpu_zero(ctypes.c_longlong(self.N), _p(a))
        for sub in range(1, self.rkm_endstep + 1):
            self.substep(dtime * self.kdt[sub - 1], self.kco[sub - 1] if sub < self.rkm_endstep else None)
