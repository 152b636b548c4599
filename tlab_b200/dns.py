"""Host-side mirror of tools/dns' time advance (TIME_RUNGEKUTTA and friends) over the C ABI."""
import ctypes

import numpy as np

from . import lib as _lib

RKM_EXP3, RKM_EXP4 = 3, 4
DNS_BCS_DIRICHLET, DNS_BCS_NEUMANN = 3, 4
MAX_SCAL = 8
_BUOY = {"none": 0, "homogeneous": 1, "linear": 2}


class DnsParams(ctypes.Structure):
    """tlab_dns_params of include/tlab_gpu.h"""
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
                ("nscal", ctypes.c_int), ("rkm_mode", ctypes.c_int), ("buoyancy_type", ctypes.c_int),
                ("scal_limit", ctypes.c_int),
                ("bcs_flow_jmin", ctypes.c_int * 3), ("bcs_flow_jmax", ctypes.c_int * 3),
                ("bcs_scal_jmin", ctypes.c_int * MAX_SCAL), ("bcs_scal_jmax", ctypes.c_int * MAX_SCAL),
                ("visc", ctypes.c_double), ("schmidt", ctypes.c_double * MAX_SCAL),
                ("buoyancy_params", ctypes.c_double * 2), ("buoyancy_vector", ctypes.c_double * 3),
                ("scal_min", ctypes.c_double * MAX_SCAL), ("scal_max", ctypes.c_double * MAX_SCAL)]


class Dns:
    """Device-resident q, s, hq, hs and the explicit low-storage Runge-Kutta advance."""

    def __init__(self, g, visc, schmidt, rkm_mode=RKM_EXP4, buoyancy_type="none", buoyancy_params=(0.0, 0.0),
                 buoyancy_vector=(0.0, 0.0, 0.0), bbackground=None,
                 bcs_flow_jmin=(DNS_BCS_DIRICHLET,) * 3, bcs_flow_jmax=(DNS_BCS_DIRICHLET,) * 3,
                 bcs_scal_jmin=None, bcs_scal_jmax=None, scal_limit=True, scal_min=0.0, scal_max=1.0, kmax=None):
        """kmax: thickness of this rank's z slab (None: the whole domain, single process)."""
        L = _lib.load()
        self.g = g
        self.nx, self.ny, self.nz = g[0].size, g[1].size, (g[2].size if kmax is None else int(kmax))
        self.inb_scal = len(schmidt)
        p = DnsParams()
        p.nx, p.ny, p.nz, p.nscal, p.rkm_mode = self.nx, self.ny, self.nz, self.inb_scal, rkm_mode
        p.buoyancy_type = _BUOY[buoyancy_type]
        p.scal_limit = int(bool(scal_limit))
        p.visc = visc
        for i in range(3):
            p.bcs_flow_jmin[i] = bcs_flow_jmin[i]
            p.bcs_flow_jmax[i] = bcs_flow_jmax[i]
            p.buoyancy_vector[i] = buoyancy_vector[i]
        for i in range(self.inb_scal):
            p.schmidt[i] = schmidt[i]
            p.bcs_scal_jmin[i] = (bcs_scal_jmin or (DNS_BCS_DIRICHLET,) * MAX_SCAL)[i]
            p.bcs_scal_jmax[i] = (bcs_scal_jmax or (DNS_BCS_DIRICHLET,) * MAX_SCAL)[i]
            p.scal_min[i] = scal_min
            p.scal_max[i] = scal_max
        p.buoyancy_params[0], p.buoyancy_params[1] = buoyancy_params[0], buoyancy_params[1]
        bb = None
        if bbackground is not None:
            bb = np.ascontiguousarray(bbackground, dtype=np.float64)
            assert bb.size == self.ny
        h = ctypes.c_void_p()
        _lib.check(L.tlab_dns_create(ctypes.byref(p), g[0].handle, g[1].handle, g[2].handle,
                                     bb.ctypes.data_as(ctypes.c_void_p) if bb is not None else None, ctypes.byref(h)))
        self.handle = h
        self.params = p
        self.shape = (self.nz, self.ny, self.nx)

    def set(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        assert a.shape == self.shape
        _lib.check(_lib.load().tlab_dns_upload_host(self.handle, name.encode(), a.ctypes.data_as(ctypes.c_void_p)))

    def get(self, name):
        a = np.empty(self.shape)
        _lib.check(_lib.load().tlab_dns_download_host(self.handle, name.encode(), a.ctypes.data_as(ctypes.c_void_p)))
        return a

    def read_restart(self, flow_name, scal_name=None, koff=0, nz_total=None):
        """Load tlab restart files (`flow.<it>.1..3`, `scal.<it>.1..ns`, tlab_b200/io.py) into q and s; with `nz_total`
        this rank takes its z-slab `[koff, koff + nz)` of the global field.  Returns (nt, rtime)."""
        from . import io as tio
        nzt = self.nz if nz_total is None else int(nz_total)
        q, nt, params = tio.read_fields(flow_name, self.nx, self.ny, nzt, 3, koff=koff, kmax=self.nz)
        for i in range(3):
            self.set("q%d" % (i + 1), q[i])
        if scal_name is not None and self.inb_scal > 0:
            s, _, _ = tio.read_fields(scal_name, self.nx, self.ny, nzt, self.inb_scal, koff=koff, kmax=self.nz)
            for i in range(self.inb_scal):
                self.set("s%d" % (i + 1), s[i])
        return nt, (float(params[0]) if len(params) else 0.0)

    def write_restart(self, flow_name, scal_name, nt, rtime, visc, schmidt=(), koff=0, nz_total=None):
        """Write q and s as tlab restart files (IO_Write_Fields with the headers of tlab_consistency_check.f90:148-163)."""
        from . import io as tio
        tio.write_fields(flow_name, nt, [self.get("q%d" % (i + 1)) for i in range(3)], tio.flow_params(rtime, visc),
                         koff=koff, nz_total=nz_total)
        if self.inb_scal > 0:
            tio.write_fields(scal_name, nt, [self.get("s%d" % (i + 1)) for i in range(self.inb_scal)],
                             [tio.scal_params(rtime, visc, schmidt[i]) for i in range(self.inb_scal)],
                             koff=koff, nz_total=nz_total)

    def device_ptr(self, name):
        p = ctypes.c_void_p()
        _lib.check(_lib.load().tlab_dns_field(self.handle, name.encode(), ctypes.byref(p)))
        return p.value

    def rhs(self, dte):
        _lib.check(_lib.load().tlab_rhs_global_incompressible_1(self.handle, float(dte)))

    def substep(self, dte, kco=0.0, scale_h=False):
        _lib.check(_lib.load().tlab_time_substep(self.handle, float(dte), float(kco), int(scale_h)))

    def runge_kutta_stage(self, dtime, stage):
        _lib.check(_lib.load().tlab_time_rungekutta_stage(self.handle, float(dtime), int(stage)))

    def runge_kutta(self, dtime):
        """TIME_RUNGEKUTTA (time.f90:185-333)"""
        _lib.check(_lib.load().tlab_time_rungekutta(self.handle, float(dtime)))

    def runge_kutta_host(self, dtime, q_host, s_host):
        """One full step from/to host arrays q(3,N), s(nscal,N) (numpy or pinned torch memory)."""
        _lib.check(_lib.load().tlab_time_rungekutta_host(self.handle, float(dtime), ctypes.c_void_p(q_host),
                                                         ctypes.c_void_p(s_host)))

    def courant(self, cfla, cfld=None, prandtl=1.0, dtime=0.0):
        """TIME_COURANT (time.f90:365-548): returns (dtime, CFL number, diffusion number)."""
        cfld = 0.25 * cfla if cfld is None else cfld          # dns_read_local.f90:72
        dt, c1, c2 = ctypes.c_double(dtime), ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.load().tlab_time_courant(self.handle, float(cfla), float(cfld), float(prandtl), ctypes.byref(dt),
                                                 ctypes.byref(c1), ctypes.byref(c2)))
        return dt.value, c1.value, c2.value

    def bounds_control(self):
        """DNS_BOUNDS_CONTROL dilatation: (DilMin, DilMax) = (min div u, max div u)."""
        a, b = ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.load().tlab_dns_bounds_control(self.handle, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def launch_count(self):
        c = ctypes.c_longlong()
        _lib.check(_lib.load().tlab_dns_launch_count(self.handle, ctypes.byref(c)))
        return c.value

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().tlab_dns_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rk_coefficients(mode):
    kdt = (ctypes.c_double * 5)()
    ktime = (ctypes.c_double * 5)()
    kco = (ctypes.c_double * 5)()
    n = ctypes.c_int()
    _lib.check(_lib.load().tlab_time_rk_coefficients(mode, kdt, ktime, kco, ctypes.byref(n)))
    return list(kdt)[:n.value], list(ktime)[:n.value], list(kco)[:n.value - 1]
