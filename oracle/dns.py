"""Oracle: incompressible/Boussinesq RHS and low-storage Runge-Kutta (numpy).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows /root/reference/src:
  tools/dns/time.f90   TIME_INITIALIZE (:79-135), TIME_RUNGEKUTTA (:185-333),
                       TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT (:559-670)
  tools/dns/rhs_global_incompressible_1.f90 (:15-405), combined mode, remove_divergence,
                       no stagger / IBM / anelastic / filters / towers
  physics/tlab_sources.f90 TLab_Sources_Flow (:36-131) buoyancy only;
  physics/gravity.f90  Gravity_Buoyancy (:232-342) EQNS_BOD_LINEAR (one scalar) and HOMOGENEOUS
  tools/dns/dns_local.f90 DNS_BOUNDS_LIMIT (:67-90)

Not restated: BOUNDARY_BCS_SURFACE_Y is called by the reference for every scalar
(rhs_global_incompressible_1.f90:386-389 compares a DNS_BCS_* type with DNS_SFC_STATIC) but,
for the default SfcType=static, it only computes a derivative that is discarded.
"""
import numpy as np

from . import fdm
from .fdm import BCS_NN
from .operators import (OPR_P1, Burgers, Elliptic, opr_partial, opr_poisson, boundary_bcs_neumann_y)

RKM_EXP3, RKM_EXP4 = 3, 4
DNS_BCS_DIRICHLET, DNS_BCS_NEUMANN = 3, 4


def rk_coefficients(mode):
    """time.f90:86-112.  Returns kdt, ktime, kco (0-based lists)."""
    if mode == RKM_EXP3:
        kdt = [1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0]
        ktime = [0.0, 1.0 / 3.0, 3.0 / 4.0]
        kco = [-5.0 / 9.0, -153.0 / 128.0]
    elif mode == RKM_EXP4:
        kdt = [1432997174477.0 / 9575080441755.0,
               5161836677717.0 / 13612068292357.0,
               1720146321549.0 / 2090206949498.0,
               3134564353537.0 / 4481467310338.0,
               2277821191437.0 / 14882151754819.0]
        ktime = [0.0, kdt[0],
                 2526269341429.0 / 6820363962896.0,
                 2006345519317.0 / 3224310063776.0,
                 2802321613138.0 / 2924317926251.0]
        kco = [-567301805773.0 / 1357537059087.0,
               -2404267990393.0 / 2016746695238.0,
               -3550918686646.0 / 2091501179385.0,
               -1275806237668.0 / 842570457699.0]
    else:
        raise ValueError(mode)
    return kdt, ktime, kco


class Dns:
    """State and parameters of the incompressible path (module variables of the reference)."""

    def __init__(self, g, visc, schmidt, rkm_mode=RKM_EXP4,
                 buoyancy_type='none', buoyancy_params=(0.0, 0.0), buoyancy_vector=(0.0, 0.0, 0.0),
                 bbackground=None,
                 bcs_flow_jmin=(DNS_BCS_DIRICHLET,) * 3, bcs_flow_jmax=(DNS_BCS_DIRICHLET,) * 3,
                 bcs_scal_jmin=None, bcs_scal_jmax=None,
                 scal_limit=True, scal_min=0.0, scal_max=1.0):
        self.g = g
        self.nx, self.ny, self.nz = g[0].size, g[1].size, g[2].size
        self.visc = visc
        self.schmidt = list(schmidt)
        self.inb_scal = len(self.schmidt)
        self.burgers = Burgers(g, visc, self.schmidt)
        self.elliptic = Elliptic(g)
        self.rkm_mode = rkm_mode
        self.kdt, self.ktime, self.kco = rk_coefficients(rkm_mode)
        self.rkm_endstep = len(self.kdt)
        self.buoyancy_type = buoyancy_type
        self.buoyancy_params = buoyancy_params      # (c1, c0) for 'linear'; (b,) for 'homogeneous'
        self.buoyancy_vector = buoyancy_vector      # [Gravity].Vector / Froude
        self.bbackground = np.zeros(self.ny) if bbackground is None else np.asarray(bbackground, float)
        self.bcs_flow_jmin = tuple(bcs_flow_jmin)
        self.bcs_flow_jmax = tuple(bcs_flow_jmax)
        self.bcs_scal_jmin = tuple(bcs_scal_jmin) if bcs_scal_jmin else (DNS_BCS_DIRICHLET,) * self.inb_scal
        self.bcs_scal_jmax = tuple(bcs_scal_jmax) if bcs_scal_jmax else (DNS_BCS_DIRICHLET,) * self.inb_scal
        self.scal_limit = scal_limit
        self.scal_min, self.scal_max = scal_min, scal_max
        shape = (self.nz, self.ny, self.nx)
        self.q = [np.zeros(shape) for _ in range(3)]
        self.s = [np.zeros(shape) for _ in range(self.inb_scal)]
        self.hq = [np.zeros(shape) for _ in range(3)]
        self.hs = [np.zeros(shape) for _ in range(self.inb_scal)]
        self.dte = 0.0
        self.last_pressure = None

    # -----------------------------------------------------------------------
    def gravity_buoyancy(self):
        """Gravity_Buoyancy, gravity.f90:232-342."""
        if self.buoyancy_type == 'homogeneous':
            return np.full_like(self.q[0], self.buoyancy_params[0])
        if self.buoyancy_type == 'linear':
            c1, c0 = self.buoyancy_params
            dummy = self.bbackground - c0
            return c1 * self.s[0] - dummy[None, :, None]
        raise ValueError(self.buoyancy_type)

    def sources_flow(self):
        """TLab_Sources_Flow, tlab_sources.f90:36-131 (buoyancy only)."""
        if self.buoyancy_type == 'none':
            return
        for iq in range(3):
            if abs(self.buoyancy_vector[iq]) > 0.0:
                tmp1 = self.gravity_buoyancy()
                self.hq[iq] = self.hq[iq] + self.buoyancy_vector[iq] * tmp1

    # -----------------------------------------------------------------------
    def rhs_global_incompressible_1(self):
        """rhs_global_incompressible_1.f90:15-405."""
        g = self.g
        u, v, w = self.q
        hq, hs, s = self.hq, self.hs, self.s
        bcs = [[0, 0], [0, 0]]
        B = self.burgers
        SELF = None

        tmp1 = B.apply(0, 0, bcs, u, u)
        tmp2 = B.apply(1, 0, bcs, v, v)
        tmp3 = B.apply(2, 0, bcs, w, w)
        tmp7 = B.apply(1, 0, bcs, u, v)
        tmp8 = B.apply(2, 0, bcs, u, w)
        hq[0] = hq[0] + tmp1 + tmp7 + tmp8
        tmp7 = B.apply(0, 0, bcs, v, u)
        tmp8 = B.apply(2, 0, bcs, v, w)
        hq[1] = hq[1] + tmp2 + tmp7 + tmp8
        tmp7 = B.apply(0, 0, bcs, w, u)
        tmp8 = B.apply(1, 0, bcs, w, v)
        hq[2] = hq[2] + tmp3 + tmp7 + tmp8
        for is_ in range(self.inb_scal):
            tmp1 = B.apply(0, is_ + 1, bcs, s[is_], u)
            tmp2 = B.apply(1, is_ + 1, bcs, s[is_], v)
            tmp3 = B.apply(2, is_ + 1, bcs, s[is_], w)
            hs[is_] = hs[is_] + tmp1 + tmp2 + tmp3

        # remove residual divergence
        dummy = 1.0 / self.dte
        tmp2 = hq[1] + v * dummy
        tmp3 = hq[0] + u * dummy
        tmp4 = hq[2] + w * dummy
        tmp1 = opr_partial(1, OPR_P1, bcs, g[1], tmp2)
        tmp2 = opr_partial(0, OPR_P1, bcs, g[0], tmp3)
        tmp3 = opr_partial(2, OPR_P1, bcs, g[2], tmp4)
        tmp1 = tmp1 + tmp2 + tmp3

        # Neumann BCs for the pressure
        ny = self.ny
        bcs_hb = hq[1][:, 0, :].copy()
        bcs_ht = hq[1][:, ny - 1, :].copy()
        tmp1, tmp3 = opr_poisson(self.elliptic, tmp1, bcs_hb, bcs_ht, BCS_NN)
        self.last_pressure = tmp1

        tmp2 = opr_partial(0, OPR_P1, bcs, g[0], tmp1)
        tmp4 = opr_partial(2, OPR_P1, bcs, g[2], tmp1)
        hq[0] = hq[0] - tmp2
        hq[1] = hq[1] - tmp3
        hq[2] = hq[2] - tmp4

        # boundary conditions
        for iq in range(3):
            ibc = 0
            if self.bcs_flow_jmin[iq] == DNS_BCS_NEUMANN:
                ibc += 1
            if self.bcs_flow_jmax[iq] == DNS_BCS_NEUMANN:
                ibc += 2
            hb = np.zeros((self.nz, self.nx))
            ht = np.zeros((self.nz, self.nx))
            if ibc > 0:
                hb, ht = boundary_bcs_neumann_y(ibc, g[1], hq[iq])
            hq[iq][:, 0, :] = hb
            hq[iq][:, ny - 1, :] = ht
        for is_ in range(self.inb_scal):
            ibc = 0
            if self.bcs_scal_jmin[is_] == DNS_BCS_NEUMANN:
                ibc += 1
            if self.bcs_scal_jmax[is_] == DNS_BCS_NEUMANN:
                ibc += 2
            hb = np.zeros((self.nz, self.nx))
            ht = np.zeros((self.nz, self.nx))
            if ibc > 0:
                hb, ht = boundary_bcs_neumann_y(ibc, g[1], hs[is_])
            hs[is_][:, 0, :] = hb
            hs[is_][:, ny - 1, :] = ht

    # -----------------------------------------------------------------------
    def substep_incompressible_explicit(self):
        """TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT, time.f90:559-670 (EQNS_CONVECTIVE, RHS combined)."""
        self.sources_flow()
        self.rhs_global_incompressible_1()
        for iq in range(3):
            self.q[iq] = self.q[iq] + self.dte * self.hq[iq]
        for is_ in range(self.inb_scal):
            self.s[is_] = self.s[is_] + self.dte * self.hs[is_]

    def bounds_limit(self):
        """DNS_BOUNDS_LIMIT, dns_local.f90:67-90."""
        if self.scal_limit:
            for is_ in range(self.inb_scal):
                self.s[is_] = np.minimum(np.maximum(self.s[is_], self.scal_min), self.scal_max)

    def courant(self, cfla, cfld=None, prandtl=1.0, dtime=0.0):
        """TIME_COURANT, time.f90:365-548 (incompressible) with the grid factors of TIME_INITIALIZE :136-178.
        Returns (dtime, CFL number, diffusion number)."""
        g = self.g
        cfld = 0.25 * cfla if cfld is None else cfld
        ods = [1.0 / gi.jac[1:, 1] for gi in g]
        u, v, w = self.q
        wrk = np.abs(u) * ods[0][None, None, :] + np.abs(v) * ods[1][None, :, None]
        if g[2].size > 1:
            wrk = wrk + np.abs(w) * ods[2][:, None, None]
        pmax1 = wrk.max()
        dx2i = 0.0
        for gi, o in zip(g, ods):
            if gi.size > 1:
                dx2i = dx2i + (o * o).max()
        sf = 1.0
        sf = max(sf, 1.0 / prandtl)
        if self.inb_scal:
            sf = max(sf, 1.0 / min(self.schmidt))
        pmax2 = sf * self.visc * dx2i
        dtc = dtd = 1.0e20
        if pmax1 > 0.0:
            dtc = cfla / pmax1
        if pmax2 > 0.0:
            dtd = cfld / pmax2
        if cfla > 0.0:
            dtime = min(dtc, dtd)
        return dtime, dtime * pmax1, dtime * pmax2

    def bounds_control(self):
        """DilMin, DilMax of DNS_BOUNDS_CONTROL (dns_local.f90:166-189) through FI_INVARIANT_P."""
        bcs = [[0, 0], [0, 0]]
        res = opr_partial(0, OPR_P1, bcs, self.g[0], self.q[0])
        res = res + opr_partial(1, OPR_P1, bcs, self.g[1], self.q[1])
        res = -(res + opr_partial(2, OPR_P1, bcs, self.g[2], self.q[2]))
        amn, amx = res.min(), res.max()
        return -amx, -amn

    def runge_kutta(self, dtime, hook=None):
        """TIME_RUNGEKUTTA, time.f90:185-333 (explicit low-storage branch)."""
        for a in self.hq + self.hs:
            a[...] = 0.0
        for sub in range(1, self.rkm_endstep + 1):
            self.dte = dtime * self.kdt[sub - 1]
            self.substep_incompressible_explicit()
            self.bounds_limit()
            if sub < self.rkm_endstep:
                alpha = self.kco[sub - 1]
                for i in range(3):
                    self.hq[i] = alpha * self.hq[i]
                for i in range(self.inb_scal):
                    self.hs[i] = alpha * self.hs[i]
            if hook is not None:
                hook(sub, self)
