"""Host-side mirror of the reference's operator interface on top of the C ABI.

Same names, argument order and meaning as the Fortran module procedures
(src/operators/opr_partial.f90, src/physics/opr_burgers.f90, src/operators/opr_elliptic.f90,
src/tools/dns/boundary_bcs.f90), so that the parity tests read like the reference's own
validation programs (src/valid/).  Arrays are torch float64 CUDA tensors holding the Fortran
layout a(nx,ny,nz) -- i.e. a C-ordered tensor of shape (nz, ny, nx); torch only provides device
memory here, every operator is a call into libtlab_gpu.so.
"""
import ctypes

import numpy as np
import torch

from . import lib as _lib

OPR_P1, OPR_P2, OPR_P2_P1 = 1, 2, 3
OPR_B_SELF, OPR_B_U_IN = 0, 1
BCS_DD, BCS_ND, BCS_DN, BCS_NN = 0, 1, 2, 3

_SCHEMES = {"compactjacobian4": 4, "compactjacobian6": 6, "compactjacobian6hyper": 7, "compactdirect6": 16}


def _ptr(t):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise TypeError("expected a contiguous float64 CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def _bcs(bcs):
    b = np.asarray(bcs, dtype=np.int32)
    if b.shape == (2, 2):
        b = b.T.reshape(4)          # Fortran column-major: bcs(1,1), bcs(2,1), bcs(1,2), bcs(2,2)
    b = np.ascontiguousarray(b.reshape(4), dtype=np.int32)
    return (ctypes.c_int * 4)(*[int(x) for x in b])


def _sync():
    torch.cuda.current_stream().synchronize()


def _call(fn, *args):
    _sync()
    _lib.check(fn(*args))


class FdmPlan:
    """type(fdm_dt) (src/fdm/fdm.f90:14-29) created by FDM_CreatePlan (:143-252)."""

    def __init__(self, nodes, periodic, uniform=None, der1="compactjacobian6", der2="compactjacobian6hyper",
                 name="x", host_only=False):
        L = _lib.load()
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.size = int(self.nodes.size)
        self.periodic = bool(periodic)
        self.uniform = bool(periodic) if uniform is None else bool(uniform)
        self.name = name
        self.dir = {"x": 1, "y": 2, "z": 3}.get(name, 0)
        h = ctypes.c_void_p()
        create = L.tlab_fdm_plan_create_host if host_only else L.tlab_fdm_plan_create
        _lib.check(create(self.dir, self.size, self.nodes.ctypes.data_as(ctypes.c_void_p), int(self.periodic),
                          int(self.uniform), _SCHEMES[der1.lower()], _SCHEMES[der2.lower()], ctypes.byref(h)))
        self.handle = h

    def table(self, what):
        L = _lib.load()
        cap = max(self.size, 8) * 32
        buf = np.zeros(cap)
        cnt = ctypes.c_int()
        _lib.check(L.tlab_fdm_plan_get(self.handle, what.encode(), buf.ctypes.data_as(ctypes.c_void_p), cap,
                                       ctypes.byref(cnt)))
        return buf[:cnt.value].copy()

    @property
    def jac(self):
        return self.table("jac1")

    @property
    def mwn1(self):
        return self.table("mwn1")

    def __del__(self):
        try:
            if self.handle:
                _lib.load().tlab_fdm_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _partial(idir, type_, nx, ny, nz, bcs, g, u, result, tmp1):
    _call(_lib.load().tlab_opr_partial, idir, type_, nx, ny, nz, _bcs(bcs), g.handle, _ptr(u), _ptr(result), _ptr(tmp1))


def OPR_Partial_X(type_, nx, ny, nz, bcs, g, u, result, tmp1=None):
    """opr_partial.f90:31-150"""
    _partial(1, type_, nx, ny, nz, bcs, g, u, result, tmp1)


def OPR_Partial_Y(type_, nx, ny, nz, bcs, g, u, result, tmp1=None):
    """opr_partial.f90:266-377"""
    _partial(2, type_, nx, ny, nz, bcs, g, u, result, tmp1)


def OPR_Partial_Z(type_, nx, ny, nz, bcs, g, u, result, tmp1=None):
    """opr_partial.f90:154-262"""
    _partial(3, type_, nx, ny, nz, bcs, g, u, result, tmp1)


def OPR_Burgers_Initialize(g, visc, schmidt):
    """opr_burgers.f90:52-115; g = (gx, gy, gz)"""
    sc = np.ascontiguousarray(schmidt, dtype=np.float64)
    _lib.check(_lib.load().tlab_opr_burgers_init(g[0].handle, g[1].handle, g[2].handle, float(visc), int(sc.size),
                                                 sc.ctypes.data_as(ctypes.c_void_p)))


def _burgers(idir, ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1, u_t):
    _call(_lib.load().tlab_opr_burgers, idir, ivel, is_, nx, ny, nz, _bcs(bcs), _ptr(s), _ptr(u), _ptr(result),
          _ptr(tmp1), _ptr(u_t))


def OPR_Burgers_X(ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1=None, u_t=None):
    """opr_burgers.f90:190-273"""
    _burgers(1, ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1, u_t)


def OPR_Burgers_Y(ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1=None, u_t=None):
    """opr_burgers.f90:277-355"""
    _burgers(2, ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1, u_t)


def OPR_Burgers_Z(ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1=None, u_t=None):
    """opr_burgers.f90:359-431"""
    _burgers(3, ivel, is_, nx, ny, nz, bcs, s, u, result, tmp1, u_t)


def FDM_Der1_Solve(nlines, ibc, g, u, result):
    """fdm_derivative.f90:218-278 on u(nlines, n)"""
    _call(_lib.load().tlab_fdm_der1_solve, g.handle, nlines, ibc, _ptr(u), _ptr(result))


def FDM_Der2_Solve(nlines, g, u, result, du=None):
    """fdm_derivative.f90:413-459 on u(nlines, n)"""
    _call(_lib.load().tlab_fdm_der2_solve, g.handle, nlines, -1, _ptr(u), _ptr(du), _ptr(result))


def BOUNDARY_BCS_NEUMANN_Y(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht):
    """boundary_bcs.f90:368-473"""
    _call(_lib.load().tlab_boundary_bcs_neumann_y, ibc, nx, ny, nz, g.handle, _ptr(u), _ptr(bcs_hb), _ptr(bcs_ht))


def OPR_Elliptic_Initialize(g, kmax=0):
    """opr_elliptic.f90:86-250 (FourierXZ_Factorize); g = (gx, gy, gz); kmax = local slab thickness (0: whole domain)"""
    _sync()
    _lib.check(_lib.load().tlab_opr_elliptic_init(g[0].handle, g[1].handle, g[2].handle, int(kmax)))


def OPR_Poisson(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy=None):
    """opr_elliptic.f90:263-364; tmp1, tmp2 need (nx+2)*ny*nz elements"""
    need = (nx + 2) * ny * nz
    if tmp1.numel() < need or tmp2.numel() < need:
        raise ValueError("tmp1/tmp2 must hold (nx+2)*ny*nz doubles")
    _call(_lib.load().tlab_opr_poisson, nx, ny, nz, ibc, _ptr(p), _ptr(tmp1), _ptr(tmp2), _ptr(bcs_hb), _ptr(bcs_ht),
          _ptr(dpdy))


def OPR_Fourier_X_Forward(nx, ny, nz, in_, out):
    """opr_fourier.f90:219-273: out = c(nx/2+1, ny, nz) as (nx+2)*ny*nz doubles (re, im interleaved)"""
    if out.numel() < (nx + 2) * ny * nz:
        raise ValueError("out must hold (nx+2)*ny*nz doubles")
    _call(_lib.load().tlab_opr_fourier_x_forward, nx, ny, nz, _ptr(in_), _ptr(out))


def OPR_Fourier_X_Backward(nx, ny, nz, in_, out):
    """opr_fourier.f90:277-329 (unnormalised; may overwrite in_)"""
    _call(_lib.load().tlab_opr_fourier_x_backward, nx, ny, nz, _ptr(in_), _ptr(out))


def OPR_Fourier_Z_Forward(nx, ny, nz, in_, out):
    """opr_fourier.f90:333-381 (the reference reads the sizes from module variables)"""
    _call(_lib.load().tlab_opr_fourier_z_forward, nx, ny, nz, _ptr(in_), _ptr(out))


def OPR_Fourier_Z_Backward(nx, ny, nz, in_, out):
    """opr_fourier.f90:385-433 (unnormalised)"""
    _call(_lib.load().tlab_opr_fourier_z_backward, nx, ny, nz, _ptr(in_), _ptr(out))
