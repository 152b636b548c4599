"""No-GPU checks of the product: the C-ABI library loads and exports every symbol of include/tlab_gpu.h,
its host-side plan builder agrees with the oracle, and device entry points fail loudly without a GPU."""
import ctypes
import os

import numpy as np
import pytest

from common import grid_periodic, grid_tanh, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from tlab_b200 import build, lib
    build.build()
    return lib.load()


def test_library_exports_every_declared_symbol(L):
    from tlab_b200 import lib
    protos = lib.parse_header()
    assert len(protos) >= 40
    for name in protos:
        assert hasattr(L, name), name
    hdr = open(os.path.join(ROOT, "include", "tlab_gpu.h")).read()
    for name in ("tlab_opr_partial", "tlab_opr_burgers", "tlab_opr_poisson", "tlab_time_substep", "tlab_trp_exec_k_forward",
                 "tlab_tridss", "tlab_pentadss", "tlab_transpose"):
        assert name in hdr and name in protos


@pytest.mark.parametrize("periodic", [True, False])
def test_host_plan_tables_match_oracle(L, periodic):
    from oracle import fdm
    from tlab_b200 import opr
    nodes = grid_periodic(64) if periodic else grid_tanh(65)
    n = len(nodes)
    g = fdm.Plan(nodes, periodic, periodic)
    p = opr.FdmPlan(nodes, periodic, periodic, host_only=True)
    assert rel_l2(p.table("jac1"), g.jac[1:, 1]) < 1e-13
    assert rel_l2(p.table("lhs1").reshape(3, n).T, g.der1.lhs[1:, 1:4]) < 1e-13
    assert rel_l2(p.table("rhs1").reshape(5, n).T, g.der1.rhs[1:, 1:6]) < 1e-13
    nc = g.der1.lu.shape[1] - 1
    assert rel_l2(p.table("lu1").reshape(nc, n).T, g.der1.lu[1:, 1:]) < 1e-12
    assert rel_l2(p.table("lhs2").reshape(3, n).T, g.der2.lhs[1:, 1:4]) < 1e-13
    assert rel_l2(p.table("rhs2").reshape(10, n).T, g.der2.rhs[1:, 1:11]) < 1e-12
    nc = g.der2.lu.shape[1] - 1
    assert rel_l2(p.table("lu2").reshape(nc, n).T, g.der2.lu[1:, 1:]) < 1e-12
    if periodic:
        assert rel_l2(p.table("mwn1"), g.der1.mwn) < 1e-14
        assert rel_l2(p.table("mwn2"), g.der2.mwn) < 1e-14
    else:
        assert rel_l2(p.table("rhs1_b").reshape(8, 4).T, g.der1.rhs_b[1:5, 0:8]) < 1e-13
        assert rel_l2(p.table("rhs1_t").reshape(7, 5).T, g.der1.rhs_t[0:5, 1:8]) < 1e-13


def test_integral_system_split_matches_oracle(L):
    from oracle import fdm, integral as I
    from tlab_b200 import opr, lib
    n = 65
    y = grid_tanh(n)
    g = fdm.Plan(y, False, False)
    p = opr.FdmPlan(y, False, False, name="y", host_only=True)

    def P(a):
        return a.ctypes.data_as(ctypes.c_void_p)
    for ibc in (1, 2):
        for lam in (0.0, 3.7, 250.0):
            ls = lam if ibc == 1 else -lam
            fi = I.int1_create_system(g.der1, ls, ibc)
            lhs, rhs, rb, rt = np.zeros(n * 5), np.zeros(n * 3), np.zeros(40), np.zeros(40)
            lib.check(L.tlab_fdm_int1_system_host(p.handle, ibc, ls, P(lhs), P(rhs), P(rb), P(rt)))
            assert rel_l2(lhs.reshape(5, n).T, fi.lhs[1:, 1:, 0]) < 1e-13
            assert rel_l2(rhs.reshape(3, n).T, fi.rhs[1:, 1:, 0]) < 1e-13
            assert np.abs(rb.reshape(8, 5).T - fi.rhs_b[1:6, 0:8, 0]).max() < 1e-13
            assert np.abs(rt.reshape(8, 5).T - fi.rhs_t[0:5, 1:9, 0]).max() < 1e-13


def test_rk_tables_match_reference_values(L):
    from tlab_b200 import dns as GD
    from oracle import dns as OD
    for mode in (3, 4):
        a = GD.rk_coefficients(mode)
        b = OD.rk_coefficients(mode)
        for u, v in zip(a, b):
            assert list(u) == list(v)
    kdt, _, kco = GD.rk_coefficients(4)
    assert kdt[0] == 1432997174477.0 / 9575080441755.0 and kco[3] == -1275806237668.0 / 842570457699.0   # time.f90:97-112


def test_error_conventions(L):
    """Reference error codes (dns_error.h) at the boundary; no CPU fallback."""
    import torch
    from tlab_b200 import lib, opr
    h = ctypes.c_void_p()
    nodes = np.linspace(0, 1, 8)
    rc = L.tlab_fdm_plan_create_host(2, 8, nodes.ctypes.data_as(ctypes.c_void_p), 0, 0, 6, 7, ctypes.byref(h))
    assert rc == 48                                   # DNS_ERROR_DIMGRID: too few points for the closures
    nodes = grid_tanh(32)
    rc = L.tlab_fdm_plan_create_host(2, 32, nodes.ctypes.data_as(ctypes.c_void_p), 1, 0, 6, 7, ctypes.byref(h))
    assert rc == 85                                   # DNS_ERROR_OPTION: periodic direction must be uniform
    rc = L.tlab_fdm_plan_create_host(2, 32, nodes.ctypes.data_as(ctypes.c_void_p), 0, 0, 5, 7, ctypes.byref(h))
    assert rc == 104                                  # DNS_ERROR_UNDEVELOP: penta scheme not on the GPU path
    assert b"not implemented" in L.tlab_gpu_last_error()
    if not torch.cuda.is_available():
        rc = L.tlab_fdm_plan_create(2, 32, nodes.ctypes.data_as(ctypes.c_void_p), 0, 0, 6, 7, ctypes.byref(h))
        assert rc == 200 and b"no CPU fallback" in L.tlab_gpu_last_error()
        with pytest.raises(lib.TlabError):
            opr.FdmPlan(nodes, False, False)


def test_oracle_is_not_imported_by_the_product():
    import re
    pkg = os.path.join(ROOT, "tlab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


@pytest.mark.parametrize("n", [17, 65, 128])
def test_host_plan_direct_second_derivative_matches_oracle(L, n):
    """SpaceOrder2 = CompactDirect6 (src/fdm/fdm_comx_direct.f90:305-412): per-row lhs/rhs from the node positions."""
    from oracle import fdm
    from tlab_b200 import opr
    nodes = grid_tanh(n)
    g = fdm.Plan(nodes, False, False, mode2=fdm.FDM_COM6_DIRECT)
    p = opr.FdmPlan(nodes, False, False, der2="compactdirect6", host_only=True)
    assert not g.der2.need_1der
    assert rel_l2(p.table("lhs2").reshape(3, n).T, g.der2.lhs[1:, 1:4]) < 1e-13
    assert rel_l2(p.table("rhs2").reshape(-1, n).T[:, :5], g.der2.rhs[1:, 1:6]) < 1e-12
    nc = g.der2.lu.shape[1] - 1
    assert rel_l2(p.table("lu2").reshape(nc, n).T, g.der2.lu[1:, 1:]) < 1e-12
    # the first derivative is untouched by the choice of the second
    assert rel_l2(p.table("lhs1").reshape(3, n).T, g.der1.lhs[1:, 1:4]) < 1e-13
    # a periodic direction falls back to the hyper scheme (fdm.f90:158)
    x = grid_periodic(32)
    gp = fdm.Plan(x, True, True, mode2=fdm.FDM_COM6_DIRECT)
    pp = opr.FdmPlan(x, True, True, der2="compactdirect6", host_only=True)
    assert rel_l2(pp.table("rhs2").reshape(-1, 32).T, gp.der2.rhs[1:, 1:11]) < 1e-13
    assert rel_l2(pp.table("mwn2"), gp.der2.mwn) < 1e-14
