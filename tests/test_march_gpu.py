"""Marching-panel line kernels (csrc/march.cu: panels of 32 lines, rounds of 4 chunks, corrections one round later) against
the oracle: OPR_Partial P1 for every boundary code, OPR_Burgers for both diffusivities, along y (non-periodic, uniform and
stretched: Jacobian correction) and z (circulant closure, rounds visited 1..R-1, 0), from the shortest eligible line (two rounds)
to the C3 / C4 line lengths, with load+store and with red.global.add accumulation; and one RK substep whose RHS runs through them
(second input, accumulation into hq).  Tolerance: the north_star's 1e-12 per operator call."""
import ctypes
import itertools

import numpy as np
import pytest

from common import grid_periodic, grid_tanh, smooth_field, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _counter(name):
    from tlab_b200 import lib as tl
    L = tl.load()
    c = ctypes.c_longlong()
    tl.check(L.tlab_gpu_get_counter(name.encode(), ctypes.byref(c)))
    return c.value


def _plans(nx, ny, nz, ykind):
    from oracle import fdm
    from tlab_b200 import opr
    x, z = grid_periodic(nx), grid_periodic(nz)
    y = grid_tanh(ny) if ykind == "tanh" else np.linspace(0.0, 1.0, ny)
    yuni = ykind == "uniform"
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, yuni, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, yuni, name="y"), opr.FdmPlan(z, True, True, name="z")]
    return (x, y, z), go, gg


# (nx, ny, nz, y grid): y lines of 128 ... 1024 points, z lines of 128 ... 2048 points
CASES = [(32, 128, 256, "tanh"), (64, 256, 128, "uniform"), (32, 512, 16, "tanh"), (32, 64, 1024, "tanh"),
         (32, 1024, 16, "tanh"), (32, 16, 2048, "uniform")]


@pytest.mark.parametrize("red", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_marching_operators(cuda, case, red):
    import torch
    from oracle import operators as O
    from tlab_b200 import opr, lib as tl
    L = tl.load()
    nx, ny, nz, ykind = case
    grids, go, gg = _plans(nx, ny, nz, ykind)
    rng = np.random.default_rng(7)
    a = smooth_field((nz, ny, nx), grids, seed=5) + 0.05 * rng.standard_normal((nz, ny, nx))
    v = smooth_field((nz, ny, nx), grids, seed=6) + 0.05 * rng.standard_normal((nz, ny, nx))
    u, vel = torch.from_numpy(a).to(cuda), torch.from_numpy(v).to(cuda)
    visc, schmidt = 1.0 / 5000.0, [0.7]
    B = O.Burgers(go, visc, schmidt)
    opr.OPR_Burgers_Initialize(gg, visc, schmidt)
    P = [None, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    Bg = [None, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    tl.check(L.tlab_gpu_set_tuning(b"march", 2))
    tl.check(L.tlab_gpu_set_tuning(b"march_red", red))
    n0 = _counter("march_launches")
    bad, expected = [], 0
    try:
        for idir in (1, 2):
            n = (ny, nz)[idir - 1]
            eligible = n >= 128 and n % 64 == 0
            per = go[idir].periodic
            bcs_list = [[[0, 0], [0, 0]]] if per else [[[b1, 0], [b2, 0]] for b1, b2 in itertools.product((0, 1), (0, 1))]
            for bcs in bcs_list:
                res = torch.full_like(u, float("nan"))
                P[idir](O.OPR_P1, nx, ny, nz, bcs, gg[idir], u, res)
                e = rel_l2(res.cpu().numpy(), O.opr_partial(idir, O.OPR_P1, bcs, go[idir], a))
                expected += int(eligible)
                if not e <= TOL:
                    bad.append(("P1", idir, bcs, e))
            bcs = [[0, 0], [0, 0]]
            for is_ in (0, 1):
                res = torch.full_like(u, float("nan"))
                Bg[idir](opr.OPR_B_U_IN, is_, nx, ny, nz, bcs, u, vel, res)
                e = rel_l2(res.cpu().numpy(), B.apply(idir, is_, bcs, a, v))
                expected += int(eligible)
                if not e <= TOL:
                    bad.append(("burgers", idir, is_, e))
            res = torch.full_like(u, float("nan"))
            Bg[idir](opr.OPR_B_SELF, 0, nx, ny, nz, bcs, u, u, res)
            e = rel_l2(res.cpu().numpy(), B.apply(idir, 0, bcs, a, a))
            expected += int(eligible)
            if not e <= TOL:
                bad.append(("burgers self", idir, e))
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"march_red", 0))
        tl.check(L.tlab_gpu_set_tuning(b"march", 1))
    assert not bad, bad
    assert _counter("march_launches") - n0 == expected, (_counter("march_launches") - n0, expected)


@pytest.mark.parametrize("red", [0, 1])
@pytest.mark.parametrize("shape", [(32, 128, 128), (64, 256, 64), (32, 64, 256)])
def test_substep_through_the_marching_kernels(cuda, shape, red):
    """sources + RHS (Burgers accumulated into hq/hs, divergence of hq + q/dte, pressure gradient subtracted) + update against
    the oracle, with the y and/or z operators on marching panels."""
    from oracle import dns as OD
    from tlab_b200 import dns as GD, lib as tl
    L = tl.load()
    nx, ny, nz = shape
    (x, y, z), go, gg = _plans(nx, ny, nz, "tanh")
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[1.0], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
              buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(D,), bcs_scal_jmax=(N,))
    tl.check(L.tlab_gpu_set_tuning(b"march", 2))
    tl.check(L.tlab_gpu_set_tuning(b"march_red", red))
    try:
        o, g = OD.Dns(go, **kw), GD.Dns(gg, **kw)
        wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
        for i in range(3):
            f = 0.05 * smooth_field((nz, ny, nx), (x, y, z), seed=31 + i) * wall
            o.q[i][...] = f
            g.set("q%d" % (i + 1), f)
        sc = 0.5 + 0.02 * smooth_field((nz, ny, nx), (x, y, z), seed=40) * wall
        o.s[0][...] = sc
        g.set("s1", sc)
        n0 = _counter("march_launches")
        for _ in range(2):
            o.runge_kutta(1e-3)
            g.runge_kutta(1e-3)
        assert _counter("march_launches") > n0
        for i in range(3):
            assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-10
        assert rel_l2(g.get("s1"), o.s[0]) <= 1e-10
        g.close()
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"march_red", 0))
        tl.check(L.tlab_gpu_set_tuning(b"march", 1))
