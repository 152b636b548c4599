"""The C++/OpenMP CPU baseline (oracle/cpp/tlab_cpu.cpp, driven by oracle/cpu_baseline.py) against the numpy oracle it restates:
one RK4-5 step of the bench configuration (no-slip bottom, free-slip top, linear buoyancy, one scalar) on small grids, 3-D and
2-D, <= 1e-12 on every field.  Both are test infrastructure; the product never loads either."""
import numpy as np
import pytest

from common import grid_periodic, grid_tanh, smooth_field, rel_l2


@pytest.mark.parametrize("shape", [(32, 48, 16), (64, 33, 1), (16, 40, 32)])
def test_cpp_baseline_matches_the_numpy_oracle(shape):
    from oracle import fdm, dns as OD
    from oracle import cpu_baseline as CB
    nx, ny, nz = shape
    x, y = grid_periodic(nx), grid_tanh(ny)
    z = grid_periodic(nz) if nz > 1 else np.zeros(1)
    g = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[0.7], buoyancy_type="linear", buoyancy_params=(1.0, 0.1),
              buoyancy_vector=(0.0, 1.0, 0.0), bbackground=0.3 * y, bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(D,), bcs_scal_jmax=(N,))
    o, c = OD.Dns(g, **kw), CB.CpuDns(g, **kw)
    wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
    for i in range(3):
        f = 0.05 * smooth_field((nz, ny, nx), (x, y, z), seed=31 + i) * wall
        if nz == 1 and i == 2:
            f = 0.0 * f
        o.q[i][...] = f
        c.q[i][...] = f
    sc = 0.5 + 0.02 * smooth_field((nz, ny, nx), (x, y, z), seed=40) * wall
    o.s[0][...] = sc
    c.s[0][...] = sc
    for _ in range(2):
        o.runge_kutta(1e-3)
        c.runge_kutta(1e-3)
    for i in range(3):
        if np.abs(o.q[i]).max() > 0:
            assert rel_l2(c.q[i], o.q[i]) <= 1e-12, i
        else:
            assert np.abs(c.q[i]).max() == 0.0
        # hq carries the pressure gradient of a forcing ~ q/dte: round-off is amplified by 1/dte there, not in the fields
        assert rel_l2(c.hq[i], o.hq[i]) <= 1e-9 or np.abs(o.hq[i]).max() == 0.0
    assert rel_l2(c.s[0], o.s[0]) <= 1e-12
