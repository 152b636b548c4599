"""Size-independent properties at the benchmark sizes (no oracle: it would take minutes on a CPU):
the reference's vburgers identity, linearity, and the Poisson round trip on 512^3 / 256^3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grids(nx, ny, nz):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import grid_periodic, grid_tanh
    return grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)


def _rel(a, b):
    import torch
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


def test_burgers_identity_and_linearity_512(cuda):
    import torch
    from tlab_b200 import opr
    from bench import synth_field
    nx = ny = nz = 512
    x, y, z = _grids(nx, ny, nz)
    g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    visc = 1.0 / 5000.0
    opr.OPR_Burgers_Initialize(g, visc, [1.0])
    a = synth_field(torch, cuda, (nz, ny, nx), x, y, z, 3, 1.0)
    b = synth_field(torch, cuda, (nz, ny, nx), x, y, z, 4, 1.0)
    bcs = [[0, 0], [0, 0]]
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    B = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    r, d2, d1, t = (torch.empty_like(a) for _ in range(4))
    for idir in range(3):
        B[idir](opr.OPR_B_SELF, 0, nx, ny, nz, bcs, a, a, r)
        P[idir](opr.OPR_P2_P1, nx, ny, nz, bcs, g[idir], a, d2, d1)
        assert _rel(r, visc * d2 - a * d1) <= 1e-11          # vburgers.f90:76-153
        # linearity of the derivative: d(a + 2 b) = d a + 2 d b
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, g[idir], b, t)
        comb = d1 + 2.0 * t
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, g[idir], a + 2.0 * b, t)
        assert _rel(t, comb) <= 1e-12
        # the mean of a periodic derivative vanishes
        if g[idir].periodic:
            assert abs(float(d1.mean())) <= 1e-11 * float(d1.abs().max())


def test_poisson_round_trip_256(cuda):
    import torch
    from tlab_b200 import opr
    from bench import synth_field
    nx, ny, nz = 256, 256, 256
    x, y, z = _grids(nx, ny, nz)
    g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    a = synth_field(torch, cuda, (nz, ny, nx), x, y, z, 9, 1.0)       # |k| <= 8: no Nyquist content
    bcs = [[0, 0], [0, 0]]
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    f = torch.zeros_like(a)
    d1, d2, ay = (torch.empty_like(a) for _ in range(3))
    for idir in range(3):
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, g[idir], a, d1)
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, g[idir], d1, d2)
        f += d2
        if idir == 1:
            ay.copy_(d1)
    opr.OPR_Elliptic_Initialize(g)
    t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=cuda)
    t2 = torch.zeros_like(t1)
    dpdy = torch.zeros_like(a)
    opr.OPR_Poisson(nx, ny, nz, opr.BCS_NN, f, t1, t2, ay[:, 0, :].contiguous(), ay[:, ny - 1, :].contiguous(), dpdy)
    ref = a - a[:, 0, :].mean()
    assert _rel(f, ref) <= 1e-10
    assert _rel(dpdy, ay) <= 1e-10
