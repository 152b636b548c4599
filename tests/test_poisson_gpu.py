"""GPU parity of OPR_Poisson against the oracle and the reference's round-trip recipe
(src/valid/elliptic/vpoisson.f90:162-248)."""
import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched, smooth_field, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(nx, ny, nz, ykind):
    from oracle import fdm
    from tlab_b200 import opr
    x = grid_periodic(nx)
    z = grid_periodic(nz) if nz > 1 else np.zeros(1)
    y = {"stretched": grid_stretched(ny), "tanh": grid_tanh(ny), "uniform": np.linspace(0, 1, ny)}[ykind]
    yuni = ykind == "uniform"
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, yuni, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, yuni, name="y"), opr.FdmPlan(z, True, True, name="z")]
    return (x, y, z), go, gg


@pytest.mark.parametrize("case", [(32, 33, 16, "stretched"), (64, 48, 32, "tanh"), (34, 20, 18, "uniform"), (64, 40, 1, "stretched")])
def test_poisson_matches_oracle(cuda, case):
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz, ykind = case
    grids, go, gg = _setup(nx, ny, nz, ykind)
    rng = np.random.default_rng(3)
    f = smooth_field((nz, ny, nx), grids, seed=11) + 0.05 * rng.standard_normal((nz, ny, nx))
    hb = 0.3 * rng.standard_normal((nz, nx))
    ht = 0.3 * rng.standard_normal((nz, nx))
    ell = O.Elliptic(go)
    p_ref, dpdy_ref = O.opr_poisson(ell, f, hb, ht)
    opr.OPR_Elliptic_Initialize(gg)
    p = torch.from_numpy(f).to(cuda)
    t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=cuda)
    t2 = torch.zeros_like(t1)
    dpdy = torch.zeros_like(p)
    opr.OPR_Poisson(nx, ny, nz, opr.BCS_NN, p, t1, t2, torch.from_numpy(hb).to(cuda), torch.from_numpy(ht).to(cuda), dpdy)
    e1, e2 = rel_l2(p.cpu().numpy(), p_ref), rel_l2(dpdy.cpu().numpy(), dpdy_ref)
    assert e1 <= TOL and e2 <= TOL, (e1, e2)


def test_poisson_round_trip(cuda):
    """f = dxdx a + dydy a + dzdz a (OPR_P1 twice per direction), Neumann data from dy a:
    OPR_Poisson returns a - <a>(bottom mean mode pinned to 0) and dpdy = dy a to round-off."""
    import torch
    from tlab_b200 import opr
    nx, ny, nz = 64, 48, 32
    grids, go, gg = _setup(nx, ny, nz, "tanh")
    a_np = smooth_field((nz, ny, nx), grids, seed=21)
    a = torch.from_numpy(a_np).to(cuda)
    bcs = [[0, 0], [0, 0]]
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    f = torch.zeros_like(a)
    d1 = torch.empty_like(a)
    d2 = torch.empty_like(a)
    ay = torch.empty_like(a)
    for idir in range(3):
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, gg[idir], a, d1)
        P[idir](opr.OPR_P1, nx, ny, nz, bcs, gg[idir], d1, d2)
        f += d2
        if idir == 1:
            ay.copy_(d1)
    opr.OPR_Elliptic_Initialize(gg)
    t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=cuda)
    t2 = torch.zeros_like(t1)
    dpdy = torch.zeros_like(a)
    hb = ay[:, 0, :].contiguous()
    ht = ay[:, ny - 1, :].contiguous()
    opr.OPR_Poisson(nx, ny, nz, opr.BCS_NN, f, t1, t2, hb, ht, dpdy)
    ref = a_np - a_np[:, 0, :].mean()
    assert rel_l2(f.cpu().numpy(), ref) <= 1e-11
    assert rel_l2(dpdy.cpu().numpy(), ay.cpu().numpy()) <= 1e-11
