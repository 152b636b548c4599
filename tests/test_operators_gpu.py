"""GPU parity of the directional operators against the oracle (rel. L2 <= 1e-12 per operator call)."""
import itertools

import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched, smooth_field, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _plans(nx, ny, nz, ykind="stretched", xper=True, zper=True):
    from oracle import fdm
    from tlab_b200 import opr
    x = grid_periodic(nx) if xper else grid_stretched(nx, 2.0)
    z = grid_periodic(nz) if zper else grid_stretched(nz, 3.0)
    y = {"stretched": grid_stretched(ny), "tanh": grid_tanh(ny), "uniform": np.linspace(0, 1, ny)}[ykind]
    yuni = ykind == "uniform"
    go = [fdm.Plan(x, xper, xper, name="x"), fdm.Plan(y, False, yuni, name="y"), fdm.Plan(z, zper, zper, name="z")]
    gg = [opr.FdmPlan(x, xper, xper, name="x"), opr.FdmPlan(y, False, yuni, name="y"), opr.FdmPlan(z, zper, zper, name="z")]
    return (x, y, z), go, gg


CASES = [(64, 48, 32, "stretched", True, True),
         (50, 33, 20, "tanh", True, True),          # ragged: chunk sizes 16/17, lines not a multiple of the tile
         (32, 16, 17, "uniform", False, False),      # biased schemes in all directions, single-chunk y
         (130, 40, 36, "stretched", True, False),
         (256, 32, 16, "tanh", True, True),          # fast kernels: constant interior chunks + circulant closure in x
         (16, 48, 256, "stretched", True, True),     # the same in z; single-chunk x
         (32, 256, 16, "uniform", True, True),       # uniform non-periodic y: constant interior chunks
         (64, 64, 64, "tanh", False, False),         # biased schemes on full chunks in all directions
         (16, 32, 1024, "tanh", True, True)]         # z lines of 64 chunks: TMA kernel with 8 lines per CTA (64-byte rows)


@pytest.mark.parametrize("case", CASES)
def test_opr_partial(cuda, case):
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz, ykind, xper, zper = case
    grids, go, gg = _plans(nx, ny, nz, ykind, xper, zper)
    a = smooth_field((nz, ny, nx), grids)
    u = torch.from_numpy(a).to(cuda)
    fns = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    bad = []
    for idir in range(3):
        per = go[idir].periodic
        bcs_list = [[[0, 0], [0, 0]]] if per else [[[b1, 0], [b2, 0]] for b1, b2 in itertools.product((0, 1), (0, 1))]
        for bcs in bcs_list:
            for type_ in (O.OPR_P1, O.OPR_P2, O.OPR_P2_P1):
                res = torch.full_like(u, float("nan"))
                tmp = torch.full_like(u, float("nan"))
                fns[idir](type_, nx, ny, nz, bcs, gg[idir], u, res, tmp if type_ == O.OPR_P2_P1 else None)
                ref = O.opr_partial(idir, type_, bcs, go[idir], a)
                if type_ == O.OPR_P2_P1:
                    e2, e1 = rel_l2(res.cpu().numpy(), ref[0]), rel_l2(tmp.cpu().numpy(), ref[1])
                    if not (e2 <= TOL and e1 <= TOL):
                        bad.append((idir, bcs, type_, e2, e1))
                else:
                    e = rel_l2(res.cpu().numpy(), ref)
                    if not e <= TOL:
                        bad.append((idir, bcs, type_, e))
    assert not bad, bad


@pytest.mark.parametrize("case", CASES)
def test_opr_burgers(cuda, case):
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz, ykind, xper, zper = case
    grids, go, gg = _plans(nx, ny, nz, ykind, xper, zper)
    visc, schmidt = 1.0 / 5000.0, [1.0, 0.7]
    B = O.Burgers(go, visc, schmidt)
    opr.OPR_Burgers_Initialize(gg, visc, schmidt)
    s_np = smooth_field((nz, ny, nx), grids, seed=1)
    v_np = smooth_field((nz, ny, nx), grids, seed=2)
    s, v = torch.from_numpy(s_np).to(cuda), torch.from_numpy(v_np).to(cuda)
    fns = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    bcs = [[0, 0], [0, 0]]
    for idir in range(3):
        for is_ in range(3):
            res = torch.full_like(s, float("nan"))
            fns[idir](opr.OPR_B_U_IN, is_, nx, ny, nz, bcs, s, v, res)
            e = rel_l2(res.cpu().numpy(), B.apply(idir, is_, bcs, s_np, v_np))
            assert e <= TOL, ("u_in", idir, is_, e)
        res = torch.full_like(s, float("nan"))
        fns[idir](opr.OPR_B_SELF, 0, nx, ny, nz, bcs, s, s, res)
        e = rel_l2(res.cpu().numpy(), B.apply(idir, 0, bcs, s_np, s_np))
        assert e <= TOL, ("self", idir, e)
        # the reference's own consistency check (src/valid/burgers/vburgers.f90:76-153)
        d2 = torch.empty_like(s)
        d1 = torch.empty_like(s)
        [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z][idir](opr.OPR_P2_P1, nx, ny, nz, bcs, gg[idir], s, d2, d1)
        ident = (visc * d2 - s * d1).cpu().numpy()
        assert rel_l2(res.cpu().numpy(), ident) <= 1e-11


@pytest.mark.parametrize("shape", [(40, 37, 24), (32, 256, 16), (64, 512, 16)])     # general kernel / fast kernel: CTAs of the wall chunks only, 32 and 64 lines
def test_boundary_bcs_neumann_y(cuda, shape):
    import torch
    from oracle import operators as O
    from tlab_b200 import opr
    nx, ny, nz = shape
    grids, go, gg = _plans(nx, ny, nz, "tanh")
    a = smooth_field((nz, ny, nx), grids, seed=5)
    u = torch.from_numpy(a).to(cuda)
    for ibc in (1, 2, 3):
        hb = torch.full((nz, nx), float("nan"), dtype=torch.float64, device=cuda)
        ht = torch.full((nz, nx), float("nan"), dtype=torch.float64, device=cuda)
        opr.BOUNDARY_BCS_NEUMANN_Y(ibc, nx, ny, nz, gg[1], u, hb, ht)
        rb, rt = O.boundary_bcs_neumann_y(ibc, go[1], a)
        if ibc in (1, 3):
            assert rel_l2(hb.cpu().numpy(), rb) <= TOL
        if ibc in (2, 3):
            assert rel_l2(ht.cpu().numpy(), rt) <= TOL


def test_fdm_solve_lines_first(cuda):
    """FDM_Der1_Solve / FDM_Der2_Solve on the (nlines, n) view, as src/valid/fdm/vpartial.f90 drives them."""
    import torch
    from oracle import fdm
    from tlab_b200 import opr
    n, nlines = 256, 6
    y = grid_tanh(n)
    go = fdm.Plan(y, False, False)
    gg = opr.FdmPlan(y, False, False, name="y")
    rng = np.random.default_rng(7)
    u_np = np.exp(-((y[:, None] - 0.5) / 0.1) ** 2) * rng.uniform(0.5, 1.5, (1, nlines))   # (n, nlines)
    u = torch.from_numpy(np.ascontiguousarray(u_np)).to(cuda)
    for ibc in range(4):
        r = torch.empty_like(u)
        opr.FDM_Der1_Solve(nlines, ibc, gg, u, r)
        assert rel_l2(r.cpu().numpy(), fdm.der1_solve(ibc, go.der1, go.der1.lu, u_np)) <= TOL
    r = torch.empty_like(u)
    opr.FDM_Der2_Solve(nlines, gg, u, r)
    d1 = fdm.der1_solve(0, go.der1, go.der1.lu, u_np)
    assert rel_l2(r.cpu().numpy(), fdm.der2_solve(go.der2, go.der2.lu, u_np, d1)) <= TOL


@pytest.mark.parametrize("shape", [(128, 96, 64), (1024, 32, 16), (16, 64, 1024)])
def test_fast_and_general_kernels_agree(cuda, shape):
    """The fast kernels (lines2.cu) and the general ones (lines.cu) are two formulations of the same solves.  On long
    uniform lines the fast kernels replace the reference's LU factors in the interior by their converged values; the
    factors carry round-off noise of relative size ~eps*n from the numerically differentiated Jacobian, hence 5e-13."""
    import torch
    from tlab_b200 import lib as tl, opr
    nx, ny, nz = shape
    grids, go, gg = _plans(nx, ny, nz, "tanh")
    visc = 1.0 / 5000.0
    opr.OPR_Burgers_Initialize(gg, visc, [1.0])
    u = torch.from_numpy(smooth_field((nz, ny, nx), grids, seed=3)).to(cuda)
    v = torch.from_numpy(smooth_field((nz, ny, nx), grids, seed=4)).to(cuda)
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    B = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    bcs = [[0, 0], [0, 0]]
    out = {}
    import ctypes
    cnt = {}
    try:
        for fast in (0, 1):
            tl.check(tl.load().tlab_gpu_set_tuning(b"fast", fast))
            c = ctypes.c_longlong()
            tl.check(tl.load().tlab_gpu_get_counter(b"fast_launches", ctypes.byref(c)))
            cnt[fast] = c.value
            res = []
            for d in range(3):
                r1, r2, r3, r4 = (torch.full_like(u, float("nan")) for _ in range(4))
                P[d](opr.OPR_P2_P1, nx, ny, nz, bcs, gg[d], u, r1, r2)
                B[d](opr.OPR_B_U_IN, 0, nx, ny, nz, bcs, u, v, r3)
                P[d](opr.OPR_P1, nx, ny, nz, bcs, gg[d], u, r4)
                res += [r1, r2, r3, r4]
            out[fast] = res
    finally:
        tl.check(tl.load().tlab_gpu_set_tuning(b"fast", 1))
    c = ctypes.c_longlong()
    tl.check(tl.load().tlab_gpu_get_counter(b"fast_launches", ctypes.byref(c)))
    assert cnt[1] == cnt[0] and c.value == cnt[1] + 9, "the fast kernels did not run"
    for a, b in zip(out[0], out[1]):
        assert float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(a)) <= (1e-13 if max(shape) <= 128 else 5e-13)


DIRECT_CASES = [(32, 65, 16, "tanh", True, True),        # the usual use: stretched wall-normal direction, periodic x/z fall back to the hyper scheme
                (48, 40, 33, "stretched", False, False),  # per-row coefficients in all three directions, ragged chunks
                (16, 17, 16, "tanh", True, True),         # a single chunk (n = 17: the boundary and interior row classes all touch)
                (32, 256, 32, "tanh", True, True)]        # 16 chunks: lines the fast kernels would have taken


@pytest.mark.parametrize("case", DIRECT_CASES)
def test_compact_direct6_second_derivative(cuda, case):
    """SpaceOrder2 = CompactDirect6 (src/fdm/fdm_comx_direct.f90:305-412, MatMul_5d fdm_matmul.f90:266-320):
    OPR_Partial P2 / P2_P1 and OPR_Burgers against the oracle; the first derivative keeps CompactJacobian6."""
    import torch
    from oracle import fdm, operators as O
    from tlab_b200 import opr
    nx, ny, nz, ykind, xper, zper = case
    x = grid_periodic(nx) if xper else grid_stretched(nx, 2.0)
    z = grid_periodic(nz) if zper else grid_stretched(nz, 3.0)
    y = grid_tanh(ny) if ykind == "tanh" else grid_stretched(ny)
    grids = (x, y, z)
    go = [fdm.Plan(x, xper, xper, name="x", mode2=fdm.FDM_COM6_DIRECT), fdm.Plan(y, False, False, name="y", mode2=fdm.FDM_COM6_DIRECT),
          fdm.Plan(z, zper, zper, name="z", mode2=fdm.FDM_COM6_DIRECT)]
    gg = [opr.FdmPlan(x, xper, xper, name="x", der2="compactdirect6"), opr.FdmPlan(y, False, False, name="y", der2="compactdirect6"),
          opr.FdmPlan(z, zper, zper, name="z", der2="compactdirect6")]
    a = smooth_field((nz, ny, nx), grids)
    u = torch.from_numpy(a).to(cuda)
    fns = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    bcs = [[0, 0], [0, 0]]
    for idir in range(3):
        for type_ in (O.OPR_P1, O.OPR_P2, O.OPR_P2_P1):
            res = torch.full_like(u, float("nan"))
            tmp = torch.full_like(u, float("nan"))
            fns[idir](type_, nx, ny, nz, bcs, gg[idir], u, res, tmp if type_ == O.OPR_P2_P1 else None)
            ref = O.opr_partial(idir, type_, bcs, go[idir], a)
            if type_ == O.OPR_P2_P1:
                assert rel_l2(res.cpu().numpy(), ref[0]) <= TOL, (idir, type_)
                assert rel_l2(tmp.cpu().numpy(), ref[1]) <= TOL, (idir, type_)
            else:
                assert rel_l2(res.cpu().numpy(), ref) <= TOL, (idir, type_)
    # the scheme really is a different one from the Jacobian form on a stretched grid
    gj = fdm.Plan(y, False, False, name="y")
    d2j = O.opr_partial(1, O.OPR_P2, bcs, gj, a)
    d2d = O.opr_partial(1, O.OPR_P2, bcs, go[1], a)
    assert 1e-10 < rel_l2(d2d, d2j) < 0.5
    visc, schmidt = 1.0 / 5000.0, [1.0]
    B = O.Burgers(go, visc, schmidt)
    opr.OPR_Burgers_Initialize(gg, visc, schmidt)
    s_np = smooth_field((nz, ny, nx), grids, seed=1)
    v_np = smooth_field((nz, ny, nx), grids, seed=2)
    s, v = torch.from_numpy(s_np).to(cuda), torch.from_numpy(v_np).to(cuda)
    bfn = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    for idir in range(3):
        for is_ in range(2):
            res = torch.full_like(s, float("nan"))
            bfn[idir](opr.OPR_B_U_IN, is_, nx, ny, nz, bcs, s, v, res)
            assert rel_l2(res.cpu().numpy(), B.apply(idir, is_, bcs, s_np, v_np)) <= TOL, ("u_in", idir, is_)


@pytest.mark.parametrize("shape", [(256, 32, 128), (1024, 16, 32), (128, 16, 512)])
def test_circulant_form_and_closure_form_agree_with_the_oracle(cuda, shape):
    """Periodic directions: the fast kernels solve the circulant systems either with the reference's rank-one closure
    (TRIDPFS / TRIDPSS, linear3.f90:321-442: per-point tables in the chunks at the two ends of a line) or in circulant form
    (constant chunks everywhere, windows wrapping around the line; the default).  Both against the oracle."""
    import torch
    from oracle import operators as O
    from tlab_b200 import lib as tl, opr
    nx, ny, nz = shape
    grids, go, gg = _plans(nx, ny, nz, "tanh")
    visc = 1.0 / 5000.0
    Bo = O.Burgers(go, visc, [1.0])
    opr.OPR_Burgers_Initialize(gg, visc, [1.0])
    a = smooth_field((nz, ny, nx), grids, seed=7)
    b = smooth_field((nz, ny, nx), grids, seed=8)
    u, v = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    P = [opr.OPR_Partial_X, None, opr.OPR_Partial_Z]
    B = [opr.OPR_Burgers_X, None, opr.OPR_Burgers_Z]
    bcs = [[0, 0], [0, 0]]
    L = tl.load()
    try:
        for circ in (0, 1):
            tl.check(L.tlab_gpu_set_tuning(b"circ", circ))
            tl.check(L.tlab_gpu_set_tuning(b"march", 0))          # the line kernels of lines2.cu in both directions
            for d in (0, 2):
                r1, r2, r3, r4 = (torch.full_like(u, float("nan")) for _ in range(4))
                P[d](opr.OPR_P2_P1, nx, ny, nz, bcs, gg[d], u, r1, r2)
                ref = O.opr_partial(d, O.OPR_P2_P1, bcs, go[d], a)
                assert rel_l2(r1.cpu().numpy(), ref[0]) <= TOL and rel_l2(r2.cpu().numpy(), ref[1]) <= TOL, (circ, d)
                P[d](opr.OPR_P1, nx, ny, nz, bcs, gg[d], u, r4)
                assert rel_l2(r4.cpu().numpy(), O.opr_partial(d, O.OPR_P1, bcs, go[d], a)) <= TOL, (circ, d)
                B[d](opr.OPR_B_U_IN, 1, nx, ny, nz, bcs, u, v, r3)
                assert rel_l2(r3.cpu().numpy(), Bo.apply(d, 1, bcs, a, b)) <= TOL, (circ, d)
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"circ", 1))
        tl.check(L.tlab_gpu_set_tuning(b"march", 1))
