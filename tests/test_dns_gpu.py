"""GPU parity of the RHS and the Runge-Kutta advance: rel. L2 <= 1e-12 on hq/hs after one RHS call,
<= 1e-10 on the fields after 10 RK steps (BASELINE.json north_star)."""
import numpy as np
import pytest

from common import grid_periodic, grid_tanh, grid_stretched, smooth_field, rel_l2

pytestmark = pytest.mark.gpu


def _pair(nx, ny, nz, ykind="tanh", rkm=4, free_slip_top=True, buoy=True):
    from oracle import fdm, dns as OD
    from tlab_b200 import opr, dns as GD
    x = grid_periodic(nx)
    z = grid_periodic(nz) if nz > 1 else np.zeros(1)
    y = {"stretched": grid_stretched(ny), "tanh": grid_tanh(ny), "uniform": np.linspace(0, 1, ny)}[ykind]
    yuni = ykind == "uniform"
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, yuni, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, yuni, name="y"), opr.FdmPlan(z, True, True, name="z")]
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[1.0], rkm_mode=rkm,
              bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N) if free_slip_top else (D, D, D),
              bcs_scal_jmin=(D,), bcs_scal_jmax=(N,) if free_slip_top else (D,))
    if buoy:
        kw.update(buoyancy_type="linear", buoyancy_params=(1.0, 0.0), buoyancy_vector=(0.0, 1.0, 0.0),
                  bbackground=0.1 * y)
    o = OD.Dns(go, **kw)
    g = GD.Dns(gg, **kw)
    grids = (x, y, z)
    shape = (nz, ny, nx)
    wall = np.sin(np.pi * y / y[-1] / 2.0)[None, :, None] if free_slip_top else np.sin(np.pi * y / y[-1])[None, :, None]
    fields = [0.5 * smooth_field(shape, grids, seed=31 + i) * wall for i in range(3)]
    if nz == 1:
        fields[2][:] = 0.0
    sc = 0.5 + 0.2 * smooth_field(shape, grids, seed=40) * wall / 3.0
    for i in range(3):
        o.q[i][...] = fields[i]
        g.set("q%d" % (i + 1), fields[i])
    o.s[0][...] = sc
    g.set("s1", sc)
    return o, g


@pytest.mark.parametrize("case", [(32, 33, 16, "tanh", True), (48, 40, 20, "stretched", False), (64, 32, 1, "uniform", True)])
def test_rhs_single_call(cuda, case):
    nx, ny, nz, ykind, fs = case
    o, g = _pair(nx, ny, nz, ykind, free_slip_top=fs)
    dte = 1e-3
    o.dte = dte
    o.sources_flow()
    o.rhs_global_incompressible_1()
    g.substep(dte, 0.0, False)       # sources + rhs + q update; compare hq (unchanged by the update) and q
    for i in range(3):
        e = rel_l2(g.get("hq%d" % (i + 1)), o.hq[i])
        assert e <= 1e-12, ("hq", i, e)
    assert rel_l2(g.get("hs1"), o.hs[0]) <= 1e-12
    assert rel_l2(g.get("p"), o.last_pressure) <= 1e-11


@pytest.mark.parametrize("rkm", [3, 4])
def test_ten_rk_steps(cuda, rkm):
    o, g = _pair(32, 33, 16, "tanh", rkm=rkm)
    dt = 2e-3
    for it in range(10):
        o.runge_kutta(dt)
        g.runge_kutta(dt)
    for i in range(3):
        e = rel_l2(g.get("q%d" % (i + 1)), o.q[i])
        assert e <= 1e-10, ("q", i, e)
    assert rel_l2(g.get("s1"), o.s[0]) <= 1e-10


def test_rk_step_from_host_buffers(cuda):
    import torch
    o, g = _pair(32, 33, 16, "tanh")
    N = 32 * 33 * 16
    qh = torch.empty(3 * N, dtype=torch.float64).pin_memory()
    sh = torch.empty(N, dtype=torch.float64).pin_memory()
    for i in range(3):
        qh[i * N:(i + 1) * N] = torch.from_numpy(o.q[i].ravel())
    sh[:] = torch.from_numpy(o.s[0].ravel())
    o.runge_kutta(1e-3)
    g.runge_kutta_host(1e-3, qh.data_ptr(), sh.data_ptr())
    for i in range(3):
        assert rel_l2(qh[i * N:(i + 1) * N].numpy(), o.q[i].ravel()) <= 1e-11
    assert rel_l2(sh.numpy(), o.s[0].ravel()) <= 1e-11
    assert g.launch_count() > 0


def test_scalar_clipping(cuda):
    """DNS_BOUNDS_LIMIT: default [0,1] clipping of the scalar is part of the substep."""
    o, g = _pair(32, 33, 16, "tanh")
    big = o.s[0] * 3.0 - 0.7
    o.s[0][...] = big
    g.set("s1", big)
    o.runge_kutta(1e-3)
    g.runge_kutta(1e-3)
    s = g.get("s1")
    assert s.min() >= 0.0 and s.max() <= 1.0
    assert rel_l2(s, o.s[0]) <= 1e-11


def test_courant_and_dilatation(cuda):
    """TIME_COURANT and the dilatation bounds of DNS_BOUNDS_CONTROL (the dns.out columns dt, CFL#, D#, DilMin, DilMax)."""
    o, g = _pair(32, 33, 16, "tanh")
    for cfla in (1.2, -1.0):
        ro = o.courant(cfla, dtime=1e-3)
        rg = g.courant(cfla, dtime=1e-3)
        for a, b in zip(rg, ro):
            assert abs(a - b) <= 1e-13 * max(1.0, abs(b)), (cfla, rg, ro)
    do, dg = o.bounds_control(), g.bounds_control()
    scale = max(abs(do[0]), abs(do[1]))
    assert abs(dg[0] - do[0]) <= 1e-12 * scale and abs(dg[1] - do[1]) <= 1e-12 * scale
    dt = ro[0]
    o.runge_kutta(1e-3)
    g.runge_kutta(1e-3)
    do, dg = o.bounds_control(), g.bounds_control()
    # after a step the interior is divergence free to round-off; the extrema sit at the walls
    assert abs(dg[0] - do[0]) <= 1e-9 * max(1.0, abs(do[0])) and abs(dg[1] - do[1]) <= 1e-9 * max(1.0, abs(do[1]))


def test_case01_shape_two_dimensional_step(cuda):
    """BASELINE config 1: the 2-D shape of examples/Case01 (512x256x1), one RK4-5 step compared field by field.
    (The reference's broadband initial condition needs its Fortran RNG; a deterministic shear layer is used.)"""
    from oracle import fdm, dns as OD
    from tlab_b200 import opr, dns as GD
    nx, ny = 512, 256
    x = grid_periodic(nx, 2.0)
    y = np.linspace(0.0, 1.0, ny)
    z = np.zeros(1)
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, True, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, True, name="y"), opr.FdmPlan(z, True, True, name="z")]
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 1600.0, schmidt=[1.0], bcs_flow_jmin=(N, D, N), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(N,), bcs_scal_jmax=(N,))
    o, g = OD.Dns(go, **kw), GD.Dns(gg, **kw)
    Y, X = np.meshgrid(y, x, indexing="ij")
    u = 0.5 * np.tanh((Y - 0.5) / 0.05) + 0.01 * np.sin(2 * np.pi * X) * np.exp(-((Y - 0.5) / 0.1) ** 2)
    v = 0.01 * np.cos(2 * np.pi * X) * np.exp(-((Y - 0.5) / 0.1) ** 2)
    s = 0.5 * (1.0 + np.tanh((Y - 0.5) / 0.05))
    for i, f in enumerate((u, v, np.zeros_like(u))):
        o.q[i][...] = f[None]
        g.set("q%d" % (i + 1), f[None])
    o.s[0][...] = s[None]
    g.set("s1", s[None])
    dt = o.courant(1.2)[0]
    o.runge_kutta(dt)
    g.runge_kutta(dt)
    for i in range(2):
        assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
    assert np.abs(g.get("q3")).max() == 0.0
    assert rel_l2(g.get("s1"), o.s[0]) <= 1e-11


@pytest.mark.parametrize("tune", [{"fuse": 1}, {"fuse": 1, "pf_next": 1}, {"persist": 1}, {"pf_dist": 3}, {"fast": 0},
                                  {"tma": 1}, {"poisson_split": 0}, {"poisson_split": 1},
                                  {"neu_compact": 0}, {"fuse_update": 0}, {"lazy_scale": 0}, {"lazy_scale": 1, "fuse": 1},
                                  {"lazy_scale": 1, "fast": 0}, {"march": 0}, {"march": 2}])
def test_tuning_variants_give_the_same_step(cuda, tune):
    """The optional kernel variants (fused multi-field Burgers launch, next-field / next-tile L2 prefetch, persistent
    cp.async staging, general kernels) are alternative schedules of the same arithmetic: one RK step on full chunks
    (64 x 64 x 32) must agree with the oracle like the default path does."""
    from tlab_b200 import lib as tl
    L = tl.load()
    defaults = {"fuse": 0, "pf_next": 0, "persist": 0, "pf_dist": -1, "fast": 1, "tma": 0, "poisson_split": -1, "neu_compact": 1, "fuse_update": 1,
                "lazy_scale": 1, "march": 1}
    try:
        for k, v in tune.items():
            tl.check(L.tlab_gpu_set_tuning(k.encode(), v))
        o, g = _pair(64, 64, 32, "tanh")
        o.runge_kutta(1e-3)
        g.runge_kutta(1e-3)
        for i in range(3):
            assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
        assert rel_l2(g.get("s1"), o.s[0]) <= 1e-11
    finally:
        for k, v in defaults.items():
            tl.check(L.tlab_gpu_set_tuning(k.encode(), v))


@pytest.mark.parametrize("tune", [{"fuse": 1}, {"persist": 1, "march": 0}, {"tma": 1, "march": 0}, {"pair": 1, "march": 0}, {"march": 0},
                                  {"march": 2}, {"march": 2, "march_peel": 0}, {"circ": 0}, {"fast": 0}, {"lines_x": 8}, {"lines_yz": 8, "march": 0}])
def test_tuning_variants_on_lines_long_enough_for_the_circulant_form(cuda, tune):
    """The same on 128 x 192 x 128: x and z lines of 8 chunks (circulant form in every whole-line variant: fused launch,
    cp.async staging, TMA tiles, CTA pairs, other tilings; and its closure-form fallback circ = 0), y lines of 12 chunks
    (unscaled constant chunks; marching kernel with and without the peeled constant-only steps)."""
    from tlab_b200 import lib as tl
    L = tl.load()
    defaults = {"fuse": 0, "persist": 0, "tma": 0, "pair": 0, "march": 1, "march_peel": 1, "circ": 1, "fast": 1, "lines_x": 0, "lines_yz": 0}
    try:
        for k, v in tune.items():
            tl.check(L.tlab_gpu_set_tuning(k.encode(), v))
        o, g = _pair(128, 192, 128, "tanh")
        if "ref" not in _LONG_REF:                 # the oracle's step once for all variants (it is the slow part)
            o.runge_kutta(1e-3)
            _LONG_REF["ref"] = [q.copy() for q in o.q] + [o.s[0].copy()]
        g.runge_kutta(1e-3)
        for name, ref in zip(("q1", "q2", "q3", "s1"), _LONG_REF["ref"]):
            assert rel_l2(g.get(name), ref) <= 1e-11
    finally:
        for k, v in defaults.items():
            tl.check(L.tlab_gpu_set_tuning(k.encode(), v))


_LONG_REF = {}


@pytest.mark.parametrize("shape,lines", [((16, 64, 128), 0), ((16, 32, 1024), 0), ((16, 32, 1024), 4), ((16, 512, 16), 4),
                                         ((16, 512, 16), 8)])
def test_step_through_the_tma_kernels(cuda, shape, lines):
    """y and z operators as persistent CTAs fed by the TMA unit (tensor-map loads, reduce-add stores into hq/hs):
    32/16/8/4 lines per CTA (row-order permutations in shared memory), 3-D (y) and 2-D (z) tensor maps, second input
    and accumulation.  The step must agree with the oracle and the TMA kernels must actually have run."""
    import ctypes
    from tlab_b200 import lib as tl
    L = tl.load()
    before = ctypes.c_longlong(0)
    tl.check(L.tlab_gpu_get_counter(b"tma_launches", ctypes.byref(before)))
    try:
        tl.check(L.tlab_gpu_set_tuning(b"tma", 1))
        tl.check(L.tlab_gpu_set_tuning(b"lines_yz", lines))
        o, g = _pair(*shape, "tanh")
        o.runge_kutta(1e-3)
        g.runge_kutta(1e-3)
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"lines_yz", 0))
        tl.check(L.tlab_gpu_set_tuning(b"tma", 0))
    after = ctypes.c_longlong(0)
    tl.check(L.tlab_gpu_get_counter(b"tma_launches", ctypes.byref(after)))
    assert after.value - before.value >= 5 * 10, "the TMA line kernels did not run"
    for i in range(3):
        assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
    assert rel_l2(g.get("s1"), o.s[0]) <= 1e-11


@pytest.mark.parametrize("march,circ", [(1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("shape,pv", [((32, 32, 256), 2), ((16, 32, 192), 2), ((16, 32, 512), 4), ((64, 16, 384), 3),
                                      ((32, 16, 512), 2), ((32, 32, 768), 3)])
def test_split_z_operators_on_virtual_slabs(cuda, shape, pv, march, circ):
    """The split-z kernels of splitz.cu (z operators of a z-split domain without transposes: halo planes and chunk ends
    exchanged between neighbouring slabs) run over pv virtual slabs of one field on one GPU: the RK step must agree with
    the oracle, and with the whole-line kernels to round-off.  Slab thicknesses 96 (the minimum: 6 chunks) to 256 planes; with
    march = 1 the finishing phase runs as a march over panels of 32 lines (splitz_march_kernel) when the slab holds a multiple
    of 4 chunks, seeded from the neighbours' chunk ends -- in circulant form (circ = 1: constant chunks, windows wrapping from rank
    to rank, every rank the same work) or with the rank-one closure of the reference's periodic solver (circ = 0: closure terms
    traded between the first and the last rank)."""
    import ctypes
    from tlab_b200 import lib as tl
    L = tl.load()
    o, g = _pair(*shape, "tanh")
    o.runge_kutta(1e-3)
    g.runge_kutta(1e-3)
    whole = [g.get("q%d" % (i + 1)) for i in range(3)] + [g.get("s1")]
    before, mbefore = ctypes.c_longlong(0), ctypes.c_longlong(0)
    tl.check(L.tlab_gpu_get_counter(b"splitz_ops", ctypes.byref(before)))
    tl.check(L.tlab_gpu_get_counter(b"splitz_march_ops", ctypes.byref(mbefore)))
    try:
        tl.check(L.tlab_gpu_set_tuning(b"split_emulate", pv))
        tl.check(L.tlab_gpu_set_tuning(b"march", march))
        tl.check(L.tlab_gpu_set_tuning(b"circ", circ))
        _, g2 = _pair(*shape, "tanh")
        g2.runge_kutta(1e-3)
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"split_emulate", 0))
        tl.check(L.tlab_gpu_set_tuning(b"march", 1))
        tl.check(L.tlab_gpu_set_tuning(b"circ", 1))
    after, mafter = ctypes.c_longlong(0), ctypes.c_longlong(0)
    tl.check(L.tlab_gpu_get_counter(b"splitz_ops", ctypes.byref(after)))
    tl.check(L.tlab_gpu_get_counter(b"splitz_march_ops", ctypes.byref(mafter)))
    assert after.value - before.value == 5 * 6, "the split-z kernels did not run"
    chunks = shape[2] // pv // 16
    eligible = march == 1 and chunks % 4 == 0 and chunks >= 8 and (shape[0] * shape[1]) % 32 == 0
    assert mafter.value - mbefore.value == (5 * 6 * pv if eligible else 0), (mafter.value - mbefore.value, eligible)
    split = [g2.get("q%d" % (i + 1)) for i in range(3)] + [g2.get("s1")]
    ref = o.q + o.s
    for a, b, c in zip(split, whole, ref):
        assert rel_l2(a, c) <= 1e-11
        assert rel_l2(a, b) <= 1e-13


def test_restart_files_round_trip(cuda, tmp_path):
    """q, s written as tlab restart files (flow.<it>.<n>, scal.<it>.<n>; tlab_b200/io.py) and read back into a second
    instance: the next RK step of both instances is identical."""
    o, g = _pair(32, 33, 16, "tanh")
    g.runge_kutta(1e-3)
    flow, scal = str(tmp_path / "flow.1"), str(tmp_path / "scal.1")
    g.write_restart(flow, scal, nt=1, rtime=1e-3, visc=1.0 / 5000.0, schmidt=[1.0])
    _, g2 = _pair(32, 33, 16, "tanh")
    nt, rtime = g2.read_restart(flow, scal)
    assert nt == 1 and rtime == 1e-3
    g.runge_kutta(1e-3)
    g2.runge_kutta(1e-3)
    for name in ("q1", "q2", "q3", "s1"):
        assert np.array_equal(g.get(name), g2.get(name))


@pytest.mark.parametrize("pair", [1, 0])
def test_long_periodic_lines_cta_pairs(cuda, pair):
    """z lines of 64 chunks (nz = 1024): with pair = 1 a line is shared by a cluster of 2 CTAs exchanging chunk ends through
    distributed shared memory (16 lines per tile, 128-byte rows); pair = 0 (default) keeps one CTA of 8 lines.  Same step."""
    from tlab_b200 import lib as tl
    L = tl.load()
    try:
        tl.check(L.tlab_gpu_set_tuning(b"pair", pair))
        o, g = _pair(16, 32, 1024, "tanh")
        o.runge_kutta(1e-3)
        g.runge_kutta(1e-3)
    finally:
        tl.check(L.tlab_gpu_set_tuning(b"pair", 0))
    for i in range(3):
        assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
    assert rel_l2(g.get("s1"), o.s[0]) <= 1e-11


def test_two_scalars_with_different_diffusivities(cuda):
    """inb_scal = 2: the "further scalars" branch of the RHS (rhs_global_incompressible_1.f90:149-162 loops over every
    scalar with its own diffusivity visc/Sc_is and its own boundary conditions).  Two RK steps against the oracle.
    (The reference examples with two prognostic scalars -- Case11/12/17-19 -- all add physics outside SURVEY section 8:
    buffer-zone relaxation, AirWaterLinear thermodynamics, radiation; hence the oracle, not a dns.out.ref golden.)"""
    from oracle import fdm, dns as OD
    from tlab_b200 import opr, dns as GD
    nx, ny, nz = 48, 64, 32
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    D, N = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[1.0, 0.7], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
              buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(D, N), bcs_scal_jmax=(N, D))
    o, g = OD.Dns(go, **kw), GD.Dns(gg, **kw)
    shape = (nz, ny, nx)
    wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
    for i in range(3):
        f = 0.5 * smooth_field(shape, (x, y, z), seed=31 + i) * wall
        o.q[i][...] = f
        g.set("q%d" % (i + 1), f)
    for i in range(2):
        sc = 0.5 + 0.1 * smooth_field(shape, (x, y, z), seed=40 + i) * wall
        o.s[i][...] = sc
        g.set("s%d" % (i + 1), sc)
    o.dte = 1e-3
    o.sources_flow()
    o.rhs_global_incompressible_1()
    g.substep(1e-3, 0.0, False)
    for i in range(2):
        assert rel_l2(g.get("hs%d" % (i + 1)), o.hs[i]) <= 1e-12
    _, g = None, GD.Dns(gg, **kw)
    o = OD.Dns(go, **kw)
    for i in range(3):
        f = 0.5 * smooth_field(shape, (x, y, z), seed=31 + i) * wall
        o.q[i][...] = f
        g.set("q%d" % (i + 1), f)
    for i in range(2):
        sc = 0.5 + 0.1 * smooth_field(shape, (x, y, z), seed=40 + i) * wall
        o.s[i][...] = sc
        g.set("s%d" % (i + 1), sc)
    for _ in range(2):
        o.runge_kutta(1e-3)
        g.runge_kutta(1e-3)
    for i in range(3):
        assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
    for i in range(2):
        assert rel_l2(g.get("s%d" % (i + 1)), o.s[i]) <= 1e-11


def test_two_live_states_of_different_size(cuda):
    """Two tlab_dns_t handles alive at once, different grids and different viscosities, stepped alternately: each owns its
    Poisson solver (cuFFT plans, eigenvalues, per-mode planes) and re-points the diffusion-scaled LU sets of its plans
    before every RHS, so neither disturbs the other (the reference has one set of module variables per process; a library
    with handles must not)."""
    o1, g1 = _pair(32, 33, 16, "tanh")
    o2, g2 = _pair(64, 48, 32, "stretched", free_slip_top=False)
    for _ in range(2):
        o1.runge_kutta(1e-3)
        o2.runge_kutta(2e-3)
        g1.runge_kutta(1e-3)
        g2.runge_kutta(2e-3)
    for o, g in ((o1, g1), (o2, g2)):
        for i in range(3):
            assert rel_l2(g.get("q%d" % (i + 1)), o.q[i]) <= 1e-11
        assert rel_l2(g.get("s1"), o.s[0]) <= 1e-11
    # the stand-alone operator API keeps its own (process-wide) solver: initialising it must not disturb the handles either
    from tlab_b200 import opr
    opr.OPR_Elliptic_Initialize(g2.g)
    o1.runge_kutta(1e-3)
    g1.runge_kutta(1e-3)
    assert rel_l2(g1.get("q2"), o1.q[1]) <= 1e-11


def test_pending_rk_factor_is_invisible_to_the_host(cuda):
    """`hq = hq*kco` is not written by the update kernel but folded into the first accumulation of the next substep; a host that
    reads hq / hs between stages (tlab_dns_download_host, tlab_dns_field) must still see the scaled arrays, bit for bit the ones
    of the eager schedule, and the next stage must not apply the factor twice."""
    from tlab_b200 import lib as tl
    L = tl.load()
    out = {}
    for lazy in (0, 1):
        tl.check(L.tlab_gpu_set_tuning(b"lazy_scale", lazy))
        try:
            o, g = _pair(64, 64, 32, "tanh")
            g.runge_kutta_stage(1e-3, 0)
            g.runge_kutta_stage(1e-3, 1)
            mid = [g.get(n) for n in ("hq1", "hq2", "hq3", "hs1")]       # flushes the pending factor
            g.runge_kutta_stage(1e-3, 2)
            end = [g.get(n) for n in ("hq1", "hq2", "hq3", "hs1", "q1", "q2", "q3", "s1")]
            out[lazy] = (mid, end)
            g.close()
        finally:
            tl.check(L.tlab_gpu_set_tuning(b"lazy_scale", 1))
    for a, b in zip(out[0][0] + out[0][1], out[1][0] + out[1][1]):
        assert np.array_equal(a, b)
