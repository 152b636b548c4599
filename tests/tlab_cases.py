"""Deterministic example cases of the reference restated from their tlab.ini, so that the `dns.out.ref` the reference
build logged for each of them -- time, dt, CFL number, diffusion number, min/max of the dilatation over ten iterations --
is a golden vector.  All of them: 2-D, incompressible Boussinesq, CompactJacobian6 +
CompactJacobian6Hyper, RK4-5, TimeCFL 1.2, linear buoyancy b = s along y with the background profile subtracted
(physics/tlab_background.f90:216-223), scalar clipped to [0, 1] (dns_read_local.f90:175-204).

  grid     tools/initialize/grid/grid_main.f90:55-110; grid_local.f90:41-66 (BLD_TANH, up to three modes)
  scalar   tools/initialize/scal/scal_local.f90:244-340 (SCAL_FLUCTUATION_PLANE): PlaneDiscrete displaces the profile,
           DeltaDiscrete perturbs its thickness, by sum_m A_m cos(m 2 pi x / scale_x + phi_m); utils/profiles.f90:169-232
           (Erf: mean + delta/2 erf(-(y - ymean) / (2 thick)), ymean = y_1 + scale_y * YMeanRelative, default 0.5);
           utils/discrete.f90:43-85
  loop     tools/dns/dns_main.f90: TIME_RUNGEKUTTA, TIME_COURANT, DNS_BOUNDS_CONTROL, one log line per iteration

Case10: convective boundary layer, 512 x 257, one tanh mode, no-slip / free-slip, scalar Dirichlet / Neumann.
Case06 / Case07: unstable / stable density interface, 512 x 256 (full 16-point chunks in y: the fast line kernels),
two tanh modes, free-slip walls, Neumann scalar."""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CFL = 1.2

CASES = {
    "case10": dict(nx=512, ny=257, scale_x=2.0, tanh=[(0.9375, 2.0, 0.0078125)], reynolds=2000.0,
                   thick=0.02, delta=2.0, mean=1.0, ymean_rel=0.0, pert="delta", ampl=[0.0, 0.0, 0.001],
                   flow_jmin="noslip", flow_jmax="freeslip", scal_jmin="dirichlet", scal_jmax="neumann"),
    "case06": dict(nx=512, ny=256, scale_x=2.0, tanh=[(0.0625, 2.0, -0.0078125), (0.9375, 2.0, 0.0078125)], reynolds=5000.0,
                   thick=0.005859375, delta=-1.0, mean=0.5, ymean_rel=0.5, pert="plane", ampl=[0.0, 0.0, 0.0, 0.029296875],
                   flow_jmin="freeslip", flow_jmax="freeslip", scal_jmin="neumann", scal_jmax="neumann"),
    "case07": dict(nx=512, ny=256, scale_x=2.0, tanh=[(0.0625, 2.0, -0.0078125), (0.9375, 2.0, 0.0078125)], reynolds=5000.0,
                   thick=0.005859375, delta=1.0, mean=0.5, ymean_rel=0.5, pert="plane", ampl=[0.0, 0.0, 0.0, 0.029296875],
                   flow_jmin="freeslip", flow_jmax="freeslip", scal_jmin="neumann", scal_jmax="neumann"),
}


def grids(c):
    nx, ny = c["nx"], c["ny"]
    x = np.arange(nx + 1) * (c["scale_x"] / nx)       # nx + 1 points on [0, scale], the last one dropped (periodic)
    s = np.arange(ny) * (1.0 / (ny - 1))
    work = np.zeros(ny)
    for st, f, delta in c["tanh"]:
        work = work + (f - 1.0) * delta * np.log(np.exp((s - st) / delta) + 1.0)
    y = s + (work - work[0])
    return x[:nx].copy(), y, np.zeros(1)


def erf_profile(c, y, thick, ymean):
    xi = (y - ymean) / thick
    return c["mean"] + c["delta"] * 0.5 * np.vectorize(math.erf)(-0.5 * xi)


def _ymean(c, y):
    return y[0] + (y[-1] - y[0]) * c["ymean_rel"]


def initial_scalar(c, x, y):
    disp = np.zeros_like(x)
    for m, a in enumerate(c["ampl"], start=1):
        if a != 0.0:
            disp = disp + a * np.cos(m * (2.0 * np.pi / c["scale_x"]) * x)
    ym = _ymean(c, y)
    if c["pert"] == "delta":
        s = erf_profile(c, y[:, None], (c["thick"] + disp)[None, :], ym)
    else:
        s = erf_profile(c, y[:, None] - disp[None, :], c["thick"], ym)
    return s[None]                                     # (1, ny, nx)


def background(c, y):
    return erf_profile(c, y, c["thick"], _ymean(c, y))


def reference_log(name):
    rows = []
    for line in open(os.path.join(HERE, "golden", "%s_dns.out.ref" % name)):
        if line.startswith("#"):
            continue
        t = line.split()
        rows.append(dict(it=int(t[1]), time=float(t[2]), dt=float(t[3]), cfl=float(t[4]), dif=float(t[5]), visc=float(t[6]),
                         dilmin=float(t[7]), dilmax=float(t[8])))
    return rows


def dns_kwargs(c, mod, y):
    D, N = mod.DNS_BCS_DIRICHLET, mod.DNS_BCS_NEUMANN
    flow = {"noslip": (D, D, D), "freeslip": (N, D, N)}
    scal = {"dirichlet": (D,), "neumann": (N,)}
    return dict(visc=1.0 / c["reynolds"], schmidt=[1.0], rkm_mode=mod.RKM_EXP4, buoyancy_type="linear",
                buoyancy_params=(1.0, 0.0), buoyancy_vector=(0.0, 1.0, 0.0), bbackground=background(c, y),
                bcs_flow_jmin=flow[c["flow_jmin"]], bcs_flow_jmax=flow[c["flow_jmax"]],
                bcs_scal_jmin=scal[c["scal_jmin"]], bcs_scal_jmax=scal[c["scal_jmax"]])


def run(sim, niter=10):
    """The loop of dns_main.f90; returns the log rows (it, time, dt, cfl, dif, dilmin, dilmax)."""
    rows = []
    rtime = 0.0
    dt, cfl, dif = sim.courant(CFL)
    rows.append(dict(it=0, time=rtime, dt=dt, cfl=cfl, dif=dif, dilmin=0.0, dilmax=0.0))
    for it in range(1, niter + 1):
        sim.runge_kutta(dt)
        rtime += dt
        dt, cfl, dif = sim.courant(CFL)
        dmin, dmax = sim.bounds_control()
        rows.append(dict(it=it, time=rtime, dt=dt, cfl=cfl, dif=dif, dilmin=dmin, dilmax=dmax))
    return rows


def matches_printed(value, printed, digits):
    """`printed` is `value` as the reference logged it with `digits` significant digits (Fortran E format): equal within
    half a unit of the last printed digit (plus 2 % of it for the rounding of the printed value itself)."""
    if printed == 0.0:
        return abs(value) < 0.5 * 10.0 ** (-digits)
    unit = 10.0 ** (math.floor(math.log10(abs(printed))) + 1 - digits)
    return abs(value - printed) <= 0.52 * unit


def compare_with_reference_log(name, rows, dil_digits=6):
    """-> list of mismatches (empty = the run reproduces examples/<Case>/dns.out.ref digit by digit)."""
    bad = []
    for a, b in zip(rows, reference_log(name)):
        assert a["it"] == b["it"]
        for key, digits in (("time", 6), ("dt", 3), ("cfl", 3), ("dif", 3), ("dilmin", dil_digits), ("dilmax", dil_digits)):
            if not matches_printed(a[key], b[key], digits):
                bad.append((a["it"], key, a[key], b[key]))
    return bad
