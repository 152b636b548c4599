"""examples/Case10 of the reference (2-D convective boundary layer, 512 x 257, Boussinesq, tanh-stretched y, RK4-5,
CFL-controlled time step) restated from its tlab.ini so that `examples/Case10/dns.out.ref` -- the reference build's own
log of ten iterations: time, dt, CFL number, diffusion number, min/max of the dilatation -- becomes a golden vector.

  grid     tools/initialize/grid/grid_main.f90:55-110, grid_local.f90:41-66 (BLD_TANH); [IniGridOx/Oy] of tlab.ini
  scalar   tools/initialize/scal/scal_local.f90:244-340 (SCAL_FLUCTUATION_PLANE, DeltaDiscrete: the thickness of the Erf
           profile is perturbed by 0.001 cos(3 * 2 pi x / scale_x)), utils/profiles.f90:169-232 (Erf), utils/discrete.f90:43-85
  physics  Reynolds 2000, Schmidt 1, Froude 1, linear buoyancy b = s along y with the background profile subtracted
           (physics/tlab_background.f90:216-223), no-slip bottom, free-slip top, scalar Dirichlet / Neumann, clipping to [0, 1]
  loop     tools/dns/dns_main.f90: TIME_RUNGEKUTTA, TIME_COURANT, DNS_BOUNDS_CONTROL, one log line per iteration
The reference's tlab.ini (copied values only): Imax 512, Jmax 257, scales 2 and 1, tanh (0.9375, 2, 0.0078125),
ThickScalar1 0.02, DeltaScalar1 2, MeanScalar1 1, 2DAmpl 0,0,0.001, TimeCFL 1.2."""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NX, NY = 512, 257
SCALE_X = 2.0
CFL = 1.2
VISC = 1.0 / 2000.0


def grids():
    x = np.arange(NX + 1) * (SCALE_X / NX)            # 513 points on [0, 2], last one dropped (periodic)
    s = np.arange(NY) * (1.0 / (NY - 1))
    st, f, delta = 0.9375, 2.0, 0.0078125
    work = (f - 1.0) * delta * np.log(np.exp((s - st) / delta) + 1.0)
    y = s + (work - work[0])
    return x[:NX].copy(), y, np.zeros(1)


def erf_profile(y, thick, mean=1.0, delta=2.0, ymean=0.0):
    xi = (y - ymean) / thick
    return mean + delta * 0.5 * np.vectorize(math.erf)(-0.5 * xi)


def initial_scalar(x, y):
    disp = 0.001 * np.cos(3.0 * (2.0 * np.pi / SCALE_X) * x)       # modes 1, 2 have zero amplitude
    thick = 0.02 + disp
    return erf_profile(y[:, None], thick[None, :])[None]            # (1, ny, nx)


def background(y):
    return erf_profile(y, 0.02)


def reference_log():
    rows = []
    for line in open(os.path.join(HERE, "golden", "case10_dns.out.ref")):
        if line.startswith("#"):
            continue
        t = line.split()
        rows.append(dict(it=int(t[1]), time=float(t[2]), dt=float(t[3]), cfl=float(t[4]), dif=float(t[5]), visc=float(t[6]),
                         dilmin=float(t[7]), dilmax=float(t[8])))
    return rows


def dns_kwargs(mod, y):
    D, N = mod.DNS_BCS_DIRICHLET, mod.DNS_BCS_NEUMANN
    return dict(visc=VISC, schmidt=[1.0], rkm_mode=mod.RKM_EXP4, buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
                buoyancy_vector=(0.0, 1.0, 0.0), bbackground=background(y),
                bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N), bcs_scal_jmin=(D,), bcs_scal_jmax=(N,))


def run(sim, niter=10):
    """The loop of dns_main.f90; returns the log rows (it, time, dt, cfl, dif, dilmin, dilmax)."""
    rows = []
    rtime = 0.0
    dt, cfl, dif = sim.courant(CFL)
    dmin, dmax = 0.0, 0.0
    rows.append(dict(it=0, time=rtime, dt=dt, cfl=cfl, dif=dif, dilmin=dmin, dilmax=dmax))
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


def compare_with_reference_log(rows):
    """-> list of mismatches (empty = the run reproduces examples/Case10/dns.out.ref digit by digit)."""
    bad = []
    for a, b in zip(rows, reference_log()):
        assert a["it"] == b["it"]
        for key, digits in (("time", 6), ("dt", 3), ("cfl", 3), ("dif", 3), ("dilmin", 6), ("dilmax", 6)):
            if not matches_printed(a[key], b[key], digits):
                bad.append((a["it"], key, a[key], b[key]))
    return bad
