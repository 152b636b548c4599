#!/usr/bin/env python
"""Benchmark of the tlab hot path: RK-substep throughput (Gpts/s) on synthetic fields.

A "step" is one low-storage Runge-Kutta substep (sources + RHS_GLOBAL_INCOMPRESSIBLE_1 + update) of the
incompressible/Boussinesq equations with one scalar on the BASELINE.json configuration
"convective boundary layer 1024x512x1024, RK + FFT Poisson" (C3 of SURVEY.md section 8(d)).

  python bench.py --gpus 1 --steps K --warmup W            GPU arm (this library)
  python bench.py --impl reference ...                     reference arm: the CPU restatement of the
                                                           reference's algorithm (oracle/, "port"), all host cores

One JSON line is printed by rank 0.  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, nz)
    "c3": (1024, 512, 1024),      # BASELINE.json configs[2], single GPU
    "c4": (2048, 1024, 2048),     # BASELINE.json configs[3], 8 GPUs (2^29 points per GPU)
    "c3-half": (1024, 512, 512),
    "c2": (512, 512, 512),
    "small": (256, 128, 256),
    "tiny": (64, 48, 32),
}
ALG_BYTES_PER_PT_SUBSTEP = 995.0          # 124.4 sweeps of 8 B
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel classes from the committed
# ncu --set full captures (profiles/), scaled to the C3 grid (bytes per point x points); None if not captured
TRAFFIC_BYTES_PER_PT = {"burgers_y": 24.07, "burgers_z": 23.79}   # profiles/ncu_full_burgers_strided_r01.json (U_IN, 512^3)


def load_traffic():
    """Measured DRAM bytes per point and launch of the line-kernel classes in the RHS of this bench (they include the
    read-modify-write of hq that the launches fuse): profiles/ncu_dram_bench_r01.json, written by tools/ncu_dram_summary.py
    from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass over one substep."""
    path = os.path.join(ROOT, "profiles", "ncu_dram_bench_r02.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "ncu_dram_bench_r01.json")
    try:
        return {k: float(v) for k, v in json.load(open(path))["bytes_per_point_per_launch"].items()}
    except Exception:
        return dict(TRAFFIC_BYTES_PER_PT)
PHYS = dict(visc=1.0 / 5000.0, schmidt=[1.0], dtime=1.0e-3)


def grid_periodic(n, length=2.0 * np.pi):
    return np.arange(n) * length / n


def grid_tanh(n, length=1.0, st=0.9375, f=2.0, delta=0.0078125):
    s = np.arange(n) * length / (n - 1)
    y = s + (f - 1.0) * delta * np.logaddexp((s - st) / delta, 0.0)
    return y - y[0]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs.  The reference cannot be built here (Fortran 2008 + FFTW3, no Fortran compiler in the image), so the CPU arm is
# the C++17/OpenMP restatement of its algorithm (oracle/cpp/tlab_cpu.cpp, checked against the numpy oracle, kind "port"),
# on all host cores, on a bounded sample of the bench workload: a slab of the C3 grid with the full x and y extents
# (1024 x 512 points per plane, the same line lengths, schemes, boundary conditions and physics) and CPU_SAMPLE_NZ planes.
CPU_SAMPLE_NZ = 128


def cpu_leg(nx, ny, nz, warm, steps):
    """Runs in a process of its own (clean OpenMP environment): W + K substeps of the sample, each one timed."""
    from oracle import fdm, dns as OD, cpu_baseline as CB
    t0 = time.perf_counter()
    x, z, y = grid_periodic(nx), grid_periodic(nz), grid_tanh(ny)
    g = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
    D, Nn = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
    c = CB.CpuDns(g, visc=PHYS["visc"], schmidt=PHYS["schmidt"], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
                  buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(Nn, D, Nn), bcs_scal_jmin=(D,),
                  bcs_scal_jmax=(Nn,))
    init_s = time.perf_counter() - t0
    Z, Y, X = np.meshgrid(z, y / y[-1], x, indexing="ij", sparse=True)
    wall = np.sin(0.5 * np.pi * Y)
    rng = np.random.default_rng(20261017)
    for i in range(3):
        c.q[i][...] = 0.05 * np.sin(rng.integers(1, 5) * X + rng.integers(1, 5) * Z + rng.uniform(0, 6.28)) * wall
    c.s[0][...] = 0.5 + 0.05 * np.sin(X + 2 * Z) * wall
    kdt, _, kco = OD.rk_coefficients(OD.RKM_EXP4)
    times = []
    for i in range(warm + steps):
        sub = i % 5
        if sub == 0:
            for a in c.hq + c.hs:
                a[...] = 0.0
        t1 = time.perf_counter()
        c.substep(PHYS["dtime"] * kdt[sub], kco[sub] if sub < 4 else None)
        times.append(time.perf_counter() - t1)
    ok = bool(np.isfinite(c.q[0]).all() and np.isfinite(c.s[0]).all())
    print(json.dumps({"cpu_leg": True, "seconds": times, "init_s": init_s, "threads": c.threads, "finite": ok,
                      "grid": [nx, ny, nz]}))


def run_cpu_leg(nx, ny, nz, warm, steps):
    """-> (seconds per substep of the sample [list of the timed ones], threads, description)"""
    import subprocess
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)            # torchrun pins it to 1; the baseline is meant to use every core
    env["OMP_PROC_BIND"] = "false"
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-leg", "%d,%d,%d,%d,%d" % (nx, ny, nz, warm, steps)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{") and "cpu_leg" in l]
    if r.returncode != 0 or not lines:
        raise RuntimeError("cpu leg failed: " + r.stderr[-800:])
    d = json.loads(lines[-1])
    if not d["finite"]:
        raise RuntimeError("cpu leg produced non-finite fields")
    desc = ("%dx%dx%d slab of the %s grid (full x and y lines, %d of its z planes), %d warm-up + %d timed RK4-5 substeps, "
            "C++17/OpenMP restatement oracle/cpp/tlab_cpu.cpp, tables initialised in %.0f s (untimed)"
            % (nx, ny, nz, "1024x512x1024", nz, warm, steps, d["init_s"]))
    return d["seconds"][warm:], d["threads"], desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, nz = WORKLOADS[args.workload]
    snz = min(CPU_SAMPLE_NZ, nz)
    secs, threads, desc = run_cpu_leg(nx, ny, snz, args.warmup, args.steps)
    sec = float(np.mean(secs))
    value = nx * ny * snz / sec / 1e9
    out = {"impl": "reference", "metric": "rk_substep_throughput", "value": value, "unit": "Gpts/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "incompressible Boussinesq CBL %dx%dx%d, RK4-5 substep, 1 scalar, CompactJacobian6 + "
                                  "CompactJacobian6Hyper, tanh-stretched y" % (nx, ny, nz),
                      "sample": "each step = one substep of a %dx%dx%d slab of that grid (ms_per_step is the measured time of "
                                "such a step; value = its points / that time)" % (nx, ny, snz),
                      "note": "CPU restatement of the reference algorithm in C++/OpenMP on all host cores (the image has no "
                              "Fortran compiler and no FFTW: the reference itself cannot be built)"},
           "cpu_baseline": {"value": value, "unit": "Gpts/s", "cores": threads, "kind": "port", "sample": desc},
           "e2e": {"value": value, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
def synth_field(torch, dev, shape, x, y, z, seed, amp):
    """sum of 8 smooth modes A sin(kx x + kz z + phi) g(y), built on the device."""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    xt = torch.from_numpy(x).to(dev)
    zt = torch.from_numpy(z).to(dev)
    yt = torch.from_numpy(y / (y[-1] if y[-1] != 0 else 1.0)).to(dev)
    out = torch.zeros(shape, dtype=torch.float64, device=dev)
    gs = [torch.ones_like(yt), torch.cos(np.pi * yt), yt]
    wall = torch.sin(0.5 * np.pi * yt)
    for m in range(8):
        kx, kz = int(rng.integers(-8, 9)), int(rng.integers(-8, 9))
        a, ph = float(rng.uniform(0.1, 1.0)), float(rng.uniform(0, 2 * np.pi))
        # sin(kx x + kz z + ph) = sin(kx x + ph) cos(kz z) + cos(kx x + ph) sin(kz z)
        sx, cx = torch.sin(kx * xt + ph), torch.cos(kx * xt + ph)
        sz, cz = torch.sin(kz * zt), torch.cos(kz * zt)
        gy = (a * amp) * gs[m % 3] * wall
        out += gy[None, :, None] * (cz[:, None, None] * sx[None, None, :] + sz[:, None, None] * cx[None, None, :])
    return out


def make_sim(GD, opr, mpi, nx, ny, nz, world, rank):
    """Plans + device state of the bench configuration (CBL-like: no-slip bottom, free-slip top, linear buoyancy, 1 scalar)."""
    kmax, koff = nz, 0
    if world > 1:
        kmax, koff = mpi.slab(nz, rank, world)
    x, z, y = grid_periodic(nx), grid_periodic(nz), grid_tanh(ny)
    g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    D, Nn = GD.DNS_BCS_DIRICHLET, GD.DNS_BCS_NEUMANN
    sim = GD.Dns(g, visc=PHYS["visc"], schmidt=PHYS["schmidt"], rkm_mode=GD.RKM_EXP4, buoyancy_type="linear",
                 buoyancy_params=(1.0, 0.0), buoyancy_vector=(0.0, 1.0, 0.0),
                 bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(Nn, D, Nn), bcs_scal_jmin=(D,), bcs_scal_jmax=(Nn,),
                 kmax=(kmax if world > 1 else None))
    return sim, g, (x, y, z), kmax, koff


def fill_fields(torch, tl, L, sim, dev, grids, kmax, koff):
    import ctypes
    x, y, z = grids
    nx, ny = len(x), len(y)
    N = nx * ny * kmax
    for i, nm in enumerate(["q1", "q2", "q3", "s1"]):
        f = synth_field(torch, dev, (kmax, ny, nx), x, y, z[koff:koff + kmax], 20261017 + i, 0.05)
        if nm == "s1":
            f = 0.5 + f
        torch.cuda.synchronize()
        tl.check(L.tlab_gpu_copy(ctypes.c_void_p(sim.device_ptr(nm)), ctypes.c_void_p(f.data_ptr()), N * 8))
        del f
    torch.cuda.empty_cache()


def parity_block(torch, dist, tl, L, opr, GD, mpi, dev, world, rank):
    """Correctness record of the multi-GPU paths, taken before anything is timed: one RK step of a small split-eligible grid
    (32 x 32 x 128 P: slabs of 8 chunks, the thinnest the marching split-z kernels take) on the P ranks against the single-domain oracle on rank 0 -- once with the default path
    (split-z operators over peer memory + kx-split Poisson stage) and once with splitz = 0 (K-transposes for every z operator).
    The oracle is the checker here, never the thing measured."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import smooth_field, rel_l2
    nx, ny, nz = 32, 32, 128 * world
    out = {"grid": [nx, ny, nz], "paths": {}}
    ref = None
    for label, tune in (("default", None), ("splitz=0", ("splitz", 0))):
        if tune:
            tl.check(L.tlab_gpu_set_tuning(tune[0].encode(), tune[1]))
        cnt0 = {}
        for key in ("p2p_exchanges", "nccl_exchanges", "splitz_ops", "splitz_march_ops"):
            c = ctypes.c_longlong()
            tl.check(L.tlab_gpu_get_counter(key.encode(), ctypes.byref(c)))
            cnt0[key] = c.value
        sim, g, (x, y, z), kmax, koff = make_sim(GD, opr, mpi, nx, ny, nz, world, rank)
        wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
        full = [0.5 * smooth_field((nz, ny, nx), (x, y, z), seed=31 + i) * wall for i in range(3)]
        full.append(0.5 + 0.1 * smooth_field((nz, ny, nx), (x, y, z), seed=40) * wall)
        for nm, f in zip(["q1", "q2", "q3", "s1"], full):
            sim.set(nm, f[koff:koff + kmax])
        sim.runge_kutta(PHYS["dtime"])
        mine = [sim.get(nm) for nm in ["q1", "q2", "q3", "s1"]]
        gathered = []
        for f in mine:
            t = torch.from_numpy(f).to(dev)
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            gathered.append(torch.cat(parts, dim=0).cpu().numpy())
        rec = {}
        for key in cnt0:
            c = ctypes.c_longlong()
            tl.check(L.tlab_gpu_get_counter(key.encode(), ctypes.byref(c)))
            rec[key] = c.value - cnt0[key]
        if rank == 0:
            if ref is None:
                from oracle import fdm, dns as OD
                go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
                D, Nn = OD.DNS_BCS_DIRICHLET, OD.DNS_BCS_NEUMANN
                o = OD.Dns(go, visc=PHYS["visc"], schmidt=PHYS["schmidt"], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
                           buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(Nn, D, Nn),
                           bcs_scal_jmin=(D,), bcs_scal_jmax=(Nn,))
                for i in range(3):
                    o.q[i][...] = full[i]
                o.s[0][...] = full[3]
                o.runge_kutta(PHYS["dtime"])
                ref = o.q + o.s
            rec["max_rel_l2"] = float(max(rel_l2(a, b) for a, b in zip(gathered, ref)))
        out["paths"][label] = rec
        sim.close()
        if tune:
            tl.check(L.tlab_gpu_set_tuning(tune[0].encode(), 1))
    if rank == 0:
        out["max_rel_l2"] = max(r["max_rel_l2"] for r in out["paths"].values())
        out["tolerance"] = 1e-11
        out["ok"] = bool(out["max_rel_l2"] <= 1e-11)
    return out


def time_substeps(torch, dist, tl, L, sim, stream, dev, world, steps, warmup, dtime=None, nstage=5):
    """W untimed + K timed substeps, CUDA events on the library stream, max over ranks.  Returns ms per substep."""
    dtime = PHYS["dtime"] if dtime is None else dtime
    tl.check(L.tlab_gpu_set_async(1))
    for i in range(warmup):
        sim.runge_kutta_stage(dtime, i % nstage)
    tl.check(L.tlab_gpu_synchronize())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record(stream)
    for i in range(warmup, warmup + steps):
        sim.runge_kutta_stage(dtime, i % nstage)
    e1.record(stream)
    tl.check(L.tlab_gpu_synchronize())
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    tl.check(L.tlab_gpu_set_async(0))
    return ms / steps


def extra_records(torch, tl, L, opr, dev, stream, peak):
    """BASELINE.json configs[1] and configs[4], driver-run: OPR_Partial X/Y/Z (P1, P2, P2_P1) on 512^3 with the non-uniform y
    of the bench, and OPR_Poisson alone on 256^3 ... 1024^3 (120 B/pt stage-streaming model, the y stage against 24 B/pt)."""
    import ctypes
    out = {"opr_partial_512": [], "poisson_sweep": []}

    def timeit(fn, iters=10, warm=3):
        tl.check(L.tlab_gpu_set_async(1))
        for _ in range(warm):
            fn()
        tl.check(L.tlab_gpu_synchronize())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        tl.check(L.tlab_gpu_synchronize())
        torch.cuda.synchronize()
        tl.check(L.tlab_gpu_set_async(0))
        return e0.elapsed_time(e1) / iters

    nx = ny = nz = 512
    N = nx * ny * nz
    x, z, y = grid_periodic(nx), grid_periodic(nz), grid_tanh(ny)
    g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    u = torch.randn(N, dtype=torch.float64, device=dev)
    r1, r2 = torch.empty_like(u), torch.empty_like(u)
    bcs = [[0, 0], [0, 0]]
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    for d, nm in enumerate("XYZ"):
        for tname, typ, nb in (("P1", opr.OPR_P1, 16), ("P2", opr.OPR_P2, 16), ("P2_P1", opr.OPR_P2_P1, 24)):
            ms = timeit(lambda: P[d](typ, nx, ny, nz, bcs, g[d], u, r1, r2 if typ == opr.OPR_P2_P1 else None))
            gbs = nb * N / (ms * 1e-3) / 1e9
            out["opr_partial_512"].append({"op": "OPR_Partial_%s %s" % (nm, tname), "ms": ms, "GBs": gbs,
                                           "frac_measured_peak": gbs / peak, "frac_8TBs": gbs / 8000.0})
    del u, r1, r2, g
    torch.cuda.empty_cache()
    for shape in ((256, 256, 256), (512, 512, 512), (1024, 512, 1024), (1024, 1024, 1024)):
        nx, ny, nz = shape
        N = nx * ny * nz
        try:
            x, z, y = grid_periodic(nx), grid_periodic(nz), grid_tanh(ny)
            g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
            opr.OPR_Elliptic_Initialize(g)
            p = torch.randn(N, dtype=torch.float64, device=dev)
            t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=dev)
            t2 = torch.zeros_like(t1)
            dp = torch.empty_like(p)
            hb = torch.zeros(nx * nz, dtype=torch.float64, device=dev)
            ht = torch.zeros_like(hb)
            tl.check(L.tlab_gpu_profile(1))
            ms = timeit(lambda: opr.OPR_Poisson(nx, ny, nz, 3, p, t1, t2, hb, ht, dp), iters=5, warm=2)
            mc, cc = (ctypes.c_double * 16)(), (ctypes.c_int * 16)()
            tl.check(L.tlab_gpu_profile_report(mc, cc, 16))
            tl.check(L.tlab_gpu_profile(0))
            yms = mc[8] / cc[8] if cc[8] else None
            gbs = 120.0 * N / (ms * 1e-3) / 1e9
            out["poisson_sweep"].append({"grid": list(shape), "ms": ms, "GBs_120B_model": gbs, "frac_measured_peak": gbs / peak,
                                         "frac_8TBs": gbs / 8000.0, "GBs_24B_floor": 24.0 * N / (ms * 1e-3) / 1e9,
                                         "y_stage_ms": yms,
                                         "y_stage_frac_measured_peak": (24.0 * N / (yms * 1e-3) / 1e9 / peak) if yms else None})
            del p, t1, t2, dp, hb, ht, g
        except Exception as ex:      # e.g. out of memory on a smaller part
            out["poisson_sweep"].append({"grid": list(shape), "error": str(ex)[:200]})
        torch.cuda.empty_cache()
    # release the process-wide solver of the stand-alone API (re-initialise on a tiny grid)
    try:
        x, z, y = grid_periodic(16), grid_periodic(16), grid_tanh(16)
        opr.OPR_Elliptic_Initialize([opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"),
                                     opr.FdmPlan(z, True, True, name="z")])
    except Exception:
        pass
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tlab_b200 import lib as tl, opr, dns as GD
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = tl.load()
    tl.check(L.tlab_gpu_init(local_rank))
    for kv in [t for t in args.tune.split(",") if t]:
        k, v = kv.split("=")
        tl.check(L.tlab_gpu_set_tuning(k.encode(), int(v)))

    nx, ny, nz = WORKLOADS[args.workload]
    if args.nx:
        nx, ny, nz = args.nx, args.ny, args.nz
    from tlab_b200 import mpi
    parity = None
    if world > 1:
        mpi.init_from_torch_distributed()
        if not args.no_parity:
            parity = parity_block(torch, dist, tl, L, opr, GD, mpi, dev, world, rank)
    sim, g, (x, y, z), kmax, koff = make_sim(GD, opr, mpi, nx, ny, nz, world, rank)
    Nglobal = nx * ny * nz
    N = nx * ny * kmax                       # points of this rank's slab
    fill_fields(torch, tl, L, sim, dev, (x, y, z), kmax, koff)

    sp = ctypes.c_void_p()
    tl.check(L.tlab_gpu_stream(ctypes.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value, device=dev)
    dtime = PHYS["dtime"]
    nstage = 5

    def substeps(k0, k):
        for i in range(k0, k0 + k):
            sim.runge_kutta_stage(dtime, i % nstage)

    tl.check(L.tlab_gpu_set_async(1))
    substeps(0, args.warmup)
    tl.check(L.tlab_gpu_synchronize())
    launches0 = sim.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    tl.check(L.tlab_gpu_profile(1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record(stream)
    substeps(args.warmup, args.steps)
    e1.record(stream)
    tl.check(L.tlab_gpu_synchronize())
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    clocks = sampler.stop()
    launches = sim.launch_count() - launches0
    ms_cls = (ctypes.c_double * 11)()
    cnt_cls = (ctypes.c_int * 11)()
    tl.check(L.tlab_gpu_profile_report(ms_cls, cnt_cls, 11))
    tl.check(L.tlab_gpu_profile(0))
    cls_names = ["burgers_x", "burgers_y", "burgers_z", "partial_x", "partial_y", "partial_z", "neumann_bcs", "fft",
                 "poisson_y", "elementwise", "transpose"]
    breakdown = {n_: {"ms_per_step": ms_cls[i] / args.steps, "launches_per_step": cnt_cls[i] / args.steps}
                 for i, n_ in enumerate(cls_names) if cnt_cls[i] > 0}
    ms_per_step = ms_total / args.steps
    value = Nglobal / (ms_per_step * 1e-3) / 1e9

    # roofline: every kernel class against the algorithmic bytes of SURVEY.md 8(d) (compulsory traffic at the reference's
    # operator surface).  Per launch: OPR_Burgers SELF 16 / U_IN 24 B/pt -> mean of the 1 + 3 launches of a direction 22;
    # OPR_Partial P1 16; the y stage of OPR_Poisson 24 (1r + 2w); the five FFT stages together 96 (6r + 6w).  The line-kernel
    # launches of the RHS also perform the reference's separate accumulation / pressure-forcing sweeps (`hq = hq + tmp`,
    # `tmp = hq + q/dte`), which fused cost one more operand: `fused_bytes` (30 / 28 / 24 B/pt) is what the launch must move.
    peak, peak_kind = load_peaks()
    contract = {"burgers_x": 22.0, "burgers_y": 22.0, "burgers_z": 22.0, "partial_x": 16.0, "partial_y": 16.0, "partial_z": 16.0,
                "poisson_y": 24.0}
    fused = {"burgers_x": 30.0, "burgers_y": 30.0, "burgers_z": 30.0, "partial_x": 28.0, "partial_y": 24.0, "partial_z": 28.0,
             "poisson_y": 24.0}
    traffic = load_traffic()
    per_class = {}
    for k in breakdown:
        i = cls_names.index(k)
        if k in contract:
            avg = ms_cls[i] / cnt_cls[i]
            per_class[k] = {"avg_launch_ms": avg, "launches_per_step": cnt_cls[i] / args.steps,
                            "achieved": contract[k] * N / (avg * 1e-3) / 1e9, "frac": contract[k] * N / (avg * 1e-3) / 1e9 / peak,
                            "achieved_fused_bytes": fused[k] * N / (avg * 1e-3) / 1e9,
                            "frac_fused_bytes": fused[k] * N / (avg * 1e-3) / 1e9 / peak}
        elif k == "fft":
            t = ms_cls[i] / args.steps          # all FFT stages of a substep together (cuFFT's own kernels)
            per_class[k] = {"ms_per_step": t, "achieved": 96.0 * N / (t * 1e-3) / 1e9, "frac": 96.0 * N / (t * 1e-3) / 1e9 / peak,
                            "note": "cuFFT (library), 12 sweeps of 8 B/pt"}
    own = [k for k in per_class if k != "fft"]
    dom = max(own, key=lambda k: breakdown[k]["ms_per_step"])            # largest share of the substep
    worst = min(own, key=lambda k: per_class[k]["frac"])                  # furthest below its roofline
    d = per_class[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": d["achieved"], "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": d["frac"], "frac_of_8TBs": d["achieved"] / 8000.0,
                "traffic": (traffic[dom] * N if dom in traffic else None),
                "traffic_source": "profiles/ncu_dram_bench_r02.json (ncu dram__bytes_read+write of this command, per launch); "
                                  "not measurable inside an unprofiled run",
                "algorithmic_bytes_per_launch": contract[dom] * N, "avg_launch_ms": d["avg_launch_ms"],
                "algorithmic_bytes_note": "SURVEY 8(d) operator surface, %g B/pt; the launch also fuses the reference's separate "
                                          "accumulation sweep (%g B/pt in total): achieved_fused_bytes" % (contract[dom], fused[dom]),
                "achieved_fused_bytes": d["achieved_fused_bytes"], "frac_fused_bytes": d["frac_fused_bytes"],
                "worst_kernel": worst, "worst_frac": per_class[worst]["frac"],
                "per_class": per_class,
                "substep": {"algorithmic_bytes_per_gpu": ALG_BYTES_PER_PT_SUBSTEP * N,
                            "achieved_per_gpu": ALG_BYTES_PER_PT_SUBSTEP * N / (ms_per_step * 1e-3) / 1e9,
                            "frac": ALG_BYTES_PER_PT_SUBSTEP * N / (ms_per_step * 1e-3) / 1e9 / peak,
                            "frac_of_8TBs": ALG_BYTES_PER_PT_SUBSTEP * N / (ms_per_step * 1e-3) / 1e9 / 8000.0}}

    # end to end: one full RK step (5 substeps) from and to pinned host buffers through the C ABI
    tl.check(L.tlab_gpu_set_async(0))
    e2e = None
    try:
        qh = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        sh = torch.empty(N, dtype=torch.float64).pin_memory()
        for i in range(3):
            tl.check(L.tlab_gpu_download(ctypes.c_void_p(qh.data_ptr() + i * N * 8), ctypes.c_void_p(sim.device_ptr("q%d" % (i + 1))), N * 8))
        tl.check(L.tlab_gpu_download(ctypes.c_void_p(sh.data_ptr()), ctypes.c_void_p(sim.device_ptr("s1")), N * 8))
        sim.runge_kutta_host(dtime, qh.data_ptr(), sh.data_ptr())      # warm-up
        reps = 2
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            sim.runge_kutta_host(dtime, qh.data_ptr(), sh.data_ptr())
        a1.record(stream)
        torch.cuda.synchronize()
        ms_e2e = a0.elapsed_time(a1) / (reps * nstage)
        if world > 1:
            tmax = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms_e2e = float(tmax.item())
        e2e = {"value": Nglobal / (ms_e2e * 1e-3) / 1e9, "unit": "Gpts/s", "h2d_bytes_per_step": 4 * Nglobal * 8 / nstage,
               "d2h_bytes_per_step": 4 * Nglobal * 8 / nstage, "ms_per_step": ms_e2e,
               "call": "tlab_time_rungekutta_host: 4 fields up, 5 substeps, 4 fields down"}
        del qh, sh
    except Exception as ex:           # e.g. not enough pinned host memory
        e2e = {"value": None, "unit": "Gpts/s", "error": str(ex)}
        if world > 1:
            raise

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            snz = min(CPU_SAMPLE_NZ, nz)
            secs, cores, desc = run_cpu_leg(nx, ny, snz, 1, 5)
            sec = float(np.mean(secs))
            cpu = {"value": nx * ny * snz / sec / 1e9, "unit": "Gpts/s", "cores": cores, "kind": "port", "sample": desc,
                   "seconds_per_sample_substep": sec}
        except Exception as ex:
            cpu = {"value": None, "unit": "Gpts/s", "cores": None, "kind": "port", "sample": "failed: %s" % str(ex)[:300]}

    zc = ctypes.c_longlong(0)
    tl.check(L.tlab_gpu_get_counter(b"splitz_ops", ctypes.byref(zc)))
    z_path = "whole lines on one GPU" if world == 1 else (
        "on the slabs: halo planes + chunk ends exchanged with the neighbours over peer memory (splitz)" if zc.value > 0
        else "K-transposes to z pencils (all-to-all)")
    sim.close()
    del sim
    torch.cuda.empty_cache()

    # BASELINE configs[3]: 2048 x 1024 x 2048 on 8 GPUs (2^29 points per GPU), timed in the same run
    c4 = None
    if world == 8 and args.workload == "c3" and not args.no_c4:
        try:
            cx, cy, cz = WORKLOADS["c4"]
            sim4, g4, grids4, kmax4, koff4 = make_sim(GD, opr, mpi, cx, cy, cz, world, rank)
            fill_fields(torch, tl, L, sim4, dev, grids4, kmax4, koff4)
            k4 = min(args.steps, 5)
            tl.check(L.tlab_gpu_profile(1))
            ms4 = time_substeps(torch, dist, tl, L, sim4, stream, dev, world, k4, 3)
            m4, n4 = (ctypes.c_double * 11)(), (ctypes.c_int * 11)()
            tl.check(L.tlab_gpu_profile_report(m4, n4, 11))
            tl.check(L.tlab_gpu_profile(0))
            c4 = {"breakdown_ms": {n_: m4[i] / (k4 + 3) for i, n_ in enumerate(cls_names) if n4[i] > 0},
                  "workload": "2048x1024x2048 (BASELINE configs[3]), z-slabs x8", "ms_per_step": ms4, "steps": k4, "warmup": 3,
                  "value": cx * cy * cz / (ms4 * 1e-3) / 1e9, "unit": "Gpts/s", "points_per_gpu": cx * cy * kmax4,
                  "roofline_substep_frac": ALG_BYTES_PER_PT_SUBSTEP * cx * cy * kmax4 / (ms4 * 1e-3) / 1e9 / peak,
                  "note": "weak-scaled against C3 on one GPU (same 2^29 points per GPU): efficiency = value / (8 x the N=1 value of "
                          "this bench); the driver computes it from its own N=1 line"}
            sim4.close()
            del sim4
            torch.cuda.empty_cache()
        except Exception as ex:
            c4 = {"error": str(ex)[:300]}

    extra = None
    if world == 1 and args.workload == "c3" and not args.no_extra:
        extra = extra_records(torch, tl, L, opr, dev, stream, peak)

    out = {"metric": "rk_substep_throughput", "value": value, "unit": "Gpts/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "incompressible Boussinesq CBL %dx%dx%d, RK4-5 substep, 1 scalar, CompactJacobian6 + "
                                  "CompactJacobian6Hyper, tanh-stretched y" % (nx, ny, nz),
                      "l2": "working set per substep >> 126 MB L2 (each field %.2f GB)" % (N * 8 / 1e9),
                      "decomposition": "z-slabs x%d" % world,
                      "z_operators": z_path},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "breakdown_ms": breakdown}
    if parity is not None:
        out["parity"] = parity
    if c4 is not None:
        out["c4"] = c4
    if extra is not None:
        out["extra"] = extra
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        mpi.finalize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-leg", default="", help=argparse.SUPPRESS)     # internal: nx,ny,nz,warm,steps of the CPU sample
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the oracle parity record taken before the timed region")
    ap.add_argument("--no-c4", action="store_true", help="N = 8: skip the additional 2048x1024x2048 timing")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the OPR_Partial 512^3 and OPR_Poisson sweep records")
    ap.add_argument("--tune", default="", help="library tuning knobs, e.g. fuse=0,pf_next=1 (tlab_gpu_set_tuning)")
    args = ap.parse_args()
    if args.cpu_leg:
        cpu_leg(*[int(v) for v in args.cpu_leg.split(",")])
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
