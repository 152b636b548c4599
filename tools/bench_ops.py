#!/usr/bin/env python
"""Operator-level bandwidth table (BASELINE.json configs[1]: compact FD d/dx, d2/dx2 on 512^3, non-uniform y).

Times OPR_Partial / OPR_Burgers / OPR_Poisson calls through the C ABI with CUDA events on the library stream
and prints achieved GB/s against the algorithmic bytes of SURVEY.md 8(d) (P1/P2 16 B/pt, P2_P1 24, Burgers
SELF 16 / U_IN 24, Poisson 120 with 24 alongside)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import grid_periodic, grid_tanh, load_peaks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,512,512")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--lines-x", default="0")
    ap.add_argument("--lines-yz", default="0")
    ap.add_argument("--poisson", action="store_true")
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="", help="regular expression selecting the operator labels to run")
    ap.add_argument("--fast", default="1", help="comma list of 0/1: general kernels (lines.cu) / fast kernels (lines2.cu)")
    ap.add_argument("--pf-dist", default="-1", help="comma list of L2 prefetch distances of the fast kernels (-1 auto, 0 off)")
    ap.add_argument("--persist", default="1", help="comma list of 0/1: persistent cp.async variant of the fast y/z kernels")
    ap.add_argument("--tune", default="", help="further tuning keys, e.g. poisson_split=1,tma=1")
    ap.add_argument("--no-lines", action="store_true", help="skip the line operators (Poisson only)")
    ap.add_argument("--prefetch", type=int, default=-1, help="1/0: force the persistent cp.async variant of the y/z kernels")
    args = ap.parse_args()
    import torch
    from tlab_b200 import lib as tl, opr
    L = tl.load()
    if args.prefetch >= 0:
        tl.check(L.tlab_gpu_set_tuning(b"prefetch", args.prefetch))
    for kv in [t for t in args.tune.split(",") if t]:
        k, v = kv.split("=")
        tl.check(L.tlab_gpu_set_tuning(k.encode(), int(v)))
    dev = torch.device("cuda:0")
    nx, ny, nz = [int(v) for v in args.shape.split(",")]
    N = nx * ny * nz
    x, z, y = grid_periodic(nx), grid_periodic(nz), grid_tanh(ny)
    g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    opr.OPR_Burgers_Initialize(g, 1.0 / 5000.0, [1.0])
    u = torch.randn(N, dtype=torch.float64, device=dev)
    v = torch.randn(N, dtype=torch.float64, device=dev)
    r1 = torch.empty_like(u)
    r2 = torch.empty_like(u)
    sp = ctypes.c_void_p()
    tl.check(L.tlab_gpu_stream(ctypes.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value, device=dev)
    peak, kind = load_peaks()
    bcs = [[0, 0], [0, 0]]
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z]
    B = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z]
    rows = []

    import re

    def timeit(fn, nbytes, label):
        if args.only and not re.search(args.only, label):
            return
        tl.check(L.tlab_gpu_set_async(1))
        for _ in range(args.warmup):
            fn()
        tl.check(L.tlab_gpu_synchronize())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        tl.check(L.tlab_gpu_synchronize())
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": label, "ms": ms, "GBs": gbs, "frac_measured_peak": gbs / peak, "frac_8TBs": gbs / 8000.0})
        print("%-34s %9.3f ms %9.1f GB/s  %5.1f%% of %s peak %.0f  (%4.1f%% of 8 TB/s)" %
              (label, ms, gbs, 100 * gbs / peak, kind, peak, 100 * gbs / 8000.0), flush=True)
        tl.check(L.tlab_gpu_set_async(0))

    combos = [(lx, lyz, fast, pf, ps) for lx in [int(s) for s in args.lines_x.split(",")]
              for lyz in [int(s) for s in args.lines_yz.split(",")] for fast in [int(s) for s in args.fast.split(",")]
              for pf in ([int(s) for s in args.pf_dist.split(",")] if fast else [0])
              for ps in ([int(s) for s in args.persist.split(",")] if fast else [0])]
    for lx, lyz, fast, pf, ps in combos:
        if not args.no_lines:
            tl.check(L.tlab_gpu_set_tuning(b"lines_x", lx))
            tl.check(L.tlab_gpu_set_tuning(b"lines_yz", lyz))
            tl.check(L.tlab_gpu_set_tuning(b"fast", fast))
            tl.check(L.tlab_gpu_set_tuning(b"pf_dist", pf))
            tl.check(L.tlab_gpu_set_tuning(b"persist", ps))
            tag = " [fast=%d pf=%d ps=%d Lx=%d Lyz=%d]" % (fast, pf, ps, lx, lyz)
            for d, nm in enumerate("xyz"):
                timeit(lambda: P[d](opr.OPR_P1, nx, ny, nz, bcs, g[d], u, r1), 16 * N, "OPR_Partial_%s P1" % nm.upper() + tag)
                timeit(lambda: P[d](opr.OPR_P2, nx, ny, nz, bcs, g[d], u, r1), 16 * N, "OPR_Partial_%s P2" % nm.upper() + tag)
                timeit(lambda: P[d](opr.OPR_P2_P1, nx, ny, nz, bcs, g[d], u, r1, r2), 24 * N, "OPR_Partial_%s P2_P1" % nm.upper() + tag)
                timeit(lambda: B[d](opr.OPR_B_SELF, 0, nx, ny, nz, bcs, u, u, r1), 16 * N, "OPR_Burgers_%s SELF" % nm.upper() + tag)
                timeit(lambda: B[d](opr.OPR_B_U_IN, 1, nx, ny, nz, bcs, u, v, r1), 24 * N, "OPR_Burgers_%s U_IN" % nm.upper() + tag)
    if args.poisson:
        t1 = torch.zeros((nx + 2) * ny * nz, dtype=torch.float64, device=dev)
        t2 = torch.zeros_like(t1)
        hb = torch.zeros(nx * nz, dtype=torch.float64, device=dev)
        ht = torch.zeros_like(hb)
        p = u.clone()
        for il in [int(v) for v in os.environ.get("POISSON_IL", "1,0").split(",")]:
            tl.check(L.tlab_gpu_set_tuning(b"poisson_il", il))
            opr.OPR_Elliptic_Initialize(g)
            for minb, split in [tuple(int(c) for c in v.split(":")) for v in os.environ.get("POISSON_CFG", "3:0,3:1,4:0,4:1").split(",")]:
                tl.check(L.tlab_gpu_set_tuning(b"poisson_minb", minb))
                tl.check(L.tlab_gpu_set_tuning(b"poisson_split", split))
                tl.check(L.tlab_gpu_profile(1))
                timeit(lambda: opr.OPR_Poisson(nx, ny, nz, 3, p, t1, t2, hb, ht, r1), 120 * N,
                       "OPR_Poisson (120 B/pt model) il=%d minb=%d split=%d" % (il, minb, split))
                ms = (ctypes.c_double * 16)()
                cn = (ctypes.c_int * 16)()
                tl.check(L.tlab_gpu_profile_report(ms, cn, 16))
                tl.check(L.tlab_gpu_profile(0))
                if rows and cn[8]:
                    rows[-1]["poisson_y_ms"] = ms[8] / cn[8]
                    rows[-1]["GBs_at_24B_floor"] = rows[-1]["GBs"] * 24.0 / 120.0
                    print("      y solves %.3f ms per call" % (ms[8] / cn[8]), flush=True)
        tl.check(L.tlab_gpu_set_tuning(b"poisson_il", 0))
        tl.check(L.tlab_gpu_set_tuning(b"poisson_minb", 3))
        tl.check(L.tlab_gpu_set_tuning(b"poisson_split", -1))
    if args.json:
        json.dump({"shape": [nx, ny, nz], "peak_GBs": peak, "peak_kind": kind, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
