#!/usr/bin/env python
"""Small RK steps that touch every kernel family, for compute-sanitizer (tools/gpu_sanitize.sh): default path (whole-line
kernels in x and y, marching panels in z, team-per-mode Poisson), marching panels in y as well, split-z kernels over two
virtual slabs, general kernels, TMA kernels.  Prints SANITIZE_STEP_OK when every variant finished."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from common import grid_periodic, grid_tanh, smooth_field
    from tlab_b200 import opr, dns as GD, lib as tl
    L = tl.load()
    variants = [("default", {}), ("march=2", {"march": 2}), ("march_red", {"march": 2, "march_red": 1}),
                ("split_emulate=2", {"split_emulate": 2}), ("fast=0", {"fast": 0}), ("tma=1", {"tma": 1, "march": 0}),
                ("lazy_scale=0", {"lazy_scale": 0}), ("fuse=1", {"fuse": 1}),
                # lines long enough for the circulant form in x and z (8 chunks) and for the peeled marching steps in y (3 rounds)
                ("circ_peel", {"shape": (128, 192, 128)}), ("split_circ", {"shape": (128, 64, 256), "split_emulate": 2})]
    only = sys.argv[1:] or None
    D, N = GD.DNS_BCS_DIRICHLET, GD.DNS_BCS_NEUMANN
    for name, tune in variants:
        if (only and name not in only) or (not only and "shape" in tune):
            continue
        tune = dict(tune)
        nx, ny, nz = tune.pop("shape", (32, 128, 192))
        x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
        for k, v in tune.items():
            tl.check(L.tlab_gpu_set_tuning(k.encode(), v))
        g = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
        sim = GD.Dns(g, visc=1.0 / 5000.0, schmidt=[1.0], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
                     buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N), bcs_scal_jmin=(D,),
                     bcs_scal_jmax=(N,))
        wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
        for i, nm in enumerate(["q1", "q2", "q3"]):
            sim.set(nm, 0.05 * smooth_field((nz, ny, nx), (x, y, z), seed=31 + i) * wall)
        sim.set("s1", 0.5 + 0.02 * smooth_field((nz, ny, nx), (x, y, z), seed=40) * wall)
        sim.runge_kutta(1e-3)
        q1 = sim.get("q1")
        assert np.isfinite(q1).all()
        print("variant %-18s ok  |q1| = %.6e" % (name, np.abs(q1).max()), flush=True)
        sim.close()
        for k in tune:
            tl.check(L.tlab_gpu_set_tuning(k.encode(), {"march": 1, "lazy_scale": 1}.get(k, 0)))
    print("SANITIZE_STEP_OK")


if __name__ == "__main__":
    main()
