#!/usr/bin/env python
"""Generate golden vectors from the reference's own numpy restatement of its compact schemes.

Runs ONLY in the authoring container (needs /root/reference): imports
/root/reference/scripts/python/compact_lib.py (the reference's C1N6 first-derivative schemes: periodic and
biased tridiagonal 3-5-6-5-3, Jacobian formulation on non-uniform grids) and stores inputs and outputs in
tests/golden/compact_lib_c1n6.npz.  The tests compare the oracle (and, on the GPU, the CUDA path) against them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/scripts/python")
import compact_lib as cl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # periodic, uniform (spacing exactly representable so that compact_lib accepts the grid as uniform)
    n = 64
    x = np.arange(n) / 16.0
    u = np.stack([np.sin(2 * np.pi * 3 * x / 4.0), np.cos(2 * np.pi * 5 * x / 4.0) + 0.3 * x * 0,
                  rng.standard_normal(n)], axis=1)
    out["per_x"], out["per_u"] = x, u
    out["per_du"] = cl.compactder(u.copy(), x, periodic=True)
    # non-periodic, non-uniform (tanh-like stretching), biased boundary closures
    n = 97
    s = np.linspace(0.0, 1.0, n)
    y = s + 0.25 * np.sin(np.pi * s) / np.pi
    v = np.stack([np.exp(-((y - 0.5) / 0.2) ** 2), np.sin(7 * y) * y, rng.standard_normal(n)], axis=1)
    out["nonuni_x"], out["nonuni_u"] = y, v
    out["nonuni_du"] = cl.compactder(v.copy(), y, periodic=False)
    # non-periodic, uniform
    n = 40
    xu = np.arange(n) / 32.0
    w = np.stack([xu ** 3 - xu, np.cos(9 * xu)], axis=1)
    out["uni_x"], out["uni_u"] = xu, w
    out["uni_du"] = cl.compactder(w.copy(), xu, periodic=False)
    np.savez_compressed(os.path.join(HERE, "compact_lib_c1n6.npz"), **out)
    print("written", os.path.join(HERE, "compact_lib_c1n6.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
