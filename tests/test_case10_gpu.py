"""examples/Case10, Case06, Case07 `dns.out.ref` reproduced by the CUDA path: ten CFL-controlled RK4-5 steps from the
restated initial conditions (tests/tlab_cases.py), time step, CFL / diffusion numbers and dilatation bounds computed on
the device; the logged digits of the reference build must come out (see test_case10_cpu.py for the digit counts).
Case10 has 257 points in y (general line kernels), Case06 / 07 have 256 (fast kernels on the stretched grid)."""
import numpy as np
import pytest

import tlab_cases as tc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dil_digits", [("case10", 6), ("case07", 5), ("case06", 5)])
def test_gpu_reproduces_reference_log(cuda, name, dil_digits):
    from tlab_b200 import opr, dns as GD
    c = tc.CASES[name]
    x, y, z = tc.grids(c)
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    g = GD.Dns(gg, **tc.dns_kwargs(c, GD, y))
    g.set("s1", tc.initial_scalar(c, x, y))
    rows = tc.run(g, 10)
    bad = tc.compare_with_reference_log(name, rows, dil_digits)
    s = g.get("s1")
    g.close()
    assert bad == []
    assert s.min() >= 0.0 and s.max() <= 1.0
