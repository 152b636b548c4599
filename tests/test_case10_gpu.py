"""examples/Case10/dns.out.ref reproduced by the CUDA path: ten CFL-controlled RK4-5 steps of the 512 x 257 Boussinesq
case from the restated initial condition (tests/case10.py), time step, CFL / diffusion numbers and dilatation bounds
computed on the device; every logged digit of the reference build must come out."""
import numpy as np
import pytest

import case10

pytestmark = pytest.mark.gpu


def test_gpu_reproduces_case10_log(cuda):
    from tlab_b200 import opr, dns as GD
    x, y, z = case10.grids()
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    g = GD.Dns(gg, **case10.dns_kwargs(GD, y))
    g.set("s1", case10.initial_scalar(x, y))
    rows = case10.run(g, 10)
    assert case10.compare_with_reference_log(rows) == []
    s = g.get("s1")
    assert s.min() >= 0.0 and s.max() <= 1.0 and np.isfinite(g.get("q2")).all()
    g.close()
