"""CPU oracle for the tlab incompressible/Boussinesq RHS + RK substep path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the
reference's Fortran algorithm (turbulencia/tlab, /root/reference/src), written
function by function with the reference file:line each one follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product (``tlab_b200``) never
does.

Parity pinning (see DESIGN.md "Oracle"): the reference cannot be built here
(Fortran 2008 + FFTW3; no Fortran compiler in the image), and ships no binary
golden fields.  The oracle is pinned against
  * the reference build's own logs of examples/Case10, Case06 and Case07
    (``dns.out.ref``, copied to tests/golden/case*_dns.out.ref): from each case's
    restated initial condition
    (tests/tlab_cases.py) the oracle reproduces all ten logged iterations -- time,
    dt, CFL and diffusion numbers, min/max dilatation -- to every printed digit
    (6 significant digits for the dilatation), tests/test_case10_cpu.py; this
    covers grid, compact schemes on the stretched grid, OPR_Burgers, the Poisson
    solver (FFT included), the RK4-5 advance, TIME_COURANT, DNS_BOUNDS_CONTROL,
  * the reference's own numpy restatement of the C1N6 schemes,
    ``scripts/python/compact_lib.py`` (imported in this container by
    ``tests/golden/make_golden.py``; outputs committed under tests/golden/),
  * the uniform-grid coefficient limits quoted in the reference sources,
  * the reference's self-consistency recipes (vburgers, vpoisson, vintegral,
    vpartial analytic convergence).
FFT values alone (FFTW3, external, un-pinned version) have no golden of their
own; they are pinned through the Poisson round-trip identity and through the
Case10 log (whose dilatation after each step is what the Poisson solve leaves).
"""
