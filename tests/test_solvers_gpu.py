"""thomas3 / thomas5 substitution stages, TLab_Transpose and the golden first-derivative vectors on the GPU."""
import ctypes
import os

import numpy as np
import pytest

from common import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _dev(torch, cuda, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_thomas_substitution_stages(cuda):
    import torch
    from oracle import fdm
    from tlab_b200 import lib as tl
    L = tl.load()
    rng = np.random.default_rng(2)
    n, m = 50, 37          # nmax, len (ragged on purpose)
    a, b, c = rng.uniform(0.1, 0.4, n), rng.uniform(1.0, 2.0, n), rng.uniform(0.1, 0.4, n)
    f = rng.standard_normal((n, m))         # C (n, len) == Fortran f(len, n)
    # TRIDSS
    la, lb, lc = a.copy(), b.copy(), c.copy()
    fdm.tridfs(la, lb, lc)
    ref = f.copy()
    fdm.tridss(la, lb, lc, ref)
    fd = _dev(torch, cuda, f)
    ds = [_dev(torch, cuda, v) for v in (la, lb, lc)]     # keep the device copies alive during the call
    torch.cuda.synchronize()
    tl.check(L.tlab_tridss(n, m, *[_p(v) for v in ds], _p(fd)))
    assert rel_l2(fd.cpu().numpy(), ref) <= 1e-13
    # TRIDPSS
    pa, pb, pc, pd, pe = a.copy(), b.copy(), c.copy(), np.zeros(n), np.zeros(n)
    fdm.tridpfs(pa, pb, pc, pd, pe)
    ref = f.copy()
    fdm.tridpss(pa, pb, pc, pd, pe, ref)
    fd = _dev(torch, cuda, f)
    wrk = torch.zeros(m, dtype=torch.float64, device=cuda)
    ds = [_dev(torch, cuda, v) for v in (pa, pb, pc, pd, pe)]
    torch.cuda.synchronize()
    tl.check(L.tlab_tridpss(n, m, *[_p(v) for v in ds], _p(fd), _p(wrk)))
    assert rel_l2(fd.cpu().numpy(), ref) <= 1e-13
    # PENTADSS / PENTADSS2
    d5 = [rng.uniform(0.05, 0.2, n), rng.uniform(0.1, 0.4, n), rng.uniform(1.5, 2.0, n), rng.uniform(0.1, 0.4, n),
          rng.uniform(0.05, 0.2, n)]
    for fs, ss, fn in ((fdm.pentadfs, fdm.pentadss, L.tlab_pentadss), (fdm.pentadfs2, fdm.pentadss2, L.tlab_pentadss2)):
        w = [v.copy() for v in d5]
        fs(*w)
        ref = f.copy()
        ss(*w, ref)
        fd = _dev(torch, cuda, f)
        ds = [_dev(torch, cuda, v) for v in w]
        torch.cuda.synchronize()
        tl.check(fn(n, m, *[_p(v) for v in ds], _p(fd)))
        assert rel_l2(fd.cpu().numpy(), ref) <= 1e-13


def test_transpose(cuda):
    import torch
    from tlab_b200 import lib as tl
    L = tl.load()
    rng = np.random.default_rng(4)
    nra, nca = 70, 45
    a = rng.standard_normal((nca, nra))             # Fortran a(nra, nca)
    ad = _dev(torch, cuda, a)
    bd = torch.zeros(nra * nca, dtype=torch.float64, device=cuda)
    torch.cuda.synchronize()
    tl.check(L.tlab_transpose(_p(ad), nra, nca, nra, _p(bd), nca))
    assert np.array_equal(bd.cpu().numpy().reshape(nra, nca), a.T)
    ac = rng.standard_normal((nca, nra, 2))
    ad = _dev(torch, cuda, ac)
    bd = torch.zeros(nra * nca * 2, dtype=torch.float64, device=cuda)
    torch.cuda.synchronize()
    tl.check(L.tlab_transpose_complex(_p(ad), nra, nca, nra, _p(bd), nca))
    assert np.array_equal(bd.cpu().numpy().reshape(nra, nca, 2), ac.transpose(1, 0, 2))


def test_golden_first_derivative(cuda):
    """The CUDA path against the vectors produced by the reference's compact_lib.py (tests/golden)."""
    import torch
    from tlab_b200 import opr
    gold = np.load(os.path.join(HERE, "golden", "compact_lib_c1n6.npz"))
    for key, per, uni in (("per", True, True), ("nonuni", False, False), ("uni", False, True)):
        x, u, du = gold[key + "_x"], gold[key + "_u"], gold[key + "_du"]
        g = opr.FdmPlan(x, per, uni, name="y")
        n, m = u.shape
        ud = _dev(torch, cuda, u)                    # (n, nlines) == Fortran u(nlines, n)
        rd = torch.empty_like(ud)
        opr.FDM_Der1_Solve(m, 0, g, ud, rd)
        assert rel_l2(rd.cpu().numpy(), du) <= 1e-12, key
