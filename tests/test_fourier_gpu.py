"""OPR_Fourier_X/Z_Forward/Backward as entry points (opr_fourier.f90:219-433): layout of the half spectrum c(nx/2+1, ny, nz)
incl. the Nyquist element, the strided z transform, FFTW's sign and (missing) normalisation -- against numpy's FFT, which is what
the oracle of OPR_Poisson uses (oracle/operators.py).  Tolerance 1e-12 relative L2 (north_star, per operator call)."""
import numpy as np
import pytest

from common import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(64, 48, 32), (30, 7, 12), (128, 16, 1), (16, 5, 100)])
def test_fourier_transforms_match_numpy(cuda, shape):
    import torch
    from tlab_b200 import opr
    nx, ny, nz = shape
    rng = np.random.default_rng(11)
    a = rng.standard_normal((nz, ny, nx))
    u = torch.from_numpy(a).to(cuda)
    nxh = nx // 2 + 1
    c = torch.full(((nx + 2) * ny * nz,), float("nan"), dtype=torch.float64, device=cuda)
    opr.OPR_Fourier_X_Forward(nx, ny, nz, u, c)
    cx = c.cpu().numpy().reshape(nz, ny, nxh, 2)
    cx = cx[..., 0] + 1j * cx[..., 1]
    ref_x = np.fft.rfft(a, axis=2)
    assert rel_l2(np.abs(cx - ref_x), 0 * np.abs(ref_x)) <= 1e-12 * np.linalg.norm(ref_x)
    # Nyquist element of every line is real and sits at index nx/2
    assert np.abs(cx[..., nxh - 1].imag).max() <= 1e-12 * np.abs(ref_x).max()
    d = torch.full_like(c, float("nan"))
    opr.OPR_Fourier_Z_Forward(nx, ny, nz, c, d)
    cz = d.cpu().numpy().reshape(nz, ny, nxh, 2)
    cz = cz[..., 0] + 1j * cz[..., 1]
    ref_z = np.fft.fft(ref_x, axis=0)
    assert np.linalg.norm(cz - ref_z) <= 1e-12 * np.linalg.norm(ref_z)
    # backward: unnormalised, z in place, x out of place
    opr.OPR_Fourier_Z_Backward(nx, ny, nz, d, d)
    back_z = d.cpu().numpy().reshape(nz, ny, nxh, 2)
    back_z = back_z[..., 0] + 1j * back_z[..., 1]
    assert np.linalg.norm(back_z - nz * ref_x) <= 1e-12 * np.linalg.norm(nz * ref_x)
    r = torch.full_like(u, float("nan"))
    opr.OPR_Fourier_X_Backward(nx, ny, nz, d, r)
    assert rel_l2(r.cpu().numpy(), nx * nz * a) <= 1e-12


def test_managed_memory_is_valid_on_both_sides(cuda):
    """tlab_gpu_malloc_managed: the array a Fortran host indexes is the array the kernels read (OPR_Partial through the same
    pointer, against the derivative of the same data in ordinary device memory)."""
    import ctypes
    import torch
    from tlab_b200 import opr, lib as tl
    from common import grid_periodic
    L = tl.load()
    nx, ny, nz = 64, 4, 4
    N = nx * ny * nz
    g = opr.FdmPlan(grid_periodic(nx), True, True, name="x")
    pu, pr = ctypes.c_void_p(), ctypes.c_void_p()
    tl.check(L.tlab_gpu_malloc_managed(ctypes.byref(pu), N * 8))
    tl.check(L.tlab_gpu_malloc_managed(ctypes.byref(pr), N * 8))
    hu = np.ctypeslib.as_array(ctypes.cast(pu, ctypes.POINTER(ctypes.c_double)), shape=(N,))
    hr = np.ctypeslib.as_array(ctypes.cast(pr, ctypes.POINTER(ctypes.c_double)), shape=(N,))
    a = np.sin(3 * np.tile(grid_periodic(nx), ny * nz)) + 0.1 * np.random.default_rng(2).standard_normal(N)
    hu[:] = a                                   # written by the host ...
    tl.check(L.tlab_gpu_prefetch(pu, N * 8, 1))
    bcs = (ctypes.c_int * 4)(0, 0, 0, 0)
    tl.check(L.tlab_opr_partial(1, 1, nx, ny, nz, bcs, g.handle, pu, pr, None))     # ... read and written by the device ...
    tl.check(L.tlab_gpu_prefetch(pr, N * 8, 0))
    got = np.array(hr)                          # ... and read back by the host, no copies
    u = torch.from_numpy(a).to(cuda)
    r = torch.empty_like(u)
    opr.OPR_Partial_X(opr.OPR_P1, nx, ny, nz, [[0, 0], [0, 0]], g, u, r)
    assert np.array_equal(got, r.cpu().numpy())
    del hu, hr
    tl.check(L.tlab_gpu_free(pu))
    tl.check(L.tlab_gpu_free(pr))
