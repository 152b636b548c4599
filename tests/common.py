"""Shared synthetic grids and fields for the parity tests."""
import numpy as np


def grid_periodic(n, length=2.0 * np.pi):
    return np.arange(n) * length / n


def grid_tanh(n, length=1.0, st=0.9375, f=2.0, delta=0.0078125):
    """tlab's tanh stretching (src/tools/initialize/grid/grid_local.f90:55-66): one segment
    y(s) = s + (f-1)*delta*ln(exp((s-st)/delta)+1) on the uniform grid s, shifted so that y(0)=0,
    with Case10's parameters by default."""
    s = np.arange(n) * length / (n - 1)
    y = s + (f - 1.0) * delta * np.logaddexp((s - st) / delta, 0.0)
    return y - y[0]


def grid_stretched(n, length=1.0, amp=0.2):
    s = np.linspace(0.0, length, n)
    return s + amp * np.sin(np.pi * s / length) * length / np.pi


def smooth_field(shape, grids, seed=20261017, nmodes=8):
    """sum_m A_m sin(k_m.x + phi_m) g_m(y), integer wave vectors |k|<=8."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    x, y, z = grids
    lx = (x[1] - x[0]) * nx if nx > 1 else 1.0
    lz = (z[1] - z[0]) * nz if nz > 1 else 1.0
    ly = y[-1] - y[0] if ny > 1 else 1.0
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    out = np.zeros(shape)
    for m in range(nmodes):
        kx = rng.integers(-8, 9)
        kz = rng.integers(-8, 9) if nz > 1 else 0
        a = rng.uniform(0.1, 1.0)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        g = [np.ones_like(Y), np.cos(np.pi * Y / ly), Y / ly][m % 3]
        out += a * np.sin(2 * np.pi * kx * X / lx + 2 * np.pi * kz * Z / lz + ph) * g
    return out


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def dense_direct2(der, n):
    """A2 (tridiagonal + wall extensions) and B2 (pentadiagonal, MatMul_5d column conventions) of the direct second derivative."""
    L, R = der.lhs, der.rhs
    A, B = np.zeros((n, n)), np.zeros((n, n))
    for i in range(1, n + 1):
        if i > 1:
            A[i - 1, i - 2] = L[i, 1]
        A[i - 1, i - 1] = L[i, 2]
        if i < n:
            A[i - 1, i] = L[i, 3]
    A[0, 2] = L[1, 1]                       # extended stencil of the wall rows (fdm_comx_direct.f90: third lhs coefficient)
    A[n - 1, n - 3] = L[n, 3]
    for i in range(3, n - 1):
        for k in range(1, 6):
            B[i - 1, i - 3 + k - 1] = R[i, k]
    B[0, 0:3] = R[1, 3:6]; B[0, 3] = R[1, 1]
    B[1, 0:4] = R[2, 2:6]
    B[n - 2, n - 4:n] = R[n - 1, 1:5]
    B[n - 1, n - 3:n] = R[n, 1:4]; B[n - 1, n - 4] = R[n, 5]
    return A, B
