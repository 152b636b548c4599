"""The mathematics behind the chunked line solves (csrc/lines2.cu) and the split-z operators (csrc/splitz.cu), on the CPU
and with the reference's own LU factors: a periodic compact line solve (TRIDPFS / TRIDPSS, src/utils/linear3.f90:269-442)
can be carried out by ranks that each own a contiguous part of the line and know, beyond their own data, only

  * the forward end values y of the 6 chunks (of 16 points) before theirs,
  * the end values y and zero-inflow first values x^_0 of the 6 chunks after theirs,
  * (first and last rank only) the closure terms of the first K0 / last K1 chunks of the line, from the ends of the last 12
    (first 6) chunks,

and the result equals the whole-line solve to round-off.  This is a numpy restatement of the algorithm, not of the CUDA
code: it pins the claims the kernels rely on (look-back / look-ahead windows of 6 chunks, closure carried by a few chunks
at the two ends, no closure correction in the interior) against the oracle's TRIDPSS."""
import numpy as np
import pytest

from common import grid_periodic

C = 16          # points per chunk
LB = 6          # look-back / look-ahead window in chunks


def _periodic_lu(n, second=False):
    """(alpha, beta, gamma, pd, pe) of the factored circulant system of the first (second) derivative, as TRIDPFS leaves them."""
    from oracle import fdm
    g = fdm.Plan(grid_periodic(n), True, True, name="z")
    lu = g.der2.lu if second else g.der1.lu
    return [lu[1:, k].copy() for k in range(1, 6)]


def _whole_line(lu, f):
    from oracle import fdm
    x = f.copy()
    fdm.tridpss(*lu, x)
    return x


def _recurrences(lu):
    """TRIDPSS as two first-order recurrences over the whole line plus a rank-one closure:
         yh_j = f_j + a_j yh_{j-1};   x_N = sum_j p_j yh_j;   x_j = d_j yh_j + g_j x_{j+1} + e_j x_N."""
    alpha, beta, gamma, pd, pe = lu
    n = alpha.size
    a = np.zeros(n)
    a[1:n - 1] = alpha[1:n - 1] * beta[0:n - 2] / beta[1:n - 1]
    p = -beta[n - 1] * pd * beta
    p[n - 1] = beta[n - 1]
    d = beta.copy(); d[n - 1] = 0.0
    g = gamma.copy(); g[n - 2] = 0.0; g[n - 1] = 0.0
    e = pe.copy(); e[n - 1] = 1.0
    return a, p, d, g, e


def _chunk_tables(a, p, d, g, e):
    n = a.size
    T = n // C
    P = np.zeros(n); Q = np.zeros(n); R = np.zeros(n)
    Af = np.zeros(T); Rb = np.zeros(T); Q0 = np.zeros(T); PP = np.zeros(T)
    for t in range(T):
        s0 = t * C
        w = 1.0
        for j in range(C):
            w *= a[s0 + j]; P[s0 + j] = w; PP[t] += p[s0 + j] * w
        Af[t] = w
        w, q = 1.0, 0.0
        for j in range(C - 1, -1, -1):
            w *= g[s0 + j]; R[s0 + j] = w
            q = d[s0 + j] * P[s0 + j] + g[s0 + j] * q; Q[s0 + j] = q
        Rb[t] = w; Q0[t] = q
    S = np.zeros(n + 1)
    for i in range(n - 1, -1, -1):
        S[i] = e[i] + g[i] * S[i + 1]
    return dict(T=T, P=P, Q=Q, R=R, S=S[:n], Af=Af, Rb=Rb, Q0=Q0, PP=PP)


def _local_sweeps(f, a, p, d, g):
    """Zero-inflow sweeps of every chunk: x^ (n values), y end, x^_0, closure part per chunk."""
    n = f.shape[0]
    T = n // C
    xh = np.zeros_like(f)
    yend = np.zeros((T,) + f.shape[1:]); xh0 = np.zeros_like(yend); part = np.zeros_like(yend)
    for t in range(T):
        s0 = t * C
        y = np.zeros((C,) + f.shape[1:])
        acc = 0.0
        for j in range(C):
            acc = f[s0 + j] + a[s0 + j] * acc
            y[j] = acc
            part[t] += p[s0 + j] * acc
        yend[t] = acc
        x = 0.0
        for j in range(C - 1, -1, -1):
            x = d[s0 + j] * y[j] + g[s0 + j] * x
            xh[s0 + j] = x
        xh0[t] = xh[s0]
    return xh, yend, xh0, part


def _carrying_chunks(tab, p):
    """Chunks whose closure weights p or closure response S are not negligible (2^-80 of the largest)."""
    T = tab["T"]
    pm, sm = np.abs(p).max(), np.abs(tab["S"]).max()
    car = [np.any(np.abs(p[t * C:(t + 1) * C]) > np.ldexp(pm, -80)) or np.any(np.abs(tab["S"][t * C:(t + 1) * C]) > np.ldexp(sm, -80))
           for t in range(T)]
    k0 = 0
    while k0 < T and car[k0]:
        k0 += 1
    k1 = 0
    while k1 < T - k0 and car[T - 1 - k1]:
        k1 += 1
    assert not any(car[k0:T - k1]), "closure carried by interior chunks"
    return k0, k1


def _split_solve(lu, f, nranks):
    a, p, d, g, e = _recurrences(lu)
    tab = _chunk_tables(a, p, d, g, e)
    T = tab["T"]
    Tl = T // nranks
    assert Tl >= LB and T >= 2 * LB
    K0, K1 = _carrying_chunks(tab, p)
    assert K0 <= LB and K1 <= LB
    xh, yend, xh0, part = _local_sweeps(f, a, p, d, g)      # every rank computes these for its own chunks only

    def A_of(t, ys):
        """A(t) = sum_k wf_k y(t-k) from a dict of available chunk ends (missing = beyond the window or before the line)."""
        A, w = 0.0, 1.0
        for k in range(1, LB + 1):
            if t - k < 0:
                break
            A = A + w * ys[t - k]
            w *= tab["Af"][t - k]
        return A

    x = np.zeros_like(f)
    for r in range(nranks):
        t0, t1 = r * Tl, (r + 1) * Tl
        prev, nxt = (r - 1) % nranks, (r + 1) % nranks
        # what this rank holds: its own ends, the previous rank's last 6 y, the next rank's first 6 (y, x^_0, part)
        ys = {t: yend[t] for t in range(t0, t1)}
        ys.update({t: yend[t] for t in range((prev + 1) * Tl - LB, (prev + 1) * Tl)})
        nys = {t: yend[t] for t in range(nxt * Tl, nxt * Tl + LB)}
        if r < nranks - 1:
            ys.update(nys)
        A = {t: A_of(t, ys) for t in range(t0, t1)}
        zeta = {t: xh0[t] + tab["Q0"][t] * A[t] for t in range(t0, t1)}
        if r < nranks - 1:
            for t in range(t1, t1 + LB):
                zeta[t] = xh0[t] + tab["Q0"][t] * A_of(t, ys)
        # closure: first rank from its own first K0 chunks and the tail it was sent; last rank from its own last K1 chunks
        # and the first chunks of the line (its periodic neighbour's); nobody else needs it
        xN = 0.0
        if r == 0 or r == nranks - 1:
            head_y = {t: yend[t] for t in range(0, LB)}
            tail_y = {t: yend[t] for t in range(T - 2 * LB, T)}
            for t in range(0, K0):
                xN = xN + (part[t] + tab["PP"][t] * A_of(t, head_y))
            for t in range(T - K1, T):
                xN = xN + (part[t] + tab["PP"][t] * A_of(t, tail_y))
        for t in range(t0, t1):
            B, w = 0.0, 1.0
            for k in range(1, LB + 1):
                if t + k > T - 1:
                    break
                B = B + w * zeta[t + k]
                w *= tab["Rb"][t + k]
            sl = slice(t * C, (t + 1) * C)
            shape = (C,) + (1,) * (f.ndim - 1)
            x[sl] = xh[sl] + tab["Q"][sl].reshape(shape) * A[t] + tab["R"][sl].reshape(shape) * B + tab["S"][sl].reshape(shape) * xN
    return x


@pytest.mark.parametrize("n,nranks,second", [(1024, 8, False), (1024, 2, False), (1024, 8, True), (192, 2, False), (2048, 8, True)])
def test_split_periodic_line_solve_equals_tridpss(n, nranks, second):
    lu = _periodic_lu(n, second)
    rng = np.random.default_rng(11)
    f = rng.standard_normal((n, 5))
    ref = _whole_line(lu, f)
    got = _split_solve(lu, f, nranks)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err <= 1e-13, err


def test_one_rank_is_the_chunked_whole_line_solve():
    """nranks = 1 is the single-GPU kernel of lines2.cu: 64 chunks, windows of 6."""
    lu = _periodic_lu(1024)
    f = np.random.default_rng(12).standard_normal((1024, 3))
    err = np.linalg.norm(_split_solve(lu, f, 1) - _whole_line(lu, f)) / np.linalg.norm(_whole_line(lu, f))
    assert err <= 1e-13, err


# ---------------------------------------------------------------------------------------------------------------------
# non-periodic lines (y on the stretched grid): TRIDSS (linear3.f90:56-150) with windows of 6 chunks, and the claim behind
# the compact BOUNDARY_BCS_NEUMANN_Y kernel: the derivative at a wall is determined by the 6 chunks next to that wall.
def _biased_lu(n):
    from oracle import fdm
    from common import grid_tanh
    g = fdm.Plan(grid_tanh(n), False, False, name="y")
    return [g.der1.lu[1:, k].copy() for k in range(1, 4)]        # ibc = 0: alpha, beta, gamma of TRIDFS


def _chunked_tridss(lu, f, only_chunks=None):
    """x with look-back / look-ahead windows of LB chunks; `only_chunks`: the right-hand side of every other chunk is
    treated as absent (zero), as the compact Neumann kernel does."""
    alpha, beta, gamma = lu
    n = alpha.size
    T = n // C
    a = alpha.copy(); a[0] = 0.0
    d = beta.copy()
    g = gamma * beta
    g[n - 1] = 0.0
    fz = f.copy()
    if only_chunks is not None:
        for t in range(T):
            if t not in only_chunks:
                fz[t * C:(t + 1) * C] = 0.0
    P = np.zeros(n); Q = np.zeros(n); R = np.zeros(n); Af = np.zeros(T); Rb = np.zeros(T); Q0 = np.zeros(T)
    xh = np.zeros_like(f); yend = np.zeros((T,) + f.shape[1:]); xh0 = np.zeros_like(yend)
    for t in range(T):
        s0 = t * C
        w = 1.0
        acc = 0.0
        y = np.zeros((C,) + f.shape[1:])
        for j in range(C):
            w *= a[s0 + j]; P[s0 + j] = w
            acc = fz[s0 + j] + a[s0 + j] * acc
            y[j] = acc
        Af[t] = w; yend[t] = acc
        w, q, x = 1.0, 0.0, 0.0
        for j in range(C - 1, -1, -1):
            w *= g[s0 + j]; R[s0 + j] = w
            q = d[s0 + j] * P[s0 + j] + g[s0 + j] * q; Q[s0 + j] = q
            x = d[s0 + j] * y[j] + g[s0 + j] * x
            xh[s0 + j] = x
        Rb[t] = w; Q0[t] = q; xh0[t] = xh[s0]
    A = np.zeros_like(yend)
    for t in range(T):
        w = 1.0
        for k in range(1, LB + 1):
            if t - k < 0:
                break
            A[t] = A[t] + w * yend[t - k]
            w *= Af[t - k]
    zeta = xh0 + Q0.reshape((T,) + (1,) * (f.ndim - 1)) * A
    x = np.zeros_like(f)
    for t in range(T):
        B, w = 0.0, 1.0
        for k in range(1, LB + 1):
            if t + k > T - 1:
                break
            B = B + w * zeta[t + k]
            w *= Rb[t + k]
        sl = slice(t * C, (t + 1) * C)
        shape = (C,) + (1,) * (f.ndim - 1)
        x[sl] = xh[sl] + Q[sl].reshape(shape) * A[t] + R[sl].reshape(shape) * B
    return x


def test_chunked_biased_line_solve_equals_tridss():
    from oracle import fdm
    n = 512
    lu = _biased_lu(n)
    f = np.random.default_rng(13).standard_normal((n, 4))
    ref = f.copy()
    fdm.tridss(*lu, ref)
    got = _chunked_tridss(lu, f)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13


def test_wall_values_feel_six_chunks_only():
    """The solution next to a wall changes by less than 1e-15 of its size when the right-hand side beyond the 6 chunks
    next to that wall is dropped (what the Neumann boundary-value kernel relies on)."""
    from oracle import fdm
    n = 512
    T = n // C
    lu = _biased_lu(n)
    f = np.random.default_rng(14).standard_normal((n, 4))
    ref = f.copy()
    fdm.tridss(*lu, ref)
    bottom = _chunked_tridss(lu, f, only_chunks=set(range(LB)))
    top = _chunked_tridss(lu, f, only_chunks=set(range(T - LB, T)))
    scale = np.abs(ref).max()
    assert np.abs(bottom[:2] - ref[:2]).max() <= 1e-15 * scale
    assert np.abs(top[-2:] - ref[-2:]).max() <= 1e-15 * scale
