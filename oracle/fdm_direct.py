"""Oracle: compact schemes derived directly on the non-uniform grid ("CompactDirect6" second derivative).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows /root/reference/src:
  fdm/fdm_base.f90          Pi (:31-44), Pi_p (:47-67), Pi_pp_3 (:70-78), Lag (:82-97), Lag_p (:100-125), Lag_pp_3 (:128-143)
  fdm/fdm_comx_direct.f90   FDM_C2N6_Direct (:305-412), coef_c2n4 (:468-495), coef_c2n3_biased (:497-552),
                            a2n6_coef, b2n6_coef, c2n6_coef (:554-626), PIp_o_PI, PIpp_o_PI, D_coef, A1D, A2D, B1D, B2D, C1D, C2D (:642-783)
Indices are the reference's 1-based ones: x is passed as a padded array xp with xp[i] = x(i).
"""
import numpy as np


def _pad(x):
    xp = np.zeros(len(x) + 1)
    xp[1:] = x
    return xp


# ---- fdm_base.f90 ---------------------------------------------------------------------------------------------------
def Pi(x, j, idx):
    f = 1.0
    for k in idx:
        f = f * (x[j] - x[k])
    return f


def Pi_p(x, j, idx):
    f = 0.0
    for k in range(len(idx)):
        dummy = 1.0
        for m in range(len(idx)):
            if m != k:
                dummy = dummy * (x[j] - x[idx[m]])
        f = f + dummy
    return f


def Pi_pp_3(x, j, idx):
    return 2.0 * (x[j] - x[idx[0]] + x[j] - x[idx[1]] + x[j] - x[idx[2]])


def Lag(x, j, i, idx):
    f = 1.0
    for k in idx:
        if k != i:
            f = f * (x[j] - x[k]) / (x[i] - x[k])
    return f


def Lag_p(x, j, i, idx):
    den = 1.0
    f = 0.0
    for k in range(len(idx)):
        if idx[k] != i:
            dummy = 1.0
            for m in range(len(idx)):
                if idx[m] != i and m != k:
                    dummy = dummy * (x[j] - x[idx[m]])
            f = f + dummy
            den = den * (x[i] - x[idx[k]])
    return f / den


def Lag_pp_3(x, j, i, idx):
    f = 2.0
    for k in idx:
        if k != i:
            f = f / (x[i] - x[k])
    return f


# ---- fdm_comx_direct.f90 -------------------------------------------------------------------------------------------
def PIp_o_PI(x, j, i):
    f = ((x[j] - x[i + 2]) * (x[j] - x[i - 2]) + (x[j] - x[i]) * (x[j] - x[i - 2]) + (x[j] - x[i]) * (x[j] - x[i + 2]))
    return f / Pi(x, j, [i - 2, i, i + 2])


def PIpp_o_PI(x, j, i):
    f = x[j] - x[i + 2] + x[j] - x[i - 2] + x[j] - x[i]
    return 2.0 * f / Pi(x, j, [i - 2, i, i + 2])


def D_coef(x, i):
    dx = x[i + 1] - x[i - 1]
    return (6.0 + 4.0 * dx * (PIp_o_PI(x, i + 1, i) - PIp_o_PI(x, i - 1, i))
            - 2.0 * dx ** 2.0 * PIp_o_PI(x, i + 1, i) * PIp_o_PI(x, i - 1, i))


def A1D(x, im, ip, i):
    dx = x[ip] - x[im]
    return (-4.0 * PIp_o_PI(x, ip, i) - 2.0 * PIp_o_PI(x, im, i)
            + 2.0 * dx * (PIp_o_PI(x, ip, i) * PIp_o_PI(x, im, i) - PIpp_o_PI(x, ip, i))
            + dx ** 2.0 * PIpp_o_PI(x, ip, i) * PIp_o_PI(x, im, i))


def A2D(x, im, ip, i):
    dx = x[ip] - x[im]
    return (4.0 * PIp_o_PI(x, ip, i) * PIp_o_PI(x, im, i) - PIpp_o_PI(x, ip, i)
            - 2.0 / dx * (PIp_o_PI(x, ip, i) - PIp_o_PI(x, im, i))
            + dx * PIpp_o_PI(x, ip, i) * PIp_o_PI(x, im, i))


def B1D(x, im, ip, i):
    dx = x[ip] - x[im]
    return -(2.0 / dx + PIp_o_PI(x, ip, i)) * dx ** 2.0


def B2D(x, im, ip, i):
    dx = x[ip] - x[im]
    return 1.0 - dx * PIp_o_PI(x, im, i)


def C1D(x, j, i):
    dx = x[i + 1] - x[i - 1]
    dxp = x[i + 1] - x[j]
    dxm = x[j] - x[i - 1]
    return ((dxp - dxm) / (dxp * dxm) * (6.0 - 4.0 * dx ** 2.0 / (dxp * dxm))
            + 2.0 * dx * (dxm / dxp - dxp / dxm) * PIp_o_PI(x, i - 1, i) * PIp_o_PI(x, i + 1, i)
            + PIp_o_PI(x, i - 1, i) * (4.0 * dx / dxp - 4.0 * dx / dxm - 2.0 * dx ** 2.0 / dxp ** 2.0)
            - PIp_o_PI(x, i + 1, i) * (4.0 * dx / dxp - 4.0 * dx / dxm + 2.0 * dx ** 2.0 / dxm ** 2.0))


def C2D(x, j, i):
    dx = x[i + 1] - x[i - 1]
    dxp = x[i + 1] - x[j]
    dxm = x[j] - x[i - 1]
    return (2.0 * (1.0 / dxp ** 2.0 + 1.0 / dxm ** 2.0 - 1.0 / (dxp * dxm))
            + 2.0 * dx ** 2.0 / (dxp * dxm) * PIp_o_PI(x, i + 1, i) * PIp_o_PI(x, i - 1, i)
            - 2.0 * PIp_o_PI(x, i + 1, i) * dx / dxm * (1.0 / dxp - 1.0 / dxm)
            - 2.0 * PIp_o_PI(x, i - 1, i) * dx / dxp * (1.0 / dxp - 1.0 / dxm))


def a2n6_coef(x, im, ip, i):
    dx = x[ip] - x[im]
    dxp = x[i] - x[ip]
    dxm = x[i] - x[im]
    f1 = B1D(x, ip, im, i) * (dxm + dxp) + B2D(x, im, ip, i) * dxp * (dxp + 2.0 * dxm)
    f1 = f1 * 2.0 * Pi_p(x, i, [i - 2, i, i + 2])
    f2 = B1D(x, ip, im, i) + B2D(x, im, ip, i) * dxp
    f2 = f2 * Pi_pp_3(x, i, [i - 2, i, i + 2]) * dxp * dxm
    return -(f1 + f2) / dx / Pi(x, ip, [i - 2, i, i + 2])


def b2n6_coef(x, im, ip, i):
    dx = x[ip] - x[im]
    dxp = x[i] - x[ip]
    dxm = x[i] - x[im]
    D = D_coef(x, i)
    f1 = 1.0 + A1D(x, im, ip, i) / D * (dxm + dxp) + A2D(x, im, ip, i) / D * dxp * (dxp + 2.0 * dxm)
    f1 = f1 * 2.0 * Pi_p(x, i, [i - 2, i, i + 2])
    f2 = 1.0 + A1D(x, im, ip, i) / D * dxp + A2D(x, im, ip, i) / D * dxp ** 2.0
    f2 = f2 * Pi_pp_3(x, i, [i - 2, i, i + 2]) * dxm
    return (f1 + f2) / dx / Pi(x, ip, [i - 2, i, i + 2])


def c2n6_coef(x, j, i):
    dx = x[i] - x[j]
    dxp = x[i] - x[i + 1]
    dxm = x[i] - x[i - 1]
    dxp2 = x[j] - x[i + 1]
    dxm2 = x[j] - x[i - 1]
    D = D_coef(x, i)
    f1 = (C1D(x, j, i) / D * (1.0 + dx / dxp + dx / dxm) + C2D(x, j, i) / D * (2.0 + dx / dxp + dx / dxm) * dx
          + 1.0 / dxp + 1.0 / dxm)
    f1 = f1 * 2.0 * Lag_p(x, i, j, [i - 2, i, i + 2]) * dxp * dxm / (dxp2 * dxm2)
    f2 = 1.0 + C1D(x, j, i) / D * dx + C2D(x, j, i) / D * dx ** 2.0
    f2 = f2 * Lag_pp_3(x, j, j, [i - 2, i, i + 2]) * dxp * dxm / (dxp2 * dxm2)
    return f1 + f2


def coef_c2n4(x, i):
    dx = x[i + 1] - x[i - 1]
    dxp = x[i + 1] - x[i]
    dxm = x[i] - x[i - 1]
    D = dxp * dxm + dx ** 2.0
    am1 = (dxm ** 2.0 - dxp ** 2.0 + dxp * dxm) * dxp / dx / D
    a = 1.0
    ap1 = (dxp ** 2.0 - dxm ** 2.0 + dxp * dxm) * dxm / dx / D
    bm1 = dxp / dx * 12.0 / D
    b = -12.0 / D
    bp1 = dxm / dx * 12.0 / D
    return [am1, a, ap1, bm1, b, bp1]


def coef_c2n3_biased(x, i, backwards=False):
    i1 = i
    if backwards:
        i2, i3, i4 = i - 1, i - 2, i - 3
    else:
        i2, i3, i4 = i + 1, i + 2, i + 3
    dx1 = x[i2] - x[i1]
    dx3 = x[i2] - x[i3]
    dx4 = x[i2] - x[i4]
    set_m = [i1, i3, i4]
    a1 = 1.0
    a2 = (0.5 * dx1 * Pi_pp_3(x, i1, set_m) - Pi_p(x, i1, set_m)) / Pi_p(x, i2, set_m)
    b2 = (Pi_pp_3(x, i1, set_m) + 0.5 * dx1 * Pi_pp_3(x, i1, set_m) * Pi_pp_3(x, i2, set_m) / Pi_p(x, i2, set_m)
          - Pi_p(x, i1, set_m) / Pi_p(x, i2, set_m) * Pi_pp_3(x, i2, set_m))
    b2 = b2 / Pi(x, i2, set_m)
    D = Lag(x, i2, i1, set_m) + dx1 * Lag_p(x, i2, i1, set_m)
    b1 = (-2.0 * Lag_p(x, i1, i1, set_m) * (Lag(x, i2, i1, set_m) + 2.0 * dx1 * Lag_p(x, i2, i1, set_m))
          + 2.0 * Lag_p(x, i2, i1, set_m))
    b1 = b1 / D / dx1 + Lag_pp_3(x, i1, i1, set_m)
    D = Lag(x, i2, i3, set_m) + dx3 * Lag_p(x, i2, i3, set_m)
    b3 = ((Lag(x, i2, i3, set_m) + dx1 * Lag_p(x, i2, i3, set_m)) * dx1 * Lag_pp_3(x, i1, i3, set_m)
          - 2.0 * (Lag(x, i2, i3, set_m) + 2.0 * dx1 * Lag_p(x, i2, i3, set_m)) * Lag_p(x, i1, i3, set_m))
    b3 = b3 / D / dx3
    D = Lag(x, i2, i4, set_m) + dx4 * Lag_p(x, i2, i4, set_m)
    b4 = ((Lag(x, i2, i4, set_m) + dx1 * Lag_p(x, i2, i4, set_m)) * dx1 * Lag_pp_3(x, i1, i4, set_m)
          - 2.0 * (Lag(x, i2, i4, set_m) + 2.0 * dx1 * Lag_p(x, i2, i4, set_m)) * Lag_p(x, i1, i4, set_m))
    b4 = b4 / D / dx4
    return [a1, a2, b1, b2, b3, b4]


def c2n6_direct(nodes):
    """FDM_C2N6_Direct (fdm_comx_direct.f90:305-412).  Returns lhs(n+1, 3+1), rhs(n+1, 5+1) (1-based padded) and nb_diag."""
    nmax = len(nodes)
    x = _pad(np.asarray(nodes, dtype=np.float64))
    lhs = np.zeros((nmax + 1, 4))
    rhs = np.zeros((nmax + 1, 6))
    # first / last points
    n = 1
    coef = coef_c2n3_biased(x, n)
    dummy = 1.0 / coef[2]
    lhs[n, 2] = coef[0] * dummy
    lhs[n, 3] = coef[1] * dummy
    for col, c in zip((3, 4, 5, 1), coef[2:6]):          # b, bp1, bp2, bp3; bp3 is saved into rhs(1)
        rhs[n, col] = c * dummy
    n = nmax
    coef = coef_c2n3_biased(x, n, backwards=True)
    dummy = 1.0 / coef[2]
    lhs[n, 2] = coef[0] * dummy
    lhs[n, 1] = coef[1] * dummy
    for col, c in zip((3, 2, 1, 5), coef[2:6]):          # b, bm1, bm2, bm3; bm3 is saved into rhs(5)
        rhs[n, col] = c * dummy
    # second / second-to-last points
    for n in (2, nmax - 1):
        coef = coef_c2n4(x, n)
        dummy = 1.0 / coef[4]
        lhs[n, 1:4] = np.array(coef[0:3]) * dummy
        rhs[n, 2:5] = np.array(coef[3:6]) * dummy
    # interior points
    for n in range(3, nmax - 1):
        D = D_coef(x, n)
        a = 1.0
        ap1 = a2n6_coef(x, n - 1, n + 1, n) / D
        am1 = a2n6_coef(x, n + 1, n - 1, n) / D
        bp1 = b2n6_coef(x, n - 1, n + 1, n)
        bm1 = b2n6_coef(x, n + 1, n - 1, n)
        dxp = x[n] - x[n + 1]
        dxm = x[n] - x[n - 1]
        b = (2.0 * C2D(x, n, n) / D + 2.0 * C1D(x, n, n) / D * ((dxm + dxp) / (dxp * dxm) + Lag_p(x, n, n, [n - 2, n, n + 2]))
             + (2.0 + 2.0 * Lag_p(x, n, n, [n - 2, n, n + 2]) * (dxm + dxp)) / (dxp * dxm) + Lag_pp_3(x, n, n, [n - 2, n, n + 2]))
        bp2 = c2n6_coef(x, n + 2, n)
        bm2 = c2n6_coef(x, n - 2, n)
        dummy = 1.0 / bp1
        lhs[n, 1:4] = np.array([am1, a, ap1]) * dummy
        rhs[n, 1:6] = np.array([bm2, bm1, b, bp1, bp2]) * dummy
    return lhs, rhs, (3, 5)
