// OPR_Poisson: Fourier in x and z (cuFFT), factorised compact integral operators in y.
//
// Replaces OPR_Elliptic_Initialize / OPR_Poisson_FourierXZ_Factorize (src/operators/opr_elliptic.f90:86-364),
// OPR_Fourier_X/Z_Forward/Backward (src/operators/opr_fourier.f90:219-433), OPR_ODE2_Factorize_NN and
// _NN_Sing/_DN_Sing (src/operators/opr_odes.f90:37-96,165-183,265-386), FDM_Int1_Initialize/Solve
// (src/fdm/fdm_integral.f90:58-314) and PENTADFS/PENTADSS (src/utils/linear5.f90:30-131).
//
// Layout: the half spectrum stays in cuFFT's natural order c(kx, y, kz), kx fastest.  One thread owns one
// (kx,kz) mode and marches along y, so consecutive threads read consecutive kx: coalesced, and the three
// complex transposes of the reference (opr_elliptic.f90:301,335,336) disappear.  The pentadiagonal system
// of each mode, (B + lambda A) with lambda = +-sqrt(kx'^2 + kz'^2), is assembled from two shared banded
// tables (L0 + lambda L1, built on the host) and factorised on the fly during the forward sweep instead
// of being stored (the reference keeps 2 x ny x 5 doubles per mode, opr_elliptic.f90:140,205-209).
// The lambda-only "fundamental" solutions (v1, e-, u1, s+, e+; opr_odes.f90:308-348) are computed once
// at initialisation and kept, since they do not depend on the forcing.
#include "../../include/tlab_gpu.h"
#include "context.h"
#include "poisson.h"
#include "trp.h"
#include <cufft.h>
#include <cmath>
#include <vector>
#include <algorithm>

#ifndef POISSON_MIN_BLOCKS
#define POISSON_MIN_BLOCKS 4
#endif

namespace tlab {

namespace {

struct LineRef {
    double* p;            // element of row r (1-based) is p[(r-1)*js]; nullptr reads as 0
    long long js;
    __device__ __forceinline__ double get(int r) const { return p ? p[(long long)(r - 1) * js] : 0.0; }
    __device__ __forceinline__ void set(int r, double v) const { p[(long long)(r - 1) * js] = v; }
};

// -------------------------------------------------------------------------------------------------
// Solve one first-order integral problem  u' + lam u = f  for NL right-hand sides of one mode.
//   P      shared banded tables of this side (BCS_MIN: condition at row 1, BCS_MAX: at row n)
//   f      forcing lines; rows 1 and n are never read, the far-end value is passed in fend
//   bc     value imposed at the near end;   res: result lines (rows 1..n written)
//   ysc, csc, dsc, esc: per-mode scratch lines (intermediate vector and the three upper factors)
//   du     (optional) derivative at the far end, FDM_Int1_Solve's du_boundary
//   FMODE  0: LU factors computed on the fly, upper factors through the scratch lines csc, dsc
//          1: the same, and the four factor lines (fa, fb lower; csc, dsc upper) are kept for later calls
//          2: factor lines read instead of recomputed (they depend on lambda only: no divisions, no table rows and a
//             shorter dependent chain per row)
template <int NL, int FMODE = 0>
__device__ void int1_solve(const Int1Dev& P, double lam, const LineRef (&f)[NL], double fscale,
                           const double (&fend)[NL], const double (&bc)[NL], const LineRef (&res)[NL],
                           const LineRef (&ysc)[NL], LineRef csc, LineRef dsc, LineRef esc, double* du,
                           LineRef fa = LineRef{nullptr, 0}, LineRef fb = LineRef{nullptr, 0}) {
    const int n = P.n;
    const bool is_min = (P.bc == BCS_MIN);
    auto Lrow = [&](int r, double (&row)[6]) {
#pragma unroll
        for (int k = 1; k <= 5; k++) row[k] = P.L0[(r - 1) * 5 + k - 1] + lam * P.L1[(r - 1) * 5 + k - 1];
    };
    auto rhs = [&](int r, int k) { return P.rhs[(r - 1) * 3 + k - 1]; };
    auto F = [&](int l, int r) { return f[l].get(r) * fscale; };

    // ---- reduction of the far-end row into its neighbours (lambda-dependent)
    double Le[6], La[6], Lb[6];           // far-end row, its neighbour, next neighbour (after reduction)
    double rb[4][4], rt[3][5];            // rhs_b(1:3, 0:3), rhs_t(0:2, 1:4)
#pragma unroll
    for (int r = 1; r <= 3; r++)
#pragma unroll
        for (int c = 0; c <= 3; c++) rb[r][c] = P.rb[r][c];
#pragma unroll
    for (int r = 0; r <= 2; r++)
#pragma unroll
        for (int c = 1; c <= 4; c++) rt[r][c] = P.rt[r][c];
    if (is_min) {
        Lrow(n, Le); Lrow(n - 1, La); Lrow(n - 2, Lb);
        const double dummy = 1.0 / Le[3];
#pragma unroll
        for (int k = 1; k <= 5; k++) Le[k] = -Le[k] * dummy;
        Le[3] = 1.0;
        La[1] = La[1] + La[4] * Le[5]; La[2] = La[2] + La[4] * Le[1]; La[3] = La[3] + La[4] * Le[2];
        Lb[2] = Lb[2] + Lb[5] * Le[5]; Lb[3] = Lb[3] + Lb[5] * Le[1]; Lb[4] = Lb[4] + Lb[5] * Le[2];
#pragma unroll
        for (int r = 0; r <= 2; r++)
#pragma unroll
            for (int c = 1; c <= 3; c++) rt[r][c] = rhs(n - 2 + r, c);
#pragma unroll
        for (int c = 1; c <= 3; c++) rt[2][c] = rt[2][c] * dummy;
        rt[1][1] = rt[1][1] - La[4] * rt[2][3]; rt[1][2] = rt[1][2] - La[4] * rt[2][1]; rt[1][3] = rt[1][3] - La[4] * rt[2][2];
        rt[0][2] = rt[0][2] - Lb[5] * rt[2][3]; rt[0][3] = rt[0][3] - Lb[5] * rt[2][1]; rt[0][4] = rt[0][4] - Lb[5] * rt[2][2];
    } else {
        Lrow(1, Le); Lrow(2, La); Lrow(3, Lb);
        const double dummy = 1.0 / Le[3];
#pragma unroll
        for (int k = 1; k <= 5; k++) Le[k] = -Le[k] * dummy;
        Le[3] = 1.0;
        La[3] = La[3] + La[2] * Le[4]; La[4] = La[4] + La[2] * Le[5]; La[5] = La[5] + La[2] * Le[1];
        Lb[2] = Lb[2] + Lb[1] * Le[4]; Lb[3] = Lb[3] + Lb[1] * Le[5]; Lb[4] = Lb[4] + Lb[1] * Le[1];
#pragma unroll
        for (int r = 1; r <= 3; r++)
#pragma unroll
            for (int c = 1; c <= 3; c++) rb[r][c] = rhs(r, c);
#pragma unroll
        for (int c = 1; c <= 3; c++) rb[1][c] = rb[1][c] * dummy;
        rb[2][1] = rb[2][1] - La[2] * rb[1][2]; rb[2][2] = rb[2][2] - La[2] * rb[1][3]; rb[2][3] = rb[2][3] - La[2] * rb[1][1];
        rb[3][0] = rb[3][0] - Lb[1] * rb[1][2]; rb[3][1] = rb[3][1] - Lb[1] * rb[1][3]; rb[3][2] = rb[3][2] - Lb[1] * rb[1][1];
    }

    double F1[NL], FN[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        F1[l] = is_min ? bc[l] : fend[l];
        FN[l] = is_min ? fend[l] : bc[l];
    }

    // ---- forward sweep: right-hand side, on-the-fly LU (PENTADFS) and forward substitution (PENTADSS)
    const int nmax = n - 2;
    double c1 = 0, c2 = 0, d1 = 0, d2 = 0, e1 = 0, e2 = 0;      // factors of rows m-1 and m-2 (unflipped)
    double y1[NL], y2[NL], um[NL], u0[NL], up[NL];              // u(r-1), u(r), u(r+1)
    double bcs_far[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) { y1[l] = y2[l] = 0.0; um[l] = 0.0; u0[l] = F(l, 2); up[l] = F(l, 3); bcs_far[l] = 0.0; }
    if (!is_min) {
#pragma unroll
        for (int l = 0; l < NL; l++) bcs_far[l] = F1[l] * rb[1][2] + u0[l] * rb[1][3] + up[l] * rb[1][1];
    }
    // The forcing rows are fetched FRB rows ahead, in blocks (memory-level parallelism: the sweep is a dependent chain
    // and would otherwise wait for one DRAM round trip every other row).
    constexpr int FRB = 8;
    for (int m0 = 1; m0 <= nmax; m0 += FRB) {
        double fblk[NL][FRB];
        double ablk[FMODE == 2 ? FRB : 1], bblk[FMODE == 2 ? FRB : 1];
#pragma unroll
        for (int j = 0; j < FRB; j++) {
            const int rr = m0 + j + 3;              // row r + 2 of step m = m0 + j
#pragma unroll
            for (int l = 0; l < NL; l++) fblk[l][j] = (rr <= n - 1) ? F(l, rr) : 0.0;
            if (FMODE == 2) {
                const bool ok = (m0 + j <= nmax);
                ablk[j] = ok ? fa.get(m0 + j + 1) : 0.0;
                bblk[j] = ok ? fb.get(m0 + j + 1) : 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < FRB; j++) {
            const int m = m0 + j;
            if (m > nmax) break;
            const int r = m + 1;
            double a = 0.0, b = 0.0, c = 1.0, d = 0.0, e = 0.0;
            if (FMODE == 2) {
                a = ablk[j]; b = bblk[j];
            } else {
                double row[6];
                if (is_min && r == n - 1) {
#pragma unroll
                    for (int k = 1; k <= 5; k++) row[k] = La[k];
                } else if (is_min && r == n - 2) {
#pragma unroll
                    for (int k = 1; k <= 5; k++) row[k] = Lb[k];
                } else if (!is_min && r == 2) {
#pragma unroll
                    for (int k = 1; k <= 5; k++) row[k] = La[k];
                } else if (!is_min && r == 3) {
#pragma unroll
                    for (int k = 1; k <= 5; k++) row[k] = Lb[k];
                } else {
                    Lrow(r, row);
                }
                a = row[1]; b = row[2]; c = row[3]; d = row[4]; e = row[5];
                if (m == 2) {
                    b = b / c1;
                    c = c - b * d1;
                    d = d - b * e1;
                } else if (m >= 3) {
                    a = a / c2;
                    b = (b - a * d2) / c1;
                    c = c - b * d1 - a * e2;
                    if (m < nmax) d = d - b * e1;
                }
                if (FMODE == 1) { fa.set(r, m >= 3 ? a : 0.0); fb.set(r, m >= 2 ? b : 0.0); }
            }
            const double nb = -b, na = -a;
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const double unext = fblk[l][j];
                double rv;
                if (r == 2) rv = F1[l] * rb[2][1] + u0[l] * rb[2][2] + up[l] * rb[2][3];
                else if (r == 3) rv = F1[l] * rb[3][0] + um[l] * rb[3][1] + u0[l] * rb[3][2] + up[l] * rb[3][3];
                else if (r == n - 2) rv = um[l] * rt[0][1] + u0[l] * rt[0][2] + up[l] * rt[0][3] + FN[l] * rt[0][4];
                else if (r == n - 1) rv = um[l] * rt[1][1] + u0[l] * rt[1][2] + FN[l] * rt[1][3];
                else rv = um[l] * rhs(r, 1) + u0[l] * rhs(r, 2) + up[l];
                if (is_min && r == n - 1) bcs_far[l] = um[l] * rt[2][3] + u0[l] * rt[2][1] + FN[l] * rt[2][2];
                double y;
                if (m == 1) y = rv;
                else if (m == 2) y = rv + y1[l] * nb;
                else y = rv + y1[l] * nb + y2[l] * na;      // (FMODE 2 reads a = 0 at m <= 2 and b = 0 at m = 1)
                ysc[l].set(r, y);
                y2[l] = y1[l]; y1[l] = y;
                um[l] = u0[l]; u0[l] = up[l]; up[l] = unext;
            }
            if (FMODE != 2) {
                csc.set(r, 1.0 / c);
                dsc.set(r, -d);
                c2 = c1; c1 = c; d2 = d1; d1 = d; e2 = e1; e1 = e;
            }
        }
    }

    // ---- backward sweep
    double x1[NL], x2[NL];            // x(m+1), x(m+2)
#pragma unroll
    for (int l = 0; l < NL; l++) x1[l] = x2[l] = 0.0;
    double r2v[NL], r3v[NL], r4v[NL]; // results at rows 2, 3, 4 (near-end closure / derivative)
    // The fifth diagonal is not touched by the elimination, so -e is rebuilt from the shared tables instead of being
    // stored (row 2 of a BCS_MAX system is the one reduced row whose e changed).
    constexpr int BRB = 4;
    for (int m0 = nmax; m0 >= 1; m0 -= BRB) {
        double cblk[BRB], dblk[BRB], yblk[NL][BRB];
#pragma unroll
        for (int j = 0; j < BRB; j++) {
            const int r = m0 - j + 1;
            const bool ok = (m0 - j >= 1);
            cblk[j] = ok ? csc.get(r) : 0.0;
            dblk[j] = ok ? dsc.get(r) : 0.0;
#pragma unroll
            for (int l = 0; l < NL; l++) yblk[l][j] = ok ? ysc[l].get(r) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < BRB; j++) {
            const int m = m0 - j;
            if (m < 1) break;
            const int r = m + 1;
            const double ci = cblk[j], nd = dblk[j];
            const double ne = (!is_min && r == 2) ? -La[5] : -(P.L0[(r - 1) * 5 + 4] + lam * P.L1[(r - 1) * 5 + 4]);
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const double y = yblk[l][j];
                double x;
                if (m == nmax) x = y * ci;
                else if (m == nmax - 1) x = (y + x1[l] * nd) * ci;
                else x = (y + x1[l] * nd + x2[l] * ne) * ci;
                res[l].set(r, x);
                if (m == nmax - 2) {
                    // rows n-1, n-2, n-3 are known: far-end closure (BCS_MIN) or far-end derivative (BCS_MAX)
                    const double xn1 = x2[l], xn2 = x1[l], xn3 = x;
                    if (is_min) {
                        double v = bcs_far[l];
                        v = v + Le[2] * xn1;
                        v = v + Le[1] * xn2;
                        v = v + Le[5] * xn3;
                        res[l].set(n, v);
                    } else {
                        res[l].set(n, bc[l]);
                        if (du) {
                            double row[6];
                            Lrow(n, row);
                            double v = row[3] * bc[l];
                            v = v + row[2] * xn1;
                            v = v + row[1] * xn2;
                            v = v + row[5] * xn3;
                            v = v + rhs(n, 1) * F(l, n - 1);
                            du[l] = v;
                        }
                    }
                }
                if (r == 4) r4v[l] = x;
                if (r == 3) r3v[l] = x;
                if (r == 2) r2v[l] = x;
                x2[l] = x1[l]; x1[l] = x;
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; l++) {
        if (is_min) {
            res[l].set(1, bc[l]);
            if (du) {
                double row[6];
                Lrow(1, row);
                double v = row[3] * bc[l];
                v = v + row[4] * r2v[l];
                v = v + row[5] * r3v[l];
                v = v + row[1] * r4v[l];
                v = v + rhs(1, 3) * F(l, 2);
                du[l] = v;
            }
        } else {
            double v = bcs_far[l];
            v = v + Le[4] * r2v[l];
            v = v + Le[5] * r3v[l];
            v = v + Le[1] * r4v[l];
            res[l].set(1, v);
        }
    }
}

template <int MINB, bool FAC, int NL>
__global__ void poisson_modes_kernel(PoissonDev D, double* __restrict__ cf, double* __restrict__ cv);

struct ModeGeom {
    int nxh, ny, nz;          // half-spectrum extent in x, lines in y, modes in z (local)
    long long nmodes;         // nxh * nz
};

__device__ __forceinline__ bool mode_is_singular(const PoissonDev& D, int i, int k) {
    return (i == D.i_sing0 || i == D.i_sing1) && (k == D.k_sing0 || k == D.k_sing1);
}

// per-mode planes (fundamental solutions, scratch, LU factors), blocked by 32 modes so that a warp marching along y
// streams through contiguous memory.  Planes that are read together are interleaved row by row inside a block (D.il):
// the five fundamental lines, the factor lines in the pairs a sweep reads, the two scratch lines of the components --
// element (row, k, m) of a group of g planes at (((m / 32) * ny + row) * g + k) * 32 + m % 32.  A warp then follows three
// address streams per sweep instead of five to seven (fewer pages and DRAM rows open at a time).
enum { P_FUND = 0, P_SCR = 1, P_FAC = 2 };
__device__ __forceinline__ LineRef plane(const PoissonDev& D, int which, int k, long long m) {
    double* base = (which == P_FUND) ? D.fund : (which == P_SCR ? D.scr : D.fac);
    int g0, gn;
    if (!D.il) { g0 = k; gn = 1; }
    else if (which == P_FUND) { g0 = 0; gn = 5; }
    else if (which == P_FAC) { g0 = k & ~1; gn = 2; }
    else if (k < 2) { g0 = 0; gn = 2; }
    else { g0 = k; gn = 1; }
    LineRef r;
    r.p = base + (long long)g0 * D.plane_sz + ((m >> 5) * ((long long)D.ny * gn) + (k - g0)) * 32 + (m & 31);
    r.js = 32LL * gn;
    return r;
}

// -------------------------------------------------------------------------------------------------
// initialisation: fundamental solutions and the 3x3 boundary system of every regular mode
__global__ void poisson_fundamental_kernel(PoissonDev D) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= D.nmodes) return;
    const int i = (int)(m % D.nxh), k = (int)(m / D.nxh);
    if (mode_is_singular(D, i, k)) return;
    const double lam = sqrt(D.lambda[m]);
    const long long NM = D.nmodes;
    const long long plane_sz = D.plane_sz;
    const int n = D.ny;
    // stage 1: v1 (forcing delta at row n, v1(1) = 0) and e- (no forcing, e-(1) = 1)
    {
        LineRef f[2] = {{nullptr, 0}, {nullptr, 0}};
        double fend[2] = {1.0, 0.0}, bc[2] = {0.0, 1.0};
        LineRef res[2] = {plane(D, P_FUND, 0, m), plane(D, P_FUND, 1, m)};
        LineRef ysc[2] = {plane(D, P_SCR, 0, m), plane(D, P_SCR, 1, m)};
        if (D.fac)
            int1_solve<2, 1>(D.smin, lam, f, 1.0, fend, bc, res, ysc, plane(D, P_FAC, 2, m),
                             plane(D, P_FAC, 3, m), plane(D, P_SCR, 5, m), nullptr,
                             plane(D, P_FAC, 0, m), plane(D, P_FAC, 1, m));
        else
            int1_solve<2>(D.smin, lam, f, 1.0, fend, bc, res, ysc, plane(D, P_SCR, 3, m),
                          plane(D, P_SCR, 4, m), plane(D, P_SCR, 5, m), nullptr);
    }
    // stage 2: u1, s+, e+ from (v1, e-, 0) with values (0, 0, 1) at row n
    double der[3];
    {
        LineRef v1 = plane(D, P_FUND, 0, m), em = plane(D, P_FUND, 1, m);
        LineRef f[3] = {v1, em, {nullptr, 0}};
        double fend[3] = {v1.get(1), em.get(1), 0.0}, bc[3] = {0.0, 0.0, 1.0};
        LineRef res[3] = {plane(D, P_FUND, 2, m), plane(D, P_FUND, 3, m),
                          plane(D, P_FUND, 4, m)};
        LineRef ysc[3] = {plane(D, P_SCR, 0, m), plane(D, P_SCR, 1, m),
                          plane(D, P_SCR, 2, m)};
        if (D.fac)
            int1_solve<3, 1>(D.smax, -lam, f, 1.0, fend, bc, res, ysc, plane(D, P_FAC, 6, m),
                             plane(D, P_FAC, 7, m), plane(D, P_SCR, 5, m), der,
                             plane(D, P_FAC, 4, m), plane(D, P_FAC, 5, m));
        else
            int1_solve<3>(D.smax, -lam, f, 1.0, fend, bc, res, ysc, plane(D, P_SCR, 3, m),
                          plane(D, P_SCR, 4, m), plane(D, P_SCR, 5, m), der);
    }
    // boundary system (opr_odes.f90:329-348), stored LU-decomposed
    const double v1n = plane(D, P_FUND, 0, m).get(n), emn = plane(D, P_FUND, 1, m).get(n);
    const double u11 = plane(D, P_FUND, 2, m).get(1), sp1 = plane(D, P_FUND, 3, m).get(1);
    const double ep1 = plane(D, P_FUND, 4, m).get(1);
    double a11 = 1.0 + lam * sp1, a21 = emn, a31 = der[1];
    double a12 = lam * ep1, a22 = lam, a32 = der[2];
    double a13 = lam * u11, a23 = v1n, a33 = der[0];
    a12 = a12 / a11;
    a22 = a22 - a21 * a12;
    a32 = a32 - a31 * a12;
    a13 = a13 / a11;
    a23 = (a23 - a21 * a13) / a22;
    a33 = a33 - a31 * a13 - a32 * a23;
    double* A = D.amat;
    A[0 * NM + m] = a11; A[1 * NM + m] = a21; A[2 * NM + m] = a31;
    A[3 * NM + m] = a12; A[4 * NM + m] = a22; A[5 * NM + m] = a32;
    A[6 * NM + m] = a13; A[7 * NM + m] = a23; A[8 * NM + m] = a33;
}

// -------------------------------------------------------------------------------------------------
// per call: regular modes, Neumann/Neumann (OPR_ODE2_Factorize_NN)
__device__ void poisson_singular_mode(const PoissonDev& D, double* __restrict__ cf, double* __restrict__ cv, int i, int k, long long slot);

// NL = 2: one thread per mode, both components (real, imaginary) of the mode in the same thread, which shares the factor
// and fundamental-solution loads.  NL = 1: one thread per component (adjacent lanes = re, im of a mode: unit-stride
// accesses of the complex lines and twice the threads) -- for few modes per GPU (kx-split stage of a split domain),
// where the per-mode marches along y are latency-bound and there are not enough of them to fill the SMs.
template <int MINB, bool FAC, int NL>
__global__ void __launch_bounds__(128, MINB) poisson_modes_kernel(PoissonDev D, double* __restrict__ cf, double* __restrict__ cv) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long m = (NL == 1) ? (h >> 1) : h;
    const int l0 = (NL == 1) ? (int)(h & 1) : 0;
    if (m >= D.nmodes) return;
    const int i = (int)(m % D.nxh), k = (int)(m / D.nxh);
    if (mode_is_singular(D, i, k)) { if (l0 == 0) poisson_singular_mode(D, cf, cv, i, k, m); return; }
    const double lam = sqrt(D.lambda[m]);
    const long long NM = D.nmodes;
    const long long plane_sz = D.plane_sz;
    const int n = D.ny;
    // complex lines of this mode inside c(kx, y, kz): re/im interleaved
    const long long off = 2 * ((long long)i + (long long)D.nxh * D.ny * k);
    const long long js = 2LL * D.nxh;
    LineRef fl[NL], vl[NL], ysc[NL];
    double bcb[NL], bct[NL], zero[NL];
    const double norm = D.norm;
#pragma unroll
    for (int l = 0; l < NL; l++) {
        fl[l] = LineRef{cf + off + l0 + l, js};
        vl[l] = LineRef{cv + off + l0 + l, js};
        ysc[l] = plane(D, P_SCR, l0 + l, m);
        bcb[l] = fl[l].get(1) * norm;      // bcs(1:2,1) = f(1:2)
        bct[l] = fl[l].get(n) * norm;      // bcs(1:2,2) = f(2ny-1:2ny)
        zero[l] = 0.0;
    }
    LineRef csc = plane(D, P_SCR, 2, m), dsc = plane(D, P_SCR, 3, m),
            esc = plane(D, P_SCR, 4, m);
    // v^(0): v' + lam v = f, f(n) = 0, v(1) = 0
    {
        if (FAC)
            int1_solve<NL, 2>(D.smin, lam, fl, norm, zero, zero, vl, ysc, plane(D, P_FAC, 2, m),
                              plane(D, P_FAC, 3, m), esc, nullptr, plane(D, P_FAC, 0, m),
                              plane(D, P_FAC, 1, m));
        else
            int1_solve<NL>(D.smin, lam, fl, norm, zero, zero, vl, ysc, csc, dsc, esc, nullptr);
    }
    // u^(0): u' - lam u = v, u(n) = 0  (written over the forcing, which is no longer needed)
    double du0[NL];
    {
        double fend[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) fend[l] = vl[l].get(1);
        if (FAC)
            int1_solve<NL, 2>(D.smax, -lam, vl, 1.0, fend, zero, fl, ysc, plane(D, P_FAC, 6, m),
                              plane(D, P_FAC, 7, m), esc, du0, plane(D, P_FAC, 4, m),
                              plane(D, P_FAC, 5, m));
        else
            int1_solve<NL>(D.smax, -lam, vl, 1.0, fend, zero, fl, ysc, csc, dsc, esc, du0);
    }
    // constraint and boundary conditions (opr_odes.f90:350-367)
    const double* A = D.amat;
    const double a11 = A[0 * NM + m], a21 = A[1 * NM + m], a31 = A[2 * NM + m];
    const double a12 = A[3 * NM + m], a22 = A[4 * NM + m], a32 = A[5 * NM + m];
    const double a13 = A[6 * NM + m], a23 = A[7 * NM + m], a33 = A[8 * NM + m];
    LineRef v1 = plane(D, P_FUND, 0, m), em = plane(D, P_FUND, 1, m);
    LineRef u1 = plane(D, P_FUND, 2, m), sp = plane(D, P_FUND, 3, m),
            ep = plane(D, P_FUND, 4, m);
    const LineRef (&ul)[NL] = fl;
    double fn[NL], v_1[NL], u_n[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        v_1[l] = (bcb[l] - lam * ul[l].get(1)) / a11;
        u_n[l] = (bct[l] - vl[l].get(n) - a21 * v_1[l]) / a22;
        fn[l] = (bct[l] - du0[l] - a31 * v_1[l] - a32 * u_n[l]) / a33;
        u_n[l] = u_n[l] - a23 * fn[l];
        v_1[l] = v_1[l] - a12 * u_n[l] - a13 * fn[l];
    }
    // rows n, n-1 .. 2, 1: the five fundamental lines are read once for both components, CRB rows per batch
    constexpr int CRB = 4;
    for (int r0 = n; r0 >= 1; r0 -= CRB) {
        double fu[5][CRB], uu0[NL][CRB], vv0[NL][CRB];
#pragma unroll
        for (int j = 0; j < CRB; j++) {
            const int r = r0 - j;
            const bool ok = (r >= 1);
            fu[0][j] = ok ? v1.get(r) : 0.0;
            fu[1][j] = ok ? em.get(r) : 0.0;
            fu[2][j] = (ok && r < n) ? u1.get(r) : 0.0;
            fu[3][j] = (ok && r < n) ? sp.get(r) : 0.0;
            fu[4][j] = (ok && r < n) ? ep.get(r) : 0.0;
#pragma unroll
            for (int l = 0; l < NL; l++) {
                uu0[l][j] = (ok && r < n) ? ul[l].get(r) : 0.0;
                vv0[l][j] = (ok && r > 1) ? vl[l].get(r) : 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < CRB; j++) {
            const int r = r0 - j;
            if (r < 1) break;
#pragma unroll
            for (int l = 0; l < NL; l++) {
                double uu, vv;
                if (r == n) {
                    uu = u_n[l];                 // u(:, nx) has been replaced by the boundary value
                    vv = vv0[l][j] + fn[l] * fu[0][j] + v_1[l] * fu[1][j] + lam * uu;
                } else if (r == 1) {
                    uu = uu0[l][j] + fn[l] * fu[2][j] + v_1[l] * fu[3][j] + u_n[l] * fu[4][j];
                    vv = v_1[l] + lam * uu;
                } else {
                    uu = uu0[l][j] + fn[l] * fu[2][j] + v_1[l] * fu[3][j] + u_n[l] * fu[4][j];
                    vv = vv0[l][j] + fn[l] * fu[0][j] + v_1[l] * fu[1][j] + lam * uu;
                }
                ul[l].set(r, uu);
                vl[l].set(r, vv);
            }
        }
    }
}

// per call: the (up to four) singular modes, OPR_ODE2_Factorize_NN_Sing -> _DN_Sing
// `slot` selects the per-mode planes (the mode index for the full planes, 0..3 for the small planes of the warp path)
// the lines a singular mode works on: forcing / p^ (in place), dp^/dy, and nine scratch lines
struct SingLines { LineRef fre, fim, vre, vim, ysc0, ysc1, csc, dsc, esc, v1, u1; };

__device__ void poisson_singular_solve(const PoissonDev& D, double lam, const SingLines& Ln) {
    const int n = D.ny;
    const LineRef fre = Ln.fre, fim = Ln.fim, vre = Ln.vre, vim = Ln.vim;
    const double norm = D.norm;
    const double bct[2] = {fre.get(n) * norm, fim.get(n) * norm};
    LineRef ysc[2] = {Ln.ysc0, Ln.ysc1};
    LineRef csc = Ln.csc, dsc = Ln.dsc, esc = Ln.esc;
    LineRef v1 = Ln.v1, u1 = Ln.u1;
    const double zero2[2] = {0.0, 0.0};
    // v^(0): v' = f with f(1) = 0, v(n) = bcs(:,2)
    {
        LineRef f[2] = {fre, fim}, res[2] = {vre, vim};
        int1_solve<2>(D.smax, -lam, f, norm, zero2, bct, res, ysc, csc, dsc, esc, nullptr);
    }
    // v^(1): forcing delta at row 1, v1(n) = 0
    {
        LineRef f[1] = {{nullptr, 0}}, res[1] = {v1}, ys[1] = {ysc[0]};
        const double fend[1] = {1.0}, bc[1] = {0.0};
        int1_solve<1>(D.smax, -lam, f, 1.0, fend, bc, res, ys, csc, dsc, esc, nullptr);
    }
    // u^(0): u' = v, u(1) = 0
    double du0[2], du1[1];
    {
        LineRef f[2] = {vre, vim}, res[2] = {fre, fim};
        const double fend[2] = {vre.get(n), vim.get(n)};
        int1_solve<2>(D.smin, lam, f, 1.0, fend, zero2, res, ysc, csc, dsc, esc, du0);
    }
    {
        LineRef f[1] = {v1}, res[1] = {u1}, ys[1] = {ysc[0]};
        const double fend[1] = {v1.get(n)}, bc[1] = {0.0};
        int1_solve<1>(D.smin, lam, f, 1.0, fend, bc, res, ys, csc, dsc, esc, du1);
    }
    const double ff = 1.0 / (du1[0] - v1.get(1));
    LineRef ul[2] = {fre, fim}, vl[2] = {vre, vim};
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const double cdu = (vl[l].get(1) - du0[l]) * ff;
        for (int r = 1; r <= n; r++) {
            ul[l].set(r, ul[l].get(r) + cdu * u1.get(r));
            vl[l].set(r, vl[l].get(r) + cdu * v1.get(r));
        }
    }
}

// Singular mode on lines in shared memory with everything that depends on lambda only kept from the first call (`cache`, global:
// ten lines of n doubles -- lower and upper LU factors of both systems, the fundamental lines v^(1), u^(1) -- and three scalars):
// later calls run two factor-free solves (no divisions, no table rows) instead of four factorising ones.  Leaves v^(0) and u^(0)
// in the lines and returns the weights of v^(1), u^(1); the caller adds them (in parallel).
__device__ void poisson_singular_cached(const PoissonDev& D, double lam, const SingLines& Ln, double* __restrict__ cache, double (&cdu)[2]) {
    const int n = D.ny;
    const LineRef famax = {cache, 1}, fbmax = {cache + n, 1}, cmax = {cache + 2 * n, 1}, dmax = {cache + 3 * n, 1};
    const LineRef famin = {cache + 4 * n, 1}, fbmin = {cache + 5 * n, 1}, cmin = {cache + 6 * n, 1}, dmin = {cache + 7 * n, 1};
    const LineRef v1 = {cache + 8 * n, 1}, u1 = {cache + 9 * n, 1};
    double* sc = cache + 10 * n;                  // [0] filled flag, [1] du1, [2] v1(1)
    const LineRef fre = Ln.fre, fim = Ln.fim, vre = Ln.vre, vim = Ln.vim;
    LineRef ysc[2] = {Ln.ysc0, Ln.ysc1};
    const double norm = D.norm;
    const double bct[2] = {fre.get(n) * norm, fim.get(n) * norm};
    const double zero2[2] = {0.0, 0.0};
    if (sc[0] == 0.0) {
        {   // v^(1): forcing delta at row 1, v1(n) = 0; keeps the factors of the BCS_MAX system
            LineRef f[1] = {{nullptr, 0}}, res[1] = {v1}, ys[1] = {ysc[0]};
            const double fend[1] = {1.0}, bc[1] = {0.0};
            int1_solve<1, 1>(D.smax, -lam, f, 1.0, fend, bc, res, ys, cmax, dmax, Ln.esc, nullptr, famax, fbmax);
        }
        {   // u^(1): u' = v1, u1(1) = 0; keeps the factors of the BCS_MIN system
            LineRef f[1] = {v1}, res[1] = {u1}, ys[1] = {ysc[0]};
            const double fend[1] = {v1.get(n)}, bc[1] = {0.0};
            double du1[1];
            int1_solve<1, 1>(D.smin, lam, f, 1.0, fend, bc, res, ys, cmin, dmin, Ln.esc, du1, famin, fbmin);
            sc[1] = du1[0];
            sc[2] = v1.get(1);
        }
        __threadfence();
        sc[0] = 1.0;
    }
    {   // v^(0): v' = f with f(1) = 0, v(n) = bcs(:,2)
        LineRef f[2] = {fre, fim}, res[2] = {vre, vim};
        int1_solve<2, 2>(D.smax, -lam, f, norm, zero2, bct, res, ysc, cmax, dmax, Ln.esc, nullptr, famax, fbmax);
    }
    double du0[2];
    {   // u^(0): u' = v, u(1) = 0
        LineRef f[2] = {vre, vim}, res[2] = {fre, fim};
        const double fend[2] = {vre.get(n), vim.get(n)};
        int1_solve<2, 2>(D.smin, lam, f, 1.0, fend, zero2, res, ysc, cmin, dmin, Ln.esc, du0, famin, fbmin);
    }
    const double ff = 1.0 / (sc[1] - sc[2]);
    cdu[0] = (vre.get(1) - du0[0]) * ff;
    cdu[1] = (vim.get(1) - du0[1]) * ff;
}

// the same on the mode's lines in global memory (thread-per-mode kernels)
__device__ void poisson_singular_mode(const PoissonDev& D, double* __restrict__ cf, double* __restrict__ cv, int i, int k, long long slot) {
    const double lam = sqrt(D.lambda[(long long)i + (long long)D.nxh * k]);
    const long long m = slot;
    const long long off = 2 * ((long long)i + (long long)D.nxh * D.ny * k);
    const long long js = 2LL * D.nxh;
    SingLines Ln;
    Ln.fre = {cf + off, js}; Ln.fim = {cf + off + 1, js};
    Ln.vre = {cv + off, js}; Ln.vim = {cv + off + 1, js};
    Ln.ysc0 = plane(D, P_SCR, 0, m); Ln.ysc1 = plane(D, P_SCR, 1, m);
    Ln.csc = plane(D, P_SCR, 2, m); Ln.dsc = plane(D, P_SCR, 3, m); Ln.esc = plane(D, P_SCR, 4, m);
    // fundamental lines of this mode live in the (otherwise unused) fund planes of the mode
    Ln.v1 = plane(D, P_FUND, 0, m); Ln.u1 = plane(D, P_FUND, 2, m);
    poisson_singular_solve(D, lam, Ln);
}

// -------------------------------------------------------------------------------------------------
// Team-per-mode y solves (ny = 8 T, T <= 128).  A team of NW = 1, 2 or 4 warps owns one (kx, kz) mode, one lane one chunk of
// 8 rows, the lines of both components (real, imaginary) live in registers.  The substitution stages of PENTADSS are
// second-order linear recurrences,
//     forward   y_r = f_r - b_r y_{r-1} - a_r y_{r-2},      backward  x_r = (y_r - d_r x_{r+1} - e_r x_{r+2}) / c_r,
// so a chunk maps its two inflow values to its two outflow values by an affine 2x2 map.  Every lane sweeps its chunk once
// with zero inflow (end values and the homogeneous map only), the true inflow of all chunks follows from a scan of the maps
// (Kogge-Stone over the lanes of a warp, 5 shuffle steps, then the warp totals through shared memory; unlike the
// tridiagonal line kernels the maps do not decay for small wavenumbers, so no window), and a second sweep produces the rows.
// Per (mode, row) the kernel reads the forcing (2 doubles), the upper LU factors 1/c, -d of both systems (4; the lower
// factors a, b are rebuilt from them and the shared tables, the fifth diagonal from the tables) and the five fundamental
// lines, and writes p^ and dp^/dy (4): 15 doubles against the 37 of the thread-per-mode kernel, which parks the
// intermediate vector in global memory and re-reads every line per sweep.
// The complex lines are strided by nx/2+1 in memory: a CTA of MW teams takes MW adjacent kx and moves the (ny x MW) tile
// through shared memory with row segments of MW * 16 bytes on the global side.
constexpr int TR = 8;                          // rows per lane
__host__ __device__ inline int tpad(int r0) { return (r0 >> 3) * 9 + (r0 & 7); }      // 0-based row -> tile slot (conflict-free chunk reads)

__device__ __forceinline__ double2 operator*(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 fma2(double s, double2 a, double2 c) { return make_double2(fma(s, a.x, c.x), fma(s, a.y, c.y)); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_dn_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double2 shfl_up_2(double2 v, int d) { return make_double2(shfl_up_d(v.x, d), shfl_up_d(v.y, d)); }
__device__ __forceinline__ double2 shfl_dn_2(double2 v, int d) { return make_double2(shfl_dn_d(v.x, d), shfl_dn_d(v.y, d)); }

// affine map of a chunk: (s1, s2) -> (m11 s1 + m12 s2 + p1, m21 s1 + m22 s2 + p2); s1 is the value next to the chunk
struct Aff { double m11, m12, m21, m22; double2 p1, p2; };

__device__ __forceinline__ void prefetch_l2g(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NW>
__device__ __forceinline__ void team_barrier(int id) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NW * 32) : "memory");
}

// Scan over the chunks of a team; DIR = +1: chunk j follows chunk j-1 (forward sweep), DIR = -1: chunk j follows chunk j+1.
// Returns the inflow of this lane's chunk, i.e. the outflow of the composition of all chunks before it (zero state first).
// xch: NW slots of shared memory of this team and this scan.
template <int DIR, int NW>
__device__ __forceinline__ void team_scan(Aff a, int lane, int wt, Aff* __restrict__ xch, int barid, double2& in1, double2& in2) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Aff e;      // the earlier group
        if (DIR > 0) {
            e.m11 = shfl_up_d(a.m11, off); e.m12 = shfl_up_d(a.m12, off); e.m21 = shfl_up_d(a.m21, off); e.m22 = shfl_up_d(a.m22, off);
            e.p1 = shfl_up_2(a.p1, off); e.p2 = shfl_up_2(a.p2, off);
        } else {
            e.m11 = shfl_dn_d(a.m11, off); e.m12 = shfl_dn_d(a.m12, off); e.m21 = shfl_dn_d(a.m21, off); e.m22 = shfl_dn_d(a.m22, off);
            e.p1 = shfl_dn_2(a.p1, off); e.p2 = shfl_dn_2(a.p2, off);
        }
        const bool has = (DIR > 0) ? (lane >= off) : (lane + off < 32);
        if (has) {      // a := a o e
            Aff r;
            r.m11 = a.m11 * e.m11 + a.m12 * e.m21; r.m12 = a.m11 * e.m12 + a.m12 * e.m22;
            r.m21 = a.m21 * e.m11 + a.m22 * e.m21; r.m22 = a.m21 * e.m12 + a.m22 * e.m22;
            r.p1 = fma2(a.m11, e.p1, fma2(a.m12, e.p2, a.p1));
            r.p2 = fma2(a.m21, e.p1, fma2(a.m22, e.p2, a.p2));
            a = r;
        }
    }
    double2 c1 = make_double2(0.0, 0.0), c2 = c1;          // state entering this warp
    double2 o1 = a.p1, o2 = a.p2;                            // state leaving this lane's chunk
    if (NW > 1) {
        if (lane == (DIR > 0 ? 31 : 0)) xch[wt] = a;
        team_barrier<NW>(barid);
#pragma unroll
        for (int s = 0; s < NW - 1; s++) {
            const int wq = (DIR > 0) ? s : NW - 1 - s;
            const bool before = (DIR > 0) ? (wq < wt) : (wq > wt);
            if (before) {
                const Aff e = xch[wq];
                const double2 n1 = fma2(e.m11, c1, fma2(e.m12, c2, e.p1));
                const double2 n2 = fma2(e.m21, c1, fma2(e.m22, c2, e.p2));
                c1 = n1; c2 = n2;
            }
        }
        o1 = fma2(a.m11, c1, fma2(a.m12, c2, a.p1));
        o2 = fma2(a.m21, c1, fma2(a.m22, c2, a.p2));
    }
    if (DIR > 0) {
        in1 = shfl_up_2(o1, 1); in2 = shfl_up_2(o2, 1);
        if (lane == 0) { in1 = c1; in2 = c2; }
    } else {
        in1 = shfl_dn_2(o1, 1); in2 = shfl_dn_2(o2, 1);
        if (lane == 31) { in1 = c1; in2 = c2; }
    }
}

// One first-order integral problem u' + lam u = f of one mode, both components, on a team.
//   X[q]     in: forcing (already scaled) at row r0 + q, r0 = 8 j + 1 (j = chunk); out: result rows (rows 1 and n included)
//   Fm1, Fp1 the forcing rows before / after the chunk
//   F1, FN   the values used instead of rows 1 and n of the forcing (see int1_solve), bc the imposed value
//   fac      upper LU factors (1/c, -d) of this mode and side in lane order
//   cs       64 doubles of shared memory of this team: the reduced boundary rows (they depend on lambda; every lane would
//            otherwise hold ~50 doubles of them in registers although only the first and the last lane use them)
//   xch      2 * NW scan slots of this team;  du: derivative at row n (BCS_MAX only; valid in the last lane)
enum { CS_LE = 0, CS_LA = 5, CS_LB = 10, CS_RB = 16, CS_RT = 28 };       // Le[1..5], La[1..5], Lb[1..5], rb[1..3][0..3], rt[0..2][1..4]
template <bool IS_MIN, int NW>
__device__ __forceinline__ void team_int1(const Int1Dev& P, const WarpSide& W, const double2* __restrict__ fac, double lam, int T, int j,
                                          int lane, int wt, int barid, double2 (&X)[TR], double2 Fm1, double2 Fp1, double2 F1, double2 FN,
                                          double2 bc, double* __restrict__ cs, Aff* __restrict__ xch, double2* du) {
    const int n = P.n;
    const bool active = j < T, first = (j == 0), last = (j == T - 1);
    const double2 zero = make_double2(0.0, 0.0);
    auto Lrow = [&](int r, double (&row)[6]) {
#pragma unroll
        for (int k = 1; k <= 5; k++) row[k] = __ldg(&P.L0[(r - 1) * 5 + k - 1]) + lam * __ldg(&P.L1[(r - 1) * 5 + k - 1]);
    };
    auto rhs = [&](int r, int k) { return __ldg(&P.rhs[(r - 1) * 3 + k - 1]); };
    // ---- reduction of the far-end row into its neighbours (as in int1_solve), by the first lane into shared memory
    if (first) {
        double Le[6], La[6], Lb[6];
        double rb[4][4], rt[3][5];
#pragma unroll
        for (int r = 1; r <= 3; r++)
#pragma unroll
            for (int c = 0; c <= 3; c++) rb[r][c] = P.rb[r][c];
#pragma unroll
        for (int r = 0; r <= 2; r++)
#pragma unroll
            for (int c = 1; c <= 4; c++) rt[r][c] = P.rt[r][c];
        if (IS_MIN) {
            Lrow(n, Le); Lrow(n - 1, La); Lrow(n - 2, Lb);
            const double dummy = 1.0 / Le[3];
#pragma unroll
            for (int k = 1; k <= 5; k++) Le[k] = -Le[k] * dummy;
            Le[3] = 1.0;
            La[1] = La[1] + La[4] * Le[5]; La[2] = La[2] + La[4] * Le[1]; La[3] = La[3] + La[4] * Le[2];
            Lb[2] = Lb[2] + Lb[5] * Le[5]; Lb[3] = Lb[3] + Lb[5] * Le[1]; Lb[4] = Lb[4] + Lb[5] * Le[2];
#pragma unroll
            for (int r = 0; r <= 2; r++)
#pragma unroll
                for (int c = 1; c <= 3; c++) rt[r][c] = rhs(n - 2 + r, c);
#pragma unroll
            for (int c = 1; c <= 3; c++) rt[2][c] = rt[2][c] * dummy;
            rt[1][1] = rt[1][1] - La[4] * rt[2][3]; rt[1][2] = rt[1][2] - La[4] * rt[2][1]; rt[1][3] = rt[1][3] - La[4] * rt[2][2];
            rt[0][2] = rt[0][2] - Lb[5] * rt[2][3]; rt[0][3] = rt[0][3] - Lb[5] * rt[2][1]; rt[0][4] = rt[0][4] - Lb[5] * rt[2][2];
        } else {
            Lrow(1, Le); Lrow(2, La); Lrow(3, Lb);
            const double dummy = 1.0 / Le[3];
#pragma unroll
            for (int k = 1; k <= 5; k++) Le[k] = -Le[k] * dummy;
            Le[3] = 1.0;
            La[3] = La[3] + La[2] * Le[4]; La[4] = La[4] + La[2] * Le[5]; La[5] = La[5] + La[2] * Le[1];
            Lb[2] = Lb[2] + Lb[1] * Le[4]; Lb[3] = Lb[3] + Lb[1] * Le[5]; Lb[4] = Lb[4] + Lb[1] * Le[1];
#pragma unroll
            for (int r = 1; r <= 3; r++)
#pragma unroll
                for (int c = 1; c <= 3; c++) rb[r][c] = rhs(r, c);
#pragma unroll
            for (int c = 1; c <= 3; c++) rb[1][c] = rb[1][c] * dummy;
            rb[2][1] = rb[2][1] - La[2] * rb[1][2]; rb[2][2] = rb[2][2] - La[2] * rb[1][3]; rb[2][3] = rb[2][3] - La[2] * rb[1][1];
            rb[3][0] = rb[3][0] - Lb[1] * rb[1][2]; rb[3][1] = rb[3][1] - Lb[1] * rb[1][3]; rb[3][2] = rb[3][2] - Lb[1] * rb[1][1];
        }
#pragma unroll
        for (int k = 1; k <= 5; k++) { cs[CS_LE + k - 1] = Le[k]; cs[CS_LA + k - 1] = La[k]; cs[CS_LB + k - 1] = Lb[k]; }
#pragma unroll
        for (int r = 1; r <= 3; r++)
#pragma unroll
            for (int c = 0; c <= 3; c++) cs[CS_RB + (r - 1) * 4 + c] = rb[r][c];
#pragma unroll
        for (int r = 0; r <= 2; r++)
#pragma unroll
            for (int c = 1; c <= 4; c++) cs[CS_RT + r * 4 + c - 1] = rt[r][c];
    }
    team_barrier<NW>(barid);
    auto Le = [&](int k) { return cs[CS_LE + k - 1]; };
    auto La = [&](int k) { return cs[CS_LA + k - 1]; };
    auto Lb = [&](int k) { return cs[CS_LB + k - 1]; };
    auto rb = [&](int r, int c) { return cs[CS_RB + (r - 1) * 4 + c]; };
    auto rt = [&](int r, int c) { return cs[CS_RT + r * 4 + c - 1]; };

    // ---- right-hand side (MatMul_3d with the reduced boundary rows) in place; rows 1 and n are not part of the system
    double2 bcs_far = zero, fn1 = zero;                       // fn1 = forcing at row n-1 (for du)
    {
        double2 s1 = zero, s2 = zero, s5 = zero, s6 = zero;
        if (first) {
            if (!IS_MIN) bcs_far = F1 * rb(1, 2) + X[1] * rb(1, 3) + X[2] * rb(1, 1);
            s1 = F1 * rb(2, 1) + X[1] * rb(2, 2) + X[2] * rb(2, 3);
            s2 = F1 * rb(3, 0) + X[1] * rb(3, 1) + X[2] * rb(3, 2) + X[3] * rb(3, 3);
        }
        if (last) {       // rows n-2 (q = 5), n-1 (q = 6), n (q = 7)
            if (IS_MIN) bcs_far = X[5] * rt(2, 3) + X[6] * rt(2, 1) + FN * rt(2, 2);
            s5 = X[4] * rt(0, 1) + X[5] * rt(0, 2) + X[6] * rt(0, 3) + FN * rt(0, 4);
            s6 = X[5] * rt(1, 1) + X[6] * rt(1, 2) + FN * rt(1, 3);
            fn1 = X[6];
        }
        double2 prev = Fm1;
#pragma unroll
        for (int q = 0; q < TR; q++) {
            const double2 cur = X[q], up = (q < TR - 1) ? X[q < TR - 1 ? q + 1 : q] : Fp1;
            const double2 c = active ? __ldg(W.rh + q * T + j) : zero;
            X[q] = prev * c.x + cur * c.y + up;
            prev = cur;
        }
        if (first) { X[0] = zero; X[1] = s1; X[2] = s2; }
        if (last) { X[5] = s5; X[6] = s6; X[7] = zero; }
        if (!active) {
#pragma unroll
            for (int q = 0; q < TR; q++) X[q] = zero;
        }
    }

    // ---- forward sweep.  The lower factors -a, -b of a row are rebuilt from the stored upper factors (1/c, -d) of the two
    // previous rows and the table rows (PENTADFS: a = A / c(r-2), b = (B - a d(r-2)) / c(r-1)).
    {
        double na[TR], nb[TR];
        double2 f1 = (active && j > 0) ? __ldg(fac + (TR - 1) * T + j - 1) : zero;      // factors of rows r0 - 1, r0 - 2
        double2 f2 = (active && j > 0) ? __ldg(fac + (TR - 2) * T + j - 1) : zero;
        Aff a;
        double2 p1 = zero, p2 = zero;
        double h11 = 1.0, h12 = 0.0, h21 = 0.0, h22 = 1.0;       // (y_{q-1}, y_{q-2}) as functions of the inflow (s1, s2)
#pragma unroll
        for (int q = 0; q < TR; q++) {
            const double2 ta = active ? __ldg(W.ab0 + q * T + j) : zero, tb = active ? __ldg(W.ab1 + q * T + j) : zero;
            double Ar = ta.x + lam * ta.y, Br = tb.x + lam * tb.y;
            if (IS_MIN && last && q == 5) { Ar = Lb(1); Br = Lb(2); }        // rows n-2, n-1: the reduced rows Lb, La
            if (IS_MIN && last && q == 6) { Ar = La(1); Br = La(2); }
            if (!IS_MIN && first && q == 2) Br = Lb(2);                        // row 3 of a BCS_MAX system: reduced row Lb
            double av = Ar * f2.x;
            if (first && q <= 2) av = 0.0;                                     // rows 2, 3 (m = 1, 2): no second sub-diagonal
            double bv = (Br + av * f2.y) * f1.x;
            if ((first && q <= 1) || (last && q == TR - 1)) { av = 0.0; bv = 0.0; }     // rows 1, n (not in the system), row 2 (m = 1)
            na[q] = -av; nb[q] = -bv;
            f2 = f1; f1 = active ? __ldg(fac + q * T + j) : zero;
            const double2 y = fma2(nb[q], p1, fma2(na[q], p2, X[q]));
            const double g1 = nb[q] * h11 + na[q] * h21, g2 = nb[q] * h12 + na[q] * h22;
            p2 = p1; p1 = y;
            h21 = h11; h22 = h12; h11 = g1; h12 = g2;
        }
        a.m11 = h11; a.m12 = h12; a.m21 = h21; a.m22 = h22; a.p1 = p1; a.p2 = p2;
        double2 s1, s2;
        team_scan<+1, NW>(a, lane, wt, xch, barid, s1, s2);
#pragma unroll
        for (int q = 0; q < TR; q++) {
            const double2 y = fma2(nb[q], s1, fma2(na[q], s2, X[q]));
            X[q] = y;
            s2 = s1; s1 = y;
        }
    }

    // ---- backward sweep.  The fifth diagonal is untouched by the elimination: -e from the tables (row 2 of a BCS_MAX system
    // is the one reduced row whose e changed).
    {
        double ic[TR], nd[TR], ne[TR];
#pragma unroll
        for (int q = 0; q < TR; q++) {
            const double2 t = active ? __ldg(fac + q * T + j) : zero;
            const double2 e = active ? __ldg(W.e + q * T + j) : zero;
            ic[q] = t.x; nd[q] = t.y;
            ne[q] = -(e.x + lam * e.y);
        }
        if (!IS_MIN && first) ne[1] = -La(5);
        Aff a;
        double2 p1 = zero, p2 = zero;
        double h11 = 1.0, h12 = 0.0, h21 = 0.0, h22 = 1.0;
#pragma unroll
        for (int q = TR - 1; q >= 0; q--) {
            const double2 x = fma2(nd[q], p1, fma2(ne[q], p2, X[q])) * ic[q];
            const double g1 = (nd[q] * h11 + ne[q] * h21) * ic[q], g2 = (nd[q] * h12 + ne[q] * h22) * ic[q];
            p2 = p1; p1 = x;
            h21 = h11; h22 = h12; h11 = g1; h12 = g2;
        }
        a.m11 = h11; a.m12 = h12; a.m21 = h21; a.m22 = h22; a.p1 = p1; a.p2 = p2;
        double2 s1, s2;
        team_scan<-1, NW>(a, lane, wt, xch + NW, barid, s1, s2);
#pragma unroll
        for (int q = TR - 1; q >= 0; q--) {
            const double2 x = fma2(nd[q], s1, fma2(ne[q], s2, X[q])) * ic[q];
            X[q] = x;
            s2 = s1; s1 = x;
        }
    }

    // ---- closure rows
    if (IS_MIN) {
        if (first) X[0] = bc;
        if (last) {
            double2 v = bcs_far;
            v = v + X[6] * Le(2);
            v = v + X[5] * Le(1);
            v = v + X[4] * Le(5);
            X[7] = v;
        }
    } else {
        if (last) {
            X[7] = bc;
            if (du) {
                double row[6];
                Lrow(n, row);
                double2 v = bc * row[3];
                v = v + X[6] * row[2];
                v = v + X[5] * row[1];
                v = v + X[4] * row[5];
                v = v + fn1 * rhs(n, 1);
                *du = v;
            }
        }
        if (first) {
            double2 v = bcs_far;
            v = v + X[1] * Le(4);
            v = v + X[2] * Le(5);
            v = v + X[3] * Le(1);
            X[0] = v;
        }
    }
}

// shared memory of one team, after the tile: reduced boundary rows, scan slots (4 scans), broadcast values
struct TeamShared {
    double cs[64];
    Aff xch[4 * 4];
    double2 bcb, bct, fend, v0n, u01, du0;
};

template <int NW, int MW>
__global__ void __launch_bounds__(32 * NW * MW, (32 * NW * MW <= 256) ? 2 : 1)
poisson_team_kernel(PoissonDev D, double* __restrict__ cf, double* __restrict__ cv) {
    extern __shared__ double2 wtile[];                 // [MW][TP] tiles, then [MW] TeamShared
    const int T = D.T, n = D.ny;
    const int TP = (T * 9) | 1;                        // odd: adjacent modes of a row land in adjacent 16-byte slots
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int mt = w / NW, wt = w - mt * NW;           // team in the CTA, warp in the team
    const int j = wt * 32 + lane;                      // chunk
    const int barid = 1 + mt;
    constexpr int NT = 32 * NW * MW;
    if (blockIdx.y == 0) {
        // The (up to four) singular modes are marched by one thread each (OPR_ODE2_Factorize_NN_Sing: a different, rank-deficient
        // problem) on small planes of their own.  They take as long as ~1000 dependent loads; the first CTA of the grid starts
        // them so that they run beside the regular modes instead of behind them.
        // one CTA per singular mode; its lines live in shared memory while the thread marches (a march through global memory
        // waits for a loaded memory system at every row: 3.7 ms for four modes against 1.4 ms for the 66 560 regular modes of the
        // per-GPU share at 8 GPUs, i.e. the critical path of the stage)
        if (blockIdx.x < 4) {
            const int t = blockIdx.x;
            const int is = (t & 1) ? D.i_sing1 : D.i_sing0, ks = (t & 2) ? D.k_sing1 : D.k_sing0;
            const bool dup = ((t & 2) && D.k_sing1 == D.k_sing0) || ((t & 1) && D.i_sing1 == D.i_sing0);
            if (!dup && is >= 0 && is < D.nxh && ks >= 0 && ks < D.nz) {
                double* sl = reinterpret_cast<double*>(wtile);               // 11 n doubles
                double2* cf2s = reinterpret_cast<double2*>(cf) + ((size_t)is + (size_t)D.nxh * n * ks);
                double2* cv2s = reinterpret_cast<double2*>(cv) + ((size_t)is + (size_t)D.nxh * n * ks);
                for (int row = threadIdx.x; row < n; row += NT) reinterpret_cast<double2*>(sl)[row] = cf2s[(size_t)D.nxh * row];
                __syncthreads();
                double* cache = D.sing + (size_t)t * (10 * (size_t)n + 8);
                if (threadIdx.x == 0) {
                    SingLines Ln;
                    Ln.fre = {sl, 2}; Ln.fim = {sl + 1, 2};
                    Ln.vre = {sl + 2 * n, 2}; Ln.vim = {sl + 2 * n + 1, 2};
                    Ln.ysc0 = {sl + 4 * n, 1}; Ln.ysc1 = {sl + 5 * n, 1};
                    Ln.csc = {sl + 6 * n, 1}; Ln.dsc = {sl + 7 * n, 1}; Ln.esc = {sl + 8 * n, 1};
                    Ln.v1 = {sl + 9 * n, 1}; Ln.u1 = {sl + 10 * n, 1};
                    double cdu[2];
                    poisson_singular_cached(D, sqrt(D.lambda[(long long)is + (long long)D.nxh * ks]), Ln, cache, cdu);
                    sl[6 * n] = cdu[0]; sl[6 * n + 1] = cdu[1];
                }
                __syncthreads();
                {
                    // u = u^(0) + cdu u^(1), v = v^(0) + cdu v^(1) (opr_odes.f90:90-95), all rows at once
                    const double c0 = sl[6 * n], c1 = sl[6 * n + 1];
                    const double* v1 = cache + 8 * (size_t)n;
                    const double* u1 = cache + 9 * (size_t)n;
                    for (int row = threadIdx.x; row < n; row += NT) {
                        const double a = u1[row], b = v1[row];
                        sl[2 * row] = sl[2 * row] + c0 * a; sl[2 * row + 1] = sl[2 * row + 1] + c1 * a;
                        sl[2 * n + 2 * row] = sl[2 * n + 2 * row] + c0 * b; sl[2 * n + 2 * row + 1] = sl[2 * n + 2 * row + 1] + c1 * b;
                    }
                }
                __syncthreads();
                for (int row = threadIdx.x; row < n; row += NT) {
                    cf2s[(size_t)D.nxh * row] = reinterpret_cast<double2*>(sl)[row];
                    cv2s[(size_t)D.nxh * row] = reinterpret_cast<double2*>(sl + 2 * n)[row];
                }
            }
        }
        return;
    }
    const int k = blockIdx.y - 1;
    const int i0 = blockIdx.x * MW;
    const size_t plane0 = (size_t)D.nxh * n * k;       // complex elements before this z plane
    double2* cf2 = reinterpret_cast<double2*>(cf);
    double2* cv2 = reinterpret_cast<double2*>(cv);
    const double norm = D.norm;
    const int i = i0 + mt;
    const bool work = (i < D.nxh) && !mode_is_singular(D, i, k);
    // ---- L2 prefetch: what this team reads later (upper factors of both systems, fundamental lines), and the forcing tile of
    // the CTA that takes this CTA's place on the SM (pf_dist CTAs later), so that its first wait is an L2 hit
    if (work) {
        const long long m = (long long)i + (long long)D.nxh * k;
        const char* f0 = reinterpret_cast<const char*>(D.wfac_min + (size_t)m * n);
        const char* f1 = reinterpret_cast<const char*>(D.wfac_max + (size_t)m * n);
        const char* f2 = reinterpret_cast<const char*>(D.wfund + (size_t)m * 5 * n);
        const int tt = wt * 32 + lane;
        for (int l = tt; l < n / 8; l += 32 * NW) { prefetch_l2g(f0 + (size_t)l * 128); prefetch_l2g(f1 + (size_t)l * 128); }
        for (int l = tt; l < (5 * n) / 16; l += 32 * NW) prefetch_l2g(f2 + (size_t)l * 128);
    }
    if (D.pf_dist > 0) {
        const unsigned lin = (blockIdx.y - 1) * gridDim.x + blockIdx.x + (unsigned)D.pf_dist;
        const unsigned ky = lin / gridDim.x, bx = lin - ky * gridDim.x;
        if (ky < gridDim.y - 1) {
            const double2* base = cf2 + (size_t)D.nxh * n * ky + (size_t)bx * MW;
            const int last_mw = min(MW, D.nxh - (int)bx * MW) - 1;
            for (int row = threadIdx.x; row < n; row += NT) {
                prefetch_l2g(base + (size_t)D.nxh * row);
                prefetch_l2g(base + (size_t)D.nxh * row + last_mw);
            }
        }
    }
    // ---- forcing tile in (row segments of MW complex numbers), scaled
    for (int idx = threadIdx.x; idx < n * MW; idx += NT) {
        const int row = idx / MW, mw = idx - row * MW;
        if (i0 + mw < D.nxh) {
            const double2 v = __ldcs(cf2 + plane0 + (size_t)D.nxh * row + i0 + mw);
            wtile[mw * TP + tpad(row)] = make_double2(v.x * norm, v.y * norm);
        }
    }
    __syncthreads();
    double2* tile = wtile + mt * TP;
    TeamShared& sh = reinterpret_cast<TeamShared*>(wtile + MW * TP)[mt];
    double2 X[TR];
    const bool active = j < T, first = (j == 0), last = (j == T - 1);
    const int c9 = j * 9;
    const double2 zero = make_double2(0.0, 0.0);
    if (work) {
        const long long m = (long long)i + (long long)D.nxh * k;
        const double lam = sqrt(D.lambda[m]);
#pragma unroll
        for (int q = 0; q < TR; q++) X[q] = active ? tile[c9 + q] : zero;
        double2 Fm1 = (active && j > 0) ? tile[c9 - 2] : zero;               // slot of row r0 - 1: (j-1)*9 + 7
        double2 Fp1 = (j < T - 1) ? tile[c9 + 9] : zero;
        if (first) sh.bcb = X[0];                                             // bcs(1:2,1) = f(1:2)
        if (last) sh.bct = X[TR - 1];                                         // bcs(1:2,2) = f(2ny-1:2ny)
        // v^(0): v' + lam v = f, f(n) = 0, v(1) = 0   (in place)
        team_int1<true, NW>(D.smin, D.wmin, D.wfac_min + (size_t)m * n, lam, T, j, lane, wt, barid, X, Fm1, Fp1, zero, zero, zero,
                            sh.cs, sh.xch, nullptr);
        // v^(0) is parked in the tile (needed again in the correction); u^(0): u' - lam u = v, u(n) = 0
        if (active) {
#pragma unroll
            for (int q = 0; q < TR; q++) tile[c9 + q] = X[q];
        }
        if (first) sh.fend = X[0];
        if (last) sh.v0n = X[TR - 1];
        team_barrier<NW>(barid);
        Fm1 = (active && j > 0) ? tile[c9 - 2] : zero;
        Fp1 = (j < T - 1) ? tile[c9 + 9] : zero;
        double2 du0 = zero;
        team_int1<false, NW>(D.smax, D.wmax, D.wfac_max + (size_t)m * n, -lam, T, j, lane, wt, barid, X, Fm1, Fp1, sh.fend, zero, zero,
                             sh.cs, sh.xch + 2 * NW, &du0);
        if (first) sh.u01 = X[0];
        if (last) sh.du0 = du0;
        team_barrier<NW>(barid);
        // constraint and boundary conditions (opr_odes.f90:350-367)
        const long long NM = D.nmodes;
        const double* A = D.amat;
        const double a11 = A[0 * NM + m], a21 = A[1 * NM + m], a31 = A[2 * NM + m];
        const double a12 = A[3 * NM + m], a22 = A[4 * NM + m], a32 = A[5 * NM + m];
        const double a13 = A[6 * NM + m], a23 = A[7 * NM + m], a33 = A[8 * NM + m];
        const double2 bcb = sh.bcb, bct = sh.bct, u0_1 = sh.u01, v0_n = sh.v0n;
        du0 = sh.du0;
        double2 v_1, u_n, fn;
        v_1 = make_double2((bcb.x - lam * u0_1.x) / a11, (bcb.y - lam * u0_1.y) / a11);
        u_n = make_double2((bct.x - v0_n.x - a21 * v_1.x) / a22, (bct.y - v0_n.y - a21 * v_1.y) / a22);
        fn = make_double2((bct.x - du0.x - a31 * v_1.x - a32 * u_n.x) / a33, (bct.y - du0.y - a31 * v_1.y - a32 * u_n.y) / a33);
        u_n = make_double2(u_n.x - a23 * fn.x, u_n.y - a23 * fn.y);
        v_1 = make_double2(v_1.x - a12 * u_n.x - a13 * fn.x, v_1.y - a12 * u_n.y - a13 * fn.y);
        const double* fu = D.wfund + (size_t)m * 5 * n;
        if (active) {
            // rows of this chunk: v^(0) comes back from the tile row by row and dp^/dy takes its place; p^ stays in registers
#pragma unroll
            for (int q = 0; q < TR; q++) {
                const int o = q * T + j;
                const double f0 = __ldcs(fu + o), f1 = __ldcs(fu + n + o), f2 = __ldcs(fu + 2 * n + o), f3 = __ldcs(fu + 3 * n + o),
                             f4 = __ldcs(fu + 4 * n + o);
                const int r = j * TR + q + 1;
                const double2 v0 = tile[c9 + q];
                double2 uu, vv;
                if (r == n) {
                    uu = u_n;
                    vv = make_double2(v0.x + fn.x * f0 + v_1.x * f1 + lam * uu.x, v0.y + fn.y * f0 + v_1.y * f1 + lam * uu.y);
                } else {
                    uu = make_double2(X[q].x + fn.x * f2 + v_1.x * f3 + u_n.x * f4, X[q].y + fn.y * f2 + v_1.y * f3 + u_n.y * f4);
                    if (r == 1) vv = make_double2(v_1.x + lam * uu.x, v_1.y + lam * uu.y);
                    else vv = make_double2(v0.x + fn.x * f0 + v_1.x * f1 + lam * uu.x, v0.y + fn.y * f0 + v_1.y * f1 + lam * uu.y);
                }
                X[q] = uu;
                tile[c9 + q] = vv;
            }
        }
    }
    // ---- results out through the tile: dp^/dy into the second array, then p^ over the forcing
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * MW; idx += NT) {
        const int row = idx / MW, mw = idx - row * MW;
        if (i0 + mw < D.nxh && !mode_is_singular(D, i0 + mw, k))
            __stcs(cv2 + plane0 + (size_t)D.nxh * row + i0 + mw, wtile[mw * TP + tpad(row)]);
    }
    __syncthreads();
    if (work && active) {
#pragma unroll
        for (int q = 0; q < TR; q++) tile[c9 + q] = X[q];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * MW; idx += NT) {
        const int row = idx / MW, mw = idx - row * MW;
        if (i0 + mw < D.nxh && !mode_is_singular(D, i0 + mw, k))
            __stcs(cf2 + plane0 + (size_t)D.nxh * row + i0 + mw, wtile[mw * TP + tpad(row)]);
    }
}

// full planes (blocked by 32 modes, see plane()) -> lane order of the warp kernel, one plane per launch.  One CTA per group
// of 32 modes: coalesced reads along the modes, transposed through shared memory, coalesced writes along the rows of a mode.
// dst element of (mode m, 0-based row r): dst[(m * n + (r % 8) * T + r / 8) * dstride + doff]
__global__ void poisson_relayout_kernel(PoissonDev D, int which, int kplane, bool inner_only, double* __restrict__ dst,
                                        long long mode_stride, int dstride, int doff) {
    __shared__ double t[32][33];
    const int n = D.ny, T = D.T;
    const long long m0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int r0 = 0; r0 < n; r0 += 32) {
        for (int rr = ty; rr < 32; rr += 8) {
            const int r = r0 + rr + 1;
            const long long m = m0 + tx;
            bool ok = (m < D.nmodes) && (r <= n);
            if (inner_only) ok = ok && r > 1 && r < n;
            t[rr][tx] = ok ? plane(D, which, kplane, m).get(r) : 0.0;
        }
        __syncthreads();
        for (int mm = ty; mm < 32; mm += 8) {
            const long long m = m0 + mm;
            const int rb = r0 + tx;                 // 0-based row
            if (m < D.nmodes && rb < n)
                dst[((size_t)m * mode_stride + (size_t)((rb & 7) * T + (rb >> 3))) * dstride + doff] = t[tx][mm];
        }
        __syncthreads();
    }
}

// boundary-condition planes into rows 1 and ny of the forcing (opr_elliptic.f90:285-286)
__global__ void poisson_set_bcs_kernel(double* __restrict__ p, const double* __restrict__ hb, const double* __restrict__ ht,
                                       int nx, int ny, int nz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * nz) return;
    const int i = (int)(idx % nx);
    const long long k = idx / nx;
    p[i + (long long)nx * ny * k] = hb[idx];
    p[i + (long long)nx * ((ny - 1) + (long long)ny * k)] = ht[idx];
}

// ---- kx-split exchanges over peer memory (complex numbers as double2) ---------------------------------------
// slab c(kx, y, zl) of this rank  ->  pencil_p(kx - kx0[p], y, rank*nzl + zl) of every rank p   (stores into the peers)
struct KxTab { double2* p[8]; int kx0[9]; };

constexpr int KXR = 4;          // rows in flight per CTA pass (independent loads first, then the stores)

__global__ void kx_push_kernel(const double2* __restrict__ c, KxTab t, int nxh, int ny, int nzl, int rank, int P, double scale) {
    const int rows = ny * nzl;
    for (int row0 = blockIdx.x * KXR; row0 < rows; row0 += gridDim.x * KXR) {
        for (int j = threadIdx.x; j < nxh; j += blockDim.x) {
            int p = 0;
#pragma unroll
            for (int q = 1; q < 8; q++) if (q < P && j >= t.kx0[q]) p = q;
            const int kxl = t.kx0[p + 1] - t.kx0[p];
            double2 v[KXR];
#pragma unroll
            for (int r = 0; r < KXR; r++) if (row0 + r < rows) v[r] = __ldcs(c + (size_t)nxh * (row0 + r) + j);
#pragma unroll
            for (int r = 0; r < KXR; r++) {
                const int row = row0 + r;
                if (row < rows) {
                    const int y = row % ny, zl = row / ny;
                    const size_t prow = (size_t)y + (size_t)ny * ((size_t)rank * nzl + zl);
                    t.p[p][(size_t)(j - t.kx0[p]) + (size_t)kxl * prow] = v[r];
                }
            }
        }
    }
}
// pencil_p(kx - kx0[p], y, rank*nzl + zl) of every rank p  ->  slab c(kx, y, zl) of this rank   (loads from the peers)
__global__ void kx_pull_kernel(double2* __restrict__ c, KxTab t, int nxh, int ny, int nzl, int rank, int P) {
    const int rows = ny * nzl;
    for (int row0 = blockIdx.x * KXR; row0 < rows; row0 += gridDim.x * KXR) {
        for (int j = threadIdx.x; j < nxh; j += blockDim.x) {
            int p = 0;
#pragma unroll
            for (int q = 1; q < 8; q++) if (q < P && j >= t.kx0[q]) p = q;
            const int kxl = t.kx0[p + 1] - t.kx0[p];
            double2 v[KXR];
#pragma unroll
            for (int r = 0; r < KXR; r++) {
                const int row = row0 + r;
                if (row < rows) {
                    const int y = row % ny, zl = row / ny;
                    const size_t prow = (size_t)y + (size_t)ny * ((size_t)rank * nzl + zl);
                    v[r] = t.p[p][(size_t)(j - t.kx0[p]) + (size_t)kxl * prow];
                }
            }
#pragma unroll
            for (int r = 0; r < KXR; r++) if (row0 + r < rows) c[(size_t)nxh * (row0 + r) + j] = v[r];
        }
    }
}

const double* up(std::vector<void*>& allocs, const double* h, size_t count) {
    double* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(double)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h, count * sizeof(double), cudaMemcpyHostToDevice);
    allocs.push_back(d);
    return d;
}

int make_side(const HostDer& der1, int bc, Int1Dev& S, std::vector<void*>& allocs) {
    HostInt1 H;
    int rc = int1_create_base(der1, bc, H);
    if (rc) return rc;
    const int n = H.n;
    std::vector<double> L0((size_t)n * 5), L1((size_t)n * 5), R((size_t)n * 3);
    for (int r = 1; r <= n; r++) {
        for (int k = 1; k <= 5; k++) { L0[(size_t)(r - 1) * 5 + k - 1] = H.L0(r, k); L1[(size_t)(r - 1) * 5 + k - 1] = H.L1(r, k); }
        for (int k = 1; k <= 3; k++) R[(size_t)(r - 1) * 3 + k - 1] = H.rhs(r, k);
    }
    S.n = n; S.bc = bc;
    S.L0 = up(allocs, L0.data(), L0.size());
    S.L1 = up(allocs, L1.data(), L1.size());
    S.rhs = up(allocs, R.data(), R.size());
    for (int r = 1; r <= 3; r++) for (int c = 0; c <= 3; c++) S.rb[r][c] = H.rhs_b0(r, c);
    for (int r = 0; r <= 2; r++) for (int c = 1; c <= 4; c++) S.rt[r][c] = H.rhs_t0(r, c);
    return (S.L0 && S.L1 && S.rhs) ? 0 : TLAB_ERR_ALLOC;
}

// lane-order copies of the shared tables of one side (see WarpSide): entry of row r = 8 j + q + 1 at [q * T + j]
int make_warp_side(const HostDer& der1, int bc, int T, WarpSide& Ws, std::vector<void*>& allocs) {
    HostInt1 H;
    int rc = int1_create_base(der1, bc, H);
    if (rc) return rc;
    const int n = H.n;
    std::vector<double> t[4];
    for (auto& v : t) v.assign((size_t)2 * n, 0.0);
    for (int r = 1; r <= n; r++) {
        const size_t o = (size_t)(((r - 1) & 7) * T + ((r - 1) >> 3)) * 2;
        t[0][o] = H.rhs(r, 1); t[0][o + 1] = H.rhs(r, 2);
        t[1][o] = H.L0(r, 1); t[1][o + 1] = H.L1(r, 1);
        t[2][o] = H.L0(r, 2); t[2][o + 1] = H.L1(r, 2);
        t[3][o] = H.L0(r, 5); t[3][o + 1] = H.L1(r, 5);
    }
    const double* d[4];
    for (int k = 0; k < 4; k++) { d[k] = up(allocs, t[k].data(), t[k].size()); if (!d[k]) return TLAB_ERR_ALLOC; }
    Ws.rh = reinterpret_cast<const double2*>(d[0]);
    Ws.ab0 = reinterpret_cast<const double2*>(d[1]);
    Ws.ab1 = reinterpret_cast<const double2*>(d[2]);
    Ws.e = reinterpret_cast<const double2*>(d[3]);
    return 0;
}

}  // namespace

// One thread per mode, or (tuning key poisson_split: 1 always, 0 never, -1 when this GPU holds fewer than 200 000 modes)
// one thread per component; the latter needs the stored factor lines.
static void launch_modes(const PoissonDev& D, double* cf, double* cv, cudaStream_t st) {
    if (D.T > 0) {
        // team per mode: MW adjacent kx per CTA (row segments of MW * 16 bytes), one z plane per blockIdx.y
        const int NW = D.T <= 32 ? 1 : (D.T <= 64 ? 2 : 4);
        int MW = ctx().tune_poisson_warp;
        if (MW != 2 && MW != 4 && MW != 8) MW = 4;
        if (NW * MW > 16) MW = 16 / NW;
        const int TP = (D.T * 9) | 1;
        const size_t smem = std::max((size_t)MW * TP * sizeof(double2) + (size_t)MW * sizeof(TeamShared), (size_t)11 * D.ny * sizeof(double));
        const dim3 grid((unsigned)std::max((D.nxh + MW - 1) / MW, 4), (unsigned)D.nz + 1);      // row 0: singular modes, one CTA each
        PoissonDev Dl = D;
        Dl.pf_dist = ctx().tune_poisson_pf < 0 ? 2 * 148 : ctx().tune_poisson_pf;
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, 32 * NW * MW, smem, st>>>(Dl, cf, cv);
        };
        if (NW == 1) { if (MW == 8) go(poisson_team_kernel<1, 8>); else if (MW == 4) go(poisson_team_kernel<1, 4>); else go(poisson_team_kernel<1, 2>); }
        else if (NW == 2) { if (MW == 8) go(poisson_team_kernel<2, 8>); else if (MW == 4) go(poisson_team_kernel<2, 4>); else go(poisson_team_kernel<2, 2>); }
        else { if (MW == 4) go(poisson_team_kernel<4, 4>); else go(poisson_team_kernel<4, 2>); }
        return;
    }
    const int threads = 128;
    int split = ctx().tune_poisson_split;
    // measured (profiles/ncu_full_poisson_r01.json): 66 560 modes (C3 on 8 GPUs) 5.6 -> 5.1 ms with one thread per component and
    // 4 CTAs/SM; 525 312 modes (C3 on one GPU) 17.4 -> 18.7 ms: only geometries that leave the SMs short of threads split
    if (split < 0) split = (D.nmodes < 200000) ? 1 : 0;
    if (!D.fac) split = 0;
    const int minb = split ? std::max(ctx().tune_poisson_minb, 4) : ctx().tune_poisson_minb;
    const long long work = split ? 2 * D.nmodes : D.nmodes;
    const unsigned blocks = (unsigned)((work + threads - 1) / threads);
    if (split) {
        if (minb == 4) poisson_modes_kernel<4, true, 1><<<blocks, threads, 0, st>>>(D, cf, cv);
        else if (minb == 2) poisson_modes_kernel<2, true, 1><<<blocks, threads, 0, st>>>(D, cf, cv);
        else poisson_modes_kernel<3, true, 1><<<blocks, threads, 0, st>>>(D, cf, cv);
    } else if (D.fac) {
        if (minb == 4) poisson_modes_kernel<4, true, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
        else if (minb == 2) poisson_modes_kernel<2, true, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
        else poisson_modes_kernel<3, true, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
    } else {
        if (minb == 4) poisson_modes_kernel<4, false, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
        else if (minb == 2) poisson_modes_kernel<2, false, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
        else poisson_modes_kernel<3, false, 2><<<blocks, threads, 0, st>>>(D, cf, cv);
    }
}

static int cufft_check(cufftResult r, const char* what) {
    if (r == CUFFT_SUCCESS) return 0;
    return fail(TLAB_ERR_CUDA, std::string(what) + ": cuFFT error " + std::to_string((int)r));
}

void Poisson::release() {
    if (plan_fx) cufftDestroy(plan_fx);
    if (plan_bx) cufftDestroy(plan_bx);
    if (plan_z) cufftDestroy(plan_z);
    plan_fx = plan_bx = plan_z = 0;
    if (c3) trp().unregister_buffer(c3);
    c3 = nullptr;
    if (cpa) trp().unregister_buffer(cpa);
    if (cpb) trp().unregister_buffer(cpb);
    cpa = cpb = nullptr;
    kxsplit = false;
    for (void* a : allocs) cudaFree(a);
    allocs.clear();
    ready = false;
}

int Poisson::init(tlab_plan_s* gx, tlab_plan_s* gy, tlab_plan_s* gz, int nz_local) {
    release();
    if (!gx || !gy || !gz) return fail(TLAB_ERR_OPTION, "OPR_Elliptic_Initialize: null plan");
    if (!gx->p.periodic || (gz->p.n > 1 && !gz->p.periodic))
        return fail(TLAB_ERR_OPTION, "OPR_Poisson (Fourier) needs periodic x and z");
    if (gy->p.periodic || gy->p.n < 16) return fail(TLAB_ERR_OPTION, "OPR_Poisson needs a non-periodic y with >= 16 points");
    nx = gx->p.n; ny = gy->p.n; nzg = gz->p.n;
    nz = (nz_local > 0) ? nz_local : nzg;                                    // local slab thickness (kmax)
    P = trp().P;
    if (nz * P != nzg) return fail(TLAB_ERR_PARPARTITION, "OPR_Poisson: kmax times the number of ranks differs from the grid size in z");
    const int koff = trp().rank * nz;                                        // ims_offset_k
    if (nx % 2) return fail(TLAB_ERR_DIMGRID, "OPR_Poisson needs an even number of points in x");
    nxh = nx / 2 + 1;
    // split domain: try the kx-split spectral stage (needs peer-mapped pencils), else the reference's y-line split
    kxsplit = false;
    const int rank = trp().rank;
    if (P > 1 && nzg > 1 && P <= 8 && trp().p2p_enabled && ctx().tune_kxsplit && nxh >= P) {
        for (int p = 0; p <= P; p++) kx0[p] = (int)(((long long)p * nxh) / P);
        int kxl_max = 0;
        for (int p = 0; p < P; p++) kxl_max = std::max(kxl_max, kx0[p + 1] - kx0[p]);
        const size_t pen = (size_t)2 * kxl_max * ny * nzg;
        double *a = nullptr, *b = nullptr;
        if (cudaMalloc(&a, pen * sizeof(double)) != cudaSuccess || cudaMalloc(&b, pen * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(TLAB_ERR_ALLOC, "OPR_Elliptic_Initialize: out of device memory (kx pencils)");
        }
        allocs.push_back(a); allocs.push_back(b);
        cpa = a; cpb = b;
        if (int rc = trp().register_buffer(cpa)) return rc;
        if (int rc = trp().register_buffer(cpb)) return rc;
        kxsplit = trp().find(cpa) != nullptr && trp().find(cpb) != nullptr;
    }
    if (P > 1 && !kxsplit && ((long long)nxh * ny) % P) return fail(TLAB_ERR_PARPARTITION, "OPR_Poisson: (nx/2+1)*ny is not a multiple of the number of ranks");
    const int kxo = kxsplit ? kx0[rank] : 0;                    // first wavenumber / modes in x / planes in z of the y solves
    const int mxh = kxsplit ? kx0[rank + 1] - kx0[rank] : nxh;
    const int mz_n = kxsplit ? nzg : nz;
    const int mko = kxsplit ? 0 : koff;
    D.nxh = mxh; D.ny = ny; D.nz = mz_n; D.nmodes = (long long)mxh * mz_n;
    D.norm = 1.0 / double((long long)nx * nzg);                              // opr_elliptic.f90:130
    D.i_sing0 = 0 - kxo; D.i_sing1 = nx / 2 - kxo;                           // opr_elliptic.f90:148-149 (0-based, local)
    D.k_sing0 = 0 - mko; D.k_sing1 = nzg / 2 - mko;                          // task-local indices (:177-178)
    // lambda(k,i) = mwn_x(i)^2 + mwn_z(k)^2 from the first-derivative modified wavenumbers (:199-203)
    std::vector<double> lam((size_t)D.nmodes);
    const std::vector<double>& mx = gx->p.h.der1.mwn;
    const std::vector<double>& mz = gz->p.h.der1.mwn;
    for (int k = 0; k < mz_n; k++)
        for (int i = 0; i < mxh; i++) {
            double l = mx[kxo + i] * mx[kxo + i];
            if (nzg > 1) l = l + mz[mko + k] * mz[mko + k];
            lam[(size_t)i + (size_t)mxh * k] = l;
        }
    D.lambda = up(allocs, lam.data(), lam.size());
    if (int rc = make_side(gy->p.h.der1, BCS_MIN, D.smin, allocs)) return fail(rc, "integral operator (BCS_MIN) setup failed");
    if (int rc = make_side(gy->p.h.der1, BCS_MAX, D.smax, allocs)) return fail(rc, "integral operator (BCS_MAX) setup failed");
    const size_t plane_sz = (size_t)((D.nmodes + 31) / 32 * 32) * ny;
    D.plane_sz = (long long)plane_sz;
    D.il = ctx().tune_poisson_il ? 1 : 0;
    double *fund = nullptr, *scr = nullptr, *amat = nullptr;
    if (cudaMalloc(&fund, 5 * plane_sz * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&scr, 6 * plane_sz * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&amat, 9 * (size_t)D.nmodes * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return fail(TLAB_ERR_ALLOC, "OPR_Elliptic_Initialize: out of device memory");
    }
    allocs.push_back(fund); allocs.push_back(scr); allocs.push_back(amat);
    D.fund = fund; D.scr = scr; D.amat = amat;
    D.fac = nullptr;
    if (ctx().tune_poisson_factors) {
        // LU factors of every mode kept (like the reference does, opr_elliptic.f90:140,205-209) when memory allows
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t need = 8 * plane_sz * sizeof(double);
        double* fac = nullptr;
        if (free_b > need + (size_t)60 * plane_sz * sizeof(double) / 2 && cudaMalloc(&fac, need) == cudaSuccess) { allocs.push_back(fac); D.fac = fac; }
        else cudaGetLastError();
    }
    cudaMemset(fund, 0, 5 * plane_sz * sizeof(double));
    // cuFFT plans (Appendix B of SURVEY.md: same geometry as the FFTW plans, opr_fourier.f90:101-171)
    cudaStream_t st = ctx().stream;
    int n1[1] = {nx};
    if (int rc = cufft_check(cufftPlanMany(&plan_fx, 1, n1, n1, 1, nx, n1, 1, nxh, CUFFT_D2Z, ny * nz), "cufftPlanMany D2Z")) return rc;
    if (int rc = cufft_check(cufftPlanMany(&plan_bx, 1, n1, n1, 1, nxh, n1, 1, nx, CUFFT_Z2D, ny * nz), "cufftPlanMany Z2D")) return rc;
    cufftSetStream(plan_fx, st);
    cufftSetStream(plan_bx, st);
    if (nzg > 1) {
        // slab layout (P = 1) or z-pencil layout after the K-transpose (P > 1): lines interleaved, stride = howmany
        int n3[1] = {nzg};
        const int howmany = kxsplit ? (kx0[rank + 1] - kx0[rank]) * ny : (int)(((long long)nxh * ny) / P);
        if (int rc = cufft_check(cufftPlanMany(&plan_z, 1, n3, n3, howmany, 1, n3, howmany, 1, CUFFT_Z2Z, howmany), "cufftPlanMany Z2Z")) return rc;
        cufftSetStream(plan_z, st);
    }
    if (P > 1 && !kxsplit) {
        double* c3buf = nullptr;
        if (cudaMalloc(&c3buf, (size_t)2 * nxh * ny * nz * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(TLAB_ERR_ALLOC, "OPR_Elliptic_Initialize: out of device memory (pencil buffer)");
        }
        allocs.push_back(c3buf);
        c3 = c3buf;
        if (int rc = trp().register_buffer(c3)) return rc;
    }
    const int threads = 128;
    const unsigned blocks = (unsigned)((D.nmodes + threads - 1) / threads);
    poisson_fundamental_kernel<<<blocks, threads, 0, st>>>(D);
    if (int rc = cuda_check(cudaStreamSynchronize(st), "poisson fundamental solutions")) return rc;
    D.T = 0;
    if (ctx().tune_poisson_warp && D.fac && ny % 8 == 0 && ny / 8 <= 128) {
        // warp-per-mode kernel: keep the upper factors and the fundamental lines in lane order, drop the full planes
        auto drop = [&](double* b) {
            cudaFree(b);
            allocs.erase(std::remove(allocs.begin(), allocs.end(), (void*)b), allocs.end());
        };
        drop(scr); D.scr = nullptr;
        D.T = ny / 8;
        const size_t per_mode = (size_t)ny;
        double *wmin = nullptr, *wmax = nullptr, *wfund = nullptr, *sing = nullptr;
        if (cudaMalloc(&wmin, 2 * per_mode * D.nmodes * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&wmax, 2 * per_mode * D.nmodes * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&wfund, 5 * per_mode * D.nmodes * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&sing, (size_t)11 * 32 * ny * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            for (double* b : {wmin, wmax, wfund, sing}) if (b) cudaFree(b);
            return fail(TLAB_ERR_ALLOC, "OPR_Elliptic_Initialize: out of device memory (lane-order planes)");
        }
        for (double* b : {wmin, wmax, wfund, sing}) allocs.push_back(b);
        cudaMemsetAsync(sing, 0, (size_t)11 * 32 * ny * sizeof(double), st);
        const unsigned groups = (unsigned)((D.nmodes + 31) / 32);
        poisson_relayout_kernel<<<groups, 256, 0, st>>>(D, P_FAC, 2, true, wmin, (long long)ny, 2, 0);
        poisson_relayout_kernel<<<groups, 256, 0, st>>>(D, P_FAC, 3, true, wmin, (long long)ny, 2, 1);
        poisson_relayout_kernel<<<groups, 256, 0, st>>>(D, P_FAC, 6, true, wmax, (long long)ny, 2, 0);
        poisson_relayout_kernel<<<groups, 256, 0, st>>>(D, P_FAC, 7, true, wmax, (long long)ny, 2, 1);
        for (int f = 0; f < 5; f++)
            poisson_relayout_kernel<<<groups, 256, 0, st>>>(D, P_FUND, f, false, wfund + (size_t)f * ny, 5LL * ny, 1, 0);
        if (int rc = cuda_check(cudaStreamSynchronize(st), "poisson lane-order planes")) return rc;
        drop(fund); drop(D.fac);
        D.fund = nullptr; D.fac = nullptr;
        D.wfac_min = reinterpret_cast<const double2*>(wmin);
        D.wfac_max = reinterpret_cast<const double2*>(wmax);
        D.wfund = wfund; D.sing = sing;
        if (int rc = make_warp_side(gy->p.h.der1, BCS_MIN, D.T, D.wmin, allocs)) return fail(rc, "lane-order tables (BCS_MIN)");
        if (int rc = make_warp_side(gy->p.h.der1, BCS_MAX, D.T, D.wmax, allocs)) return fail(rc, "lane-order tables (BCS_MAX)");
    }
    ready = true;
    return 0;
}

// p: forcing in, solution out (nx,ny,nz); c1, c2: complex work arrays of (nx/2+1)*ny*nz; dpdy optional
int Poisson::solve(double* p, double* c1, double* c2, const double* hb, const double* ht, double* dpdy) {
    if (!ready) return fail(TLAB_ERR_OPTION, "OPR_Poisson called before OPR_Elliptic_Initialize");
    cudaStream_t st = ctx().stream;
    const long long nlc = 2LL * nxh * ny;          // doubles per z-plane of the half spectrum
    // z transform of a slab spectrum held in `c`, using `w` as pencil work space when the domain is split
    auto fft_z = [&](double* c, double* w, int dir) -> int {
        if (nzg <= 1) return 0;
        if (P == 1) {
            ProfScope ps(PC_FFT);
            return cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)c, (cufftDoubleComplex*)c, dir), "cufftExecZ2Z");
        }
        if (int rc = trp().forward(c, nullptr, 0.0, w, nlc, nz)) return rc;
        {
            ProfScope ps(PC_FFT);
            if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)w, (cufftDoubleComplex*)w, dir), "cufftExecZ2Z")) return rc;
        }
        return trp().backward(w, c, nlc, nz, 0);
    };
    {
        ProfScope ps(PC_FFT);
        const long long np = (long long)nx * nz;
        poisson_set_bcs_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(p, hb, ht, nx, ny, nz);
        if (int rc = cufft_check(cufftExecD2Z(plan_fx, p, (cufftDoubleComplex*)c1), "cufftExecD2Z")) return rc;
    }
    if (kxsplit) {
        Trp& T = trp();
        KxTab ta, tb;
        const Trp::PeerTab* pa = T.find(cpa);
        const Trp::PeerTab* pb = T.find(cpb);
        // the registry is emptied when a later registration finds that not every rank can map peer memory (trp.cu)
        if (!pa || !pb) return fail(TLAB_ERR_OPTION, "OPR_Poisson: the peer mappings of the kx-split exchange are gone; re-initialise the elliptic solver");
        for (int q = 0; q < 8; q++) { ta.p[q] = (double2*)pa->p[q]; tb.p[q] = (double2*)pb->p[q]; }
        for (int q = 0; q < 9; q++) { ta.kx0[q] = kx0[q]; tb.kx0[q] = kx0[q]; }
        const unsigned ctas = (unsigned)std::min<long long>(((long long)ny * nz + KXR - 1) / KXR, T.p2p_ctas * 4);
        {
            ProfScope ps(PC_TRANSPOSE);
            kx_push_kernel<<<ctas, 256, 0, st>>>((const double2*)c1, ta, nxh, ny, nz, T.rank, P, 1.0);
            T.launches++; T.p2p_exchanges++;
            if (int rc = T.barrier()) return rc;
        }
        {
            ProfScope ps(PC_FFT);
            if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)cpa, (cufftDoubleComplex*)cpa, CUFFT_FORWARD), "cufftExecZ2Z")) return rc;
        }
        {
            ProfScope ps(PC_POISSON_Y);
            launch_modes(D, cpa, cpb, st);
            if (int rc = cuda_check(cudaGetLastError(), "poisson mode kernels")) return rc;
        }
        if (ctx().tune_pull_overlap && T.zstream) {
            // The way back on two streams: the pull of p^ (second stream) runs beside the inverse z transform of dp^/dy, the pull of
            // dp^/dy beside the inverse x transform of p.  All NCCL barriers of this section are issued on the second stream.
            cudaStream_t s0 = st, s1 = T.zstream;
            for (cudaEvent_t& e : ev) if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            {
                ProfScope ps(PC_FFT);
                if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)cpa, (cufftDoubleComplex*)cpa, CUFFT_INVERSE), "cufftExecZ2Z")) return rc;
                cudaEventRecord(ev[0], s0);
                if (dpdy) {
                    if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)cpb, (cufftDoubleComplex*)cpb, CUFFT_INVERSE), "cufftExecZ2Z")) return rc;
                    cudaEventRecord(ev[1], s0);
                }
            }
            int rc = 0;
            ctx().stream = s1;
            cudaStreamWaitEvent(s1, ev[0], 0);
            {
                ProfScope ps(PC_TRANSPOSE);
                rc = T.barrier();
                if (!rc) kx_pull_kernel<<<ctas, 256, 0, s1>>>((double2*)c1, ta, nxh, ny, nz, T.rank, P);
                cudaEventRecord(ev[2], s1);
            }
            if (!rc && dpdy) {
                cudaStreamWaitEvent(s1, ev[1], 0);
                ProfScope ps(PC_TRANSPOSE);
                rc = T.barrier();
                if (!rc) kx_pull_kernel<<<ctas, 256, 0, s1>>>((double2*)c2, tb, nxh, ny, nz, T.rank, P);
                cudaEventRecord(ev[3], s1);
            }
            if (!rc) rc = T.barrier();
            cudaEventRecord(ev[4], s1);
            T.launches += dpdy ? 2 : 1; T.p2p_exchanges += dpdy ? 2 : 1;
            ctx().stream = s0;
            if (rc) return rc;
            cudaStreamWaitEvent(s0, ev[2], 0);
            {
                ProfScope ps(PC_FFT);
                if (int r2 = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c1, p), "cufftExecZ2D")) return r2;
            }
            if (dpdy) {
                cudaStreamWaitEvent(s0, ev[3], 0);
                ProfScope ps(PC_FFT);
                if (int r2 = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c2, dpdy), "cufftExecZ2D")) return r2;
            }
            cudaStreamWaitEvent(s0, ev[4], 0);      // the peers have read this rank's pencils: they may be overwritten again
            return 0;
        }
        {
            ProfScope ps(PC_FFT);
            if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)cpa, (cufftDoubleComplex*)cpa, CUFFT_INVERSE), "cufftExecZ2Z")) return rc;
            if (dpdy) { if (int rc = cufft_check(cufftExecZ2Z(plan_z, (cufftDoubleComplex*)cpb, (cufftDoubleComplex*)cpb, CUFFT_INVERSE), "cufftExecZ2Z")) return rc; }
        }
        {
            ProfScope ps(PC_TRANSPOSE);
            if (int rc = T.barrier()) return rc;
            kx_pull_kernel<<<ctas, 256, 0, st>>>((double2*)c1, ta, nxh, ny, nz, T.rank, P);
            if (dpdy) kx_pull_kernel<<<ctas, 256, 0, st>>>((double2*)c2, tb, nxh, ny, nz, T.rank, P);
            T.launches += dpdy ? 2 : 1; T.p2p_exchanges += dpdy ? 2 : 1;
            if (int rc = T.barrier()) return rc;
        }
        ProfScope ps(PC_FFT);
        if (int rc = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c1, p), "cufftExecZ2D")) return rc;
        if (dpdy) { if (int rc = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c2, dpdy), "cufftExecZ2D")) return rc; }
        return 0;
    }
    if (int rc = fft_z(c1, P > 1 ? c2 : nullptr, CUFFT_FORWARD)) return rc;
    {
        ProfScope ps(PC_POISSON_Y);
        launch_modes(D, c1, c2, st);
        if (int rc = cuda_check(cudaGetLastError(), "poisson mode kernels")) return rc;
    }
    if (int rc = fft_z(c1, c3, CUFFT_INVERSE)) return rc;
    {
        ProfScope ps(PC_FFT);
        if (int rc = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c1, p), "cufftExecZ2D")) return rc;
    }
    if (dpdy) {
        if (int rc = fft_z(c2, c3, CUFFT_INVERSE)) return rc;
        ProfScope ps(PC_FFT);
        if (int rc = cufft_check(cufftExecZ2D(plan_bx, (cufftDoubleComplex*)c2, dpdy), "cufftExecZ2D")) return rc;
    }
    return 0;
}

Poisson& poisson() {
    static Poisson P;
    return P;
}

}  // namespace tlab

using namespace tlab;

extern "C" {

int tlab_opr_elliptic_init(tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz, int kmax_local) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    return poisson().init(gx, gy, gz, kmax_local);
}

int tlab_opr_poisson(int nx, int ny, int nz, int ibc, double* p, double* tmp1, double* tmp2, const double* bcs_hb,
                     const double* bcs_ht, double* dpdy) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    Poisson& P = poisson();
    if (!P.ready) return fail(TLAB_ERR_OPTION, "OPR_Poisson called before OPR_Elliptic_Initialize");
    if (nx != P.nx || ny != P.ny || nz != P.nz) return fail(TLAB_ERR_DIMGRID, "OPR_Poisson: extents differ from the initialised ones");
    if (ibc != TLAB_BCS_NN) return fail(TLAB_ERR_UNDEVELOP, "OPR_Poisson: only BCS_NN is implemented on the GPU path");
    if (!p || !tmp1 || !tmp2 || !bcs_hb || !bcs_ht) return fail(TLAB_ERR_OPTION, "OPR_Poisson: null argument");
    if (int rc = P.solve(p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy)) return rc;
    return finish();
}

}  // extern "C"
