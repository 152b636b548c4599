// Fast path of the batched compact-scheme line operators for sm_100a (lines whose length is a multiple of CHUNK).
//
// Same operator surface as lines.cu -- OPR_Partial (P1, P2, P2_P1), OPR_Burgers and BOUNDARY_BCS_NEUMANN_Y of the
// reference (src/operators/opr_partial.f90:31-377, src/physics/opr_burgers.f90:190-521,
// src/tools/dns/boundary_bcs.f90:368-473; banded products src/fdm/fdm_matmul.f90; Thomas sweeps
// src/utils/linear3.f90:56-150,321-442) -- with a cheaper formulation of the factored tridiagonal solves:
//
//   * a line of n points is owned by T = n/16 threads, 16 consecutive points each, in registers;
//   * each thread runs the forward and the backward substitution of its chunk ONCE, with zero inflow:
//         y_j = f_j + a_j y_{j-1},      x^_j = d_j y_j + g_j x^_{j+1}
//     (beta of the circulant factorisation is folded into a and d on the host);
//   * both sweeps are linear, so the true solution of the chunk is
//         x_j = x^_j + Q_j A + R_j B + S_j x_N
//     with A the true forward value at the end of the previous chunk, B the (x_N-free) true solution at the start of
//     the next chunk and x_N the rank-one closure of the circulant system; Q, R, S are precomputed per point
//     (constants of the scheme in the interior of a uniform direction, tables elsewhere);
//   * A and B are weighted sums of at most 6 published chunk ends (the multipliers of a chunk are < 1e-6, the
//     dropped tail is < 2^-80): two block barriers per solve instead of four, two dependent chains instead of four,
//     5 flops per point and system instead of 8.
//
// Memory: y and z lines are strided in memory and contiguous across lines, so the threads of L = 4 adjacent lines
// load their chunks straight into registers (32-byte sectors); x lines are contiguous, a tile of L lines is staged
// through shared memory with 16-byte accesses on both sides (chunk blocks padded to 18 doubles: conflict-free).
// Every CTA also prefetches into L2 the tile that the CTA `pf_dist` places later will work on, so that the DRAM
// latency of a tile is paid while earlier tiles are being solved.
#include "lines2_dev.cuh"
#include <cuda.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>

namespace cg = cooperative_groups;

namespace tlab {

namespace {

// A = sum_k wf[k] y(t-k)
__device__ __forceinline__ double look_back(const double* __restrict__ y, const double2* __restrict__ cr, const ChunkCtx& c) {
    const double2 w01 = ldg2(cr + 0), w23 = ldg2(cr + 1), w45 = ldg2(cr + 2);
    double A = w01.x * y[max(c.t - 1, 0) * c.L + c.l];
    A = fma(w01.y, y[max(c.t - 2, 0) * c.L + c.l], A);
    A = fma(w23.x, y[max(c.t - 3, 0) * c.L + c.l], A);
    A = fma(w23.y, y[max(c.t - 4, 0) * c.L + c.l], A);
    A = fma(w45.x, y[max(c.t - 5, 0) * c.L + c.l], A);
    A = fma(w45.y, y[max(c.t - 6, 0) * c.L + c.l], A);
    return A;
}
// B = sum_k wb[k] z(t+k)
__device__ __forceinline__ double look_ahead(const double* __restrict__ z, const double2* __restrict__ cr, const ChunkCtx& c) {
    const double2 w01 = ldg2(cr + 3), w23 = ldg2(cr + 4), w45 = ldg2(cr + 5);
    const int last = c.T - 1;
    double B = w01.x * z[min(c.t + 1, last) * c.L + c.l];
    B = fma(w01.y, z[min(c.t + 2, last) * c.L + c.l], B);
    B = fma(w23.x, z[min(c.t + 3, last) * c.L + c.l], B);
    B = fma(w23.y, z[min(c.t + 4, last) * c.L + c.l], B);
    B = fma(w45.x, z[min(c.t + 5, last) * c.L + c.l], B);
    B = fma(w45.y, z[min(c.t + 6, last) * c.L + c.l], B);
    return B;
}
// circulant form: constant weights, windows wrapping around the line (T >= LB2)
__device__ __forceinline__ double look_back_circ(const double* __restrict__ y, const Sys2& S, const ChunkCtx& c) {
    double A = 0.0;
#pragma unroll
    for (int k = LB2 - 1; k >= 0; k--) {
        int tm = c.t - 1 - k;
        tm += (tm < 0) ? c.T : 0;
        A = fma(S.cwf[k], y[tm * c.L + c.l], A);
    }
    return A;
}
__device__ __forceinline__ double look_ahead_circ(const double* __restrict__ z, const Sys2& S, const ChunkCtx& c) {
    double B = 0.0;
#pragma unroll
    for (int k = LB2 - 1; k >= 0; k--) {
        int tp = c.t + 1 + k;
        tp -= (tp >= c.T) ? c.T : 0;
        B = fma(S.cwb[k], z[tp * c.L + c.l], B);
    }
    return B;
}
__device__ __forceinline__ double closure(const double* __restrict__ w, const Sys2& S, const ChunkCtx& c) {
    double xN = 0.0;
    for (int k = 0; k < S.K0; k++) xN += w[k * c.L + c.l];
    for (int k = c.T - S.K1; k < c.T; k++) xN += w[k * c.L + c.l];
    return xN;
}

// one system; returns B (the true solution at the start of the next chunk, up to the x_N part) for the caller
template <bool PER>
__device__ __forceinline__ void solve_one(double (&f)[C], const Sys2& S, const ChunkCtx& c, double* sm) {
    const int TL = c.T * c.L;
    double* y = sm;
    double* z = sm + TL;
    double* w = sm + 2 * TL;
    if (PER && S.circ) {
        double ye;
        local_const(f, S, ye);
        publish(&y[c.t * c.L + c.l], ye, c);
        exchange_barrier(c);
        const double A = look_back_circ(y, S, c);
        publish(&z[c.t * c.L + c.l], fma(S.cQ[0], A, f[0]), c);
        exchange_barrier(c);
        finish_const(f, S, A, look_ahead_circ(z, S, c));
        scale_rho(f, S, c.t);
        return;
    }
    const double2* cr = reinterpret_cast<const double2*>(S.crec) + c.t * 8;
    const double2 q0pp = ldg2(cr + 6);
    const double2 fl = ldg2(cr + 7);
    const bool isc = fl.x != 0.0;
    // non-periodic constant chunk in the unscaled variable w = x / rho (plan.cu): its chunk start goes out as x, the inflow B
    // comes back as w, the solution is scaled at the end
    const bool unsc = !PER && isc && S.rho != nullptr;
    const double2* tp = tab_ptr(S, c.t);
    double yend, part = 0.0;
    if (isc) local_const(f, S, yend);
    else local_tab<PER>(f, tp, yend, part);
    publish(&y[c.t * c.L + c.l], yend, c);
    exchange_barrier(c);
    const double A = look_back(y, cr, c);
    publish(&z[c.t * c.L + c.l], fma(q0pp.x, A, unsc ? f[0] * rho_first(S, c.t) : f[0]), c);
    if (PER) publish(&w[c.t * c.L + c.l], fma(q0pp.y, A, part), c);
    exchange_barrier(c);
    const double B = look_ahead(z, cr, c);
    if (isc) {
        finish_const(f, S, A, unsc ? B * fl.y : B);
        if (unsc) scale_rho(f, S, c.t);
    } else {
        const double xN = PER ? closure(w, S, c) : 0.0;
        finish_tab<PER>(f, tp, A, B, xN);
    }
}

// two independent systems sharing the barriers
template <bool PER>
__device__ __forceinline__ void solve_two(double (&f0)[C], double (&f1)[C], const Sys2& S0, const Sys2& S1,
                                          const ChunkCtx& c, double* sm) {
    const int TL = c.T * c.L;
    double* y0 = sm;
    double* z0 = sm + TL;
    double* w0 = sm + 2 * TL;
    double* y1 = sm + 3 * TL;
    double* z1 = sm + 4 * TL;
    double* w1 = sm + 5 * TL;
    if (PER && S0.circ && S1.circ) {
        const int me = c.t * c.L + c.l;
        double ye0, ye1;
        local_const2(f0, f1, S0, S1, ye0, ye1);
        publish(&y0[me], ye0, c);
        publish(&y1[me], ye1, c);
        exchange_barrier(c);
        const double A0 = look_back_circ(y0, S0, c), A1 = look_back_circ(y1, S1, c);
        publish(&z0[me], fma(S0.cQ[0], A0, f0[0]), c);
        publish(&z1[me], fma(S1.cQ[0], A1, f1[0]), c);
        exchange_barrier(c);
        const double B0 = look_ahead_circ(z0, S0, c), B1 = look_ahead_circ(z1, S1, c);
#pragma unroll
        for (int j = 0; j < C; j++) {
            f0[j] = fma(S0.cQ[j], A0, fma(S0.cR[j], B0, f0[j]));
            f1[j] = fma(S1.cQ[j], A1, fma(S1.cR[j], B1, f1[j]));
        }
        scale_rho(f0, S0, c.t);
        scale_rho(f1, S1, c.t);
        return;
    }
    const double2* cr0 = reinterpret_cast<const double2*>(S0.crec) + c.t * 8;
    const double2* cr1 = reinterpret_cast<const double2*>(S1.crec) + c.t * 8;
    const double2 q0 = ldg2(cr0 + 6), q1 = ldg2(cr1 + 6);
    const double2 fl0 = ldg2(cr0 + 7), fl1 = ldg2(cr1 + 7);
    const bool c0 = fl0.x != 0.0, c1 = fl1.x != 0.0;
    const bool u0 = !PER && c0 && S0.rho != nullptr, u1 = !PER && c1 && S1.rho != nullptr;      // see solve_one
    const double2* tp0 = tab_ptr(S0, c.t);
    const double2* tp1 = tab_ptr(S1, c.t);
    double ye0, ye1, p0 = 0.0, p1 = 0.0;
    if (c0 && c1) local_const2(f0, f1, S0, S1, ye0, ye1);
    else {
        if (c0) local_const(f0, S0, ye0); else local_tab<PER>(f0, tp0, ye0, p0);
        if (c1) local_const(f1, S1, ye1); else local_tab<PER>(f1, tp1, ye1, p1);
    }
    const int me = c.t * c.L + c.l;
    publish(&y0[me], ye0, c);
    publish(&y1[me], ye1, c);
    exchange_barrier(c);
    const double A0 = look_back(y0, cr0, c), A1 = look_back(y1, cr1, c);
    publish(&z0[me], fma(q0.x, A0, u0 ? f0[0] * rho_first(S0, c.t) : f0[0]), c);
    publish(&z1[me], fma(q1.x, A1, u1 ? f1[0] * rho_first(S1, c.t) : f1[0]), c);
    if (PER) { publish(&w0[me], fma(q0.y, A0, p0), c); publish(&w1[me], fma(q1.y, A1, p1), c); }
    exchange_barrier(c);
    double B0 = look_ahead(z0, cr0, c), B1 = look_ahead(z1, cr1, c);
    if (u0) B0 *= fl0.y;
    if (u1) B1 *= fl1.y;
    if (c0 && c1) {
#pragma unroll
        for (int j = 0; j < C; j++) {
            f0[j] = fma(S0.cQ[j], A0, fma(S0.cR[j], B0, f0[j]));
            f1[j] = fma(S1.cQ[j], A1, fma(S1.cR[j], B1, f1[j]));
        }
    } else {
        if (c0) finish_const(f0, S0, A0, B0);
        else finish_tab<PER>(f0, tp0, A0, B0, PER ? closure(w0, S0, c) : 0.0);
        if (c1) finish_const(f1, S1, A1, B1);
        else finish_tab<PER>(f1, tp1, A1, B1, PER ? closure(w1, S1, c) : 0.0);
    }
    if (u0) scale_rho(f0, S0, c.t);
    if (u1) scale_rho(f1, S1, c.t);
}

// Jacobian term of the second derivative on non-uniform grids (fdm_derivative.f90:437-440, `f2 += rhs_d1 * du` before the
// solve): the reference's lhs is A0 diag(dx1^2) and rhs_d1 = -A0 diag(dx2) (fdm_com2_jacobian.f90:263-274), so the term is a
// diagonal correction of the SOLUTION, d2u = A^-1 B u - (dx2/dx1^2) du (times the diffusivity for a scaled Burgers system).
// Both systems are therefore solved together and corrected afterwards (identical to round-off: 1e-15 against the oracle).
__device__ __forceinline__ void jacobian_correction(double (&d2)[C], const double (&d1)[C], const Line2Args& a, const Sys2& S2,
                                                    const ChunkCtx& c) {
    const double* cp = a.cjac + ((size_t)(c.t >> 3) * C) * 8 + (c.t & 7);
#pragma unroll
    for (int j = 0; j < C; j++) d2[j] = fma(-(S2.jscale * __ldg(cp + j * 8)), d1[j], d2[j]);
}

__host__ __device__ inline size_t exch2_doubles(int T, int L) { return (size_t)6 * T * L + (size_t)2 * T * L; }

// right-hand sides + solves: u (chunk + halos) -> d1, d2
template <int MODE, bool PER, bool NEED1>
__device__ __forceinline__ void line_core2(const double (&u)[C + 6], const Line2Args& a, const Sys2& S2, const ChunkCtx& c,
                                           double* sm, double (&d1)[C], double (&d2)[C]) {
    constexpr bool WANT1 = (MODE == MODE_P1) || (MODE == MODE_P2_P1) || (MODE == MODE_BURGERS) || (MODE == MODE_NEUMANN) || NEED1;
    constexpr bool WANT2 = (MODE == MODE_P2) || (MODE == MODE_P2_P1) || (MODE == MODE_BURGERS);
    if (WANT1) {
        rhs_interior<false>(u, d1, a.rhs1);
        if (!PER) {
            if (c.t == 0) rhs_bottom(u, d1, a.rhs1);
            if (c.t == c.T - 1) rhs_top(u, d1, a.rhs1);
        }
    }
    if (WANT2) {
        rhs_interior<true>(u, d2, a.rhs2);
        if (!PER) {
            if (c.t == 0) rhs_bottom(u, d2, a.rhs2);
            if (c.t == c.T - 1) rhs_top(u, d2, a.rhs2);
        }
    }
    if (WANT1 && WANT2) {
        solve_two<PER>(d1, d2, a.s1, S2, c, sm);
        if (NEED1) jacobian_correction(d2, d1, a, S2, c);
    } else {
        if (WANT1) solve_one<PER>(d1, a.s1, c, sm);
        if (WANT2) solve_one<PER>(d2, S2, c, sm + 3 * c.T * c.L);
    }
}

// ------------------------------------------------------------------------------------------------
// y / z directions: lines strided in memory, contiguous across lines.  grid = (inner / L, nlines / inner)
// One field of one tile: fu = field, fo = result (out1), S2 = second-derivative system, KEEPV: the velocity is read with
// the default cache policy (it is shared by the fields of a fused Burgers launch and should stay in L1).
template <int MODE, bool PER, bool NEED1, bool KEEPV>
__device__ __forceinline__ void strided_field(const Line2Args& a, const ChunkCtx& c, double* sm, const double* __restrict__ fu,
                                              double* __restrict__ fo, const Sys2& S2) {
    const long long st = a.stride;
    const long long tile0 = (long long)blockIdx.y * a.outer_stride + (long long)c.bx * a.L;
    const long long lbase = tile0 + c.l;
    const int n = a.n;
    const bool has_u2 = (a.u2 != nullptr);
    const bool has_acc = (a.accumulate != 0) && (MODE == MODE_BURGERS || MODE == MODE_P1);

    // ---- L2 prefetch of a later tile (one row segment per request); not for the Neumann kernel, which reads a few rows only
    if (a.pf_dist > 0 && MODE != MODE_NEUMANN) {
        const unsigned tile = blockIdx.y * gridDim.x + (unsigned)c.bx + (unsigned)a.pf_dist;
        if (tile < gridDim.x * gridDim.y) {
            const unsigned ty = tile / gridDim.x, tx = tile - ty * gridDim.x;
            const long long pb = (long long)ty * a.outer_stride + (long long)tx * a.L;
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                const long long o = pb + (long long)r * st;
                prefetch_l2(fu + o);
                if (has_u2) prefetch_l2(a.u2 + o);
                if (MODE == MODE_BURGERS && a.vel != fu) prefetch_l2(a.vel + o);
                if (has_acc) prefetch_l2(fo + o);
            }
        }
    }

    // ---- chunk + two 3-point halos (wrapped or zero)
    double u[C + 6];
    const long long coff = lbase + (long long)(c.t * C) * st;
    if (a.pf_l1 && (MODE == MODE_BURGERS || (MODE == MODE_P1 && has_acc))) {
        // the velocity and the accumulation target are needed after the solve: ask for them now (L1, no registers)
#pragma unroll
        for (int j = 0; j < C; j++) {
            if (MODE == MODE_BURGERS && a.vel != fu) prefetch_l1(a.vel + coff + j * st);
            if (has_acc) prefetch_l1(fo + coff + j * st);
        }
    }
    const bool lok = PER || c.t > 0, rok = PER || c.t < c.T - 1;
    const long long loff = (c.t > 0) ? -3 * st : (long long)(n - 3) * st;
    const long long roff = (c.t < c.T - 1) ? (long long)C * st : -(long long)(c.t * C) * st;
    // BOUNDARY_BCS_NEUMANN_Y needs the derivative next to the walls only, and the solution there feels the right-hand
    // side of the first LB2 chunks only (the same 2^-80 truncation as the look-back): the rest of the line is not read
    const bool skip = (MODE == MODE_NEUMANN) && !((a.bcs_hb != nullptr && c.t < LB2) || (a.bcs_ht != nullptr && c.t >= c.T - LB2));
    if (skip) {
#pragma unroll
        for (int k = 0; k < C + 6; k++) u[k] = 0.0;
    } else {
        const double* __restrict__ pc = fu + coff;
#pragma unroll
        for (int j = 0; j < C; j++) u[j + 3] = __ldcs(pc + j * st);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? __ldcs(pc + loff + k * st) : 0.0;
            u[C + 3 + k] = rok ? __ldcs(pc + roff + k * st) : 0.0;
        }
        if (has_u2) {
            const double* __restrict__ p2 = a.u2 + coff;
#pragma unroll
            for (int j = 0; j < C; j++) u[j + 3] = u[j + 3] + __ldcs(p2 + j * st) * a.scale;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (lok) u[k] = u[k] + __ldcs(p2 + loff + k * st) * a.scale;
                if (rok) u[C + 3 + k] = u[C + 3 + k] + __ldcs(p2 + roff + k * st) * a.scale;
            }
        }
    }
    double nb_sum = 0.0, nt_sum = 0.0;
    if (MODE == MODE_NEUMANN) {
#pragma unroll
        for (int k = 0; k < BROW_W; k++) {
            nb_sum = nb_sum + a.neu_bot[k] * u[3 + k];
            nt_sum = nt_sum + a.neu_top[k] * u[3 + C - 1 - k];
        }
    }
    double d1[C], d2[C];
    line_core2<MODE, PER, NEED1>(u, a, S2, c, sm, d1, d2);

    if (MODE == MODE_NEUMANN) {
        // boundary values such that the normal derivative vanishes (BOUNDARY_BCS_NEUMANN_Y)
        const long long line = (long long)blockIdx.y * a.inner + (long long)c.bx * a.L + c.l;
        if (c.t == 0 && a.bcs_hb != nullptr) a.bcs_hb[line] = nb_sum + a.neu_lu_bot * d1[1];
        if (c.t == c.T - 1 && a.bcs_ht != nullptr) a.bcs_ht[line] = nt_sum + a.neu_lu_top * d1[C - 2];
        return;
    }
    double* __restrict__ o1 = fo + coff;
    if (MODE == MODE_BURGERS) {
        const double* __restrict__ vp = a.vel + coff;
        double vv[C];
#pragma unroll
        for (int j = 0; j < C; j++) vv[j] = KEEPV ? __ldg(vp + j * st) : __ldcs(vp + j * st);
        if (has_acc) {
            double oo[C];
#pragma unroll
            for (int j = 0; j < C; j++) oo[j] = __ldcs(o1 + j * st);
#pragma unroll
            for (int j = 0; j < C; j++) {
                const double r = d2[j] - vv[j] * d1[j];
                d2[j] = (a.accumulate > 0) ? oo[j] + r : oo[j] - r;
            }
        } else {
#pragma unroll
            for (int j = 0; j < C; j++) d2[j] = d2[j] - vv[j] * d1[j];
        }
    } else if (MODE == MODE_P1 && has_acc) {
#pragma unroll
        for (int j = 0; j < C; j++) {
            const double o = __ldcs(o1 + j * st);
            d1[j] = (a.accumulate > 0) ? o + d1[j] : o - d1[j];
        }
    }
#pragma unroll
    for (int j = 0; j < C; j++) {
        if (MODE == MODE_P1) __stcs(o1 + j * st, d1[j]);
        if (MODE == MODE_P2 || MODE == MODE_BURGERS) __stcs(o1 + j * st, d2[j]);
        if (MODE == MODE_P2_P1) { __stcs(o1 + j * st, d2[j]); __stcs(a.out2 + coff + j * st, d1[j]); }
    }
}

template <int MODE, bool PER, bool NEED1>
__global__ void __launch_bounds__(512, 1) lines2_strided(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = threadIdx.x >> a.lshift;
    c.bx = blockIdx.x;
    if (MODE == MODE_NEUMANN && a.neu_nb + a.neu_nt > 0) {
        // only the chunks next to the walls are resident (the wall derivative feels nothing else, see strided_field): thread
        // row t' -> chunk t' at the bottom, T - nt + (t' - nb) at the top; the exchange slots of the absent chunks read as zero
        for (int i = threadIdx.x; i < 2 * a.T * a.L; i += blockDim.x) sm[i] = 0.0;
        __syncthreads();
        if (c.t >= a.neu_nb) c.t = a.T - a.neu_nt + (c.t - a.neu_nb);
    }
    strided_field<MODE, PER, NEED1, false>(a, c, sm, a.u, a.out1, a.s2);
}

// Long periodic lines (T > 32 chunks, z at C3): a CTA of 512 threads can hold only 8 lines of 64 chunks, i.e. 64-byte row
// segments, which costs a fifth of the bandwidth (burgers_z 47 % against 58 % of the peak with 128-byte rows, measured).
// Here a line is shared by a thread-block cluster of 2 CTAs, 16 lines x 32 chunks each (128-byte rows): the chunk ends
// are published into both CTAs' exchange areas through distributed shared memory and the two block barriers of a solve
// become cluster barriers; everything else is strided_field.  Periodic, uniform directions only (no Jacobian term).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) lines2_strided_pair(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = (threadIdx.x >> a.lshift) + (int)rank * (a.T / 2);
    c.bx = blockIdx.x >> 1;
    c.pair = true;
    // generic address of the peer's exchange area (same offset in its shared-memory window)
    c.rsm = cluster.map_shared_rank(sm, rank ^ 1u);
    c.sm0 = sm;
    cluster.sync();             // both CTAs are running: their shared memory may be written from now on
    strided_field<MODE, true, false, false>(a, c, sm, a.u, a.out1, a.s2);
}

// fused Burgers launch: the fields fu[0..nf) of a tile are advected by the same velocity (OPR_Burgers_X/Y/Z of u, v, w
// and the scalars in RHS_GLOBAL_INCOMPRESSIBLE_1): the velocity comes from DRAM once per tile instead of once per field
template <bool PER, bool NEED1>
__global__ void __launch_bounds__(512, 1) lines2_strided_multi(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = threadIdx.x >> a.lshift;
    c.bx = blockIdx.x;
    for (int f = 0; f < a.nf; f++) {
        if (f > 0) __syncthreads();                 // the exchange areas of the previous field are free again
        if (a.pf_next && f + 1 < a.nf) {
            // the next field of this tile goes to L2 while this one is being solved
            const long long tile0 = (long long)blockIdx.y * a.outer_stride + (long long)blockIdx.x * a.L;
            const bool two = (a.L * 8 > 128);       // rows longer than one 128-byte line
            for (int r = threadIdx.x; r < a.n; r += blockDim.x) {
                const long long o = tile0 + (long long)r * a.stride;
                prefetch_l2(a.fu[f + 1] + o);
                prefetch_l2(a.fo[f + 1] + o);
                if (two) { prefetch_l2(a.fu[f + 1] + o + 16); prefetch_l2(a.fo[f + 1] + o + 16); }
            }
        }
        const Sys2& S2 = a.fsys[f] ? a.s2b : a.s2;
        strided_field<MODE_BURGERS, PER, NEED1, true>(a, c, sm, a.fu[f], a.fo[f], S2);
    }
}

// ------------------------------------------------------------------------------------------------
// y / z directions, persistent variant with asynchronous staging.  Each CTA loops over tiles of L lines.  The field
// (and the velocity, or the second input) of a tile is brought into shared memory with cp.async, 16 bytes = 2 adjacent
// lines per request; a tile is consumed into registers right after it has arrived, so the buffer is free again and
// the copies of the NEXT tile are in flight while this one is being solved: the DRAM latency is paid behind the
// arithmetic instead of in front of it, without a second buffer.  The accumulation target is prefetched into L2 at
// the start of the tile and read directly at the end.
// Tile layout: point i of line l at (i/16) * CS + (i%16) * L + l, CS = 16 L + 8: the chunk reads of a warp
// (L lines x 32/L chunks) are conflict-free.
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ inline int pa_chunk_stride(int L) { return C * L + 8; }

__device__ __forceinline__ void pa_issue_tile(double* tile, const double* __restrict__ g, long long st, int n, int lshift, int CS) {
    const int sh = lshift - 1;                     // log2 of the 16-byte pieces per row
    const int ppr = 1 << sh;
    const int total = n << sh;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        const int i = q >> sh, lp = (q & (ppr - 1)) * 2;
        cp_async16(tile + (i >> 4) * CS + ((i & 15) << lshift) + lp, g + (long long)i * st + lp);
    }
}

template <int MODE, bool PER, bool NEED1>
__global__ void __launch_bounds__(512, 1) lines2_strided_pa(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = threadIdx.x >> a.lshift;
    const long long st = a.stride;
    const int n = a.n, L = a.L, T = a.T;
    const int CS = pa_chunk_stride(L);
    double* bufU = sm + exch2_doubles(T, L);
    double* bufV = bufU + (size_t)T * CS;
    const bool has_u2 = (a.u2 != nullptr);
    const bool has_vel = (MODE == MODE_BURGERS) && (a.vel != a.u);
    const bool has_acc = (a.accumulate != 0) && (MODE == MODE_BURGERS || MODE == MODE_P1);
    const double* __restrict__ vsrc = has_u2 ? a.u2 : a.vel;
    const bool stage_v = has_u2 || has_vel;
    const unsigned ntiles = a.ntiles, tiles_x = a.tiles_x;

    auto tile_base = [&](unsigned tile) {
        const unsigned ty = tile / tiles_x, tx = tile - ty * tiles_x;
        return (long long)ty * a.outer_stride + (long long)tx * L;
    };
    unsigned tile = blockIdx.x;
    if (tile < ntiles) {
        const long long gb = tile_base(tile);
        pa_issue_tile(bufU, a.u + gb, st, n, a.lshift, CS);
        cp_async_commit();
        if (stage_v) pa_issue_tile(bufV, vsrc + gb, st, n, a.lshift, CS);
        cp_async_commit();
    }
    for (; tile < ntiles; tile += gridDim.x) {
        const long long gb = tile_base(tile);
        const unsigned next = tile + gridDim.x;
        const long long gn = (next < ntiles) ? tile_base(next) : 0;
        const long long coff = gb + c.l + (long long)(c.t * C) * st;
        if (has_acc) {
            // L2 prefetch of this tile's accumulation target (read at the end of the iteration)
            for (int r = threadIdx.x; r < n; r += blockDim.x) prefetch_l2(a.out1 + gb + (long long)r * st);
        }
        cp_async_wait<1>();                        // the field of this tile has arrived (the velocity may still be in flight)
        __syncthreads();
        double u[C + 6];
        {
            const bool lok = PER || c.t > 0, rok = PER || c.t < T - 1;
            const double* pc = bufU + c.t * CS + c.l;
            const double* pl = bufU + ((c.t > 0) ? (c.t - 1) : (T - 1)) * CS + (C - 3) * L + c.l;
            const double* pr = bufU + ((c.t < T - 1) ? (c.t + 1) : 0) * CS + c.l;
#pragma unroll
            for (int j = 0; j < C; j++) u[j + 3] = pc[j * L];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                u[k] = lok ? pl[k * L] : 0.0;
                u[C + 3 + k] = rok ? pr[k * L] : 0.0;
            }
        }
        if (has_u2) {
            cp_async_wait<0>();
            __syncthreads();
            const bool lok = PER || c.t > 0, rok = PER || c.t < T - 1;
            const double* qc = bufV + c.t * CS + c.l;
            const double* ql = bufV + ((c.t > 0) ? (c.t - 1) : (T - 1)) * CS + (C - 3) * L + c.l;
            const double* qr = bufV + ((c.t < T - 1) ? (c.t + 1) : 0) * CS + c.l;
#pragma unroll
            for (int j = 0; j < C; j++) u[j + 3] = u[j + 3] + qc[j * L] * a.scale;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (lok) u[k] = u[k] + ql[k * L] * a.scale;
                if (rok) u[C + 3 + k] = u[C + 3 + k] + qr[k * L] * a.scale;
            }
        }
        __syncthreads();                            // every thread has taken its chunk and halos: bufU (and bufV for u2) are free
        if (next < ntiles) pa_issue_tile(bufU, a.u + gn, st, n, a.lshift, CS);
        cp_async_commit();
        if (has_u2) {
            if (next < ntiles) pa_issue_tile(bufV, vsrc + gn, st, n, a.lshift, CS);
            cp_async_commit();
        }
        if (MODE == MODE_BURGERS && !has_vel) {
            // SELF: the advecting velocity is the field itself; park this thread's chunk in its own slots of bufV
            double* vq = bufV + c.t * CS + c.l;
#pragma unroll
            for (int j = 0; j < C; j++) vq[j * L] = u[j + 3];
        }
        double d1[C], d2[C];
        line_core2<MODE, PER, NEED1>(u, a, a.s2, c, sm, d1, d2);

        double* __restrict__ o1 = a.out1 + coff;
        if (MODE == MODE_BURGERS) {
            if (has_vel) {
                cp_async_wait<1>();                // this tile's velocity (older than the next tile's field) has arrived
                __syncthreads();
            }
            const double* vq = bufV + c.t * CS + c.l;
#pragma unroll
            for (int j = 0; j < C; j++) d2[j] = d2[j] - vq[j * L] * d1[j];
            if (has_vel) {
                __syncthreads();                    // bufV is free again
                if (next < ntiles) pa_issue_tile(bufV, vsrc + gn, st, n, a.lshift, CS);
                cp_async_commit();
            }
            if (has_acc) {
                double oo[C];
#pragma unroll
                for (int j = 0; j < C; j++) oo[j] = __ldcs(o1 + j * st);
#pragma unroll
                for (int j = 0; j < C; j++) d2[j] = (a.accumulate > 0) ? oo[j] + d2[j] : oo[j] - d2[j];
            }
        } else if (MODE == MODE_P1 && has_acc) {
            double oo[C];
#pragma unroll
            for (int j = 0; j < C; j++) oo[j] = __ldcs(o1 + j * st);
#pragma unroll
            for (int j = 0; j < C; j++) d1[j] = (a.accumulate > 0) ? oo[j] + d1[j] : oo[j] - d1[j];
        }
#pragma unroll
        for (int j = 0; j < C; j++) {
            if (MODE == MODE_P1) __stcs(o1 + j * st, d1[j]);
            if (MODE == MODE_P2 || MODE == MODE_BURGERS) __stcs(o1 + j * st, d2[j]);
            if (MODE == MODE_P2_P1) { __stcs(o1 + j * st, d2[j]); __stcs(a.out2 + coff + j * st, d1[j]); }
        }
        if (!stage_v) cp_async_commit();          // keep two groups per iteration (the wait counts above assume it)
        // exchange areas of line_core2 are reused by the next tile: its first barrier (after cp_async_wait) orders them
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// y / z directions, persistent CTAs fed by the TMA unit.  Nothing of the field traffic goes through the LSU: a tile of
// L lines (n rows of L*8 bytes, rows `stride` apart) is brought into shared memory by cp.async.bulk.tensor (boxes of
// L x RB points, completion on an mbarrier), the result tile is staged in shared memory and leaves through a bulk
// tensor store -- a reduce-add store when the operator accumulates (hq += ..., hq -= dp/dx), so the accumulation target
// is never loaded by the SM: the read-modify-write happens in L2.  Buffers: U (field), V (velocity or second input, or
// the parked field of a SELF call), R (result).  Per tile:
//     wait U [and V: second input] -> registers | barrier S1 | TMA U(next) [V(next)] | solve | wait V: velocity |
//     combine -> R | fence.proxy.async, barrier S2 | TMA store R, TMA V(next)
// so the loads of the next tile are in flight while this one is solved and stored.  Tile layout in shared memory is
// dense, point i of line l at i*L + l (what a box copy produces).  With L = 8 a row is 64 bytes = half of the banks and
// the four chunks of a warp start 16 rows apart: odd chunks visit the rows of a pair in swapped order, so that every
// access of a warp covers all banks once per 128 bytes.
struct TmaMaps {
    CUtensorMap u, v, o;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
template <bool THREE>
__device__ __forceinline__ void tma_load_box(unsigned dst, const CUtensorMap* m, unsigned bar, int c0, int c1, int c2) {
    if (THREE)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
template <bool THREE, bool ADD>
__device__ __forceinline__ void tma_store_box(const CUtensorMap* m, unsigned src, int c0, int c1, int c2) {
    if (THREE) {
        if (ADD) asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                              ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                          ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
    } else {
        if (ADD) asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                              ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
        else asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                          ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
    }
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// warp 0: one tile of one array -> shared memory (n / RB boxes, one per lane), completion on `bar`
template <bool THREE>
__device__ __forceinline__ void tma_tile_in(const CUtensorMap* m, double* buf, unsigned bar, int x0, int outer, int n, int L, int RB) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(n * L) * 8u);
    __syncwarp();
    for (int r = lane * RB; r < n; r += 32 * RB) tma_load_box<THREE>(smem_u32(buf + (size_t)r * L), m, bar, x0, r, outer);
}
template <bool THREE>
__device__ __forceinline__ void tma_tile_out(const CUtensorMap* m, const double* buf, bool add, int x0, int outer, int n, int L, int RB) {
    const int lane = threadIdx.x & 31;
    for (int r = lane * RB; r < n; r += 32 * RB) {
        if (add) tma_store_box<THREE, true>(m, smem_u32(buf + (size_t)r * L), x0, r, outer);
        else tma_store_box<THREE, false>(m, smem_u32(buf + (size_t)r * L), x0, r, outer);
    }
    bulk_commit();
}

// the 16 points of a chunk <-> registers.  A row of the tile is L*8 bytes; with L = 8 (4) it covers a half (a quarter) of
// the banks and the chunks of a warp start 16 rows apart, i.e. in the same banks: chunk t visits the rows of each aligned
// pair (quad) in the order j ^ (t & 1) (j ^ (t & 3)), so that a warp access covers all banks once per 128 bytes; the values
// are put back in order with conditional swaps.
template <int LL>
__device__ __forceinline__ void xor_permute(double* v, int s) {
    if (LL <= 8) {
#pragma unroll
        for (int k = 0; k < C; k += 2) {
            const double a = v[k], b = v[k + 1];
            v[k] = (s & 1) ? b : a;
            v[k + 1] = (s & 1) ? a : b;
        }
    }
    if (LL == 4) {
#pragma unroll
        for (int k = 0; k < C; k += 4) {
            const double a = v[k], b = v[k + 1], c = v[k + 2], d = v[k + 3];
            v[k] = (s & 2) ? c : a;
            v[k + 1] = (s & 2) ? d : b;
            v[k + 2] = (s & 2) ? a : c;
            v[k + 3] = (s & 2) ? b : d;
        }
    }
}
template <int LL>
__device__ __forceinline__ void chunk_get(const double* buf, int t, int l, double* v) {
    const int s = (LL == 8) ? (t & 1) : ((LL == 4) ? (t & 3) : 0);
    const double* pc = buf + (size_t)(t * C) * LL + l;
#pragma unroll
    for (int k = 0; k < C; k++) v[k] = pc[(k ^ s) * LL];
    xor_permute<LL>(v, s);
}
template <int LL>
__device__ __forceinline__ void chunk_put(double* buf, int t, int l, const double* v) {
    const int s = (LL == 8) ? (t & 1) : ((LL == 4) ? (t & 3) : 0);
    double* pc = buf + (size_t)(t * C) * LL + l;
    double w[C];
#pragma unroll
    for (int k = 0; k < C; k++) w[k] = v[k];
    xor_permute<LL>(w, s);
#pragma unroll
    for (int k = 0; k < C; k++) pc[(k ^ s) * LL] = w[k];
}

template <int MODE, bool PER, bool NEED1, int LL, bool THREE>
__global__ void __launch_bounds__(512, 1) lines2_strided_tma(const __grid_constant__ Line2Args a, const __grid_constant__ TmaMaps maps) {
    extern __shared__ __align__(128) double smt[];
    ChunkCtx c;
    c.L = LL; c.T = a.T;
    c.l = threadIdx.x & (LL - 1);
    c.t = threadIdx.x / LL;
    const int n = a.n, T = a.T, RB = a.tma_rb;
    double* bufU = smt + exch2_doubles(T, LL);
    double* bufV = bufU + (size_t)n * LL;
    double* bufR = bufV + (size_t)n * LL;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(bufR + (size_t)n * LL);
    const unsigned barU = smem_u32(bars), barV = smem_u32(bars + 1);
    const bool has_u2 = (a.u2 != nullptr);
    const bool has_vel = (MODE == MODE_BURGERS) && (a.vel != a.u);
    const bool has_acc = (a.accumulate != 0) && (MODE == MODE_BURGERS || MODE == MODE_P1);
    const bool warp0 = threadIdx.x < 32;
    const unsigned ntiles = a.ntiles, tiles_x = a.tiles_x;

    if (threadIdx.x == 0) {
        mbar_init(barU, 1);
        mbar_init(barV, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned tile = blockIdx.x;
    if (warp0 && tile < ntiles) {
        const unsigned ty = tile / tiles_x, tx = tile - ty * tiles_x;
        tma_tile_in<THREE>(&maps.u, bufU, barU, (int)tx * LL, (int)ty, n, LL, RB);
        if (has_u2 || has_vel) tma_tile_in<THREE>(&maps.v, bufV, barV, (int)tx * LL, (int)ty, n, LL, RB);
    }
    unsigned parity = 0;
    for (; tile < ntiles; tile += gridDim.x, parity ^= 1u) {
        const unsigned ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const unsigned next = tile + gridDim.x;
        const bool more = next < ntiles;
        const unsigned nty = more ? next / tiles_x : 0, ntx = more ? next - nty * tiles_x : 0;

        // ---- this tile's chunk + halos: shared memory -> registers
        double u[C + 6];
        const bool lok = PER || c.t > 0, rok = PER || c.t < T - 1;
        const int rl = ((c.t > 0) ? c.t * C : n) - 3;             // first row of the left halo
        const int rr = (c.t < T - 1) ? (c.t + 1) * C : 0;         // first row of the right halo
        mbar_wait(barU, parity);
        chunk_get<LL>(bufU, c.t, c.l, u + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? bufU[(size_t)(rl + k) * LL + c.l] : 0.0;
            u[C + 3 + k] = rok ? bufU[(size_t)(rr + k) * LL + c.l] : 0.0;
        }
        if (has_u2) {
            mbar_wait(barV, parity);
            double w[C];
            chunk_get<LL>(bufV, c.t, c.l, w);
#pragma unroll
            for (int j = 0; j < C; j++) u[j + 3] = u[j + 3] + w[j] * a.scale;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (lok) u[k] = u[k] + bufV[(size_t)(rl + k) * LL + c.l] * a.scale;
                if (rok) u[C + 3 + k] = u[C + 3 + k] + bufV[(size_t)(rr + k) * LL + c.l] * a.scale;
            }
        }
        if (warp0) bulk_wait_read0();              // the store of the previous tile has left R
        __syncthreads();                            // S1: U (and V with a second input) consumed, R free
        if (warp0 && more) {
            tma_tile_in<THREE>(&maps.u, bufU, barU, (int)ntx * LL, (int)nty, n, LL, RB);
            if (has_u2) tma_tile_in<THREE>(&maps.v, bufV, barV, (int)ntx * LL, (int)nty, n, LL, RB);
        }
        if (MODE == MODE_BURGERS && !has_vel) chunk_put<LL>(bufV, c.t, c.l, u + 3);   // SELF: park the advecting field

        double d1[C], d2[C];
        line_core2<MODE, PER, NEED1>(u, a, a.s2, c, smt, d1, d2);

        if (MODE == MODE_BURGERS) {
            if (has_vel) mbar_wait(barV, parity);
            double vv[C];
            chunk_get<LL>(bufV, c.t, c.l, vv);
#pragma unroll
            for (int j = 0; j < C; j++) d2[j] = d2[j] - vv[j] * d1[j];
        }
        if (has_acc && a.accumulate < 0) {
#pragma unroll
            for (int j = 0; j < C; j++) { d1[j] = -d1[j]; d2[j] = -d2[j]; }
        }
        chunk_put<LL>(bufR, c.t, c.l, (MODE == MODE_P1) ? d1 : d2);
        fence_async_smem();
        __syncthreads();                            // S2: R complete, V consumed
        if (warp0) {
            tma_tile_out<THREE>(&maps.o, bufR, has_acc, (int)tx * LL, (int)ty, n, LL, RB);
            if (has_vel && more) tma_tile_in<THREE>(&maps.v, bufV, barV, (int)ntx * LL, (int)nty, n, LL, RB);
        }
    }
    if (warp0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------------
// x direction: lines contiguous in memory; the tile of L lines (L*n contiguous doubles) is staged through shared
// memory.  Tile layout: line ll at ll*T*XB, point i at (i>>4)*XB + (i&15): 16-byte accesses are conflict-free both for
// the coalesced side (a lane owns 2 consecutive points) and for the chunk side (a lane owns 16 consecutive points).
// Each thread moves exactly 8 double2 per array (L in {1,2,4,8}).
template <bool ADD2>
__device__ __forceinline__ void tile_load(double* tile, const double* __restrict__ g, const double* __restrict__ g2,
                                          double scale, int n, int T, int L, int LS) {
    const int nth = L * T;
    const int per_line = (L == 8) ? 1 : (L == 4 ? 2 : (L == 2 ? 4 : 8));   // passes per line
    const int tid = threadIdx.x;
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
        v[k] = __ldcs(reinterpret_cast<const double2*>(g + (size_t)ll * n) + i2);
    }
    if (ADD2) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
            const double2 w = __ldcs(reinterpret_cast<const double2*>(g2 + (size_t)ll * n) + i2);
            v[k].x = v[k].x + w.x * scale;
            v[k].y = v[k].y + w.y * scale;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
        *reinterpret_cast<double2*>(tile + ll * LS + (i2 >> 3) * XB + (i2 & 7) * 2) = v[k];
    }
}

// One field of one tile.  MULTI: fused Burgers launch, the velocity tile has been staged by the caller.
// two tiles at once: all 16 loads are in flight before the first shared-memory store
__device__ __forceinline__ void tile_load_pair(double* tile, double* tile2, const double* __restrict__ g, const double* __restrict__ g2,
                                               int n, int T, int L, int LS) {
    const int nth = L * T;
    const int per_line = (L == 8) ? 1 : (L == 4 ? 2 : (L == 2 ? 4 : 8));
    const int tid = threadIdx.x;
    double2 v[8], w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
        v[k] = __ldcs(reinterpret_cast<const double2*>(g + (size_t)ll * n) + i2);
        w[k] = __ldcs(reinterpret_cast<const double2*>(g2 + (size_t)ll * n) + i2);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
        const int o = ll * LS + (i2 >> 3) * XB + (i2 & 7) * 2;
        *reinterpret_cast<double2*>(tile + o) = v[k];
        *reinterpret_cast<double2*>(tile2 + o) = w[k];
    }
}

template <int MODE, bool PER, bool NEED1, bool MULTI>
__device__ __forceinline__ void contig_field(const Line2Args& a, const ChunkCtx& c, double* sm, const double* __restrict__ fu,
                                             double* __restrict__ fo, const Sys2& S2) {
    const int n = a.n, T = a.T, L = a.L, LS = a.xls;
    const size_t tile_off = (size_t)blockIdx.x * L * n;
    double* tile = sm + exch2_doubles(T, L);
    double* vtile = tile + (size_t)L * LS;
    const bool two = MULTI || ((MODE == MODE_BURGERS) && (a.vel != fu));
    const bool has_acc = (a.accumulate != 0) && (MODE == MODE_BURGERS || MODE == MODE_P1);

    if (a.pf_dist > 0) {
        const unsigned tilei = blockIdx.x + (unsigned)a.pf_dist;
        if (tilei < gridDim.x) {
            const size_t po = (size_t)tilei * L * n;
            for (int r = threadIdx.x * 16; r < L * n; r += blockDim.x * 16) {      // one request per 128-byte line
                prefetch_l2(fu + po + r);
                if (a.u2 != nullptr) prefetch_l2(a.u2 + po + r);
                if (two && !MULTI) prefetch_l2(a.vel + po + r);
                if (has_acc) prefetch_l2(fo + po + r);
            }
        }
    }
    if (a.u2 != nullptr) tile_load<true>(tile, fu + tile_off, a.u2 + tile_off, a.scale, n, T, L, LS);
    else if (two && !MULTI) tile_load_pair(tile, vtile, fu + tile_off, a.vel + tile_off, n, T, L, LS);
    else tile_load<false>(tile, fu + tile_off, nullptr, 0.0, n, T, L, LS);
    if (two && !MULTI && a.u2 != nullptr) tile_load<false>(vtile, a.vel + tile_off, nullptr, 0.0, n, T, L, LS);
    __syncthreads();

    const double* row = tile + c.l * LS;
    double u[C + 6];
    {
        const double* pc = row + c.t * XB;
#pragma unroll
        for (int j = 0; j < C; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(pc + j);
            u[j + 3] = v.x; u[j + 4] = v.y;
        }
        const bool lok = PER || c.t > 0, rok = PER || c.t < T - 1;
        const double* pl = row + ((c.t > 0) ? c.t - 1 : T - 1) * XB + (C - 3);
        const double* pr = row + ((c.t < T - 1) ? c.t + 1 : 0) * XB;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? pl[k] : 0.0;
            u[C + 3 + k] = rok ? pr[k] : 0.0;
        }
    }
    double d1[C], d2[C];
    line_core2<MODE, PER, NEED1>(u, a, S2, c, sm, d1, d2);
    // all halo reads of the tile happened before the first barrier inside line_core2: results may overwrite it

    const int nth = L * T;
    const int per_line = (L == 8) ? 1 : (L == 4 ? 2 : (L == 2 ? 4 : 8));
    const int tid = threadIdx.x;
    double2 acc[8];
    if (has_acc) {
        // in flight while the results are staged and the barrier is crossed
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
            acc[k] = __ldcs(reinterpret_cast<const double2*>(fo + tile_off + (size_t)ll * n) + i2);
        }
    }
    {
        double* wrow = tile + c.l * LS + c.t * XB;
        const double* vrow = (two ? vtile : tile) + c.l * LS + c.t * XB;
#pragma unroll
        for (int j = 0; j < C; j += 2) {
            double2 r;
            if (MODE == MODE_P1) { r.x = d1[j]; r.y = d1[j + 1]; }
            else if (MODE == MODE_P2 || MODE == MODE_P2_P1) { r.x = d2[j]; r.y = d2[j + 1]; }
            else {
                const double2 v = *reinterpret_cast<const double2*>(vrow + j);
                r.x = d2[j] - v.x * d1[j];
                r.y = d2[j + 1] - v.y * d1[j + 1];
            }
            *reinterpret_cast<double2*>(wrow + j) = r;
            if (MODE == MODE_P2_P1) {
                double2 q; q.x = d1[j]; q.y = d1[j + 1];
                *reinterpret_cast<double2*>(vtile + c.l * LS + c.t * XB + j) = q;
            }
        }
    }
    __syncthreads();
    {
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
            v[k] = *reinterpret_cast<const double2*>(tile + ll * LS + (i2 >> 3) * XB + (i2 & 7) * 2);
        }
        if (has_acc) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double2 o = acc[k];
                if (a.acc_scale != 1.0) { o.x = DMUL(a.acc_scale, o.x); o.y = DMUL(a.acc_scale, o.y); }   // hq = hq*kco, then the sum
                if (a.accumulate > 0) { v[k].x = o.x + v[k].x; v[k].y = o.y + v[k].y; }
                else { v[k].x = o.x - v[k].x; v[k].y = o.y - v[k].y; }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
            __stcs(reinterpret_cast<double2*>(fo + tile_off + (size_t)ll * n) + i2, v[k]);
        }
    }
    if (MODE == MODE_P2_P1) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ll = (k * L) >> 3, i2 = tid + (k & (per_line - 1)) * nth;
            const double2 v = *reinterpret_cast<const double2*>(vtile + ll * LS + (i2 >> 3) * XB + (i2 & 7) * 2);
            __stcs(reinterpret_cast<double2*>(a.out2 + tile_off + (size_t)ll * n) + i2, v);
        }
    }
}

template <int MODE, bool PER, bool NEED1>
__global__ void __launch_bounds__(512, 1) lines2_contig(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = threadIdx.x >> a.lshift;
    contig_field<MODE, PER, NEED1, false>(a, c, sm, a.u, a.out1, a.s2);
}

// fused Burgers launch along x (see lines2_strided_multi): the velocity tile is staged once per tile
template <bool PER, bool NEED1>
__global__ void __launch_bounds__(512, 1) lines2_contig_multi(const __grid_constant__ Line2Args a) {
    extern __shared__ double sm[];
    ChunkCtx c;
    c.L = a.L; c.T = a.T;
    c.l = threadIdx.x & (a.L - 1);
    c.t = threadIdx.x >> a.lshift;
    {
        double* vtile = sm + exch2_doubles(a.T, a.L) + (size_t)a.L * a.xls;
        const size_t tile_off = (size_t)blockIdx.x * a.L * a.n;
        if (a.pf_dist > 0 && blockIdx.x + (unsigned)a.pf_dist < gridDim.x) {
            const size_t po = (size_t)(blockIdx.x + (unsigned)a.pf_dist) * a.L * a.n;
            for (int r = threadIdx.x * 16; r < a.L * a.n; r += blockDim.x * 16) prefetch_l2(a.vel + po + r);
        }
        tile_load<false>(vtile, a.vel + tile_off, nullptr, 0.0, a.n, a.T, a.L, a.xls);
    }
    for (int f = 0; f < a.nf; f++) {
        if (f > 0) __syncthreads();                 // tile and exchange areas of the previous field are free again
        if (a.pf_next && f + 1 < a.nf) {
            const size_t to = (size_t)blockIdx.x * a.L * a.n;
            for (int r = threadIdx.x * 16; r < a.L * a.n; r += blockDim.x * 16) {
                prefetch_l2(a.fu[f + 1] + to + r);
                prefetch_l2(a.fo[f + 1] + to + r);
            }
        }
        const Sys2& S2 = a.fsys[f] ? a.s2b : a.s2;
        contig_field<MODE_BURGERS, PER, NEED1, true>(a, c, sm, a.fu[f], a.fo[f], S2);
    }
}

// prefetch distance: the number of CTAs resident on the device (the tile that far ahead starts when this one ends)
template <class K>
int auto_pf_dist(K k, int threads, size_t smem) {
    int occ = 0, dev = 0, sms = 148;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return std::max(occ, 1) * sms;
}

long long g_tma_launches = 0;

// ---- tensor maps of the TMA kernel: one per (array, geometry), cached (the fields of a run live at fixed addresses)
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tensor_map_encoder() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
        cudaGetLastError();
    }
    return fn;
}

size_t tma_smem_bytes(int T, int L, int n) { return (exch2_doubles(T, L) + (size_t)3 * n * L) * sizeof(double) + 64; }

// field viewed as (inner, n, nouter) with strides (1, stride, outer_stride); box = L lines x rb rows
bool tensor_map_for(const double* base, const Line2Args& a, long long nouter, CUtensorMap* out) {
    if (!base) { std::memset(out, 0, sizeof(*out)); return true; }
    using Key = std::tuple<const void*, long long, int, long long, long long, long long, int, int, int>;
    static std::map<Key, CUtensorMap> cache;
    const Key key{base, a.inner, a.n, nouter, a.stride, a.outer_stride, a.L, a.tma_rb, a.tma_l2};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    const bool three = nouter > 1;
    cuuint64_t gdim[3] = {(cuuint64_t)a.inner, (cuuint64_t)a.n, (cuuint64_t)nouter};
    cuuint64_t gstr[2] = {(cuuint64_t)a.stride * 8, (cuuint64_t)a.outer_stride * 8};
    cuuint32_t box[3] = {(cuuint32_t)a.L, (cuuint32_t)a.tma_rb, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, three ? 3 : 2, const_cast<double*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           a.tma_l2 == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (a.tma_l2 == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B :
                           (a.tma_l2 == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE)),
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
    *out = m;
    return true;
}

template <int MODE, bool PER, bool NEED1, int LL, bool THREE>
cudaError_t launch2_tma_k(const Line2Args& a, const TmaMaps& maps, cudaStream_t stream) {
    auto k = lines2_strided_tma<MODE, PER, NEED1, LL, THREE>;
    const size_t smem = tma_smem_bytes(a.T, LL, a.n);
    static size_t set = 0;
    if (smem > set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        set = smem;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // CTAs per SM: as many as the shared memory of this geometry allows (1 for 512-thread tiles)
    static int occ = 0, occ_threads = 0;
    static size_t occ_smem = 0;
    const int threads = LL * a.T;
    if (occ_threads != threads || occ_smem != smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
        occ = std::max(occ, 1);
        occ_threads = threads; occ_smem = smem;
    }
    const unsigned g = std::min<unsigned>(a.ntiles, (unsigned)(occ * sms));
    k<<<g, threads, smem, stream>>>(a, maps);
    return cudaGetLastError();
}

// returns cudaErrorNotSupported when the tensor maps cannot be built (the caller then uses the LSU kernels)
template <int MODE, bool PER, bool NEED1>
cudaError_t launch2_tma(const Line2Args& a_in, dim3 grid, cudaStream_t stream) {
    Line2Args a = a_in;
    a.tiles_x = grid.x;
    a.ntiles = grid.x * grid.y;
    const long long nouter = grid.y;
    TmaMaps maps;
    const double* second = a.u2 ? a.u2 : ((MODE == MODE_BURGERS && a.vel != a.u) ? a.vel : nullptr);
    if (!tensor_map_for(a.u, a, nouter, &maps.u) || !tensor_map_for(second, a, nouter, &maps.v) ||
        !tensor_map_for(a.out1, a, nouter, &maps.o))
        return cudaErrorNotSupported;
    const bool three = nouter > 1;
    g_tma_launches++;
    if (a.L == 4) return three ? launch2_tma_k<MODE, PER, NEED1, 4, true>(a, maps, stream) : launch2_tma_k<MODE, PER, NEED1, 4, false>(a, maps, stream);
    if (a.L == 32) return three ? launch2_tma_k<MODE, PER, NEED1, 32, true>(a, maps, stream) : launch2_tma_k<MODE, PER, NEED1, 32, false>(a, maps, stream);
    if (a.L == 8) return three ? launch2_tma_k<MODE, PER, NEED1, 8, true>(a, maps, stream) : launch2_tma_k<MODE, PER, NEED1, 8, false>(a, maps, stream);
    return three ? launch2_tma_k<MODE, PER, NEED1, 16, true>(a, maps, stream) : launch2_tma_k<MODE, PER, NEED1, 16, false>(a, maps, stream);
}

template <int MODE, bool PER, bool NEED1>
cudaError_t launch2(const Line2Args& a_in, bool contig, dim3 grid, cudaStream_t stream) {
    Line2Args a = a_in;
    const int threads = (MODE == MODE_NEUMANN && a.neu_nb + a.neu_nt > 0) ? a.L * (a.neu_nb + a.neu_nt) : a.L * a.T;
    if (!contig && a.tma && (MODE == MODE_P1 || MODE == MODE_P2 || MODE == MODE_BURGERS)) {
        constexpr int M = (MODE == MODE_P1 || MODE == MODE_P2 || MODE == MODE_BURGERS) ? MODE : MODE_P1;
        const cudaError_t e = launch2_tma<M, PER, NEED1>(a, grid, stream);
        if (e != cudaErrorNotSupported) return e;
    }
    size_t smem = exch2_doubles(a.T, a.L) * sizeof(double);
    if (contig) {
        const bool two = (MODE == MODE_P2_P1) || ((MODE == MODE_BURGERS) && (a.vel != a.u));
        smem += (size_t)a.L * a.xls * sizeof(double) * (two ? 2 : 1);
        auto k = lines2_contig<MODE, PER, NEED1>;
        static size_t set = 0;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        static int pf = 0, pf_threads = 0;
        static size_t pf_smem = 0;
        if (a.pf_dist < 0) {
            if (pf_threads != threads || pf_smem != smem) { pf = auto_pf_dist(k, threads, smem); pf_threads = threads; pf_smem = smem; }
            a.pf_dist = pf;
        }
        k<<<grid, threads, smem, stream>>>(a);
    } else if (a.persist && MODE != MODE_NEUMANN) {
        auto k = lines2_strided_pa<(MODE == MODE_NEUMANN ? MODE_P1 : MODE), PER, NEED1>;
        smem += (size_t)2 * a.T * pa_chunk_stride(a.L) * sizeof(double);
        static size_t set = 48 * 1024;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        static int ctas = 0, c_threads = 0;
        static size_t c_smem = 0;
        if (c_threads != threads || c_smem != smem) { ctas = auto_pf_dist(k, threads, smem); c_threads = threads; c_smem = smem; }
        a.tiles_x = grid.x;
        a.ntiles = grid.x * grid.y;
        const unsigned g = std::min<unsigned>(a.ntiles, (unsigned)ctas);
        k<<<g, threads, smem, stream>>>(a);
    } else if (a.pair && PER && !NEED1 && (MODE == MODE_P1 || MODE == MODE_P2 || MODE == MODE_BURGERS)) {
        constexpr int M = (MODE == MODE_P1 || MODE == MODE_P2 || MODE == MODE_BURGERS) ? MODE : MODE_P1;
        auto k = lines2_strided_pair<M>;
        static size_t set = 48 * 1024;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        a.pf_dist = 0;
        k<<<dim3(grid.x * 2, grid.y, 1), threads / 2, smem, stream>>>(a);
    } else {
        auto k = lines2_strided<MODE, PER, NEED1>;
        static size_t set = 48 * 1024;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        static int pf = 0, pf_threads = 0;
        static size_t pf_smem = 0;
        if (a.pf_dist < 0) {
            if (pf_threads != threads || pf_smem != smem) { pf = auto_pf_dist(k, threads, smem); pf_threads = threads; pf_smem = smem; }
            a.pf_dist = pf;
        }
        k<<<grid, threads, smem, stream>>>(a);
    }
    return cudaGetLastError();
}

template <bool PER, bool NEED1>
cudaError_t launch2_multi(const Line2Args& a_in, bool contig, dim3 grid, cudaStream_t stream) {
    Line2Args a = a_in;
    const int threads = a.L * a.T;
    size_t smem = exch2_doubles(a.T, a.L) * sizeof(double);
    if (contig) {
        smem += (size_t)a.L * a.xls * sizeof(double) * 2;
        auto k = lines2_contig_multi<PER, NEED1>;
        static size_t set = 0;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        static int pf = 0, pf_threads = 0;
        static size_t pf_smem = 0;
        if (a.pf_dist < 0) {
            if (pf_threads != threads || pf_smem != smem) { pf = auto_pf_dist(k, threads, smem); pf_threads = threads; pf_smem = smem; }
            a.pf_dist = pf;
        }
        k<<<grid, threads, smem, stream>>>(a);
    } else {
        auto k = lines2_strided_multi<PER, NEED1>;
        static size_t set = 48 * 1024;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        static int pf = 0, pf_threads = 0;
        static size_t pf_smem = 0;
        if (a.pf_dist < 0) {
            if (pf_threads != threads || pf_smem != smem) { pf = auto_pf_dist(k, threads, smem); pf_threads = threads; pf_smem = smem; }
            a.pf_dist = pf;
        }
        k<<<grid, threads, smem, stream>>>(a);
    }
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch2_mode(const Line2Args& a, bool per, bool need1, bool contig, dim3 grid, cudaStream_t s) {
    if (per) return launch2<MODE, true, false>(a, contig, grid, s);
    if (need1) return launch2<MODE, false, true>(a, contig, grid, s);
    return launch2<MODE, false, false>(a, contig, grid, s);
}

}  // namespace

int lines2_xstride(int T, int L) {
    int LS = T * XB;
    const int want = (L >= 16) ? 0 : (16 / L) & 15;      // LS mod 16 == 16/L: the L lines of a quarter-warp hit distinct banks
    while ((LS & 15) != want) LS += 2;
    return LS;
}

bool lines2_eligible(const DevPlan& p, const Sys2& s1, const Sys2* s2, int n, long long nlines, long long inner, bool contig,
                     int L_override, int* L_out) {
    if (n < CHUNK || n % CHUNK != 0 || p.crem != 0 || p.cbase != CHUNK) return false;
    if (!s1.ok || (s2 && !s2->ok)) return false;
    const int T = n / CHUNK;
    if (T > 128) return false;
    int L;
    if (contig) {
        L = 4;
        if (L_override == 1 || L_override == 2 || L_override == 4 || L_override == 8) L = L_override;
        while (L < 8 && L * T < 64) L <<= 1;
        while (L > 1 && (L * T > 512 || nlines % L != 0)) L >>= 1;
        if (((size_t)L * n) % 2) return false;
    } else {
        L = 16;                                        // 128-byte rows when the CTA stays within 512 threads
        if (L_override == 4 || L_override == 8 || L_override == 16 || L_override == 32) L = L_override;
        while (L < 32 && L * T < 128 && inner % (2 * L) == 0) L <<= 1;
        while (L > 4 && (L * T > 512 || inner % L != 0)) L >>= 1;
        if (L * T > 512 || inner % L != 0 || nlines % inner != 0) return false;
        if (inner / L > 0x7fffffffLL || nlines / inner > 65535) return false;
    }
    if (L * T < 1) return false;
    *L_out = L;
    return true;
}

bool lines2_tma_eligible(int mode, const Line2Args& a) {
    if (mode != MODE_P1 && mode != MODE_P2 && mode != MODE_BURGERS) return false;
    if (a.L != 4 && a.L != 8 && a.L != 16 && a.L != 32) return false;
    if (a.u2 && mode == MODE_BURGERS) return false;
    if (a.L * a.T > 512 || a.n % CHUNK != 0) return false;
    auto al = [](const void* q) { return (reinterpret_cast<size_t>(q) & 15) == 0; };
    if (!(al(a.u) && al(a.u2) && al(a.vel) && al(a.out1))) return false;
    if (a.stride % 2 != 0 || a.outer_stride % 2 != 0 || a.inner % a.L != 0) return false;
    if (a.inner >= (1LL << 31) || (a.stride * 8) >= (1LL << 40) || (a.outer_stride * 8) >= (1LL << 40)) return false;
    if (tma_smem_bytes(a.T, a.L, a.n) > 227 * 1024) return false;
    return tensor_map_encoder() != nullptr;
}

long long lines2_tma_launches() { return g_tma_launches; }

size_t lines2_persist_smem(int T, int L) { return (exch2_doubles(T, L) + (size_t)2 * T * pa_chunk_stride(L)) * sizeof(double); }

cudaError_t launch_lines2(int mode, const Line2Args& a, bool periodic, bool need1, bool contig, long long nlines,
                          long long inner, cudaStream_t s) {
    dim3 grid;
    if (contig) grid = dim3((unsigned)(nlines / a.L), 1, 1);
    else grid = dim3((unsigned)(inner / a.L), (unsigned)(nlines / inner), 1);
    if (mode == MODE_BURGERS && a.nf > 0) {
        if (periodic) return launch2_multi<true, false>(a, contig, grid, s);
        if (need1) return launch2_multi<false, true>(a, contig, grid, s);
        return launch2_multi<false, false>(a, contig, grid, s);
    }
    switch (mode) {
        case MODE_P1: return launch2_mode<MODE_P1>(a, periodic, false, contig, grid, s);
        case MODE_P2: return launch2_mode<MODE_P2>(a, periodic, need1, contig, grid, s);
        case MODE_P2_P1: return launch2_mode<MODE_P2_P1>(a, periodic, need1, contig, grid, s);
        case MODE_BURGERS: return launch2_mode<MODE_BURGERS>(a, periodic, need1, contig, grid, s);
        case MODE_NEUMANN: return launch2_mode<MODE_NEUMANN>(a, periodic, false, false, grid, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace tlab
