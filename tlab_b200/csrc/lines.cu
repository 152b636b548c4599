// Batched compact-scheme line operators for sm_100a: banded right-hand side + factored tridiagonal
// solve, fused (first derivative, second derivative, both, or the Burgers operator nu*s'' - u*s')
// in one pass over the pencil: every data array is read once and written once.
//
// Replaces, per call, the reference chain  TLab_Transpose -> g%matmul -> TRIDSS/TRIDPSS -> TLab_Transpose
// (src/operators/opr_partial.f90:31-377, src/physics/opr_burgers.f90:190-521,
//  src/fdm/fdm_derivative.f90:218-278,413-459, src/fdm/fdm_matmul.f90, src/utils/linear3.f90:56-150,321-442).
//
// Parallel decomposition ("chunked substitution"):  a line of n points is owned by T threads,
// each holding a chunk of <= CHUNK consecutive points in registers.  Thomas' forward and backward
// substitutions are first-order linear recurrences  y_i = s_i + m_i y_{i-1};  each thread first runs its
// chunk with zero inflow, publishes the chunk-end value to shared memory, rebuilds its true inflow from
// the preceding chunks' ends and their precomputed multiplier products (look-back window W chosen on the
// host so that the dropped factor is < 2^-80), and then repeats the recurrence with the true inflow --
// the same operations, in the same order, as the sequential algorithm.
//
// Thread mapping: tid = l + L*t, l = line within the CTA tile (fastest, so that a warp touches L
// consecutive lines x 32/L chunks), t = chunk.  For the y and z directions lines are contiguous in l and
// global accesses are coalesced directly into registers.  For the x direction (lines contiguous in
// memory) the tile is staged through padded shared memory.
#include "lines.h"
#include <cstdio>
#include <algorithm>

namespace tlab {

namespace {

struct Chunk {
    int t, T;        // chunk index, chunks per line
    int s0, cnt;     // first point, number of points
    int j0;          // register slot of the first point (the last chunk is right-aligned: its end is slot CHUNK-1)
};

__device__ __forceinline__ Chunk make_chunk(int t, int T, int cbase, int crem) {
    Chunk c;
    c.t = t; c.T = T;
    c.s0 = t * cbase + min(t, crem);
    c.cnt = cbase + (t < crem ? 1 : 0);
    c.j0 = (t == T - 1) ? CHUNK - c.cnt : 0;
    return c;
}

__device__ __forceinline__ double ldro(const double* p) { return __ldg(p); }
__device__ __forceinline__ double2 ldro2(const double2* p) { return __ldg(p); }

// ------------------------------------------------------------------------------------------------
// NS factored tridiagonal systems solved together on register chunks (shared barriers, independent
// dependency chains).  Exchange area per system: sm_f[T*L], sm_b[T*L], sm_p[T*L], sm_q[(T8+1)*L].
__host__ __device__ inline int exch_per_system(int T, int L) { return 3 * T * L + (((T + 7) >> 3) + 1) * L; }

template <bool PER, bool FULL, int NS>
__device__ __forceinline__ void solve_tri(double (&f)[NS][CHUNK], const Chunk& c, const SolveTab* S, int l, int L,
                                          double* sm) {
    const int j1 = c.j0 + c.cnt;
    const int TL = c.T * L;
    const int T8 = (c.T + 7) >> 3;
    const int per_sys = exch_per_system(c.T, L);
    // record of register slot j: rp[j * 2*REC_GROUP] = {alpha, beta}, rp[j * 2*REC_GROUP + 1] = {gamma, delta|pe}
    const size_t rbase = ((size_t)(c.t / REC_GROUP) * CHUNK) * REC_GROUP + c.t % REC_GROUP;
    constexpr int RS = 2 * REC_GROUP;
    ChunkDesc cd[NS];
    const double2* rp[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        cd[s] = S[s].cd[c.t];
        rp[s] = S[s].rec + 2 * rbase;
    }

    // ---- forward substitution: local recurrence with zero inflow
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double e = 0.0;
        if (cd[s].flags & CD_FWD_CONST) {
            const double a = cd[s].a, b = cd[s].b;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                if (PER) e = f[s][j] * b + a * e;
                else e = f[s][j] + a * e;
                f[s][j] = e;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                if (FULL || (j >= c.j0 && j < j1)) {
                    const double2 ab = ldro2(rp[s] + j * RS);
                    if (PER) e = f[s][j] * ab.y + ab.x * e;
                    else e = f[s][j] + ab.x * e;
                    f[s][j] = e;
                }
            }
        }
        sm[s * per_sys + c.t * L + l] = e;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const double* sm_f = sm + s * per_sys;
        double cin = 0.0;
        for (int k = max(0, c.t - S[s].Wf); k < c.t; k++) cin = sm_f[k * L + l] + ldro(&S[s].cd[k].Af) * cin;
        if (c.t > 0) {
            // inflow correction: corr_j = alpha_j corr_{j-1}, corr_{-1} = inflow
            double corr = cin;
            if (cd[s].flags & CD_FWD_CONST) {
#pragma unroll
                for (int j = 0; j < CHUNK; j++) { corr = cd[s].a * corr; f[s][j] = f[s][j] + corr; }
            } else {
#pragma unroll
                for (int j = 0; j < CHUNK; j++)
                    if (FULL || (j >= c.j0 && j < j1)) { corr = ldro(&rp[s][j * RS].x) * corr; f[s][j] = f[s][j] + corr; }
            }
        }
    }
    double xN[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) xN[s] = 0.0;
    if (PER) {
        // ---- rank-one closure of the circulant system: x_N = (y_N - sum d_i y_i) * b_N
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double part = 0.0;
            if (!(cd[s].flags & CD_PD_ZERO)) {
                const double* pdp = S[s].pd + rbase;
#pragma unroll
                for (int j = 0; j < CHUNK; j++)
                    if (FULL || (j >= c.j0 && j < j1)) part = part + ldro(pdp + j * REC_GROUP) * f[s][j];
            }
            double* sm_p = sm + s * per_sys + 2 * TL;
            double* sm_q = sm_p + TL;
            sm_p[c.t * L + l] = part;
            if (c.t == c.T - 1) sm_q[T8 * L + l] = f[s][CHUNK - 1];
        }
        __syncthreads();
        if (c.t < T8) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                double* sm_p = sm + s * per_sys + 2 * TL;
                double* sm_q = sm_p + TL;
                double q = 0.0;
                for (int k = c.t * 8; k < min(c.T, c.t * 8 + 8); k++) q = q + sm_p[k * L + l];
                sm_q[c.t * L + l] = q;
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const double* sm_q = sm + s * per_sys + 3 * TL;
            double wrk = 0.0;
            for (int k = 0; k < T8; k++) wrk = wrk + sm_q[k * L + l];
            xN[s] = (sm_q[T8 * L + l] - wrk) * S[s].bN;
            if (c.t == c.T - 1) f[s][CHUNK - 1] = xN[s];
        }
    }
    // ---- backward substitution
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double e = 0.0;
        if (cd[s].flags & CD_BWD_CONST) {
            const double g = cd[s].g, d = cd[s].d;
#pragma unroll
            for (int j = CHUNK - 1; j >= 0; j--) {
                if (PER) e = f[s][j] + g * e;
                else e = (f[s][j] + g * e) * d;
                f[s][j] = e;
            }
        } else {
#pragma unroll
            for (int j = CHUNK - 1; j >= 0; j--) {
                if (FULL || (j >= c.j0 && j < j1)) {
                    const double2 gd = ldro2(rp[s] + j * RS + 1);
                    if (PER) e = (f[s][j] + gd.x * e) + gd.y * xN[s];
                    else e = (f[s][j] + gd.x * e) * gd.y;
                    f[s][j] = e;
                }
            }
        }
        sm[s * per_sys + TL + c.t * L + l] = e;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const double* sm_b = sm + s * per_sys + TL;
        double cin = 0.0;
        for (int k = min(c.T - 1, c.t + S[s].Wb); k > c.t; k--) cin = sm_b[k * L + l] + ldro(&S[s].cd[k].Ab) * cin;
        if (c.t < c.T - 1) {
            double corr = cin;
            if (cd[s].flags & CD_BWD_CONST) {
                const double m = PER ? cd[s].g : cd[s].g * cd[s].d;
#pragma unroll
                for (int j = CHUNK - 1; j >= 0; j--) { corr = m * corr; f[s][j] = f[s][j] + corr; }
            } else {
#pragma unroll
                for (int j = CHUNK - 1; j >= 0; j--) {
                    if (FULL || (j >= c.j0 && j < j1)) {
                        const double2 gd = ldro2(rp[s] + j * RS + 1);
                        corr = (PER ? gd.x : gd.x * gd.y) * corr;
                        f[s][j] = f[s][j] + corr;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// banded right-hand sides from the chunk + 3-point halo  (u[k] holds point s0 - 3 + (k - j0))
// The right-hand sides are sums with heavy cancellation (B u = O(h^2 u'')), so they are evaluated with
// explicitly rounded multiplies and adds in the reference's association order (no FMA contraction):
// a differently-rounded B u would be amplified by 1/(h k)^2 in the result.
#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

template <bool SECOND>
__device__ __forceinline__ void rhs_interior(const double (&u)[CHUNK + 6], double (&f)[CHUNK], const RhsTab& R) {
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        if (SECOND) {
            // r4*u(n) + u(n+1) + u(n-1) + r6*(u(n+2) + u(n-2)) + r7*(u(n+3) + u(n-3)), fdm_matmul.f90:608-612
            double s = DADD(DADD(DMUL(R.rc, u[j + 3]), u[j + 4]), u[j + 2]);
            s = DADD(s, DMUL(R.r2, DADD(u[j + 5], u[j + 1])));
            if (R.r3 != 0.0) s = DADD(s, DMUL(R.r3, DADD(u[j + 6], u[j])));
            f[j] = s;
        } else {
            // u(n+1) - u(n-1) + r5*(u(n+2) - u(n-2)), fdm_matmul.f90:396-398
            double s = DSUB(u[j + 4], u[j + 2]);
            if (R.r2 != 0.0) s = DADD(s, DMUL(R.r2, DSUB(u[j + 5], u[j + 1])));
            f[j] = s;
        }
    }
}

// special rows at the two ends; wb[k] = u_k (k from the bottom), wt[k] = u_{n-1-k}.  Zero coefficients
// are skipped so that the sum runs over the same terms, in the same order, as the reference's rows.
__device__ __forceinline__ void rhs_bottom(const double (&wb)[BROW_W], double (&f)[CHUNK], const RhsTab& R) {
#pragma unroll
    for (int i = 0; i < MAX_BROWS; i++) {
        if (i < R.nb) {
            double s = 0.0;
            bool first = true;
#pragma unroll
            for (int k = 0; k < BROW_W; k++) {
                const double cf = R.bot[i][k];
                if (cf != 0.0) {
                    const double term = DMUL(cf, wb[k]);
                    s = first ? term : DADD(s, term);
                    first = false;
                }
            }
            f[i] = s;
        }
    }
}
__device__ __forceinline__ void rhs_top(const double (&wt)[BROW_W], double (&f)[CHUNK], const RhsTab& R) {
#pragma unroll
    for (int q = 0; q < MAX_BROWS; q++) {
        if (q < R.nb) {
            double s = 0.0;
            bool first = true;
#pragma unroll
            for (int k = BROW_W - 1; k >= 0; k--) {      // ascending column order, as the reference
                const double cf = R.top[q][k];
                if (cf != 0.0) {
                    const double term = DMUL(cf, wt[k]);
                    s = first ? term : DADD(s, term);
                    first = false;
                }
            }
            f[CHUNK - 1 - q] = s;
        }
    }
}

// CompactDirect6 second derivative (fdm_comx_direct.f90:305-412): pentadiagonal rhs with per-row coefficients, MatMul_5d
// (fdm_matmul.f90:266-320) term by term -- ascending columns, the first upper diagonal of the interior rows (5 .. n-4) is 1
// by normalisation and not multiplied, the fourth coefficient of the first / last row sits in column 1 / 5 of its row.
__device__ __forceinline__ void rhs_direct(const double (&u)[CHUNK + 6], double (&f)[CHUNK], const double* __restrict__ rows,
                                           const Chunk& c, int n) {
    const int base = c.s0 - c.j0;
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        const int i = base + j;
        if (j < c.j0 || j >= c.j0 + c.cnt) continue;
        const double* r = rows + 5 * (size_t)i;
        const double r1 = ldro(r), r2 = ldro(r + 1), r3 = ldro(r + 2), r4 = ldro(r + 3), r5 = ldro(r + 4);
        const double um3 = u[j], um2 = u[j + 1], um1 = u[j + 2], u0 = u[j + 3], up1 = u[j + 4], up2 = u[j + 5], up3 = u[j + 6];
        double s;
        if (i == 0) s = DADD(DADD(DADD(DMUL(u0, r3), DMUL(up1, r4)), DMUL(up2, r5)), DMUL(up3, r1));
        else if (i == 1) s = DADD(DADD(DADD(DMUL(um1, r2), DMUL(u0, r3)), DMUL(up1, r4)), DMUL(up2, r5));
        else if (i == n - 1) s = DADD(DADD(DADD(DMUL(um3, r5), DMUL(um2, r1)), DMUL(um1, r2)), DMUL(u0, r3));
        else if (i == n - 2) s = DADD(DADD(DADD(DMUL(um2, r1), DMUL(um1, r2)), DMUL(u0, r3)), DMUL(up1, r4));
        else if (i < 4 || i > n - 5) s = DADD(DADD(DADD(DADD(DMUL(um2, r1), DMUL(um1, r2)), DMUL(u0, r3)), DMUL(up1, r4)), DMUL(up2, r5));
        else s = DADD(DADD(DADD(DADD(DMUL(um2, r1), DMUL(um1, r2)), DMUL(u0, r3)), up1), DMUL(up2, r5));
        f[j] = s;
    }
}

// Jacobian correction of the second derivative on non-uniform grids:  f2 += A2*jac2 * du  (tridiagonal,
// extended stencil in the first and last row)
__device__ __forceinline__ void add_jacobian_term(double (&f2)[CHUNK], const double (&d1)[CHUNK], const Chunk& c,
                                                  const double* __restrict__ rd1, int n, int l, int L, double* sm_h) {
    const int j1 = c.j0 + c.cnt;
    double first = 0.0, last = 0.0;
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        if (j == c.j0) first = d1[j];
        if (j == j1 - 1) last = d1[j];
    }
    sm_h[(2 * c.t) * L + l] = first;
    sm_h[(2 * c.t + 1) * L + l] = last;
    __syncthreads();
    const double left = (c.t > 0) ? sm_h[(2 * (c.t - 1) + 1) * L + l] : 0.0;
    const double right = (c.t < c.T - 1) ? sm_h[(2 * (c.t + 1)) * L + l] : 0.0;
    const int base = c.s0 - c.j0;
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        if (j >= c.j0 && j < j1) {
            const int i = base + j;
            double um = (j == c.j0) ? left : d1[j > 0 ? j - 1 : 0];
            double up = (j == j1 - 1) ? right : d1[j < CHUNK - 1 ? j + 1 : CHUNK - 1];
            if (i == 0) um = d1[j < CHUNK - 2 ? j + 2 : CHUNK - 1];           // extended stencil, first row
            if (i == n - 1) up = d1[j > 1 ? j - 2 : 0];                       // extended stencil, last row
            const double r1 = ldro(rd1 + 3 * i), r2 = ldro(rd1 + 3 * i + 1), r3 = ldro(rd1 + 3 * i + 2);
            if (i == n - 1) f2[j] = DADD(DADD(DADD(f2[j], DMUL(up, r3)), DMUL(um, r1)), DMUL(d1[j], r2));
            else f2[j] = DADD(DADD(DADD(f2[j], DMUL(um, r1)), DMUL(d1[j], r2)), DMUL(up, r3));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory exchange area (doubles): two systems' buffers + [sm_h 2*T*L] for the Jacobian-term halo
__host__ __device__ inline size_t exch_doubles(int T, int L) {
    return (size_t)2 * exch_per_system(T, L) + (size_t)2 * T * L;
}

template <int MODE, bool PER, bool NEED1, bool FULL>
__device__ __forceinline__ void line_core(double (&u)[CHUNK + 6], const double (&wb)[BROW_W], const double (&wt)[BROW_W],
                                          const Chunk& c, const LineArgs& a, int l, int L, double* sm,
                                          double (&d)[2][CHUNK]) {
    // d[0]: first derivative, d[1]: second derivative
    double* sm_h = sm + 2 * exch_per_system(c.T, L);
    constexpr bool WANT1 = (MODE == MODE_P1) || (MODE == MODE_P2_P1) || (MODE == MODE_BURGERS) ||
                           (MODE == MODE_NEUMANN) || NEED1;
    constexpr bool WANT2 = (MODE == MODE_P2) || (MODE == MODE_P2_P1) || (MODE == MODE_BURGERS);
    if (WANT1) {
        rhs_interior<false>(u, d[0], a.rhs1);
        if (!PER) {
            if (c.t == 0) rhs_bottom(wb, d[0], a.rhs1);
            if (c.t == c.T - 1) rhs_top(wt, d[0], a.rhs1);
        }
    }
    if (WANT2) {
        if (!PER && a.rhs2_rows != nullptr) {
            rhs_direct(u, d[1], a.rhs2_rows, c, a.n);
        } else {
            rhs_interior<true>(u, d[1], a.rhs2);
            if (!PER) {
                if (c.t == 0) rhs_bottom(wb, d[1], a.rhs2);
                if (c.t == c.T - 1) rhs_top(wt, d[1], a.rhs2);
            }
        }
    }
    if (WANT1 && WANT2 && !NEED1) {
        const SolveTab S[2] = {a.lu1, a.lu2};
        solve_tri<PER, FULL, 2>(d, c, S, l, L, sm);
    } else {
        if (WANT1) {
            double (&d1)[1][CHUNK] = reinterpret_cast<double (&)[1][CHUNK]>(d[0]);
            solve_tri<PER, FULL, 1>(d1, c, &a.lu1, l, L, sm);
        }
        if (WANT2) {
            if (NEED1) add_jacobian_term(d[1], d[0], c, a.rhs_d1, a.n, l, L, sm_h);
            double (&d2)[1][CHUNK] = reinterpret_cast<double (&)[1][CHUNK]>(d[1]);
            solve_tri<PER, FULL, 1>(d2, c, &a.lu2, l, L, sm);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// y / z directions: lines strided in memory, contiguous across lines
template <int MODE, bool PER, bool NEED1, bool FULL>
__global__ void __launch_bounds__(512) line_kernel_strided(LineArgs a) {
    extern __shared__ double sm[];
    const int L = a.L;
    const int l = threadIdx.x % L;
    const int t = threadIdx.x / L;
    const Chunk c = make_chunk(t, a.T, a.cbase, a.crem);
    long long line = (long long)blockIdx.x * L + l;
    const bool active = line < a.nlines;
    if (!active) line = a.nlines - 1;
    const long long lbase = (line / a.inner) * a.outer_stride + (line % a.inner);
    const double* __restrict__ up = a.u + lbase;
    const long long st = a.stride;
    const int n = a.n;

    double u[CHUNK + 6];
    if (FULL) {
        // every chunk holds CHUNK points: 16 loads at a fixed stride plus two 3-point halos (wrapped or zero)
        const double* __restrict__ pc = up + (long long)c.s0 * st;
        const double* __restrict__ p2 = (a.u2 != nullptr) ? a.u2 + lbase + (long long)c.s0 * st : nullptr;
        const bool lok = PER || c.t > 0, rok = PER || c.t < c.T - 1;
        const long long loff = (c.t > 0) ? -3 * st : (long long)(n - 3 - c.s0) * st;
        const long long roff = (c.t < c.T - 1) ? (long long)CHUNK * st : -(long long)c.s0 * st;
#pragma unroll
        for (int j = 0; j < CHUNK; j++) u[j + 3] = __ldcs(pc + j * st);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? __ldcs(pc + loff + k * st) : 0.0;
            u[CHUNK + 3 + k] = rok ? __ldcs(pc + roff + k * st) : 0.0;
        }
        if (p2 != nullptr) {
#pragma unroll
            for (int j = 0; j < CHUNK; j++) u[j + 3] = u[j + 3] + __ldcs(p2 + j * st) * a.scale;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (lok) u[k] = u[k] + __ldcs(p2 + loff + k * st) * a.scale;
                if (rok) u[CHUNK + 3 + k] = u[CHUNK + 3 + k] + __ldcs(p2 + roff + k * st) * a.scale;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CHUNK + 6; k++) {
            int p = c.s0 - 3 + (k - c.j0);
            const bool in_win = (k >= c.j0) && (k < c.j0 + c.cnt + 6);
            if (PER) p = (p < 0) ? p + n : (p >= n ? p - n : p);
            const bool ok = in_win && p >= 0 && p < n;
            u[k] = ok ? __ldcs(up + (long long)p * st) : 0.0;
            if (a.u2 != nullptr && ok) u[k] = u[k] + __ldcs(a.u2 + lbase + (long long)p * st) * a.scale;
        }
    }
    double wb[BROW_W], wt[BROW_W];
    if (!PER) {
#pragma unroll
        for (int k = 0; k < BROW_W; k++) {
            wb[k] = (c.t == 0) ? u[3 + k] : 0.0;
            wt[k] = (c.t == c.T - 1) ? u[3 + CHUNK - 1 - k] : 0.0;
        }
    }
    double dd[2][CHUNK];
    line_core<MODE, PER, NEED1, FULL>(u, wb, wt, c, a, l, L, sm, dd);
    double (&d1)[CHUNK] = dd[0];
    double (&d2)[CHUNK] = dd[1];

    const int j1 = c.j0 + c.cnt;
    if (MODE == MODE_NEUMANN) {
        // boundary values such that the normal derivative vanishes (BOUNDARY_BCS_NEUMANN_Y)
        if (active && c.t == 0 && a.bcs_hb != nullptr) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < BROW_W; k++) s = s + a.neu_bot[k] * wb[k];
            a.bcs_hb[line] = s + a.neu_lu_bot * d1[1];
        }
        if (active && c.t == c.T - 1 && a.bcs_ht != nullptr) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < BROW_W; k++) s = s + a.neu_top[k] * wt[k];
            a.bcs_ht[line] = s + a.neu_lu_top * d1[CHUNK - 2];
        }
        return;
    }
    double* __restrict__ o1 = a.out1 + lbase + (long long)(c.s0 - c.j0) * st;
    double* __restrict__ o2 = (MODE == MODE_P2_P1) ? a.out2 + lbase + (long long)(c.s0 - c.j0) * st : nullptr;
    const double* __restrict__ vp = (MODE == MODE_BURGERS) ? a.vel + lbase + (long long)(c.s0 - c.j0) * st : nullptr;
    if (!active) return;
    double vv[CHUNK];
    if (MODE == MODE_BURGERS) {
#pragma unroll
        for (int j = 0; j < CHUNK; j++)
            if (FULL || (j >= c.j0 && j < j1)) vv[j] = __ldcs(vp + j * st);
        if (a.accumulate != 0) {
#pragma unroll
            for (int j = 0; j < CHUNK; j++)
                if (FULL || (j >= c.j0 && j < j1)) {
                    const double r = d2[j] - vv[j] * d1[j];
                    vv[j] = __ldcs(o1 + j * st);
                    d2[j] = (a.accumulate > 0) ? vv[j] + r : vv[j] - r;
                }
        } else {
#pragma unroll
            for (int j = 0; j < CHUNK; j++)
                if (FULL || (j >= c.j0 && j < j1)) d2[j] = d2[j] - vv[j] * d1[j];
        }
    } else if (MODE == MODE_P1 && a.accumulate != 0) {
#pragma unroll
        for (int j = 0; j < CHUNK; j++)
            if (FULL || (j >= c.j0 && j < j1)) {
                const double o = __ldcs(o1 + j * st);
                d1[j] = (a.accumulate > 0) ? o + d1[j] : o - d1[j];
            }
    }
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        if (FULL || (j >= c.j0 && j < j1)) {
            if (MODE == MODE_P1) __stcs(o1 + j * st, d1[j]);
            if (MODE == MODE_P2 || MODE == MODE_BURGERS) __stcs(o1 + j * st, d2[j]);
            if (MODE == MODE_P2_P1) { __stcs(o1 + j * st, d2[j]); __stcs(o2 + j * st, d1[j]); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// y / z directions, persistent variant with asynchronous prefetch (full chunks, whole tiles only).
// Each CTA loops over tiles of L lines.  The pencils of a tile are brought into shared memory with
// cp.async (16-byte pieces, coalesced over the L contiguous lines); while a tile is being solved in
// registers, the next tile's input and this tile's velocity / accumulation target are already in flight,
// so that the DRAM latency is paid behind the arithmetic instead of in front of it.
// Tile layout in shared memory: row i of line l at (i / 16) * CS + (i % 16) * L + l, CS = 16 L + pad, which
// makes the chunk reads of a warp (lanes = L lines x 32/L chunks) bank-conflict free.
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ inline int pf_chunk_stride(int L) { return CHUNK * L + (L >= 16 ? 0 : L); }

__device__ __forceinline__ void pf_issue_tile(double* tile, const double* __restrict__ g, long long st, int n, int L, int CS) {
    const int ppr = L >> 1;                        // 16-byte pieces per row (L is a power of two)
    const int sh = __ffs(ppr) - 1;
    const int total = n * ppr;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        const int i = q >> sh, lp = (q & (ppr - 1)) * 2;
        cp_async16(tile + (i >> 4) * CS + (i & 15) * L + lp, g + (long long)i * st + lp);
    }
}

template <int MODE, bool PER, bool NEED1>
__global__ void __launch_bounds__(512) line_kernel_strided_pf(LineArgs a) {
    extern __shared__ double sm[];
    const int L = a.L;
    const int l = threadIdx.x % L;
    const int t = threadIdx.x / L;
    const Chunk c = make_chunk(t, a.T, a.cbase, a.crem);       // full chunks: s0 = 16 t, cnt = 16, j0 = 0
    const long long st = a.stride;
    const int n = a.n;
    const int CS = pf_chunk_stride(L);
    const int tile_doubles = a.T * CS;
    double* bufU = sm + exch_doubles(a.T, L);
    double* bufV = bufU + tile_doubles;            // velocity (Burgers) or second input (u + scale*u2)
    double* bufO = bufV + tile_doubles;            // accumulation target
    const bool has_u2 = (a.u2 != nullptr);
    const bool has_vel = (MODE == MODE_BURGERS) && (a.vel != a.u);
    const bool has_acc = (a.accumulate != 0) && (MODE == MODE_BURGERS || MODE == MODE_P1);
    const long long ntiles = a.nlines / L;

    auto tile_base = [&](long long tile) {
        const long long line0 = tile * L;
        return (line0 / a.inner) * a.outer_stride + (line0 % a.inner);
    };
    long long tile = blockIdx.x;
    if (tile < ntiles) {
        const long long gb = tile_base(tile);
        pf_issue_tile(bufU, a.u + gb, st, n, L, CS);
        if (has_u2) pf_issue_tile(bufV, a.u2 + gb, st, n, L, CS);
    }
    cp_async_commit();
    for (; tile < ntiles; tile += gridDim.x) {
        const long long gb = tile_base(tile);
        cp_async_wait<0>();
        __syncthreads();
        // ---- chunk and halos from shared memory
        double u[CHUNK + 6];
        {
            const double* pc = bufU + c.t * CS + l;
            const bool lok = PER || c.t > 0, rok = PER || c.t < c.T - 1;
            const double* pl = bufU + ((c.t > 0) ? (c.t - 1) : (c.T - 1)) * CS + 13 * L + l;
            const double* pr = bufU + ((c.t < c.T - 1) ? (c.t + 1) : 0) * CS + l;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) u[j + 3] = pc[j * L];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                u[k] = lok ? pl[k * L] : 0.0;
                u[CHUNK + 3 + k] = rok ? pr[k * L] : 0.0;
            }
            if (has_u2) {
                const double* qc = bufV + c.t * CS + l;
                const double* ql = bufV + ((c.t > 0) ? (c.t - 1) : (c.T - 1)) * CS + 13 * L + l;
                const double* qr = bufV + ((c.t < c.T - 1) ? (c.t + 1) : 0) * CS + l;
#pragma unroll
                for (int j = 0; j < CHUNK; j++) u[j + 3] = u[j + 3] + qc[j * L] * a.scale;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (lok) u[k] = u[k] + ql[k * L] * a.scale;
                    if (rok) u[CHUNK + 3 + k] = u[CHUNK + 3 + k] + qr[k * L] * a.scale;
                }
            }
        }
        __syncthreads();                            // bufU / bufV free again
        // ---- prefetch: this tile's velocity and accumulation target, then the next tile's input
        if (has_vel) pf_issue_tile(bufV, a.vel + gb, st, n, L, CS);
        if (has_acc) pf_issue_tile(bufO, a.out1 + gb, st, n, L, CS);
        cp_async_commit();
        const long long next = tile + gridDim.x;
        if (next < ntiles) {
            const long long gn = tile_base(next);
            pf_issue_tile(bufU, a.u + gn, st, n, L, CS);
        }
        cp_async_commit();

        double wb[BROW_W], wt[BROW_W];
        if (!PER) {
#pragma unroll
            for (int k = 0; k < BROW_W; k++) {
                wb[k] = (c.t == 0) ? u[3 + k] : 0.0;
                wt[k] = (c.t == c.T - 1) ? u[3 + CHUNK - 1 - k] : 0.0;
            }
        }
        double dd[2][CHUNK];
        if (MODE == MODE_BURGERS && !has_vel) {
            // SELF: the advecting velocity is s itself; park this thread's chunk in its own slots of bufV
            double* vq = bufV + c.t * CS + l;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) vq[j * L] = u[j + 3];
        }
        line_core<MODE, PER, NEED1, true>(u, wb, wt, c, a, l, L, sm, dd);
        double (&d1)[CHUNK] = dd[0];
        double (&d2)[CHUNK] = dd[1];

        // ---- velocity / accumulation target have arrived (all groups but the newest one are complete)
        if (has_vel || has_acc) {
            cp_async_wait<1>();
            __syncthreads();
        }
        double* __restrict__ o1 = a.out1 + gb + l + (long long)c.s0 * st;
        double* __restrict__ o2 = (MODE == MODE_P2_P1) ? a.out2 + gb + l + (long long)c.s0 * st : nullptr;
        if (MODE == MODE_BURGERS) {
            const double* vq = bufV + c.t * CS + l;
            const double* oq = bufO + c.t * CS + l;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                const double v = vq[j * L];
                double r = d2[j] - v * d1[j];
                if (has_acc) r = (a.accumulate > 0) ? oq[j * L] + r : oq[j * L] - r;
                d2[j] = r;
            }
        } else if (MODE == MODE_P1 && has_acc) {
            const double* oq = bufO + c.t * CS + l;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) d1[j] = (a.accumulate > 0) ? oq[j * L] + d1[j] : oq[j * L] - d1[j];
        }
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            if (MODE == MODE_P1) __stcs(o1 + j * st, d1[j]);
            if (MODE == MODE_P2 || MODE == MODE_BURGERS) __stcs(o1 + j * st, d2[j]);
            if (MODE == MODE_P2_P1) { __stcs(o1 + j * st, d2[j]); __stcs(o2 + j * st, d1[j]); }
        }
        // the next iteration starts with wait<0> + __syncthreads, which also orders the reads of bufV / bufO
        // above before they are overwritten
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// x direction: lines contiguous in memory; a tile of L lines is staged through padded shared memory
__device__ __forceinline__ int xpos(int i) { return i + (i >> 4); }

template <int MODE, bool PER, bool NEED1, bool FULL>
__global__ void __launch_bounds__(512) line_kernel_contig(LineArgs a) {
    extern __shared__ double sm[];
    const int L = a.L;
    const int l = threadIdx.x % L;
    const int t = threadIdx.x / L;
    const int n = a.n;
    const int S = a.xstride;                       // shared-memory stride between lines
    const Chunk c = make_chunk(t, a.T, a.cbase, a.crem);
    const long long line0 = (long long)blockIdx.x * L;
    const int nl = (int)min((long long)L, a.nlines - line0);
    double* tile = sm + exch_doubles(a.T, L);
    double* vtile = tile + (size_t)L * S;

    // cooperative coalesced load of the tile(s)
    const bool two = (MODE == MODE_BURGERS) && (a.vel != a.u);
    for (int ll = 0; ll < nl; ll++) {
        const double* __restrict__ src = a.u + (line0 + ll) * (long long)n;
        const double* __restrict__ vsrc = two ? a.vel + (line0 + ll) * (long long)n : nullptr;
        if ((n & 1) == 0) {
            for (int i = 2 * threadIdx.x; i < n; i += 2 * blockDim.x) {
                double2 v = __ldcs(reinterpret_cast<const double2*>(src + i));
                if (a.u2 != nullptr) {
                    const double2 w2 = __ldcs(reinterpret_cast<const double2*>(a.u2 + (line0 + ll) * (long long)n + i));
                    v.x = v.x + w2.x * a.scale; v.y = v.y + w2.y * a.scale;
                }
                tile[ll * S + xpos(i)] = v.x;
                tile[ll * S + xpos(i + 1)] = v.y;
                if (two) {
                    const double2 w = __ldcs(reinterpret_cast<const double2*>(vsrc + i));
                    vtile[ll * S + xpos(i)] = w.x;
                    vtile[ll * S + xpos(i + 1)] = w.y;
                }
            }
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double v = src[i];
                if (a.u2 != nullptr) v = v + a.u2[(line0 + ll) * (long long)n + i] * a.scale;
                tile[ll * S + xpos(i)] = v;
                if (two) vtile[ll * S + xpos(i)] = vsrc[i];
            }
        }
    }
    __syncthreads();

    const int lr = (l < nl) ? l : 0;               // inactive lines redo line 0 (keeps barriers uniform)
    const double* row = tile + lr * S;
    double u[CHUNK + 6];
    if (FULL) {
        // s0 = 16 t, so xpos(s0 + j) = 17 t + j: the chunk and both halos sit at fixed offsets
        const double* pc = row + c.s0 + c.t;
        const bool lok = PER || c.t > 0, rok = PER || c.t < c.T - 1;
        const double* pl = (c.t > 0) ? pc - 4 : row + (n + c.T - 4);
        const double* pr = (c.t < c.T - 1) ? pc + CHUNK + 1 : row;
#pragma unroll
        for (int j = 0; j < CHUNK; j++) u[j + 3] = pc[j];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? pl[k] : 0.0;
            u[CHUNK + 3 + k] = rok ? pr[k] : 0.0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CHUNK + 6; k++) {
            int p = c.s0 - 3 + (k - c.j0);
            const bool in_win = (k >= c.j0) && (k < c.j0 + c.cnt + 6);
            if (PER) p = (p < 0) ? p + n : (p >= n ? p - n : p);
            const bool ok = in_win && p >= 0 && p < n;
            u[k] = ok ? row[xpos(ok ? p : 0)] : 0.0;
        }
    }
    double wb[BROW_W], wt[BROW_W];
    if (!PER) {
#pragma unroll
        for (int k = 0; k < BROW_W; k++) {
            wb[k] = (c.t == 0) ? u[3 + k] : 0.0;
            wt[k] = (c.t == c.T - 1) ? u[3 + CHUNK - 1 - k] : 0.0;
        }
    }
    double dd[2][CHUNK];
    line_core<MODE, PER, NEED1, FULL>(u, wb, wt, c, a, l, L, sm, dd);
    double (&d1)[CHUNK] = dd[0];
    double (&d2)[CHUNK] = dd[1];
    // all halo reads of the tile happened before the first barrier inside line_core; results may overwrite it

    const int j1 = c.j0 + c.cnt;
    const int npass = (MODE == MODE_P2_P1) ? 2 : 1;
    for (int pass = 0; pass < npass; pass++) {
        if (pass == 1) __syncthreads();
        double* wrow = tile + lr * S;
        const double* vrow = (two ? vtile : tile) + lr * S;
        if (l < nl) {
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                if (FULL || (j >= c.j0 && j < j1)) {
                    const int pp = FULL ? (c.s0 + c.t + j) : xpos(c.s0 + j - c.j0);
                    double r;
                    if (MODE == MODE_P1) r = d1[j];
                    else if (MODE == MODE_P2) r = d2[j];
                    else if (MODE == MODE_P2_P1) r = (pass == 0) ? d2[j] : d1[j];
                    else r = d2[j] - vrow[pp] * d1[j];
                    wrow[pp] = r;
                }
            }
        }
        __syncthreads();
        double* __restrict__ obase = (pass == 0) ? a.out1 : a.out2;
        for (int ll = 0; ll < nl; ll++) {
            double* __restrict__ dst = obase + (line0 + ll) * (long long)n;
            if ((n & 1) == 0) {
                for (int i = 2 * threadIdx.x; i < n; i += 2 * blockDim.x) {
                    double2 v;
                    v.x = tile[ll * S + xpos(i)];
                    v.y = tile[ll * S + xpos(i + 1)];
                    if (a.accumulate != 0) {
                        const double2 o = *reinterpret_cast<const double2*>(dst + i);
                        if (a.accumulate > 0) { v.x = o.x + v.x; v.y = o.y + v.y; }
                        else { v.x = o.x - v.x; v.y = o.y - v.y; }
                    }
                    *reinterpret_cast<double2*>(dst + i) = v;
                }
            } else {
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    double v = tile[ll * S + xpos(i)];
                    if (a.accumulate > 0) v = dst[i] + v;
                    else if (a.accumulate < 0) v = dst[i] - v;
                    dst[i] = v;
                }
            }
        }
    }
}

bool g_prefetch = false;   // experimental persistent/cp.async variant: measured slower than the direct-load kernel (profiles/), off by default

template <int MODE, bool PER, bool NEED1>
cudaError_t launch_pf(const LineArgs& a, cudaStream_t stream) {
    const int threads = a.L * a.T;
    const long long ntiles = a.nlines / a.L;
    const size_t smem = (exch_doubles(a.T, a.L) + (size_t)3 * a.T * pf_chunk_stride(a.L)) * sizeof(double);
    auto k = line_kernel_strided_pf<MODE, PER, NEED1>;
    static int ctas_per_sm = 0;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
    ctas_per_sm = occ > 0 ? occ : 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = std::min<long long>(ntiles, (long long)sms * ctas_per_sm);
    k<<<(unsigned)grid, threads, smem, stream>>>(a);
    return cudaGetLastError();
}

inline bool pf_eligible(int mode, const LineArgs& a, bool contig, bool full) {
    if (!g_prefetch || contig || !full || mode == MODE_NEUMANN) return false;
    if (a.L < 2 || (a.L & (a.L - 1))) return false;
    if (a.nlines % a.L || a.inner % a.L || (a.stride & 1) || (a.outer_stride & 1)) return false;
    auto al = [](const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; };
    if (!al(a.u) || !al(a.u2) || !al(a.vel) || !al(a.out1)) return false;
    const size_t smem = (exch_doubles(a.T, a.L) + (size_t)3 * a.T * pf_chunk_stride(a.L)) * sizeof(double);
    return smem <= 200 * 1024;
}

template <int MODE, bool PER, bool NEED1, bool FULL>
cudaError_t launch_one(const LineArgs& a, bool contig, cudaStream_t stream) {
    if (MODE != MODE_NEUMANN && FULL && pf_eligible(MODE, a, contig, FULL)) return launch_pf<(MODE == MODE_NEUMANN ? MODE_P1 : MODE), PER, NEED1>(a, stream);
    const int threads = a.L * a.T;
    const long long blocks = (a.nlines + a.L - 1) / a.L;
    size_t smem = exch_doubles(a.T, a.L) * sizeof(double);
    if (contig) {
        const bool two = (MODE == MODE_BURGERS) && (a.vel != a.u);
        smem += (size_t)a.L * a.xstride * sizeof(double) * (two ? 2 : 1);
        auto k = line_kernel_contig<MODE, PER, NEED1, FULL>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<(unsigned)blocks, threads, smem, stream>>>(a);
    } else {
        auto k = line_kernel_strided<MODE, PER, NEED1, FULL>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<(unsigned)blocks, threads, smem, stream>>>(a);
    }
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_mode(const LineArgs& a, bool per, bool need1, bool contig, cudaStream_t s) {
    const bool full = (a.crem == 0 && a.cbase == CHUNK);      // every chunk holds exactly CHUNK points
    if (full) {
        if (per) return launch_one<MODE, true, false, true>(a, contig, s);
        if (need1) return launch_one<MODE, false, true, true>(a, contig, s);
        return launch_one<MODE, false, false, true>(a, contig, s);
    }
    if (per) return launch_one<MODE, true, false, false>(a, contig, s);
    if (need1) return launch_one<MODE, false, true, false>(a, contig, s);
    return launch_one<MODE, false, false, false>(a, contig, s);
}

}  // namespace

void set_prefetch(bool on) { g_prefetch = on; }

int pick_lines_per_cta(int T, bool contig, int override_L) {
    if (override_L > 0) return override_L;
    // tile = L lines x n points; keep the CTA at <= 512 threads and at least 4 lines (32 B segments)
    int L = contig ? 4 : 8;
    while (L > 1 && L * T > 512) L >>= 1;
    return L;
}

int xtile_stride(int n, int L) {
    int S = n + (n >> 4) + 1;
    const int want = (L >= 16) ? 1 : 16 / L;        // S mod 16 == 16/L makes chunk reads conflict-free
    while ((S & 15) != (want & 15)) S++;
    return S;
}

cudaError_t launch_lines(int mode, const LineArgs& a, bool periodic, bool need1, bool contig, cudaStream_t s) {
    switch (mode) {
        case MODE_P1: return launch_mode<MODE_P1>(a, periodic, false, contig, s);
        case MODE_P2: return launch_mode<MODE_P2>(a, periodic, need1, contig, s);
        case MODE_P2_P1: return launch_mode<MODE_P2_P1>(a, periodic, need1, contig, s);
        case MODE_BURGERS: return launch_mode<MODE_BURGERS>(a, periodic, need1, contig, s);
        case MODE_NEUMANN: return launch_mode<MODE_NEUMANN>(a, periodic, false, false, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace tlab
