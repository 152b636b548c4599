// z operators of a z-split domain WITHOUT transposes.
//
// The reference differentiates along a split direction by transposing the slabs into pencils, solving whole lines and
// transposing back (TLabMPI_Trp_ExecK_Forward / _Backward around OPR_Partial_Z and OPR_Burgers_Z,
// src/operators/opr_partial.f90:186-249, src/physics/opr_burgers.f90:387-424, src/base/tlab_mpi_transpose.f90:301-553):
// two all-to-alls of the whole field per call.  The chunked formulation of the factored tridiagonal solves (lines2.cu)
// makes that unnecessary: a chunk of 16 points only needs, beyond its own data,
//     * 3 halo points of the field on each side (banded right-hand side),
//     * the forward end values y of the 6 chunks before it             (A = sum_k wf[k] y(t-k)),
//     * the end values y and the zero-inflow solutions x^_0 of the 6 chunks after it (B = sum_k wb[k] z(t+k),
//       z(t) = Q0(t) A(t) + x^_0(t)),
//     * for the few chunks at the two ends of the (global, periodic) line, the closure x_N of the circulant system,
//       a sum of terms of the first K0 and the last K1 chunks of the line.
// So each rank keeps its slab and only exchanges, per line, 6 halo points and the end values of its first and last 6
// chunks with its two neighbours (and the ranks holding the last 12 chunks send their ends to rank 0): 54 doubles per
// line and system instead of 2 x kmax, all of it written straight into the neighbours' memory (CUDA IPC, NVLink).
//
// One call =   push of the halo planes -> barrier -> PHASE 1: right-hand sides + zero-inflow sweeps of every chunk,
//              end values stored into the neighbours' buffers -> barrier -> PHASE 2: the same sweeps again (cheaper than
//              keeping them: 8 B/pt of reads instead of 16 B/pt of writes + reads), look-back / look-ahead with the
//              neighbours' ends, correction, Burgers combination, accumulation into the result.
// The arithmetic is the one of lines2.cu term by term (same weights, same order of the sums), so a split run agrees with
// the single-GPU run to round-off.  Exchange buffers alternate between two sets, so a rank may start pushing for the
// next call while a neighbour still reads the previous one.
#include "../../include/tlab_gpu.h"
#include "splitz.h"
#include "march_dev.cuh"
#include "trp.h"
#include <algorithm>

namespace tlab {

namespace {

constexpr int TAILC = 2 * LB2;       // chunks at the end of the line whose ends go to the first rank
// exchange block, in planes of nxy doubles
constexpr int OFF_HLO = 0;                       // [3]            field planes just below the slab
constexpr int OFF_HHI = 3;                       // [3]            field planes just above
constexpr int OFF_EPREV = 6;                     // [6][2]         y of the previous rank's last 6 chunks
constexpr int OFF_ENEXT = OFF_EPREV + LB2 * 2;   // [6][2][3]      y, x^_0, p of the next rank's first 6 chunks
constexpr int OFF_TAIL = OFF_ENEXT + LB2 * 2 * 3;  // [12][2][2]   y, p of the last 12 chunks of the line (first rank only)
constexpr int BLOCK_PLANES = OFF_TAIL + TAILC * 2 * 2;

struct SplitArgs {
    int T = 0, Tl = 0, t0 = 0;        // chunks of the global line, of this slab, first global chunk of this slab
    int L = 8, lshift = 3;
    int accumulate = 0;
    double scale = 0.0;
    long long nxy = 0;
    const double* u = nullptr;
    const double* u2 = nullptr;
    const double* vel = nullptr;
    double* out = nullptr;
    const double* mine = nullptr;     // this rank's exchange block (halos, neighbours' ends)
    double* to_prev = nullptr;        // the exchange blocks of the previous / next rank and of the first rank
    double* to_next = nullptr;
    double* to_first = nullptr;
    RhsTab rhs1, rhs2;
    Sys2 s1, s2;
};

__global__ void splitz_push_kernel(const double* __restrict__ u, const double* __restrict__ u2, double scale, long long nxy,
                                   int kmax, double* __restrict__ hi_of_prev, double* __restrict__ lo_of_next) {
    const long long n3 = 3 * nxy, top = (long long)(kmax - 3) * nxy;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
        double a = u[i], b = u[top + i];
        if (u2) { a = fma(u2[i], scale, a); b = fma(u2[top + i], scale, b); }
        hi_of_prev[i] = a;            // my first planes sit above the previous slab
        lo_of_next[i] = b;            // my last planes sit below the next slab
    }
}

// A = sum_k wf[k] y(-k): y points at this chunk's slot, slots L apart
__device__ __forceinline__ double back6(const double* __restrict__ y, int L, const double2* __restrict__ cr) {
    const double2 w01 = ldg2(cr + 0), w23 = ldg2(cr + 1), w45 = ldg2(cr + 2);
    double A = w01.x * y[-1 * L];
    A = fma(w01.y, y[-2 * L], A);
    A = fma(w23.x, y[-3 * L], A);
    A = fma(w23.y, y[-4 * L], A);
    A = fma(w45.x, y[-5 * L], A);
    A = fma(w45.y, y[-6 * L], A);
    return A;
}
__device__ __forceinline__ double ahead6(const double* __restrict__ z, int L, const double2* __restrict__ cr) {
    const double2 w01 = ldg2(cr + 3), w23 = ldg2(cr + 4), w45 = ldg2(cr + 5);
    double B = w01.x * z[1 * L];
    B = fma(w01.y, z[2 * L], B);
    B = fma(w23.x, z[3 * L], B);
    B = fma(w23.y, z[4 * L], B);
    B = fma(w45.x, z[5 * L], B);
    B = fma(w45.y, z[6 * L], B);
    return B;
}

__host__ __device__ inline int split_smem_doubles(int Tl, int L) { return ((Tl + 2 * LB2) + (Tl + LB2) + Tl + LB2) * L; }

template <int MODE, int PHASE>
__global__ void __launch_bounds__(512, 1) splitz_kernel(const __grid_constant__ SplitArgs a) {
    extern __shared__ double sm[];
    constexpr int NS = (MODE == MODE_BURGERS) ? 2 : 1;
    const int L = a.L, Tl = a.Tl, T = a.T;
    const int l = threadIdx.x & (L - 1), t = threadIdx.x >> a.lshift;
    const int tg = a.t0 + t;
    const long long nxy = a.nxy;
    const long long line = (long long)blockIdx.x * L + l;
    const long long coff = line + (long long)(t * C) * nxy;

    // ---- chunk + halos (from the slab, or from the planes the neighbours pushed)
    double u[C + 6];
    {
        const double* __restrict__ pc = a.u + coff;
        const double* __restrict__ hlo = a.mine + (long long)OFF_HLO * nxy + line;
        const double* __restrict__ hhi = a.mine + (long long)OFF_HHI * nxy + line;
#pragma unroll
        for (int j = 0; j < C; j++) u[j + 3] = __ldcs(pc + j * nxy);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = (t > 0) ? __ldcs(pc + (k - 3) * nxy) : __ldcs(hlo + k * nxy);
            u[C + 3 + k] = (t < Tl - 1) ? __ldcs(pc + (C + k) * nxy) : __ldcs(hhi + k * nxy);
        }
        if (a.u2 != nullptr) {
            const double* __restrict__ p2 = a.u2 + coff;
#pragma unroll
            for (int j = 0; j < C; j++) u[j + 3] = fma(__ldcs(p2 + j * nxy), a.scale, u[j + 3]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (t > 0) u[k] = fma(__ldcs(p2 + (k - 3) * nxy), a.scale, u[k]);
                if (t < Tl - 1) u[C + 3 + k] = fma(__ldcs(p2 + (C + k) * nxy), a.scale, u[C + 3 + k]);
            }
        }
    }
    // ---- right-hand sides and zero-inflow sweeps
    double f[NS][C];
    rhs_interior<false>(u, f[0], a.rhs1);
    if (NS == 2) rhs_interior<true>(u, f[NS - 1], a.rhs2);
    double yend[NS], part[NS];
    bool isc[NS];
    const double2* cr[NS];
    const double2* tp[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const Sys2& S = (s == 0) ? a.s1 : a.s2;
        cr[s] = reinterpret_cast<const double2*>(S.crec) + (long long)tg * 8;
        isc[s] = ldg2(cr[s] + 7).x != 0.0;
        tp[s] = tab_ptr(S, tg);
        part[s] = 0.0;
        if (isc[s]) local_const(f[s], S, yend[s]);
        else local_tab<true>(f[s], tp[s], yend[s], part[s]);
    }

    if (PHASE == 1) {
        // ---- end values into the neighbours' blocks
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (t < LB2) {
                double* d = a.to_prev + ((long long)OFF_ENEXT + (long long)(t * 2 + s) * 3) * nxy + line;
                d[0] = yend[s]; d[nxy] = f[s][0]; d[2 * nxy] = part[s];
            }
            if (t >= Tl - LB2) a.to_next[((long long)OFF_EPREV + (long long)((t - (Tl - LB2)) * 2 + s)) * nxy + line] = yend[s];
            if (tg >= T - TAILC) {
                double* d = a.to_first + ((long long)OFF_TAIL + (long long)((tg - (T - TAILC)) * 2 + s) * 2) * nxy + line;
                d[0] = yend[s]; d[nxy] = part[s];
            }
        }
        return;
    }

    // ---- PHASE 2: look-back / look-ahead over own and neighbouring chunk ends
    const bool first = (a.t0 == 0), last = (a.t0 + Tl == T);
    const int per_sys = split_smem_doubles(Tl, L);
    double A[NS], B[NS], xN[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double* Y = sm + s * per_sys;            // y of chunks t0-6 .. t0+Tl+5
        Y[(LB2 + t) * L + l] = yend[s];
        if (t < LB2) {
            Y[t * L + l] = a.mine[((long long)OFF_EPREV + (long long)(t * 2 + s)) * nxy + line];
            Y[(LB2 + Tl + t) * L + l] = a.mine[((long long)OFF_ENEXT + (long long)(t * 2 + s) * 3) * nxy + line];
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const Sys2& S = (s == 0) ? a.s1 : a.s2;
        double* Y = sm + s * per_sys;
        double* Z = Y + (Tl + 2 * LB2) * L;      // z of chunks t0 .. t0+Tl+5
        double* W = Z + (Tl + LB2) * L;          // closure terms of the own chunks
        double* WS = W + Tl * L;                 // closure terms of the other end of the line
        const double2 q0 = ldg2(cr[s] + 6);
        A[s] = back6(Y + (LB2 + t) * L + l, L, cr[s]);
        Z[t * L + l] = fma(q0.x, A[s], f[s][0]);
        W[t * L + l] = fma(q0.y, A[s], part[s]);
        if (t < LB2) {
            const double* en = a.mine + ((long long)OFF_ENEXT + (long long)(t * 2 + s) * 3) * nxy + line;
            const int tgn = tg + Tl;             // the next rank's chunk t
            if (tgn < T) {
                const double2* crn = reinterpret_cast<const double2*>(S.crec) + (long long)tgn * 8;
                const double An = back6(Y + (LB2 + Tl + t) * L + l, L, crn);
                Z[(Tl + t) * L + l] = fma(ldg2(crn + 6).x, An, en[nxy]);
            } else {
                // last rank: the next chunks are the first ones of the line; their closure terms (weights vanish before chunk 0)
                Z[(Tl + t) * L + l] = 0.0;
                const double2* crw = reinterpret_cast<const double2*>(S.crec) + (long long)(tgn - T) * 8;
                const double Aw = back6(Y + (LB2 + Tl + t) * L + l, L, crw);
                WS[t * L + l] = fma(ldg2(crw + 6).y, Aw, en[2 * nxy]);
            }
            if (first && t < S.K1) {
                // first rank: closure terms of the last K1 chunks of the line from the ends their owners sent
                const int kk = T - S.K1 + t, ti = kk - (T - TAILC);
                const double2* crk = reinterpret_cast<const double2*>(S.crec) + (long long)kk * 8;
                const double* ty = a.mine + ((long long)OFF_TAIL + (long long)(ti * 2 + s) * 2) * nxy + line;   // y of chunk kk
                const long long cs = 4 * nxy;    // one chunk back in the tail
                const double2 w01 = ldg2(crk + 0), w23 = ldg2(crk + 1), w45 = ldg2(crk + 2);
                double Ak = w01.x * ty[-1 * cs];
                Ak = fma(w01.y, ty[-2 * cs], Ak);
                Ak = fma(w23.x, ty[-3 * cs], Ak);
                Ak = fma(w23.y, ty[-4 * cs], Ak);
                Ak = fma(w45.x, ty[-5 * cs], Ak);
                Ak = fma(w45.y, ty[-6 * cs], Ak);
                WS[t * L + l] = fma(ldg2(crk + 6).y, Ak, ty[nxy]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const Sys2& S = (s == 0) ? a.s1 : a.s2;
        double* Y = sm + s * per_sys;
        double* Z = Y + (Tl + 2 * LB2) * L;
        double* W = Z + (Tl + LB2) * L;
        double* WS = W + Tl * L;
        B[s] = ahead6(Z + t * L + l, L, cr[s]);
        xN[s] = 0.0;
        if (!isc[s]) {
            // x_N = sum over the first K0 chunks, then over the last K1 chunks of the line (the order of lines2.cu)
            if (first) {
                for (int k = 0; k < S.K0; k++) xN[s] += W[k * L + l];
                for (int k = 0; k < S.K1; k++) xN[s] += WS[k * L + l];
            } else if (last) {
                for (int k = 0; k < S.K0; k++) xN[s] += WS[k * L + l];
                for (int k = Tl - S.K1; k < Tl; k++) xN[s] += W[k * L + l];
            }
        }
        if (isc[s]) finish_const(f[s], S, A[s], B[s]);
        else finish_tab<true>(f[s], tp[s], A[s], B[s], xN[s]);
    }

    // ---- result
    double* __restrict__ o = a.out + coff;
    double r[C];
    if (MODE == MODE_BURGERS) {
        const double* __restrict__ vp = a.vel + coff;
        double vv[C];
#pragma unroll
        for (int j = 0; j < C; j++) vv[j] = __ldcs(vp + j * nxy);
#pragma unroll
        for (int j = 0; j < C; j++) r[j] = f[NS - 1][j] - vv[j] * f[0][j];
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) r[j] = f[0][j];
    }
    if (a.accumulate != 0) {
        double oo[C];
#pragma unroll
        for (int j = 0; j < C; j++) oo[j] = __ldcs(o + j * nxy);
#pragma unroll
        for (int j = 0; j < C; j++) r[j] = (a.accumulate > 0) ? oo[j] + r[j] : oo[j] - r[j];
    }
#pragma unroll
    for (int j = 0; j < C; j++) __stcs(o + j * nxy, r[j]);
}

template <int MODE>
cudaError_t launch_split(int phase, const SplitArgs& a, cudaStream_t st) {
    const int threads = a.L * a.Tl;
    const unsigned grid = (unsigned)(a.nxy / a.L);
    if (phase == 1) {
        splitz_kernel<MODE, 1><<<grid, threads, 0, st>>>(a);
    } else {
        const size_t smem = (size_t)split_smem_doubles(a.Tl, a.L) * ((MODE == MODE_BURGERS) ? 2 : 1) * sizeof(double);
        static size_t set = 48 * 1024;
        if (smem > set) {
            cudaError_t e = cudaFuncSetAttribute(splitz_kernel<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set = smem;
        }
        splitz_kernel<MODE, 2><<<grid, threads, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// PHASE 2 as a march (march.cu): a CTA of MW warps owns a panel of 32 lines of the slab and walks along z in rounds of MW
// chunks; the corrections of a round are applied one round later from a thread-private stash.  What a whole-line march
// gets from its own rounds comes from the neighbours here: the ring is seeded with the forward ends of the previous rank's
// last LBM chunks, the chunk starts beyond the slab are built from the next rank's first chunks, and the two ends of the
// circulant line trade their closure terms (all of it left in this rank's exchange block by PHASE 1 of the ranks
// concerned).  No rotation of the rounds is needed: the other end of the line is not ours to wait for.
// 256-byte rows whatever the slab thickness, four CTAs per SM in different phases: the finishing stage reads the
// field once, against the whole-slab kernel above that holds the slab line in one CTA (and is bound by its latency).
__host__ __device__ inline long long e_next(int t, int s, int c) { return (long long)OFF_ENEXT + (long long)(t * 2 + s) * 3 + c; }
__host__ __device__ inline long long e_prev(int t, int s) { return (long long)OFF_EPREV + (long long)(t * 2 + s); }
__host__ __device__ inline long long e_tail(int ti, int s, int c) { return (long long)OFF_TAIL + (long long)(ti * 2 + s) * 2 + c; }

// CIRC: circulant form (plan.h, Sys2::circ): constant chunks everywhere, constant window weights wrapping from rank to rank, no
// closure terms (every rank does the same work), the solution scaled by rho at the end.
template <int MODE, bool CIRC>
__global__ void __launch_bounds__(MW * ML, 4) splitz_march_kernel(const __grid_constant__ SplitArgs a) {
    constexpr bool TWO = (MODE == MODE_BURGERS);
    constexpr int NS = TWO ? 2 : 1;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int Tl = a.Tl, T = a.T, t0 = a.t0, R = Tl / MW;
    const long long nxy = a.nxy;
    const long long line = (long long)blockIdx.x * ML + lane;
    const bool first = (t0 == 0), last = (t0 + Tl == T);
    const Sys2& S1 = a.s1;
    const Sys2& S2 = a.s2;
    const MarchSm m1(sm), m2(sm + M_SYS);
    for (int i = threadIdx.x; i < NS * M_SYS; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const double* __restrict__ mine = a.mine + line;
    // ---- seeds: forward ends of the previous rank's last LBM chunks -> "previous round" half of the ring (half 1 for step 0)
    if (w < LBM) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const MarchSm& m = s ? m2 : m1;
            m.Y[(MW + MW - LBM + w) * ML + lane] = mine[e_prev(LB2 - LBM + w, s) * nxy];
        }
    }
    // ---- closure terms of the other end of the line (lines2.cu: x_N = sum over the first K0 and the last K1 chunks)
    if (!CIRC && w == 3) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const Sys2& S = s ? S2 : S1;
            const MarchSm& m = s ? m2 : m1;
            if (first) {
                for (int k = 0; k < S.K1m; k++) {
                    const int kk = T - S.K1m + k, ti = kk - (T - TAILC);
                    const double* cr = S.crec + (size_t)kk * 16;
                    double A = __ldg(cr + 0) * mine[e_tail(ti - 1, s, 0) * nxy];
                    A = fma(__ldg(cr + 1), mine[e_tail(ti - 2, s, 0) * nxy], A);
                    A = fma(__ldg(cr + 2), mine[e_tail(ti - 3, s, 0) * nxy], A);
                    m.Wc[(MW + k) * ML + lane] = fma(__ldg(cr + 13), A, mine[e_tail(ti, s, 1) * nxy]);
                }
            }
            if (last) {
                // the next rank is the first one: its chunks 0 .. K0-1 open the line (no inflow before chunk 0: weights vanish)
                for (int k = 0; k < S.K0m; k++) {
                    const double* cr = S.crec + (size_t)k * 16;
                    double A = 0.0;
                    if (k >= 1) A = __ldg(cr + 0) * mine[e_next(k - 1, s, 0) * nxy];
                    if (k >= 2) A = fma(__ldg(cr + 1), mine[e_next(k - 2, s, 0) * nxy], A);
                    if (k >= 3) A = fma(__ldg(cr + 2), mine[e_next(k - 3, s, 0) * nxy], A);
                    m.Wc[k * ML + lane] = fma(__ldg(cr + 13), A, mine[e_next(k, s, 2) * nxy]);
                }
            }
        }
    }
    __syncthreads();

    const double* __restrict__ pu = a.u + line;
    const double* __restrict__ pu2 = (a.u2 != nullptr) ? a.u2 + line : nullptr;
    const int slot = w * ML + lane;
    double o1[C], o2[C];
    double A1p = 0.0, A2p = 0.0;
    for (int s = 0; s <= R; s++) {
        const bool front = s < R, back = s > 0;
        const int h = s & 1;
        const int t = s * MW + w, tb = (s - 1) * MW + w;
        const int tg = t0 + t, tbg = t0 + tb;
        double ye1 = 0.0, ye2 = 0.0, pt1 = 0.0, pt2 = 0.0, x01 = 0.0, x02 = 0.0;
        if (front) {
            double u[C + 6], f1[C], f2[C];
            {
                // chunk + halos: from the slab, or from the planes the neighbours pushed
                const double* q = pu + (long long)(t * C) * nxy;
                const double* ql = (t > 0) ? q - 3 * nxy : mine + (long long)OFF_HLO * nxy;
                const double* qr = (t < Tl - 1) ? q + (long long)C * nxy : mine + (long long)OFF_HHI * nxy;
#pragma unroll
                for (int j = 0; j < C; j++) { u[j + 3] = __ldcs(q); q += nxy; }
#pragma unroll
                for (int k = 0; k < 3; k++) { u[k] = __ldcs(ql); u[C + 3 + k] = __ldcs(qr); ql += nxy; qr += nxy; }
                if (pu2 != nullptr) {
                    // (the pushed halo planes already hold u + scale*u2)
                    const double* p2 = pu2 + (long long)(t * C) * nxy;
                    const double* pl = p2 - 3 * nxy;
                    const double* pr = p2 + (long long)C * nxy;
#pragma unroll
                    for (int j = 0; j < C; j++) { u[j + 3] = fma(__ldcs(p2), a.scale, u[j + 3]); p2 += nxy; }
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (t > 0) u[k] = fma(__ldcs(pl), a.scale, u[k]);
                        if (t < Tl - 1) u[C + 3 + k] = fma(__ldcs(pr), a.scale, u[C + 3 + k]);
                        pl += nxy; pr += nxy;
                    }
                }
            }
            rhs_interior<false>(u, f1, a.rhs1);
            if (TWO) rhs_interior<true>(u, f2, a.rhs2);
            if (CIRC) {
                if (TWO) local_const2(f1, f2, S1, S2, ye1, ye2);
                else local_const(f1, S1, ye1);
            } else {
                march_local<true>(f1, S1, tg, ye1, pt1);
                if (TWO) march_local<true>(f2, S2, tg, ye2, pt2);
            }
            x01 = f1[0];
            if (TWO) x02 = f2[0];
#pragma unroll
            for (int j = 0; j < C; j++) {
                double* q = m1.X + j * (MW * ML) + slot;
                o1[j] = *q;
                *q = f1[j];
                if (TWO) {
                    double* q2 = m2.X + j * (MW * ML) + slot;
                    o2[j] = *q2;
                    *q2 = f2[j];
                }
            }
            m1.Y[(h * MW) * ML + slot] = ye1;
            if (TWO) m2.Y[(h * MW) * ML + slot] = ye2;
        } else {
#pragma unroll
            for (int j = 0; j < C; j++) {
                o1[j] = m1.X[j * (MW * ML) + slot];
                if (TWO) o2[j] = m2.X[j * (MW * ML) + slot];
            }
        }
        __syncthreads();
        double A1 = 0.0, A2 = 0.0;
        if (front) {
#pragma unroll
            for (int q = 0; q < NS; q++) {
                const Sys2& S = q ? S2 : S1;
                const MarchSm& m = q ? m2 : m1;
                const double* cr = S.crec + (size_t)tg * 16;
                const double A = CIRC ? march_look_back_w(m.Y, S.cwf[0], S.cwf[1], S.cwf[2], h, w, lane) : march_look_back(m.Y, cr, h, w, lane);
                m.Z[(h * MW) * ML + slot] = fma(CIRC ? S.cQ[0] : __ldg(cr + 12), A, q ? x02 : x01);
                if (!CIRC) {
                    if (tg < S.K0m) m.Wc[tg * ML + lane] = fma(__ldg(cr + 13), A, q ? pt2 : pt1);
                    if (tg >= T - S.K1m) m.Wc[(MW + tg - (T - S.K1m)) * ML + lane] = fma(__ldg(cr + 13), A, q ? pt2 : pt1);
                }
                if (q) A2 = A; else A1 = A;
            }
            if (s == R - 1 && w < LBM) {
                // chunk starts beyond the slab (the next rank's chunks 0 .. LBM-1): z = Q0 A + x^_0 with A from the ends before them,
                // this rank's last chunks (ring) and the next rank's first ones (exchange block)
                const int tgn = t0 + Tl + w;
#pragma unroll
                for (int q = 0; q < NS; q++) {
                    const Sys2& S = q ? S2 : S1;
                    const MarchSm& m = q ? m2 : m1;
                    double z = 0.0;
                    if (CIRC || tgn < T) {
                        const double* cr = S.crec + (size_t)(CIRC ? 0 : tgn) * 16;
                        auto yof = [&](int k) {            // forward end of slab chunk Tl + w - k
                            const int j = w - k;           // >= 0: the next rank's chunk j; < 0: this rank's chunk Tl + j (last round, half h)
                            return (j >= 0) ? mine[e_next(j, q, 0) * nxy] : m.Y[(h * MW + MW + j) * ML + lane];
                        };
                        double A = (CIRC ? S.cwf[0] : __ldg(cr + 0)) * yof(1);
                        A = fma(CIRC ? S.cwf[1] : __ldg(cr + 1), yof(2), A);
                        A = fma(CIRC ? S.cwf[2] : __ldg(cr + 2), yof(3), A);
                        z = fma(CIRC ? S.cQ[0] : __ldg(cr + 12), A, mine[e_next(w, q, 1) * nxy]);
                    }
                    m.Zk[w * ML + lane] = z;
                }
            }
        }
        __syncthreads();
        if (back) {
            double vv[C];
            const long long boff = line + (long long)(tb * C) * nxy;
            if (TWO) {
                const double* vp = a.vel + boff;
#pragma unroll
                for (int j = 0; j < C; j++) { vv[j] = __ldcs(vp); vp += nxy; }
            }
            const bool endr = (s == R);
            const double B1 = CIRC ? march_look_ahead_w(m1.Z + ((h ^ 1) * MW) * ML, endr ? m1.Zk : m1.Z + (h * MW) * ML,
                                                        S1.cwb[0], S1.cwb[1], S1.cwb[2], w, lane)
                                   : march_look_ahead(m1.Z + ((h ^ 1) * MW) * ML, endr ? m1.Zk : m1.Z + (h * MW) * ML,
                                                      S1.crec + (size_t)tbg * 16, w, lane);
            if (CIRC) { finish_const(o1, S1, A1p, B1); scale_rho(o1, S1, tbg); }
            else march_finish<true>(o1, S1, m1, tbg, T, A1p, B1, lane);
            if (TWO) {
                const double B2 = CIRC ? march_look_ahead_w(m2.Z + ((h ^ 1) * MW) * ML, endr ? m2.Zk : m2.Z + (h * MW) * ML,
                                                            S2.cwb[0], S2.cwb[1], S2.cwb[2], w, lane)
                                       : march_look_ahead(m2.Z + ((h ^ 1) * MW) * ML, endr ? m2.Zk : m2.Z + (h * MW) * ML,
                                                          S2.crec + (size_t)tbg * 16, w, lane);
                if (CIRC) { finish_const(o2, S2, A2p, B2); scale_rho(o2, S2, tbg); }
                else march_finish<true>(o2, S2, m2, tbg, T, A2p, B2, lane);
#pragma unroll
                for (int j = 0; j < C; j++) o1[j] = o2[j] - vv[j] * o1[j];
            }
            double* po = a.out + boff;
            if (a.accumulate == 0) {
#pragma unroll
                for (int j = 0; j < C; j++) { __stcs(po, o1[j]); po += nxy; }
            } else {
                double oo[C];
                const double* pi = po;
#pragma unroll
                for (int j = 0; j < C; j++) { oo[j] = __ldcs(pi); pi += nxy; }
#pragma unroll
                for (int j = 0; j < C; j++) { __stcs(po, (a.accumulate > 0) ? oo[j] + o1[j] : oo[j] - o1[j]); po += nxy; }
            }
        }
        A1p = A1;
        A2p = A2;
    }
}

// PHASE 1 for the march: only the chunks whose ends somebody reads -- the first MW of the slab (chunk starts and closure terms
// for the previous rank), the last LBM (forward ends for the next rank) and, on the last rank, the last LBM + MW (closure terms
// for the first rank) -- in the geometry of the march (lane = line, warp = chunk): a fraction (MW + LBM) / Tl of the field is
// read instead of all of it.
template <int MODE>
__global__ void __launch_bounds__(MW * ML, 4) splitz_ends_kernel(const __grid_constant__ SplitArgs a, int nfirst, int nlist, bool trim, bool circ) {
    constexpr bool TWO = (MODE == MODE_BURGERS);
    constexpr int NS = TWO ? 2 : 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.y * MW + w;
    if (e >= nlist) return;
    const int Tl = a.Tl, T = a.T;
    const int t = (e < nfirst) ? e : Tl - (nlist - e);
    const int tg = a.t0 + t;
    const long long nxy = a.nxy;
    const long long line = (long long)blockIdx.x * ML + lane;
    const double* __restrict__ mine = a.mine + line;
    double u[C + 6];
    {
        const double* q = a.u + line + (long long)(t * C) * nxy;
        const double* ql = (t > 0) ? q - 3 * nxy : mine + (long long)OFF_HLO * nxy;
        const double* qr = (t < Tl - 1) ? q + (long long)C * nxy : mine + (long long)OFF_HHI * nxy;
#pragma unroll
        for (int j = 0; j < C; j++) { u[j + 3] = __ldcs(q); q += nxy; }
#pragma unroll
        for (int k = 0; k < 3; k++) { u[k] = __ldcs(ql); u[C + 3 + k] = __ldcs(qr); ql += nxy; qr += nxy; }
        if (a.u2 != nullptr) {
            const double* p2 = a.u2 + line + (long long)(t * C) * nxy;
            const double* pl = p2 - 3 * nxy;
            const double* pr = p2 + (long long)C * nxy;
#pragma unroll
            for (int j = 0; j < C; j++) { u[j + 3] = fma(__ldcs(p2), a.scale, u[j + 3]); p2 += nxy; }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (t > 0) u[k] = fma(__ldcs(pl), a.scale, u[k]);
                if (t < Tl - 1) u[C + 3 + k] = fma(__ldcs(pr), a.scale, u[C + 3 + k]);
                pl += nxy; pr += nxy;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const Sys2& S = s ? a.s2 : a.s1;
        double f[C], yend, part;
        if (s) rhs_interior<true>(u, f, a.rhs2);
        else rhs_interior<false>(u, f, a.rhs1);
        if (circ) {
            // circulant form: constant chunk; (y, x^_0) of the first LBM chunks for the previous rank, y of the last LBM for the next
            local_const(f, S, yend);
            if (t < LBM) {
                double* d = a.to_prev + e_next(t, s, 0) * nxy + line;
                d[0] = yend; d[nxy] = f[0];
            }
            if (t >= Tl - LBM) a.to_next[e_prev(t - (Tl - LB2), s) * nxy + line] = yend;
            continue;
        }
        march_local<true>(f, S, tg, yend, part);
        // only what splitz_march_kernel reads: (y, x^_0) of the first LBM chunks for the previous rank -- the first rank's MW
        // chunks with their closure terms for the last one --, y of the last LBM chunks for the next rank, (y, p) of the
        // last LBM + MW chunks of the line for the first rank.  Every value crosses NVLink: 18 instead of 34 doubles per line.
        if (trim ? (t < (a.t0 == 0 ? MW : LBM)) : (t < LB2)) {
            double* d = a.to_prev + e_next(t, s, 0) * nxy + line;
            d[0] = yend; d[nxy] = f[0];
            if (!trim || a.t0 == 0) d[2 * nxy] = part;
        }
        if (t >= Tl - (trim ? LBM : LB2)) a.to_next[e_prev(t - (Tl - LB2), s) * nxy + line] = yend;
        if (tg >= T - (trim ? LBM + MW : TAILC)) {
            double* d = a.to_first + e_tail(tg - (T - TAILC), s, 0) * nxy + line;
            d[0] = yend; d[nxy] = part;
        }
    }
}

bool split_circ(const SplitArgs& a, int mode) {
    return ctx().tune_circ != 0 && a.s1.circ && (mode != MODE_BURGERS || a.s2.circ);
}

template <int MODE>
cudaError_t launch_split_ends(const SplitArgs& a, cudaStream_t st) {
    const bool last = (a.t0 + a.Tl == a.T);
    const bool trim = ctx().tune_split_trim != 0;
    const bool circ = split_circ(a, MODE);
    int nfirst = (circ || (trim && a.t0 != 0)) ? LBM : MW, nlast = (last && !circ) ? LBM + MW : LBM;
    int nlist = nfirst + nlast;
    if (nlist >= a.Tl) { nfirst = a.Tl; nlist = a.Tl; }        // the two sets meet: every chunk once
    const dim3 grid((unsigned)(a.nxy / ML), (unsigned)((nlist + MW - 1) / MW), 1);
    splitz_ends_kernel<MODE><<<grid, MW * ML, 0, st>>>(a, nfirst, nlist, trim, circ);
    return cudaGetLastError();
}

bool split_march_ok(const SplitArgs& a, int mode) {
    if (a.Tl % MW != 0 || a.Tl < 2 * MW || a.nxy % ML != 0) return false;
    if (!a.s1.march_ok || a.s1.K0m > MW || a.s1.K1m > MW) return false;
    if (mode == MODE_BURGERS && (!a.s2.march_ok || a.s2.K0m > MW || a.s2.K1m > MW)) return false;
    return true;
}

template <int MODE>
cudaError_t launch_split_march(const SplitArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)((MODE == MODE_BURGERS) ? 2 : 1) * M_SYS * sizeof(double);
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(splitz_march_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(splitz_march_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        set = true;
    }
    if (split_circ(a, MODE)) splitz_march_kernel<MODE, true><<<(unsigned)(a.nxy / ML), MW * ML, smem, st>>>(a);
    else splitz_march_kernel<MODE, false><<<(unsigned)(a.nxy / ML), MW * ML, smem, st>>>(a);
    return cudaGetLastError();
}

int pick_L(int Tl, long long nxy) {
    int L = 32;
    while (L > 4 && (L * Tl > 512 || nxy % L != 0)) L >>= 1;
    if (L * Tl > 512 || nxy % L != 0) return 0;
    return L;
}

}  // namespace

SplitZ& splitz() {
    static SplitZ s;
    return s;
}

int SplitZ::init(long long nxy_, int kmax_, int nzg_, int P_, int rank_, int emulate_) {
    if (ready && nxy == nxy_ && kmax == kmax_ && nzg == nzg_ && P == P_ && rank == rank_ && emulate == emulate_) return 0;
    destroy();
    nxy = nxy_; kmax = kmax_; nzg = nzg_; P = P_; rank = rank_; emulate = emulate_;
    if (kmax % CHUNK != 0 || kmax / CHUNK < LB2 || nzg / CHUNK < TAILC || nzg != kmax * P || P < 2) return 0;   // not eligible: ready stays false
    if (pick_L(kmax / CHUNK, nxy) == 0) return 0;
    const size_t bytes = (size_t)BLOCK_PLANES * nxy * sizeof(double);
    const int nblk = emulate > 1 ? P : 1;
    for (int par = 0; par < 2; par++) {
        for (int b = 0; b < nblk; b++) {
            double* p = nullptr;
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); destroy(); return fail(TLAB_ERR_ALLOC, "split-z exchange buffers"); }
            cudaMemsetAsync(p, 0, bytes, ctx().stream);
            block[par].push_back(p);
        }
        if (emulate <= 1) {
            // collective: every rank registers its block in the same order
            if (int rc = trp().register_buffer(block[par][0])) { destroy(); return rc; }
            if (!trp().find(block[par][0])) { destroy(); return 0; }     // no peer access: the transposes stay
        }
    }
    ready = true;
    return 0;
}

void SplitZ::destroy() {
    for (int par = 0; par < 2; par++) {
        for (double* p : block[par]) {
            if (emulate <= 1) trp().unregister_buffer(p);
            cudaFree(p);
        }
        block[par].clear();
    }
    ready = false;
}

bool SplitZ::eligible(const tlab_plan_s* g, int is) const {
    if (!ready || !g) return false;
    const DevPlan& p = g->p;
    if (!p.periodic || p.need_1der || p.n != nzg) return false;
    auto ok = [](const Sys2& s) { return s.ok && s.K0 <= LB2 && s.K1 <= LB2; };
    if (!ok(p.sys1[0])) return false;
    if (is >= 0) {
        if (g->burgers_first < 0 || is >= g->burgers_count) return false;
        if (!ok(p.sys2[g->burgers_first + is])) return false;
    }
    return true;
}

namespace {

int run_split(SplitZ& z, int mode, tlab_plan_s* g, int is, const double* u, const double* u2, double scale, const double* vel,
              double* out, int accumulate) {
    const DevPlan& p = g->p;
    const int par = (int)(z.calls++ & 1);
    z.ops++;
    cudaStream_t st = ctx().stream;
    SplitArgs a;
    a.T = z.nzg / CHUNK; a.Tl = z.kmax / CHUNK;
    a.L = pick_L(a.Tl, z.nxy);
    a.lshift = 0;
    while ((1 << a.lshift) < a.L) a.lshift++;
    a.accumulate = accumulate; a.scale = scale; a.nxy = z.nxy;
    a.rhs1 = p.rhs1[0]; a.rhs2 = p.rhs2;
    a.s1 = p.sys1[0];
    if (mode == MODE_BURGERS) a.s2 = p.sys2[g->burgers_first + is];
    // the registry is emptied when a later registration finds that not every rank can map peer memory (trp.cu)
    if (z.emulate <= 1 && !trp().find(z.block[par][0]))
        return fail(TLAB_ERR_OPTION, "split-z operators: the peer mappings of the exchange blocks are gone; re-create the DNS handle");
    const int nv = z.emulate > 1 ? z.P : 1;          // slabs handled by this process
    const long long slab = (long long)z.kmax * z.nxy;
    auto blk = [&](int r) -> double* {
        if (z.emulate > 1) return z.block[par][r];
        return trp().find(z.block[par][0])->p[r];
    };
    const unsigned push_ctas = (unsigned)std::min<long long>((3 * z.nxy + 255) / 256, 4 * 148);
    // halo planes into the neighbours' blocks
    {
        ProfScope ps(PC_TRANSPOSE);
        for (int v = 0; v < nv; v++) {
            const int r = z.emulate > 1 ? v : z.rank;
            const int prev = (r + z.P - 1) % z.P, next = (r + 1) % z.P;
            const long long fo = z.emulate > 1 ? (long long)v * slab : 0;
            splitz_push_kernel<<<push_ctas, 256, 0, st>>>(u + fo, u2 ? u2 + fo : nullptr, scale, z.nxy, z.kmax,
                                                           blk(prev) + (long long)OFF_HHI * z.nxy, blk(next) + (long long)OFF_HLO * z.nxy);
        }
        if (z.emulate <= 1) { if (int rc = trp().barrier()) return rc; }
    }
    ProfScope ps(mode == MODE_BURGERS ? PC_BURGERS_Z : PC_PARTIAL_Z);
    for (int phase = 1; phase <= 2; phase++) {
        for (int v = 0; v < nv; v++) {
            const int r = z.emulate > 1 ? v : z.rank;
            const int prev = (r + z.P - 1) % z.P, next = (r + 1) % z.P;
            const long long fo = z.emulate > 1 ? (long long)v * slab : 0;
            a.t0 = r * a.Tl;
            a.u = u + fo; a.u2 = u2 ? u2 + fo : nullptr; a.vel = vel ? vel + fo : nullptr; a.out = out + fo;
            a.mine = blk(r); a.to_prev = blk(prev); a.to_next = blk(next); a.to_first = blk(0);
            if (ctx().tune_split_local) a.to_prev = a.to_next = a.to_first = blk(r);   // timing experiment only (wrong results): no NVLink stores
            cudaError_t e;
            if (ctx().tune_march && split_march_ok(a, mode)) {
                if (phase == 1) {
                    e = (mode == MODE_BURGERS) ? launch_split_ends<MODE_BURGERS>(a, st) : launch_split_ends<MODE_P1>(a, st);
                } else {
                    z.march_ops++;
                    e = (mode == MODE_BURGERS) ? launch_split_march<MODE_BURGERS>(a, st) : launch_split_march<MODE_P1>(a, st);
                }
            } else {
                e = (mode == MODE_BURGERS) ? launch_split<MODE_BURGERS>(phase, a, st) : launch_split<MODE_P1>(phase, a, st);
            }
            if (e != cudaSuccess) return cuda_check(e, "split-z kernel");
        }
        if (phase == 1 && z.emulate <= 1) { if (int rc = trp().barrier()) return rc; }
    }
    return 0;
}

}  // namespace

int SplitZ::burgers(tlab_plan_s* g, int is, const double* s, const double* vel, double* out, int accumulate) {
    return run_split(*this, MODE_BURGERS, g, is, s, nullptr, 0.0, vel, out, accumulate);
}

int SplitZ::partial(tlab_plan_s* g, const double* a, const double* a2, double scale, double* out, int accumulate) {
    return run_split(*this, MODE_P1, g, -1, a, a2, scale, nullptr, out, accumulate);
}

}  // namespace tlab
