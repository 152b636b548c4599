// Device helpers shared by the line kernels (lines2.cu) and the split-domain z kernels (splitz.cu): banded right-hand
// sides and the zero-inflow chunk sweeps of the factored tridiagonal systems (see lines2.cu for the formulation).
#pragma once
#include "lines2.h"

namespace tlab {
namespace {

constexpr int C = CHUNK;
constexpr int XB = C + 2;          // padded chunk block of the x tile (doubles)

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

// ------------------------------------------------------------------------------------------------
// banded right-hand sides (explicitly rounded, reference association order; see lines.cu)
template <bool SECOND>
__device__ __forceinline__ void rhs_interior(const double (&u)[C + 6], double (&f)[C], const RhsTab& R) {
#pragma unroll
    for (int j = 0; j < C; j++) {
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

// special rows at the walls.  All BROW_W terms are added in the reference's column order; a zero coefficient
// contributes an exact zero, so the sum is the one of the reference's (sparser) row.
__device__ __forceinline__ void rhs_bottom(const double (&u)[C + 6], double (&f)[C], const RhsTab& R) {
#pragma unroll
    for (int i = 0; i < MAX_BROWS; i++) {
        if (i < R.nb) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < BROW_W; k++) s = DADD(s, DMUL(R.bot[i][k], u[3 + k]));
            f[i] = s;
        }
    }
}
__device__ __forceinline__ void rhs_top(const double (&u)[C + 6], double (&f)[C], const RhsTab& R) {
#pragma unroll
    for (int q = 0; q < MAX_BROWS; q++) {
        if (q < R.nb) {
            double s = 0.0;
#pragma unroll
            for (int k = BROW_W - 1; k >= 0; k--) s = DADD(s, DMUL(R.top[q][k], u[3 + C - 1 - k]));
            f[C - 1 - q] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// solves.  Shared exchange area of one system: y[T*L] | z[T*L] | w[T*L]
__host__ __device__ inline int exch2_per_system(int T, int L) { return 3 * T * L; }

struct ChunkCtx {
    int t, T, l, L;
    int bx = 0;                 // tile index along the line-index direction (blockIdx.x, or half of it for a CTA pair)
    // CTA pair (thread-block cluster of 2 sharing a line): the peer's exchange area as the generic pointer that
    // map_shared_rank returned (it must not be derived from the local array: the compiler would keep it a shared-window store)
    double* rsm = nullptr;
    const double* sm0 = nullptr;
    bool pair = false;
};

// exchange of chunk ends: every value goes to this CTA's array and, for a CTA pair, to the peer's copy as well
__device__ __forceinline__ void publish(double* slot, double v, const ChunkCtx& c) {
    *slot = v;
    if (c.pair) c.rsm[slot - c.sm0] = v;
}
__device__ __forceinline__ void exchange_barrier(const ChunkCtx& c) {
    if (c.pair) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
}

__device__ __forceinline__ const double2* tab_ptr(const Sys2& S, int t) {
    return S.tab + ((size_t)(t >> 3) * C * 4) * 8 + (t & 7);
}

// zero-inflow sweeps of one chunk with tabulated coefficients
template <bool PER>
__device__ __forceinline__ void local_tab(double (&f)[C], const double2* __restrict__ tp, double& yend, double& part) {
    double e = 0.0, pr = 0.0;
#pragma unroll
    for (int j = 0; j < C; j++) {
        const double2 ap = ldg2(tp + (j * 4 + 0) * 8);
        e = fma(ap.x, e, f[j]);
        f[j] = e;
        if (PER) pr = fma(ap.y, e, pr);
    }
    yend = e;
    part = pr;
    double xb = 0.0;
#pragma unroll
    for (int j = C - 1; j >= 0; j--) {
        const double2 dg = ldg2(tp + (j * 4 + 1) * 8);
        xb = fma(dg.y, xb, dg.x * f[j]);
        f[j] = xb;
    }
}

__device__ __forceinline__ void local_const(double (&f)[C], const Sys2& S, double& yend) {
    double e = 0.0;
#pragma unroll
    for (int j = 0; j < C; j++) { e = fma(S.ca, e, f[j]); f[j] = e; }
    yend = e;
    double xb = 0.0;
#pragma unroll
    for (int j = C - 1; j >= 0; j--) { xb = fma(S.cg, xb, S.cd * f[j]); f[j] = xb; }
}

__device__ __forceinline__ void local_const2(double (&f0)[C], double (&f1)[C], const Sys2& S0, const Sys2& S1,
                                             double& yend0, double& yend1) {
    double e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int j = 0; j < C; j++) {
        e0 = fma(S0.ca, e0, f0[j]); f0[j] = e0;
        e1 = fma(S1.ca, e1, f1[j]); f1[j] = e1;
    }
    yend0 = e0; yend1 = e1;
    double x0 = 0.0, x1 = 0.0;
#pragma unroll
    for (int j = C - 1; j >= 0; j--) {
        x0 = fma(S0.cg, x0, S0.cd * f0[j]); f0[j] = x0;
        x1 = fma(S1.cg, x1, S1.cd * f1[j]); f1[j] = x1;
    }
}


template <bool PER>
__device__ __forceinline__ void finish_tab(double (&f)[C], const double2* __restrict__ tp, double A, double B, double xN) {
#pragma unroll
    for (int j = 0; j < C; j++) {
        const double2 qr = ldg2(tp + (j * 4 + 2) * 8);
        double v = fma(qr.x, A, fma(qr.y, B, f[j]));
        if (PER) v = fma(__ldg(&tp[(j * 4 + 3) * 8].x), xN, v);
        f[j] = v;
    }
}
// circulant form: x_j *= rho_j (the column scaling of the reference's matrix, plan.h)
__device__ __forceinline__ void scale_rho(double (&x)[C], const Sys2& S, int t) {
    const double* rp = S.rho + ((size_t)(t >> 3) * C) * 8 + (t & 7);
#pragma unroll
    for (int j = 0; j < C; j++) x[j] = x[j] * __ldg(rp + j * 8);
}
__device__ __forceinline__ double rho_first(const Sys2& S, int t) { return __ldg(S.rho + ((size_t)(t >> 3) * C) * 8 + (t & 7)); }
__device__ __forceinline__ void finish_const(double (&f)[C], const Sys2& S, double A, double B) {
#pragma unroll
    for (int j = 0; j < C; j++) f[j] = fma(S.cQ[j], A, fma(S.cR[j], B, f[j]));
}


}  // namespace
}  // namespace tlab
