// Device pieces shared by the marching-panel kernels (march.cu) and their split-domain variant (splitz.cu): geometry of a
// panel, the shared-memory areas of one system, look-back / look-ahead over the rings, the finishing step.
#pragma once
#include "lines2_dev.cuh"

namespace tlab {
namespace {

constexpr int MW = 4;                       // warps per CTA = chunks per round
constexpr int LBM = 3;                      // look-back / look-ahead window in chunks
constexpr int ML = 32;                      // lines per panel (lane = line)
constexpr int M_X = C * MW * ML;            // stash of zero-inflow solutions, item (j, w, lane)
constexpr int M_R = 2 * MW * ML;            // ring of chunk ends: two halves (steps alternate) of MW chunks
constexpr int M_ZK = LBM * ML;              // circulant: z of the first LBM chunks of round 1, needed again by round 0 at the end
constexpr int M_SYS = M_X + 3 * M_R + M_ZK; // X | Y | Z | Wc | Zk   (doubles per system)

struct MarchSm {
    double *X, *Y, *Z, *Wc, *Zk;
    __device__ __forceinline__ explicit MarchSm(double* p) : X(p), Y(p + M_X), Z(p + M_X + M_R), Wc(p + M_X + 2 * M_R), Zk(p + M_X + 3 * M_R) {}
};

// chunk t of the thread's line (+ 3-point halos, wrapped or zero) -> registers
template <bool PER>
__device__ __forceinline__ void march_load(double (&u)[C + 6], const double* __restrict__ p, const double* __restrict__ p2, double scale,
                                           int t, int T, int n, long long st) {
    const bool lok = PER || t > 0, rok = PER || t < T - 1;
    const long long loff = (t > 0) ? -3 * st : (long long)(n - 3) * st;
    const long long roff = (t < T - 1) ? (long long)C * st : -(long long)(t * C) * st;
    const double* __restrict__ pc = p + (long long)(t * C) * st;
    {
        // running pointers: no table of j * stride offsets to keep in registers
        const double* q = pc;
#pragma unroll
        for (int j = 0; j < C; j++) { u[j + 3] = __ldcs(q); q += st; }
        const double* ql = pc + loff;
        const double* qr = pc + roff;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            u[k] = lok ? __ldcs(ql) : 0.0;
            u[C + 3 + k] = rok ? __ldcs(qr) : 0.0;
            ql += st; qr += st;
        }
    }
    if (p2 != nullptr) {
        const double* q = p2 + (long long)(t * C) * st;
        const double* ql = q + loff;
        const double* qr = q + roff;
#pragma unroll
        for (int j = 0; j < C; j++) { u[j + 3] = u[j + 3] + __ldcs(q) * scale; q += st; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lok) u[k] = u[k] + __ldcs(ql) * scale;
            if (rok) u[C + 3 + k] = u[C + 3 + k] + __ldcs(qr) * scale;
            ql += st; qr += st;
        }
    }
}

template <bool PER, bool SECOND>
__device__ __forceinline__ void march_rhs(const double (&u)[C + 6], double (&f)[C], const RhsTab& R, int t, int T) {
    rhs_interior<SECOND>(u, f, R);
    if (!PER) {
        if (t == 0) rhs_bottom(u, f, R);
        if (t == T - 1) rhs_top(u, f, R);
    }
}

// zero-inflow sweeps of chunk t (warp-uniform: constants or broadcast table reads)
template <bool PER>
__device__ __forceinline__ void march_local(double (&f)[C], const Sys2& S, int t, double& yend, double& part) {
    part = 0.0;
    if (__ldg(S.crec + (size_t)t * 16 + 14) != 0.0) local_const(f, S, yend);
    else local_tab<PER>(f, tab_ptr(S, t), yend, part);
}

// forward end value only (pre-step of the circulant march)
__device__ __forceinline__ double march_forward_end(const double (&f)[C], const Sys2& S, int t) {
    double e = 0.0;
    if (__ldg(S.crec + (size_t)t * 16 + 14) != 0.0) {
#pragma unroll
        for (int j = 0; j < C; j++) e = fma(S.ca, e, f[j]);
    } else {
        const double2* tp = tab_ptr(S, t);
#pragma unroll
        for (int j = 0; j < C; j++) e = fma(ldg2(tp + (j * 4 + 0) * 8).x, e, f[j]);
    }
    return e;
}

__device__ __forceinline__ double march_forward_end_const(const double (&f)[C], const Sys2& S) {
    double e = 0.0;
#pragma unroll
    for (int j = 0; j < C; j++) e = fma(S.ca, e, f[j]);
    return e;
}
// the same look-back / look-ahead with the constant weights of the circulant form
__device__ __forceinline__ double march_look_back_w(const double* __restrict__ Y, double w0, double w1, double w2, int h, int w, int lane) {
    const double* cur = Y + (h * MW) * ML + lane;
    const double* prv = Y + ((h ^ 1) * MW) * ML + lane;
    double A = w0 * ((w >= 1) ? cur[(w - 1) * ML] : prv[(MW + w - 1) * ML]);
    A = fma(w1, (w >= 2) ? cur[(w - 2) * ML] : prv[(MW + w - 2) * ML], A);
    A = fma(w2, (w >= 3) ? cur[(w - 3) * ML] : prv[(MW + w - 3) * ML], A);
    return A;
}
__device__ __forceinline__ double march_look_ahead_w(const double* __restrict__ own, const double* __restrict__ nxt,
                                                     double w0, double w1, double w2, int w, int lane) {
    double B = w0 * ((w + 1 < MW) ? own[(w + 1) * ML + lane] : nxt[(w + 1 - MW) * ML + lane]);
    B = fma(w1, (w + 2 < MW) ? own[(w + 2) * ML + lane] : nxt[(w + 2 - MW) * ML + lane], B);
    B = fma(w2, (w + 3 < MW) ? own[(w + 3) * ML + lane] : nxt[(w + 3 - MW) * ML + lane], B);
    return B;
}

// A of chunk t = round * MW + w from the forward ends of the LBM chunks before it (this round: half h, previous round: other half)
__device__ __forceinline__ double march_look_back(const double* __restrict__ Y, const double* __restrict__ cr, int h, int w, int lane) {
    const double* cur = Y + (h * MW) * ML + lane;
    const double* prv = Y + ((h ^ 1) * MW) * ML + lane;
    double A = __ldg(cr + 0) * ((w >= 1) ? cur[(w - 1) * ML] : prv[(MW + w - 1) * ML]);
    A = fma(__ldg(cr + 1), (w >= 2) ? cur[(w - 2) * ML] : prv[(MW + w - 2) * ML], A);
    A = fma(__ldg(cr + 2), (w >= 3) ? cur[(w - 3) * ML] : prv[(MW + w - 3) * ML], A);
    return A;
}
// B of chunk w of the round in `own` from the chunk starts after it (same round, then the next round in `nxt`)
__device__ __forceinline__ double march_look_ahead(const double* __restrict__ own, const double* __restrict__ nxt,
                                                   const double* __restrict__ cr, int w, int lane) {
    double B = __ldg(cr + LB2 + 0) * ((w + 1 < MW) ? own[(w + 1) * ML + lane] : nxt[(w + 1 - MW) * ML + lane]);
    B = fma(__ldg(cr + LB2 + 1), (w + 2 < MW) ? own[(w + 2) * ML + lane] : nxt[(w + 2 - MW) * ML + lane], B);
    B = fma(__ldg(cr + LB2 + 2), (w + 3 < MW) ? own[(w + 3) * ML + lane] : nxt[(w + 3 - MW) * ML + lane], B);
    return B;
}

// x = x^ + Q A + R B [+ S x_N] for chunk t (same expressions as finish_const / finish_tab)
template <bool PER>
__device__ __forceinline__ void march_finish(double (&x)[C], const Sys2& S, const MarchSm& m, int t, int T, double A, double B, int lane) {
    if (__ldg(S.crec + (size_t)t * 16 + 14) != 0.0) {
        if (!PER && S.rho != nullptr) {
            // unscaled constant chunk of a non-periodic direction (plan.cu): B as w, the solution scaled by rho
            finish_const(x, S, A, B * __ldg(S.crec + (size_t)t * 16 + 15));
            scale_rho(x, S, t);
        } else {
            finish_const(x, S, A, B);
        }
    } else {
        double xN = 0.0;
        if (PER) {
            for (int k = 0; k < S.K0m; k++) xN += m.Wc[k * ML + lane];
            for (int k = 0; k < S.K1m; k++) xN += m.Wc[(MW + k) * ML + lane];
        }
        finish_tab<PER>(x, tab_ptr(S, t), A, B, xN);
    }
}


}  // namespace
}  // namespace tlab
