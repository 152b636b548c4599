// Marching-panel variant of the strided line operators (y and z directions) for sm_100a.
//
// Same operators and the same arithmetic as lines2.cu -- OPR_Partial P1 and OPR_Burgers of the reference
// (src/operators/opr_partial.f90:31-377, src/physics/opr_burgers.f90:439-521; banded products src/fdm/fdm_matmul.f90,
// Thomas sweeps src/utils/linear3.f90:56-150,321-442) in the chunked formulation described there: every chunk of 16 points
// runs its two substitution sweeps once with zero inflow and is corrected with the true inflow values A (from the chunk ends
// before it) and B (from the chunk starts after it).
//
// What differs is who holds what.  lines2_strided keeps a whole line in the registers of one CTA (T = n/16 threads per line),
// so the number of lines per CTA -- the width of the rows the CTA reads -- shrinks with the line length (128 bytes at n = 512,
// 64 bytes at n = 1024), a tile fills the register file of an SM, and nothing overlaps the load, solve and store phases of a
// tile.  Here a CTA of MW warps owns a panel of 32 adjacent lines (lane = line: every row access of a warp is one 256-byte
// segment whatever n is) and marches along the lines in rounds of MW chunks (warp = chunk).  The corrections reach LBM = 3
// chunks (a chunk multiplies its inflow by <= 1.5e-6; the dropped fourth term is < 2^-56, checked per plan), so the chunks of a
// round are finished one round later: their zero-inflow solutions wait in a thread-private shared-memory stash (swapped
// against the new ones), and the only values that cross threads are the chunk ends (two small rings).  A CTA needs 46 KB of
// shared memory and 128 threads, four CTAs are resident per SM and are in different phases at any time: loads of one overlap
// the sweeps and barriers of the others.
//
// Circulant systems (periodic directions) are solved in circulant form (plan.h, Sys2::circ: two cyclic first-order recurrences
// with the converged LU constants -- every chunk is a constant chunk, the windows wrap around the line, no rank-one closure).
// The rounds are visited in the order 1, 2, ..., R-1, 0, so that round 0 finds its predecessors (round R-1) and round R-1 its
// successors (round 0) in the rings like any other round; the look-back of round 1 needs the forward ends of the last LBM chunks
// of round 0, which a short pre-step computes (re-reading 3 chunks of the line), and the chunk starts of round 1 are kept (Zk)
// for the look-ahead of round 0 at the end.
//
// Non-uniform grids: the Jacobian term of the second derivative is a diagonal correction of the solution (see lines2.cu,
// jacobian_correction), applied when a chunk is finished.
#include "march_dev.cuh"
#include <algorithm>

namespace tlab {

namespace {

// MODE_P1: out = d/ds (u [+ scale u2]), MODE_BURGERS: out = d2 - vel * d1 (d2 from the diffusivity-scaled system a.s2);
// accumulate = +1 / -1 adds to / subtracts from out.  JAC: non-uniform direction (Jacobian correction of d2).
// RED: the accumulation is a fire-and-forget red.global.add.f64 (one IEEE addition per element either way: same bits).
// One step of the march: the chunks of round rf enter (loads, right-hand sides, zero-inflow sweeps, ends into the rings) and
// the chunks of the round that entered one step earlier are finished.  GF / GB = false: peeled variants for a non-periodic
// direction whose entering / finishing round holds interior constant chunks only (no wall rows, no tables, no dispatch on the
// chunk records: what keeps the generic body above 128 registers).
template <int MODE, bool PER, bool JAC, bool RED, bool VPRE, bool GF, bool GB>
__device__ __forceinline__ void march_step(const Line2Args& a, const MarchSm& m1, const MarchSm& m2, const double* __restrict__ pu,
                                           const double* __restrict__ pu2, long long base, int lane, int w, int slot, int s,
                                           double (&o1)[C], double (&o2)[C], double& A1p, double& A2p) {
    constexpr bool TWO = (MODE == MODE_BURGERS);
    const int T = a.T, n = a.n, R = T / MW;
    const long long st = a.stride;
    const Sys2& S1 = a.s1;
    const Sys2& S2 = a.s2;
    {
        const bool front = s < R, back = s > 0;
        const int rf = PER ? ((s + 1 == R) ? 0 : s + 1) : s;          // chunk-order round entering the pipeline
        const int rb = PER ? ((s == R) ? 0 : s) : s - 1;              // round being finished (entered one step earlier)
        const int h = s & 1;
        const int t = rf * MW + w, tb = rb * MW + w;
        double ye1 = 0.0, ye2 = 0.0, pt1 = 0.0, pt2 = 0.0, x01 = 0.0, x02 = 0.0;
        const long long boff = base + (long long)(tb * C) * st;
        if (a.march_pf && back) {
            // what the finishing stage of this step will read (velocity, accumulation target) goes to L2 while the new round is swept
            const double* pv = a.vel + boff;
            const double* pa = a.out1 + boff;
#pragma unroll
            for (int j = 0; j < C; j++) {
                if (TWO) prefetch_l2(pv);
                if (a.accumulate != 0) prefetch_l2(pa);
                pv += st; pa += st;
            }
        }
        if (front) {
            double u[C + 6], f1[C], f2[C];
            if (!PER && !GF) {
                // peeled interior round of a non-periodic direction: halos on both sides, no wall rows, constant chunks
                march_load<true>(u, pu, pu2, a.scale, t, T, n, st);
                rhs_interior<false>(u, f1, a.rhs1);
                if (TWO) rhs_interior<true>(u, f2, a.rhs2);
            } else {
                march_load<PER>(u, pu, pu2, a.scale, t, T, n, st);
                march_rhs<PER, false>(u, f1, a.rhs1, t, T);
                if (TWO) march_rhs<PER, true>(u, f2, a.rhs2, t, T);
            }
            if (PER || !GF) {
                // circulant form: every chunk is a constant chunk
                if (TWO) local_const2(f1, f2, S1, S2, ye1, ye2);
                else local_const(f1, S1, ye1);
            } else {
                march_local<PER>(f1, S1, t, ye1, pt1);
                if (TWO) march_local<PER>(f2, S2, t, ye2, pt2);
            }
            x01 = f1[0];
            if (TWO) x02 = f2[0];
            if (!PER && !GF) {
                x01 *= rho_first(S1, t);
                if (TWO) x02 *= rho_first(S2, t);
            } else if (!PER) {
                // constant chunks of a non-periodic direction work in the unscaled variable (plan.cu): the chunk start goes out as x
                if (S1.rho != nullptr && __ldg(S1.crec + (size_t)t * 16 + 14) != 0.0) x01 *= rho_first(S1, t);
                if (TWO && S2.rho != nullptr && __ldg(S2.crec + (size_t)t * 16 + 14) != 0.0) x02 *= rho_first(S2, t);
            }
            // swap with the stash: the previous chunk's solutions come out, this chunk's go in (thread-private slots)
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
        // the advecting velocity of the chunk being finished: in flight across the two barriers
        double vv[C];
        if (TWO && back && VPRE) {
            const double* vp = a.vel + boff;
#pragma unroll
            for (int j = 0; j < C; j++) { vv[j] = __ldcs(vp); vp += st; }
        }
        __syncthreads();
        double A1 = 0.0, A2 = 0.0;
        if (front) {
            const double* cr1 = S1.crec + (size_t)t * 16;
            A1 = PER ? march_look_back_w(m1.Y, S1.cwf[0], S1.cwf[1], S1.cwf[2], h, w, lane) : march_look_back(m1.Y, cr1, h, w, lane);
            const double z1 = fma(PER ? S1.cQ[0] : __ldg(cr1 + 12), A1, x01);
            m1.Z[(h * MW) * ML + slot] = z1;
            if (PER && rf == 1 && w < LBM) m1.Zk[slot] = z1;
            if (TWO) {
                const double* cr2 = S2.crec + (size_t)t * 16;
                A2 = PER ? march_look_back_w(m2.Y, S2.cwf[0], S2.cwf[1], S2.cwf[2], h, w, lane) : march_look_back(m2.Y, cr2, h, w, lane);
                const double z2 = fma(PER ? S2.cQ[0] : __ldg(cr2 + 12), A2, x02);
                m2.Z[(h * MW) * ML + slot] = z2;
                if (PER && rf == 1 && w < LBM) m2.Zk[slot] = z2;
            }
        }
        __syncthreads();
        if (back) {
            if (TWO && !VPRE) {
                const double* vp = a.vel + boff;
#pragma unroll
                for (int j = 0; j < C; j++) { vv[j] = __ldcs(vp); vp += st; }
            }
            const bool wrap = PER && s == R;       // round 0 of a circulant line is finished last: its successors are the kept z
            const double B1 = PER ? march_look_ahead_w(m1.Z + ((h ^ 1) * MW) * ML, wrap ? m1.Zk : m1.Z + (h * MW) * ML,
                                                       S1.cwb[0], S1.cwb[1], S1.cwb[2], w, lane)
                                  : march_look_ahead(m1.Z + ((h ^ 1) * MW) * ML, m1.Z + (h * MW) * ML, S1.crec + (size_t)tb * 16, w, lane);
            if (PER) { finish_const(o1, S1, A1p, B1); scale_rho(o1, S1, tb); }
            else if (!GB) { finish_const(o1, S1, A1p, B1 * __ldg(S1.crec + (size_t)tb * 16 + 15)); scale_rho(o1, S1, tb); }
            else march_finish<PER>(o1, S1, m1, tb, T, A1p, B1, lane);
            if (TWO) {
                const double B2 = PER ? march_look_ahead_w(m2.Z + ((h ^ 1) * MW) * ML, wrap ? m2.Zk : m2.Z + (h * MW) * ML,
                                                           S2.cwb[0], S2.cwb[1], S2.cwb[2], w, lane)
                                      : march_look_ahead(m2.Z + ((h ^ 1) * MW) * ML, m2.Z + (h * MW) * ML, S2.crec + (size_t)tb * 16, w, lane);
                if (PER) { finish_const(o2, S2, A2p, B2); scale_rho(o2, S2, tb); }
                else if (!GB) { finish_const(o2, S2, A2p, B2 * __ldg(S2.crec + (size_t)tb * 16 + 15)); scale_rho(o2, S2, tb); }
                else march_finish<PER>(o2, S2, m2, tb, T, A2p, B2, lane);
                if (JAC) {
                    const double* cp = a.cjac + ((size_t)(tb >> 3) * C) * 8 + (tb & 7);
#pragma unroll
                    for (int j = 0; j < C; j++) o2[j] = fma(-(S2.jscale * __ldg(cp + j * 8)), o1[j], o2[j]);
                }
#pragma unroll
                for (int j = 0; j < C; j++) o1[j] = o2[j] - vv[j] * o1[j];
            }
            double* po = a.out1 + boff;
            if (a.accumulate == 0) {
#pragma unroll
                for (int j = 0; j < C; j++) { __stcs(po, o1[j]); po += st; }
            } else if (RED) {
#pragma unroll
                for (int j = 0; j < C; j++) { atomicAdd(po, (a.accumulate > 0) ? o1[j] : -o1[j]); po += st; }
            } else {
                double oo[C];
                const double* pi = po;
#pragma unroll
                for (int j = 0; j < C; j++) { oo[j] = __ldcs(pi); pi += st; }
#pragma unroll
                for (int j = 0; j < C; j++) { __stcs(po, (a.accumulate > 0) ? oo[j] + o1[j] : oo[j] - o1[j]); po += st; }
            }
        }
        A1p = A1;
        A2p = A2;
    }
}

template <int MODE, bool PER, bool JAC, bool RED, int MINB, bool VPRE>
__global__ void __launch_bounds__(MW * ML, MINB) lines2_march(const __grid_constant__ Line2Args a) {
    constexpr bool TWO = (MODE == MODE_BURGERS);
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int T = a.T, n = a.n, R = T / MW;
    const long long st = a.stride;
    const long long base = (long long)blockIdx.y * a.outer_stride + (long long)blockIdx.x * ML + lane;
    const Sys2& S1 = a.s1;
    const Sys2& S2 = a.s2;
    const MarchSm m1(sm), m2(sm + M_SYS);
    // stale ring slots are read with zero weights: they must hold finite numbers
    for (int i = threadIdx.x; i < (TWO ? 2 : 1) * M_SYS; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();

    const double* __restrict__ pu = a.u + base;
    const double* __restrict__ pu2 = (a.u2 != nullptr) ? a.u2 + base : nullptr;
    const int slot = w * ML + lane;                // this thread's place in a ring half / its stash column

    if (PER && w >= MW - LBM) {
        // pre-step: forward ends of chunks MW-LBM .. MW-1 of round 0, as seen by the look-back of round 1 (half 1 = "previous")
        double u[C + 6], f[C];
        march_load<PER>(u, pu, pu2, a.scale, w, T, n, st);
        march_rhs<PER, false>(u, f, a.rhs1, w, T);
        m1.Y[MW * ML + slot] = march_forward_end_const(f, S1);
        if (TWO) {
            march_rhs<PER, true>(u, f, a.rhs2, w, T);
            m2.Y[MW * ML + slot] = march_forward_end_const(f, S2);
        }
    }

    double o1[C], o2[C];                           // zero-inflow solutions of the chunk this thread handled one step earlier
    double A1p = 0.0, A2p = 0.0;                   // and its A
    if (!PER && a.march_peel) {
        // rounds 1 .. R-2 hold interior constant chunks only (checked on the host, R >= 3)
        march_step<MODE, PER, JAC, RED, VPRE, true, true>(a, m1, m2, pu, pu2, base, lane, w, slot, 0, o1, o2, A1p, A2p);
        march_step<MODE, PER, JAC, RED, VPRE, false, true>(a, m1, m2, pu, pu2, base, lane, w, slot, 1, o1, o2, A1p, A2p);
        for (int s = 2; s <= R - 2; s++)
            march_step<MODE, PER, JAC, RED, VPRE, false, false>(a, m1, m2, pu, pu2, base, lane, w, slot, s, o1, o2, A1p, A2p);
        march_step<MODE, PER, JAC, RED, VPRE, true, false>(a, m1, m2, pu, pu2, base, lane, w, slot, R - 1, o1, o2, A1p, A2p);
        march_step<MODE, PER, JAC, RED, VPRE, true, true>(a, m1, m2, pu, pu2, base, lane, w, slot, R, o1, o2, A1p, A2p);
    } else {
        for (int s = 0; s <= R; s++)
            march_step<MODE, PER, JAC, RED, VPRE, true, true>(a, m1, m2, pu, pu2, base, lane, w, slot, s, o1, o2, A1p, A2p);
    }
}

template <int MODE, bool PER, bool JAC, bool RED, int MINB, bool VPRE>
cudaError_t launch_march_v(const Line2Args& a, dim3 grid, cudaStream_t stream) {
    const size_t smem = (size_t)((MODE == MODE_BURGERS) ? 2 : 1) * M_SYS * sizeof(double);
    auto k = lines2_march<MODE, PER, JAC, RED, MINB, VPRE>;
    static bool set = false;
    if (!set) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; set = true; }
    k<<<grid, MW * ML, smem, stream>>>(a);
    return cudaGetLastError();
}

// march_cfg (tuning key "march_cfg"): CTAs per SM the kernel is compiled for (3: 168 registers, 4: 128) and whether the
// velocity of the chunk being finished is requested before the barriers (+10) or after them
template <int MODE, bool PER, bool JAC>
cudaError_t launch_march_k(const Line2Args& a, dim3 grid, cudaStream_t stream) {
    const bool red = a.march_red && a.accumulate != 0;
    const int cfg = a.march_cfg;
    if (MODE != MODE_BURGERS) {
        if (cfg % 10 == 3) return red ? launch_march_v<MODE, PER, JAC, true, 3, false>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 3, false>(a, grid, stream);
        if (cfg % 10 == 5) return red ? launch_march_v<MODE, PER, JAC, true, 5, false>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 5, false>(a, grid, stream);
        return red ? launch_march_v<MODE, PER, JAC, true, 4, false>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 4, false>(a, grid, stream);
    }
    if (cfg == 3) return red ? launch_march_v<MODE, PER, JAC, true, 3, false>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 3, false>(a, grid, stream);
    if (cfg == 13) return red ? launch_march_v<MODE, PER, JAC, true, 3, true>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 3, true>(a, grid, stream);
    if (cfg == 14) return red ? launch_march_v<MODE, PER, JAC, true, 4, true>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 4, true>(a, grid, stream);
    return red ? launch_march_v<MODE, PER, JAC, true, 4, false>(a, grid, stream) : launch_march_v<MODE, PER, JAC, false, 4, false>(a, grid, stream);
}

}  // namespace

// does the window of LBM chunks suffice for this system (dropped weights below 2^-56 = eps/8: the compact schemes of the
// reference multiply an inflow by 2e-7 (first derivative) and 1.5e-6 (second) per chunk, i.e. 8e-21 and 3.4e-18 after three) and
// do the closure chunks fit a round?
bool march_sys_ok(const std::vector<double>& crec, int T, int K0, int K1, bool periodic) {
    const double tiny = 1.3877787807814457e-17;     // 2^-56
    for (int t = 0; t < T; t++) {
        const double* c = &crec[(size_t)t * 16];
        for (int k = LBM; k < LB2; k++)
            if (std::fabs(c[k]) > tiny || std::fabs(c[LB2 + k]) > tiny) return false;
    }
    if (periodic && (K0 > MW || K1 > MW)) return false;
    return true;
}

bool march_eligible(int mode, const Line2Args& a, bool periodic, bool need1, long long nlines, long long inner) {
    if (mode != MODE_P1 && mode != MODE_BURGERS) return false;
    if (a.T % MW != 0 || a.T < 2 * MW || a.n != a.T * CHUNK) return false;
    if (inner % ML != 0 || nlines % inner != 0) return false;
    if (inner / ML > 0x7fffffffLL || nlines / inner > 65535) return false;
    if (!a.s1.ok || !a.s1.march_ok) return false;
    if (mode == MODE_BURGERS && (!a.s2.ok || !a.s2.march_ok)) return false;
    if (mode == MODE_BURGERS && a.u2 != nullptr) return false;
    if (need1 && mode == MODE_BURGERS && a.cjac == nullptr) return false;
    if (a.nf > 0) return false;
    // periodic lines march in circulant form only (constant chunks, wrapping windows; lines2.cu / plan.h, Sys2::circ)
    if (periodic && (!a.s1.circ || (mode == MODE_BURGERS && !a.s2.circ))) return false;
    return true;
}

// peeled steps: the rounds between the first and the last hold unscaled constant chunks only, in every system of the launch
bool march_peelable(int mode, const Line2Args& a, bool periodic) {
    const int R = a.T / MW;
    auto inner_const = [&](const Sys2& S) { return S.rho != nullptr && S.c_lo <= MW && S.c_hi >= a.T - MW - 1; };
    return !periodic && R >= 3 && a.march_peel >= 0 && inner_const(a.s1) && (mode != MODE_BURGERS || inner_const(a.s2));
}

cudaError_t launch_march(int mode, const Line2Args& a_in, bool periodic, bool need1, long long nlines, long long inner, cudaStream_t s) {
    const dim3 grid((unsigned)(inner / ML), (unsigned)(nlines / inner), 1);
    Line2Args a = a_in;
    a.march_peel = march_peelable(mode, a, periodic) ? 1 : 0;
    if (mode == MODE_P1) {
        return periodic ? launch_march_k<MODE_P1, true, false>(a, grid, s) : launch_march_k<MODE_P1, false, false>(a, grid, s);
    }
    if (periodic) return launch_march_k<MODE_BURGERS, true, false>(a, grid, s);
    return need1 ? launch_march_k<MODE_BURGERS, false, true>(a, grid, s) : launch_march_k<MODE_BURGERS, false, false>(a, grid, s);
}

}  // namespace tlab
