// Right-hand side of the incompressible/Boussinesq equations and the low-storage Runge-Kutta substep,
// device resident.
//
// Replaces TIME_RUNGEKUTTA / TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT (src/tools/dns/time.f90:185-333,559-670),
// RHS_GLOBAL_INCOMPRESSIBLE_1 (src/tools/dns/rhs_global_incompressible_1.f90:15-405, combined mode,
// remove_divergence), TLab_Sources_Flow + Gravity_Buoyancy (src/physics/tlab_sources.f90:36-131,
// src/physics/gravity.f90:232-342; homogeneous and linear), DNS_BOUNDS_LIMIT (src/tools/dns/dns_local.f90:67-90).
//
// Fusion relative to the reference's call sequence (values identical up to round-off, association of
// every sum kept):  each Burgers operator accumulates straight into hq/hs (no tmp + add sweep);
// the pressure forcing hq + q/dte is formed while loading the pencil of the divergence kernels and the
// three divergence terms accumulate into one array; the pressure gradient is subtracted from hq by the
// derivative kernels; q += dte*hq, the scalar clipping and hq *= kco are one sweep.
#include "../../include/tlab_gpu.h"
#include "context.h"
#include "poisson.h"
#include "trp.h"
#include "splitz.h"
#include <vector>
#include <string>
#include <cstring>
#include <cfloat>
#include <cmath>
#include <algorithm>

namespace tlab {

namespace {

constexpr int EW_THREADS = 256;

inline unsigned ew_blocks(long long n) {
    long long b = (n + EW_THREADS - 1) / EW_THREADS;
    const long long cap = 148LL * 16;     // a few waves of resident CTAs; grid-stride beyond that
    return (unsigned)(b < cap ? b : cap);
}

// hq(iq) = k(iq)*hq(iq) + vector(iq) * b,   b = c1*s1 - (ref(j) - c0)   (linear)  or  b = const (homogeneous)
__global__ void buoyancy_kernel(double* __restrict__ hq1, double* __restrict__ hq2, double* __restrict__ hq3,
                                const double* __restrict__ s1, const double* __restrict__ ref, double c1, double c0,
                                double bhom, int linear, double g1, double g2, double g3, double k1, double k2, double k3,
                                int nx, int ny, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double b;
        if (linear) {
            const int j = (int)((i / nx) % ny);
            const double dummy = ref[j] - c0;
            b = c1 * s1[i] - dummy;
        } else {
            b = bhom;
        }
        // k: pending factor of the previous substep (`hq = hq*kco`, time.f90:290-297), rounded before the sum as there; 1 otherwise
        if (g1 != 0.0) hq1[i] = __dmul_rn(k1, hq1[i]) + g1 * b;
        if (g2 != 0.0) hq2[i] = __dmul_rn(k2, hq2[i]) + g2 * b;
        if (g3 != 0.0) hq3[i] = __dmul_rn(k3, hq3[i]) + g3 * b;
    }
}

__global__ void sub_kernel(double* __restrict__ a, const double* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a[i] = a[i] - b[i];
}

// q += dte*h ; optional clipping ; h *= kco
__global__ void rk_update_kernel(double* __restrict__ q, double* __restrict__ h, double dte, double kco, int scale_h,
                                 int clip, double vmin, double vmax, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double hv = h[i];
        double v = q[i] + dte * hv;
        if (clip) v = fmin(fmax(v, vmin), vmax);
        q[i] = v;
        if (scale_h) h[i] = kco * hv;
    }
}

// the same for the wall-normal velocity with the last two sweeps of the RHS folded in (rhs_global_incompressible_1.f90:
// 334-352, 356-398 for a field with Dirichlet conditions at both walls): h = h - dpdy, h = 0 on the wall planes, then the update
__global__ void rk_update_sub_kernel(double* __restrict__ q, double* __restrict__ h, const double* __restrict__ dpdy, double dte,
                                     double kco, int scale_h, int nx, int ny, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)((i / nx) % ny);
        double hv = h[i] - dpdy[i];
        if (j == 0 || j == ny - 1) hv = 0.0;
        q[i] = q[i] + dte * hv;
        h[i] = scale_h ? kco * hv : hv;
    }
}

// planes j = 0 and j = ny-1 of a field <-> (nx, nz) arrays
__global__ void get_planes_kernel(const double* __restrict__ f, double* __restrict__ hb, double* __restrict__ ht,
                                  int nx, int ny, int nz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * nz) return;
    const int i = (int)(idx % nx);
    const long long k = idx / nx;
    hb[idx] = f[i + (long long)nx * ny * k];
    ht[idx] = f[i + (long long)nx * ((ny - 1) + (long long)ny * k)];
}

__global__ void set_planes_kernel(double* __restrict__ f, const double* __restrict__ hb, const double* __restrict__ ht,
                                  int use_b, int use_t, int nx, int ny, int nz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * nz) return;
    const int i = (int)(idx % nx);
    const long long k = idx / nx;
    f[i + (long long)nx * ny * k] = use_b ? hb[idx] : 0.0;
    f[i + (long long)nx * ((ny - 1) + (long long)ny * k)] = use_t ? ht[idx] : 0.0;
}

void rk_tables(int mode, std::vector<double>& kdt, std::vector<double>& ktime, std::vector<double>& kco) {
    if (mode == TLAB_RKM_EXP3) {        // Williamson 1980 (time.f90:87-92)
        kdt = {1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0};
        ktime = {0.0, 1.0 / 3.0, 3.0 / 4.0};
        kco = {-5.0 / 9.0, -153.0 / 128.0};
    } else {                            // Carpenter & Kennedy 1994, 5 stages (time.f90:94-112)
        kdt = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
               3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0};
        ktime = {0.0, kdt[0], 2526269341429.0 / 6820363962896.0, 2006345519317.0 / 3224310063776.0,
                 2802321613138.0 / 2924317926251.0};
        kco = {-567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
               -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    }
}

// ---- per-step diagnostics --------------------------------------------------------------------------
// Replaces TIME_COURANT (src/tools/dns/time.f90:365-548, incompressible branch; the grid factors of
// TIME_INITIALIZE :136-178) and the dilatation part of DNS_BOUNDS_CONTROL (src/tools/dns/dns_local.f90:94-234)
// with FI_INVARIANT_P (src/mappings/fi_vectorcalculus.f90:111-141) and MINMAX (src/utils/minmax.f90).
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// per-block max of |u|/dx + |v|/dy + |w|/dz
__global__ void courant_kernel(const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                               const double* __restrict__ ox, const double* __restrict__ oy, const double* __restrict__ oz,
                               int nx, int ny, int koff, int three_d, long long n, double* __restrict__ partial) {
    __shared__ double sm[RED_THREADS / 32];
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % nx);
        const long long r = i / nx;
        const int iy = (int)(r % ny);
        const int iz = (int)(r / ny);
        double val = fabs(u[i]) * ox[ix] + fabs(v[i]) * oy[iy];
        if (three_d) val = val + fabs(w[i]) * oz[koff + iz];
        m = fmax(m, val);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < RED_THREADS / 32) ? sm[threadIdx.x] : 0.0;
        m = warp_max(m);
        if (threadIdx.x == 0) partial[blockIdx.x] = m;
    }
}

__global__ void minmax_kernel(const double* __restrict__ a, long long n, double* __restrict__ pmin, double* __restrict__ pmax) {
    __shared__ double smin[RED_THREADS / 32], smax[RED_THREADS / 32];
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = a[i];
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = (threadIdx.x < RED_THREADS / 32) ? smin[threadIdx.x] : DBL_MAX;
        hi = (threadIdx.x < RED_THREADS / 32) ? smax[threadIdx.x] : -DBL_MAX;
        lo = warp_min(lo); hi = warp_max(hi);
        if (threadIdx.x == 0) { pmin[blockIdx.x] = lo; pmax[blockIdx.x] = hi; }
    }
}

__global__ void negate_kernel(double* __restrict__ a, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = -a[i];
}

// MPI_ALLREDUCE(MAX) of a few doubles among the z ranks
int allreduce_max(double* host_vals, int count) {
    Trp& t = trp();
    if (t.P == 1) return 0;
#ifdef TLAB_HAVE_NCCL
    double* d = nullptr;
    if (cudaMalloc(&d, count * sizeof(double)) != cudaSuccess) return fail(TLAB_ERR_ALLOC, "allreduce buffer");
    cudaStream_t st = ctx().stream;
    cudaMemcpyAsync(d, host_vals, count * sizeof(double), cudaMemcpyHostToDevice, st);
    ncclResult_t r = ncclAllReduce(d, d, count, ncclDouble, ncclMax, t.comm, st);
    cudaMemcpyAsync(host_vals, d, count * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d);
    if (r != ncclSuccess) return fail(TLAB_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    return 0;
#else
    return fail(TLAB_ERR_UNDEVELOP, "library built without NCCL");
#endif
}


}  // namespace

struct Dns {
    tlab_dns_params prm;
    tlab_plan_s* g[3] = {nullptr, nullptr, nullptr};
    int nx = 0, ny = 0, nz = 0, ns = 0;
    long long N = 0, Nt = 0;             // points per field; points of a work array, (nx+2)*ny*nz
    std::vector<double*> q, s, hq, hs;   // device fields
    double *tmp1 = nullptr, *tmp3 = nullptr, *c1 = nullptr, *c2 = nullptr;
    double *hb = nullptr, *ht = nullptr;
    double* bbackground = nullptr;
    // z-slab decomposition: nz is the local slab thickness, nzg the global extent, pencils hold N points each
    int P = 1, nzg = 0;
    double *zs = nullptr, *zw = nullptr, *zr = nullptr;
    // diagnostics: 1/jac per direction, max of sum 1/jac^2, diffusivity factor, reduction partials
    double* ods[3] = {nullptr, nullptr, nullptr};
    double dx2i = 0.0, schmidtfactor = 0.0;
    double* red = nullptr;
    std::vector<double> kdt, ktime, kco;
    std::vector<void*> allocs;
    double* host_stage = nullptr;        // pinned staging buffer for the *_host entry points
    size_t host_stage_bytes = 0;
    long long launches = 0;
    // Per-handle operator state (two live handles of different size must not share it): the Poisson solver with its
    // cuFFT plans, eigenvalues and per-mode planes is owned here; the diffusion-scaled LU sets of OPR_Burgers live on the
    // plans and are re-pointed by activate(); the split-z exchange blocks are re-sized by activate() when the geometry differs.
    Poisson* pois = nullptr;
    int bfirst[3] = {-1, -1, -1};
    // `hq = hq*kco` of a substep is not written by the update kernels: the factor waits here (index 0-2: hq, 3+is: hs) and is
    // applied by the first accumulation of the next substep (buoyancy source, OPR_Burgers_X) or by flush_pending()
    double pending[3 + TLAB_MAX_SCAL] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    double take_pending(int f) { const double k = pending[f]; pending[f] = 1.0; return k; }
    void flush_pending() {
        for (int f = 0; f < 3 + ns; f++) {
            const double k = take_pending(f);
            if (k != 1.0) { scale_array(f < 3 ? hq[f] : hs[f - 3], k, N, ctx().stream); launches++; }
        }
    }
    bool defer_v = false;      // substep(): `hq2 -= dpdy` and the wall planes of hq2 are left to the fused update of q2
    bool v_deferred = false;   // ... and rhs() did leave them (non-overlapped schedule, Dirichlet at both walls)

    int alloc(double** p, long long count) {
        if (cudaMalloc(p, (size_t)count * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(TLAB_ERR_ALLOC, "tlab_dns_create: out of device memory");
        }
        allocs.push_back(*p);
        return cuda_check(cudaMemsetAsync(*p, 0, (size_t)count * sizeof(double), ctx().stream), "memset");
    }

    int sources_flow() {
        if (prm.buoyancy_type == 0) return 0;
        const double g1 = prm.buoyancy_vector[0], g2 = prm.buoyancy_vector[1], g3 = prm.buoyancy_vector[2];
        if (g1 == 0.0 && g2 == 0.0 && g3 == 0.0) return 0;
        const int linear = (prm.buoyancy_type == 2);
        if (linear && ns < 1) return fail(TLAB_ERR_OPTION, "linear buoyancy needs a scalar");
        ProfScope ps(PC_ELEMENTWISE);
        const double k1 = (g1 != 0.0) ? take_pending(0) : 1.0, k2 = (g2 != 0.0) ? take_pending(1) : 1.0,
                     k3 = (g3 != 0.0) ? take_pending(2) : 1.0;
        buoyancy_kernel<<<ew_blocks(N), EW_THREADS, 0, ctx().stream>>>(
            hq[0], hq[1], hq[2], linear ? s[0] : nullptr, bbackground, prm.buoyancy_params[0], prm.buoyancy_params[1],
            prm.buoyancy_params[0], linear, g1, g2, g3, k1, k2, k3, nx, ny, N);
        launches++;
        return 0;
    }

    // Burgers operator along z, accumulated into out; with a split domain through the z pencils
    // split-z path (splitz.cu): the slabs stay where they are, neighbours exchange halo planes and chunk ends.
    // P == 1 with the tuning key split_emulate = Pv > 1 runs the same kernels over Pv virtual slabs of the field (tests).
    bool use_split(int is) {
        if (nzg <= 1 || !ctx().tune_splitz) return false;
        if (P == 1) {
            const int pv = ctx().tune_split_emulate;
            if (pv < 2 || nz % pv) return false;
            if (splitz().init((long long)nx * ny, nz / pv, nz, pv, 0, pv)) return false;
        } else if (!splitz().ready || splitz().emulate > 1) return false;
        return splitz().eligible(g[2], is);
    }

    int burgers_z(int is, const double* sf, const double* w, double* out, bool self) {
        if (use_split(is)) return splitz().burgers(g[2], is, sf, w, out, +1);
        if (P == 1) return run_burgers(3, is, nx, ny, nz, 0, g[2], sf, w, out, +1);
        const long long nxy = (long long)nx * ny;
        const int nl = (int)(nxy / P);
        int rc = 0;
        const double* zsf = zw;
        if (!self) {
            if ((rc = trp().forward(sf, nullptr, 0.0, zs, nxy, nz))) return rc;
            zsf = zs;
        }
        if ((rc = run_burgers(3, is, nl, 1, nzg, 0, g[2], zsf, zw, zr, 0))) return rc;
        return trp().backward(zr, out, nxy, nz, +1);
    }

    // out (+|-)= d/dz (a + scale*a2)
    int partial_z(const double* a, const double* a2, double scale, double* out, int accumulate) {
        if (use_split(-1)) return splitz().partial(g[2], a, a2, scale, out, accumulate);
        if (P == 1) return run_partial(3, TLAB_OPR_P1, nx, ny, nz, 0, g[2], a, out, nullptr, a2, scale, accumulate);
        const long long nxy = (long long)nx * ny;
        const int nl = (int)(nxy / P);
        int rc = 0;
        if ((rc = trp().forward(a, a2, scale, zs, nxy, nz))) return rc;
        if ((rc = run_partial(3, TLAB_OPR_P1, nl, 1, nzg, 0, g[2], zs, zr, nullptr))) return rc;
        return trp().backward(zr, out, nxy, nz, accumulate);
    }

    // ---- split domain: the z operators (transposes over NVLink + pencil kernels) run on a second, high-priority stream
    // while the x and y operators of the same fields run on the main one; a field's z contribution is pulled into hq
    // after its x and y kernels (events), so the accumulations never race.
    struct StreamSwap {
        cudaStream_t saved;
        explicit StreamSwap(cudaStream_t s) : saved(ctx().stream) { ctx().stream = s; }
        ~StreamSwap() { ctx().stream = saved; }
    };
    cudaEvent_t ev_pool[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t event(int i) {
        if (!ev_pool[i]) cudaEventCreateWithFlags(&ev_pool[i], cudaEventDisableTiming);
        return ev_pool[i];
    }

    int rhs_overlapped(double dte) {
        flush_pending();
        cudaStream_t s0 = ctx().stream, s1 = trp().zstream;
        const int b0 = 0;
        int rc = 0;
        const long long nxy = (long long)nx * ny;
        const int nl = (int)(nxy / P);
        double *u = q[0], *v = q[1], *w = q[2];
        const int nf = 3 + ns;
        auto field = [&](int f) { return f < 3 ? q[f] : s[f - 3]; };
        auto target = [&](int f) { return f < 3 ? hq[f] : hs[f - 3]; };
        // -- Burgers
        cudaEventRecord(event(0), s0);
        cudaStreamWaitEvent(s1, event(0), 0);
        { StreamSwap sw(s1); rc = trp().forward(w, nullptr, 0.0, zw, nxy, nz); }
        for (int f = 0; f < nf && !rc; f++) {
            const int is = f < 3 ? 0 : f - 2;
            {
                StreamSwap sw(s1);
                const double* zsf = zw;
                if (f != 2) { rc = trp().forward(field(f), nullptr, 0.0, zs, nxy, nz); zsf = zs; }
                if (!rc) rc = run_burgers(3, is, nl, 1, nzg, 0, g[2], zsf, zw, zr, 0);
            }
            if (!rc) rc = run_burgers(1, is, nx, ny, nz, b0, g[0], field(f), u, target(f), +1);
            if (!rc) rc = run_burgers(2, is, nx, ny, nz, b0, g[1], field(f), v, target(f), +1);
            cudaEventRecord(event(1 + (f & 1)), s0);
            {
                StreamSwap sw(s1);
                cudaStreamWaitEvent(s1, event(1 + (f & 1)), 0);
                if (!rc) rc = trp().backward(zr, target(f), nxy, nz, +1);
            }
            launches += 3;
        }
        if (rc) return rc;
        // -- pressure forcing: div(hq + q/dte); the z part needs every hq complete
        cudaEventRecord(event(3), s1);
        cudaStreamWaitEvent(s0, event(3), 0);
        cudaEventRecord(event(0), s0);
        cudaStreamWaitEvent(s1, event(0), 0);
        const double dummy = 1.0 / dte;
        {
            StreamSwap sw(s1);
            rc = trp().forward(hq[2], w, dummy, zs, nxy, nz);
            if (!rc) rc = run_partial(3, TLAB_OPR_P1, nl, 1, nzg, 0, g[2], zs, zr, nullptr);
        }
        if (!rc) rc = run_partial(2, TLAB_OPR_P1, nx, ny, nz, b0, g[1], hq[1], tmp1, nullptr, v, dummy, 0);
        if (!rc) rc = run_partial(1, TLAB_OPR_P1, nx, ny, nz, b0, g[0], hq[0], tmp1, nullptr, u, dummy, +1);
        cudaEventRecord(event(1), s0);
        {
            StreamSwap sw(s1);
            cudaStreamWaitEvent(s1, event(1), 0);
            if (!rc) rc = trp().backward(zr, tmp1, nxy, nz, +1);
        }
        launches += 3;
        cudaEventRecord(event(3), s1);
        cudaStreamWaitEvent(s0, event(3), 0);
        if (rc) return rc;
        const long long np = (long long)nx * nz;
        { ProfScope ps(PC_ELEMENTWISE);
        get_planes_kernel<<<(unsigned)((np + 255) / 256), 256, 0, s0>>>(hq[1], hb, ht, nx, ny, nz); }
        launches++;
        if ((rc = pois->solve(tmp1, c1, c2, hb, ht, tmp3))) return rc;
        launches += 2;
        // -- hq -= grad p
        cudaEventRecord(event(0), s0);
        cudaStreamWaitEvent(s1, event(0), 0);
        {
            StreamSwap sw(s1);
            rc = trp().forward(tmp1, nullptr, 0.0, zs, nxy, nz);
            if (!rc) rc = run_partial(3, TLAB_OPR_P1, nl, 1, nzg, 0, g[2], zs, zr, nullptr);
            if (!rc) rc = trp().backward(zr, hq[2], nxy, nz, -1);
        }
        if (!rc) rc = run_partial(1, TLAB_OPR_P1, nx, ny, nz, b0, g[0], tmp1, hq[0], nullptr, nullptr, 0.0, -1);
        { ProfScope ps(PC_ELEMENTWISE);
        sub_kernel<<<ew_blocks(N), EW_THREADS, 0, s0>>>(hq[1], tmp3, N); }
        launches += 3;
        cudaEventRecord(event(3), s1);
        cudaStreamWaitEvent(s0, event(3), 0);
        if (rc) return rc;
        return rhs_bcs();
    }

    int activate() {
        for (int i = 0; i < 3; i++) { g[i]->burgers_first = bfirst[i]; g[i]->burgers_count = ns + 1; }
        if (!pois || !pois->ready || pois->nx != nx || pois->ny != ny || pois->nz != nz)
            return fail(TLAB_ERR_DIMGRID, "tlab_dns: the Poisson solver of this handle is not initialised for its grid");
        if (P > 1 && ctx().tune_splitz && nzg > 1)
            return splitz().init((long long)nx * ny, nz, nzg, P, trp().rank, 0);
        return 0;
    }

    int rhs(double dte) {
        if (int rc0 = activate()) return rc0;
        bool split = true;
        for (int is = 0; is <= ns && split; is++) split = use_split(is);
        v_deferred = false;
        if (P > 1 && nzg > 1 && trp().zstream && ctx().tune_overlap && !split) return rhs_overlapped(dte);
        v_deferred = defer_v && prm.bcs_flow_jmin[1] != TLAB_DNS_BCS_NEUMANN && prm.bcs_flow_jmax[1] != TLAB_DNS_BCS_NEUMANN;
        cudaStream_t st = ctx().stream;
        const int b0 = 0;   // bcs = 0: biased, non-zero (rhs_global_incompressible_1.f90:67)
        int rc = 0;
        auto B = [&](int dir, int is, const double* sf, const double* vel, double* out, double acc_scale = 1.0) {
            if (rc) return;
            if (dir == 3) rc = burgers_z(is, sf, vel, out, sf == vel);
            else rc = run_burgers(dir, is, nx, ny, nz, b0, g[dir - 1], sf, vel, out, +1, acc_scale);
            launches++;
        };
        double *u = q[0], *v = q[1], *w = q[2];
        if (P > 1 && nzg > 1 && !split) {     // transposed w is shared by the four z operators (tmp6 of the reference)
            if ((rc = trp().forward(w, nullptr, 0.0, zw, (long long)nx * ny, nz))) return rc;
        }
        // hq_i += Bx(q_i,u) + By(q_i,v) + Bz(q_i,w), hs += ... (:98-162).  Grouped by direction, so that the fields that share
        // the advecting velocity go into one fused launch (sums reordered within round-off: x, y, z for every field).
        const int nfields = std::min(3 + ns, 4);
        const double* sf[4] = {u, v, w, ns > 0 ? s[0] : nullptr};
        double* outs[4] = {hq[0], hq[1], hq[2], ns > 0 ? hs[0] : nullptr};
        const int isv[4] = {0, 0, 0, 1};
        const bool fuse_z = (P == 1) && !split;
        if (ctx().tune_fuse) flush_pending();
        for (int dir = 1; dir <= 3 && !rc; dir++) {
            if (dir == 3 && !fuse_z) break;
            if (dir == 1 && !ctx().tune_fuse) {
                // the first accumulation into every hq / hs of this substep: it also applies the pending `h = h*kco`
                for (int f = 0; f < nfields; f++) B(1, isv[f], sf[f], u, outs[f], take_pending(f));
                continue;
            }
            rc = run_burgers_multi(dir, nfields, isv, sf, q[dir - 1], outs, nx, ny, nz, g[dir - 1], &launches);
        }
        if (!fuse_z) { B(3, 0, u, w, hq[0]); B(3, 0, v, w, hq[1]); B(3, 0, w, w, hq[2]); if (ns > 0) B(3, 1, s[0], w, hs[0]); }
        for (int is = 1; is < ns; is++) {              // further scalars: one launch each
            B(1, is + 1, s[is], u, hs[is], take_pending(3 + is)); B(2, is + 1, s[is], v, hs[is]); B(3, is + 1, s[is], w, hs[is]);
        }
        if (rc) return rc;
        // pressure forcing: div(hq + q/dte), accumulated in the order y, x, z (:177-260)
        const double dummy = 1.0 / dte;
        if ((rc = run_partial(2, TLAB_OPR_P1, nx, ny, nz, b0, g[1], hq[1], tmp1, nullptr, v, dummy, 0))) return rc;
        if ((rc = run_partial(1, TLAB_OPR_P1, nx, ny, nz, b0, g[0], hq[0], tmp1, nullptr, u, dummy, +1))) return rc;
        if ((rc = partial_z(hq[2], w, dummy, tmp1, +1))) return rc;
        launches += 3;
        // Neumann data for the pressure: hq2 at the walls (:272-281)
        const long long np = (long long)nx * nz;
        { ProfScope ps(PC_ELEMENTWISE);
        get_planes_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(hq[1], hb, ht, nx, ny, nz); }
        launches++;
        if ((rc = pois->solve(tmp1, c1, c2, hb, ht, tmp3))) return rc;     // (:284)
        launches += 2;   // boundary planes, per-mode y solves (cuFFT's own kernels not counted)
        // hq -= grad p (:319-352)
        if ((rc = run_partial(1, TLAB_OPR_P1, nx, ny, nz, b0, g[0], tmp1, hq[0], nullptr, nullptr, 0.0, -1))) return rc;
        if ((rc = partial_z(tmp1, nullptr, 0.0, hq[2], -1))) return rc;
        if (!v_deferred) {
            ProfScope ps(PC_ELEMENTWISE);
            sub_kernel<<<ew_blocks(N), EW_THREADS, 0, st>>>(hq[1], tmp3, N);
        }
        launches += 3;
        return rhs_bcs();
    }

    // boundary conditions (:356-398)
    int rhs_bcs() {
        cudaStream_t st = ctx().stream;
        const long long np = (long long)nx * nz;
        int rc = 0;
        for (int f = 0; f < 3 + ns; f++) {
            if (f == 1 && v_deferred) continue;
            double* h = (f < 3) ? hq[f] : hs[f - 3];
            const int tmin = (f < 3) ? prm.bcs_flow_jmin[f] : prm.bcs_scal_jmin[f - 3];
            const int tmax = (f < 3) ? prm.bcs_flow_jmax[f] : prm.bcs_scal_jmax[f - 3];
            int ibc = 0;
            if (tmin == TLAB_DNS_BCS_NEUMANN) ibc += 1;
            if (tmax == TLAB_DNS_BCS_NEUMANN) ibc += 2;
            if (ibc > 0) {
                if ((rc = run_neumann_y(ibc, nx, ny, nz, g[1], h, hb, ht))) return rc;
                launches++;
            }
            { ProfScope ps(PC_ELEMENTWISE);
            set_planes_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(h, hb, ht, ibc & 1, (ibc >> 1) & 1, nx, ny, nz); }
            launches++;
        }
        return cuda_check(cudaGetLastError(), "rhs");
    }

    // TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT + DNS_BOUNDS_LIMIT + hq *= kco
    int substep(double dte, double kcoef, int scale_h) {
        int rc = sources_flow();
        if (rc) return rc;
        defer_v = ctx().tune_fuse_update != 0;
        rc = rhs(dte);
        defer_v = false;
        if (rc) { v_deferred = false; return rc; }
        cudaStream_t st = ctx().stream;
        ProfScope ps(PC_ELEMENTWISE);
        // `h = h*kco` (time.f90:290-297) waits for the first accumulation of the next substep instead of being written now
        const bool lazy = scale_h && ctx().tune_lazy_scale;
        const int sh = lazy ? 0 : scale_h;
        for (int f = 0; f < 3; f++) {
            if (f == 1 && v_deferred)
                rk_update_sub_kernel<<<ew_blocks(N), EW_THREADS, 0, st>>>(q[1], hq[1], tmp3, dte, kcoef, sh, nx, ny, N);
            else
                rk_update_kernel<<<ew_blocks(N), EW_THREADS, 0, st>>>(q[f], hq[f], dte, kcoef, sh, 0, 0.0, 0.0, N);
            if (lazy) pending[f] = kcoef;
        }
        v_deferred = false;
        for (int is = 0; is < ns; is++) {
            rk_update_kernel<<<ew_blocks(N), EW_THREADS, 0, st>>>(s[is], hs[is], dte, kcoef, sh, prm.scal_limit,
                                                                  prm.scal_min[is], prm.scal_max[is], N);
            if (lazy) pending[3 + is] = kcoef;
        }
        launches += 3 + ns;
        return cuda_check(cudaGetLastError(), "rk update");
    }

    int stage(double dtime, int sub) {
        cudaStream_t st = ctx().stream;
        const int nsub = (int)kdt.size();
        if (sub < 0 || sub >= nsub) return fail(TLAB_ERR_OPTION, "Runge-Kutta stage out of range");
        if (sub == 0) {
            ProfScope ps(PC_ELEMENTWISE);
            for (double* h : hq) cudaMemsetAsync(h, 0, (size_t)N * sizeof(double), st);
            for (double* h : hs) cudaMemsetAsync(h, 0, (size_t)N * sizeof(double), st);
            for (double& k : pending) k = 1.0;
        }
        const double dte = dtime * kdt[sub];
        const bool last = (sub == nsub - 1);
        return substep(dte, last ? 0.0 : kco[sub], last ? 0 : 1);
    }

    int runge_kutta(double dtime) {
        const int nsub = (int)kdt.size();
        for (int sub = 0; sub < nsub; sub++)
            if (int rc = stage(dtime, sub)) return rc;
        return 0;
    }

    void release() {
        for (double* b : {zs, zw, zr, c2}) if (b) trp().unregister_buffer(b);
        for (void* a : allocs) cudaFree(a);
        allocs.clear();
        if (host_stage) cudaFreeHost(host_stage);
        host_stage = nullptr;
        if (pois) { pois->release(); delete pois; pois = nullptr; }
    }
};

}  // namespace tlab

struct tlab_dns_s {
    tlab::Dns d;
};

using namespace tlab;

static double* field_ptr(Dns& d, const char* name) {
    const std::string w(name ? name : "");
    auto idx = [&](size_t pre) { return w.size() > pre ? atoi(w.c_str() + pre) - 1 : -1; };
    if (w.rfind("hq", 0) == 0) { int i = idx(2); return (i >= 0 && i < 3) ? d.hq[i] : nullptr; }
    if (w.rfind("hs", 0) == 0) { int i = idx(2); return (i >= 0 && i < d.ns) ? d.hs[i] : nullptr; }
    if (w.rfind("q", 0) == 0) { int i = idx(1); return (i >= 0 && i < 3) ? d.q[i] : nullptr; }
    if (w.rfind("s", 0) == 0) { int i = idx(1); return (i >= 0 && i < d.ns) ? d.s[i] : nullptr; }
    if (w == "p") return d.tmp1;
    if (w == "dpdy") return d.tmp3;
    return nullptr;
}

extern "C" {

int tlab_dns_create(const tlab_dns_params* prm, tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz,
                    const double* bbackground_host, tlab_dns_t* out) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    if (!prm || !gx || !gy || !gz || !out) return fail(TLAB_ERR_OPTION, "tlab_dns_create: null argument");
    if (prm->nscal < 0 || prm->nscal > TLAB_MAX_SCAL) return fail(TLAB_ERR_OPTION, "tlab_dns_create: too many scalars");
    // Gravity_Buoyancy, EQNS_BOD_LINEAR (gravity.f90:248-260): this path implements the branch locProps%scalar(1) == 1,
    // b = c1 s1 - (ref - c0), for any number of prognostic scalars.  buoyancy_params = (c1, c0): the caller passes
    // c0 = parameters(inb_scal_array + 1) itself and must not select this path when scalar(1) > 1 (the Fortran shim checks).
    if (prm->rkm_mode != TLAB_RKM_EXP3 && prm->rkm_mode != TLAB_RKM_EXP4)
        return fail(TLAB_ERR_UNDEVELOP, "only the explicit RK3 / RK4(5) schemes are implemented");
    const int P = trp().P;
    if (gx->p.n != prm->nx || gy->p.n != prm->ny || gz->p.n != prm->nz * P)
        return fail(TLAB_ERR_DIMGRID, "tlab_dns_create: plan sizes differ from nx, ny, nz (nz = local slab thickness)");
    if (P > 1 && ((long long)prm->nx * prm->ny) % P)
        return fail(TLAB_ERR_PARPARTITION, "tlab_dns_create: nx*ny is not a multiple of the number of ranks");
    tlab_dns_s* h = new tlab_dns_s();
    Dns& d = h->d;
    d.prm = *prm;
    d.g[0] = gx; d.g[1] = gy; d.g[2] = gz;
    d.nx = prm->nx; d.ny = prm->ny; d.nz = prm->nz; d.ns = prm->nscal;
    d.P = P; d.nzg = gz->p.n;
    d.N = (long long)d.nx * d.ny * d.nz;
    d.Nt = (long long)(d.nx + 2) * d.ny * d.nz;
    rk_tables(prm->rkm_mode, d.kdt, d.ktime, d.kco);
    int rc = tlab_opr_burgers_init(gx, gy, gz, prm->visc, prm->nscal, prm->schmidt);
    for (int i = 0; i < 3; i++) d.bfirst[i] = d.g[i]->burgers_first;
    if (!rc) { d.pois = new Poisson(); rc = d.pois->init(gx, gy, gz, prm->nz); }
    d.q.resize(3); d.hq.resize(3); d.s.resize(d.ns); d.hs.resize(d.ns);
    for (int i = 0; i < 3 && !rc; i++) { rc = d.alloc(&d.q[i], d.N); if (!rc) rc = d.alloc(&d.hq[i], d.N); }
    for (int i = 0; i < d.ns && !rc; i++) { rc = d.alloc(&d.s[i], d.N); if (!rc) rc = d.alloc(&d.hs[i], d.N); }
    if (!rc) rc = d.alloc(&d.tmp1, d.Nt);
    if (!rc) rc = d.alloc(&d.tmp3, d.Nt);
    if (!rc) rc = d.alloc(&d.c1, d.Nt);
    if (!rc) rc = d.alloc(&d.c2, d.Nt);
    if (!rc) rc = d.alloc(&d.hb, (long long)d.nx * d.nz);
    if (!rc) rc = d.alloc(&d.ht, (long long)d.nx * d.nz);
    if (!rc) rc = d.alloc(&d.bbackground, d.ny);
    if (P > 1) {
        if (!rc) rc = d.alloc(&d.zs, d.N);
        if (!rc) rc = d.alloc(&d.zw, d.N);
        if (!rc) rc = d.alloc(&d.zr, d.N);
        // peer-memory transposes: publish the pencil buffers to the other ranks (collective, same order everywhere)
        if (!rc) rc = cuda_check(cudaStreamSynchronize(ctx().stream), "tlab_dns_create");
        for (double* b : {d.zs, d.zw, d.zr, d.c2}) if (!rc) rc = trp().register_buffer(b);
        // split-z operators: exchange blocks for halo planes and chunk ends (collective; not eligible -> transposes stay)
        if (!rc && ctx().tune_splitz && d.nzg > 1)
            rc = splitz().init((long long)d.nx * d.ny, d.nz, d.nzg, P, trp().rank, 0);
    }
    if (!rc && bbackground_host)
        rc = cuda_check(cudaMemcpyAsync(d.bbackground, bbackground_host, d.ny * sizeof(double), cudaMemcpyHostToDevice, ctx().stream), "bbackground");
    if (!rc) rc = cuda_check(cudaStreamSynchronize(ctx().stream), "tlab_dns_create");
    if (rc) { d.release(); delete h; return rc; }
    *out = h;
    return 0;
}

int tlab_dns_destroy(tlab_dns_t h) {
    if (!h) return 0;
    cudaStreamSynchronize(ctx().stream);
    h->d.release();
    delete h;
    return 0;
}

int tlab_dns_field(tlab_dns_t h, const char* name, double** dev_ptr) {
    if (!h || !dev_ptr) return fail(TLAB_ERR_OPTION, "tlab_dns_field: null argument");
    h->d.flush_pending();       // the caller may read or write hq / hs through the pointer
    double* p = field_ptr(h->d, name);
    if (!p) return fail(TLAB_ERR_OPTION, std::string("tlab_dns_field: unknown field ") + (name ? name : "(null)"));
    *dev_ptr = p;
    return 0;
}

int tlab_dns_upload_host(tlab_dns_t h, const char* name, const double* src_host) {
    if (!h || !src_host) return fail(TLAB_ERR_OPTION, "tlab_dns_upload_host: null argument");
    h->d.flush_pending();
    double* p = field_ptr(h->d, name);
    if (!p) return fail(TLAB_ERR_OPTION, "tlab_dns_upload_host: unknown field");
    return tlab_gpu_upload(p, src_host, (size_t)h->d.N * sizeof(double));
}

int tlab_dns_download_host(tlab_dns_t h, const char* name, double* dst_host) {
    if (!h || !dst_host) return fail(TLAB_ERR_OPTION, "tlab_dns_download_host: null argument");
    h->d.flush_pending();
    double* p = field_ptr(h->d, name);
    if (!p) return fail(TLAB_ERR_OPTION, "tlab_dns_download_host: unknown field");
    return tlab_gpu_download(dst_host, p, (size_t)h->d.N * sizeof(double));
}

int tlab_time_substep(tlab_dns_t h, double dte, double kco, int scale_h) {
    if (!h) return fail(TLAB_ERR_OPTION, "tlab_time_substep: null state");
    if (int rc = h->d.substep(dte, kco, scale_h)) return rc;
    return finish();
}

int tlab_rhs_global_incompressible_1(tlab_dns_t h, double dte) {
    if (!h) return fail(TLAB_ERR_OPTION, "null state");
    if (int rc = h->d.rhs(dte)) return rc;
    return finish();
}

int tlab_time_rungekutta_stage(tlab_dns_t h, double dtime, int stage) {
    if (!h) return fail(TLAB_ERR_OPTION, "tlab_time_rungekutta_stage: null state");
    if (int rc = h->d.stage(dtime, stage)) return rc;
    return finish();
}

int tlab_time_rungekutta(tlab_dns_t h, double dtime) {
    if (!h) return fail(TLAB_ERR_OPTION, "tlab_time_rungekutta: null state");
    if (int rc = h->d.runge_kutta(dtime)) return rc;
    return finish();
}

int tlab_time_rungekutta_host(tlab_dns_t h, double dtime, double* q_host, double* s_host) {
    if (!h || !q_host) return fail(TLAB_ERR_OPTION, "tlab_time_rungekutta_host: null argument");
    if (h->d.ns > 0 && !s_host) return fail(TLAB_ERR_OPTION, "tlab_time_rungekutta_host: s_host is null but the state has scalars");
    Dns& d = h->d;
    cudaStream_t st = ctx().stream;
    const size_t fb = (size_t)d.N * sizeof(double);
    for (int i = 0; i < 3; i++)
        if (int rc = cuda_check(cudaMemcpyAsync(d.q[i], q_host + (size_t)i * d.N, fb, cudaMemcpyHostToDevice, st), "upload q")) return rc;
    for (int i = 0; i < d.ns; i++)
        if (int rc = cuda_check(cudaMemcpyAsync(d.s[i], s_host + (size_t)i * d.N, fb, cudaMemcpyHostToDevice, st), "upload s")) return rc;
    if (int rc = d.runge_kutta(dtime)) return rc;
    for (int i = 0; i < 3; i++)
        if (int rc = cuda_check(cudaMemcpyAsync(q_host + (size_t)i * d.N, d.q[i], fb, cudaMemcpyDeviceToHost, st), "download q")) return rc;
    for (int i = 0; i < d.ns; i++)
        if (int rc = cuda_check(cudaMemcpyAsync(s_host + (size_t)i * d.N, d.s[i], fb, cudaMemcpyDeviceToHost, st), "download s")) return rc;
    return cuda_check(cudaStreamSynchronize(st), "tlab_time_rungekutta_host");
}

int tlab_time_rk_coefficients(int rkm_mode, double* kdt, double* ktime, double* kco, int* nsub) {
    if (rkm_mode != TLAB_RKM_EXP3 && rkm_mode != TLAB_RKM_EXP4) return fail(TLAB_ERR_UNDEVELOP, "unknown RK scheme");
    std::vector<double> a, b, c;
    rk_tables(rkm_mode, a, b, c);
    if (nsub) *nsub = (int)a.size();
    for (size_t i = 0; i < a.size(); i++) { if (kdt) kdt[i] = a[i]; if (ktime) ktime[i] = b[i]; }
    for (size_t i = 0; i < c.size(); i++) if (kco) kco[i] = c[i];
    return 0;
}

int tlab_time_courant(tlab_dns_t h, double cfla, double cfld, double prandtl, double* dtime, double* cfl_number,
                      double* diffusion_number) {
    if (!h || !dtime) return fail(TLAB_ERR_OPTION, "TIME_COURANT: null argument");
    Dns& d = h->d;
    cudaStream_t st = ctx().stream;
    // grid factors (TIME_INITIALIZE): 1/jac and the maximum of sum 1/jac^2, uploaded once
    if (!d.ods[0]) {
        double dx2i_y = 0.0;
        for (int ig = 0; ig < 3; ig++) {
            const HostPlan& hp = d.g[ig]->p.h;
            std::vector<double> o(hp.size);
            for (int i = 0; i < hp.size; i++) o[i] = 1.0 / hp.jac(i + 1, 1);
            if (int rc = d.alloc(&d.ods[ig], hp.size)) return rc;
            cudaMemcpyAsync(d.ods[ig], o.data(), o.size() * sizeof(double), cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            double mx = 0.0;
            for (double v : o) mx = std::max(mx, v * v);
            if (hp.size > 1) dx2i_y += mx;   // the maximum of the sum is the sum of the maxima on a tensor-product grid
        }
        d.dx2i = dx2i_y;
    }
    {   // the diffusivity factor depends on the caller's prandtl: recomputed on every call (time.f90:143-152)
        double sf = 1.0;
        sf = std::max(sf, 1.0 / prandtl);
        double smin = 1e300;
        for (int is = 0; is < d.ns; is++) smin = std::min(smin, d.prm.schmidt[is]);
        if (d.ns > 0) sf = std::max(sf, 1.0 / smin);
        d.schmidtfactor = sf * d.prm.visc;
    }
    const unsigned blocks = 148 * 8;
    if (!d.red) { if (int rc = d.alloc(&d.red, 2 * blocks)) return rc; }
    {
        ProfScope ps(PC_ELEMENTWISE);
        courant_kernel<<<blocks, RED_THREADS, 0, st>>>(d.q[0], d.q[1], d.q[2], d.ods[0], d.ods[1], d.ods[2], d.nx, d.ny,
                                                        trp().rank * d.nz, d.nzg > 1 ? 1 : 0, d.N, d.red);
    }
    std::vector<double> part(blocks);
    if (int rc = cuda_check(cudaMemcpyAsync(part.data(), d.red, blocks * sizeof(double), cudaMemcpyDeviceToHost, st), "courant")) return rc;
    if (int rc = cuda_check(cudaStreamSynchronize(st), "courant")) return rc;
    double pmax[2] = {0.0, d.schmidtfactor * d.dx2i};
    for (double v : part) pmax[0] = std::max(pmax[0], v);
    if (int rc = allreduce_max(pmax, 2)) return rc;
    const double big = 1.0e+20;
    double dtc = big, dtd = big;
    if (pmax[0] > 0.0) dtc = cfla / pmax[0];
    if (pmax[1] > 0.0) dtd = cfld / pmax[1];
    if (cfla > 0.0) *dtime = std::min(std::min(dtc, dtd), big);      // explicit RK: min(dtc, dtd) (time.f90:526-538)
    if (cfl_number) *cfl_number = *dtime * pmax[0];
    if (diffusion_number) *diffusion_number = *dtime * pmax[1];
    return 0;
}

int tlab_dns_bounds_control(tlab_dns_t h, double* dil_min, double* dil_max) {
    if (!h || !dil_min || !dil_max) return fail(TLAB_ERR_OPTION, "DNS_BOUNDS_CONTROL: null argument");
    Dns& d = h->d;
    cudaStream_t st = ctx().stream;
    // FI_INVARIANT_P: result = -(du/dx + dv/dy + dw/dz) in tmp1
    int rc = run_partial(1, TLAB_OPR_P1, d.nx, d.ny, d.nz, 0, d.g[0], d.q[0], d.tmp1, nullptr);
    if (!rc) rc = run_partial(2, TLAB_OPR_P1, d.nx, d.ny, d.nz, 0, d.g[1], d.q[1], d.tmp1, nullptr, nullptr, 0.0, +1);
    if (!rc) rc = d.partial_z(d.q[2], nullptr, 0.0, d.tmp1, +1);
    if (rc) return rc;
    const unsigned blocks = 148 * 8;
    if (!d.red) { if ((rc = d.alloc(&d.red, 2 * blocks))) return rc; }
    {
        ProfScope ps(PC_ELEMENTWISE);
        negate_kernel<<<blocks, RED_THREADS, 0, st>>>(d.tmp1, d.N);
        minmax_kernel<<<blocks, RED_THREADS, 0, st>>>(d.tmp1, d.N, d.red, d.red + blocks);
    }
    std::vector<double> part(2 * blocks);
    if ((rc = cuda_check(cudaMemcpyAsync(part.data(), d.red, 2 * blocks * sizeof(double), cudaMemcpyDeviceToHost, st), "minmax"))) return rc;
    if ((rc = cuda_check(cudaStreamSynchronize(st), "minmax"))) return rc;
    double amn = part[0], amx = part[blocks];
    for (unsigned b = 0; b < blocks; b++) { amn = std::min(amn, part[b]); amx = std::max(amx, part[blocks + b]); }
    double v[2] = {-amn, amx};      // MIN over ranks as MAX of the negative
    if ((rc = allreduce_max(v, 2))) return rc;
    amn = -v[0]; amx = v[1];
    // MINMAX(..., d_max_loc, d_min_loc); d_min_loc = -d_min_loc; d_max_loc = -d_max_loc  (dns_local.f90:181-182)
    *dil_max = -amn;
    *dil_min = -amx;
    return 0;
}


int tlab_dns_launch_count(tlab_dns_t h, long long* count) {
    if (!h || !count) return fail(TLAB_ERR_OPTION, "tlab_dns_launch_count: null argument");
    *count = h->d.launches + trp().launches;
    return 0;
}

}  // extern "C"
