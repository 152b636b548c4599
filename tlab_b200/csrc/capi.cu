// C ABI: runtime, plans and the directional operators (see include/tlab_gpu.h).
#include "../../include/tlab_gpu.h"
#include "context.h"
#include "trp.h"
#include "splitz.h"
#include <cstring>
#include <algorithm>
#include <cstdio>
#include <vector>

namespace tlab {

Context& ctx() {
    static Context c;
    return c;
}

int fail(int code, const std::string& msg) {
    ctx().last_error = msg;
    return code;
}

int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(TLAB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int finish() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_check(e, "kernel launch");
    if (!ctx().async) return cuda_check(cudaStreamSynchronize(ctx().stream), "stream synchronize");
    return 0;
}

ProfScope::ProfScope(int cls) {
    Context& c = ctx();
    if (!c.profiling) return;
    Context::ProfRec r;
    r.cls = cls;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!c.event_pool.empty()) { *e = c.event_pool.back(); c.event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.a, c.stream);
    c.prof.push_back(r);
    idx = (int)c.prof.size() - 1;
}

ProfScope::~ProfScope() {
    if (idx < 0) return;
    Context& c = ctx();
    cudaEventRecord(c.prof[idx].b, c.stream);
}

static int need_ready() {
    if (!ctx().ready) {
        int rc = tlab_gpu_init(-1);
        if (rc) return rc;
    }
    return 0;
}

static void fill_common(LineArgs& a, const DevPlan& p) {
    a.n = p.n; a.T = p.T; a.cbase = p.cbase; a.crem = p.crem;
    a.rhs_d1 = p.rhs_d1;
    a.rhs2_rows = p.rhs2_rows;
    a.rhs2 = p.rhs2;
}

static void set_geometry(LineArgs& a, int dir, int nx, int ny, int nz, bool& contig) {
    contig = false;
    if (dir == 1) {
        contig = true;
        a.nlines = (long long)ny * nz; a.stride = 1; a.inner = 1; a.outer_stride = nx;
    } else if (dir == 2) {
        a.nlines = (long long)nx * nz; a.stride = nx; a.inner = nx; a.outer_stride = (long long)nx * ny;
    } else {
        a.nlines = (long long)nx * ny; a.stride = (long long)nx * ny; a.inner = a.nlines; a.outer_stride = 0;
    }
    a.L = pick_lines_per_cta(a.T, contig, contig ? ctx().tune_lines_x : ctx().tune_lines_yz);
    while (a.L > 1 && a.L * a.T > 512) a.L >>= 1;
    a.xstride = contig ? xtile_stride(a.n, a.L) : 0;
}


// fast path (lines2.cu): full chunks, aligned tiles, short look-back windows; otherwise the general kernels of lines.cu
static bool build_line2(int mode, const LineArgs& a, const DevPlan& p, const Sys2& s1, const Sys2& s2, bool contig, Line2Args& b) {
    int L = 0;
    bool fast = ctx().tune_fast != 0 && lines2_eligible(p, s1, &s2, a.n, a.nlines, a.inner, contig,
                                                         contig ? ctx().tune_lines_x : ctx().tune_lines_yz, &L);
    if (fast && contig) {
        auto al = [](const void* q) { return (reinterpret_cast<size_t>(q) & 15) == 0; };
        fast = al(a.u) && al(a.u2) && al(a.vel) && al(a.out1) && al(a.out2);
    }
    if (fast && p.need_1der && !p.cjac2) fast = false;
    // per-row rhs of the CompactDirect6 second derivative: general kernels (the first derivative alone keeps the fast path)
    if (fast && p.rhs2_rows && (mode == MODE_P2 || mode == MODE_P2_P1 || mode == MODE_BURGERS)) fast = false;
    if (!fast) return false;
    b.n = a.n; b.T = a.n / CHUNK; b.L = L;
    b.xls = contig ? lines2_xstride(b.T, L) : 0;
    if (!contig && ctx().tune_persist && mode != MODE_NEUMANN && !(a.u2 && mode == MODE_BURGERS)) {
        auto al = [](const void* q) { return (reinterpret_cast<size_t>(q) & 15) == 0; };
        b.persist = al(a.u) && al(a.u2) && al(a.vel) && (a.stride % 2 == 0) && (a.outer_stride % 2 == 0) && (L % 2 == 0) &&
                    lines2_persist_smem(b.T, L) <= 220 * 1024;
    }
    b.pf_l1 = contig ? 0 : ctx().tune_pf_l1;
    b.lshift = 0;
    while ((1 << b.lshift) < L) b.lshift++;
    b.accumulate = a.accumulate; b.scale = a.scale;
    // L2 prefetch of a later tile: pays for x lines and for y lines (rows a few KB apart); with rows MBs apart (z) it
    // costs bandwidth (measured, profiles/ops_c2_r01_v9.json), so there it is on request only
    b.pf_dist = (contig || a.stride <= 4096) ? ctx().tune_pf_dist : std::max(ctx().tune_pf_dist, 0);
    b.stride = a.stride; b.inner = a.inner; b.outer_stride = a.outer_stride;
    b.u = a.u; b.u2 = a.u2; b.vel = a.vel; b.out1 = a.out1; b.out2 = a.out2; b.bcs_hb = a.bcs_hb; b.bcs_ht = a.bcs_ht;
    b.cjac = p.cjac2;
    b.rhs1 = a.rhs1; b.rhs2 = a.rhs2; b.s1 = s1; b.s2 = s2;
    if (!ctx().tune_circ) b.s1.circ = b.s2.circ = 0;
    std::memcpy(b.neu_bot, a.neu_bot, sizeof(b.neu_bot));
    std::memcpy(b.neu_top, a.neu_top, sizeof(b.neu_top));
    b.neu_lu_bot = a.neu_lu_bot; b.neu_lu_top = a.neu_lu_top;
    if (!contig && ctx().tune_pair && p.periodic && !p.need_1der && (mode == MODE_P1 || mode == MODE_P2 || mode == MODE_BURGERS) &&
        b.T > 32 && b.T % 2 == 0 && 16 * (b.T / 2) <= 512 && a.inner % 16 == 0 && !(a.u2 && mode == MODE_BURGERS)) {
        // long periodic lines: a cluster of 2 CTAs per tile of 16 lines (128-byte rows), see lines2_strided_pair
        b.L = 16; b.lshift = 4; b.pair = 1; b.persist = 0;
    }
    if (mode == MODE_NEUMANN && ctx().tune_neu_compact && !contig) {
        // wall chunks only: more lines per CTA (long row segments), no idle threads
        const int nb = a.bcs_hb ? LB2 : 0, nt = a.bcs_ht ? LB2 : 0;
        if (nb + nt > 0 && nb + nt < b.T) {
            int Lc = 64;
            while (Lc > 1 && (Lc * (nb + nt) > 512 || a.inner % Lc != 0)) Lc >>= 1;
            // the exchange area is sized for the whole line (absent chunks read as zero): long lines take fewer lines per CTA
            while (Lc > 8 && (size_t)8 * b.T * Lc * sizeof(double) > 100 * 1024) Lc >>= 1;
            if ((size_t)8 * b.T * Lc * sizeof(double) <= 200 * 1024) {
                b.neu_nb = nb; b.neu_nt = nt; b.L = Lc;
                b.lshift = 0;
                while ((1 << b.lshift) < Lc) b.lshift++;
            }
        }
    }
    b.tma_rb = 16;
    while (b.tma_rb < 256 && b.n % (b.tma_rb * 2) == 0) b.tma_rb *= 2;
    b.tma_l2 = ctx().tune_tma_l2;
    b.tma = (!contig && !b.pair && ctx().tune_tma && lines2_tma_eligible(mode, b)) ? 1 : 0;
    b.march_red = ctx().tune_march_red;
    b.march_cfg = ctx().tune_march_cfg;
    b.march_pf = ctx().tune_march_pf;
    b.march_peel = ctx().tune_march_peel ? 0 : -1;      // -1: never peel (tuning key march_peel = 0)
    // marching panels: periodic directions by default (z: 21.8 -> 14.6 ms for the four Burgers launches at C3, 8.4 -> 5.2 ms for the two
    // derivatives); non-periodic directions (all chunks read coefficient tables) only on request (march = 2) or when the lines
    // are so long that the whole-line kernels are left with 64-byte rows (more than 32 chunks)
    // ... or, for OPR_Burgers, when only the rounds at the walls read tables (unscaled constant chunks, plan.cu): the marching kernel
    // then runs constant-only steps in between (y at C3: 15.3 -> 13.9 ms for the four launches)
    const bool march_wanted = ctx().tune_march >= 2 ||
                              (ctx().tune_march == 1 && (p.periodic || b.T > 32 || (mode == MODE_BURGERS && march_peelable(mode, b, p.periodic))));
    b.march = (!contig && march_wanted && !b.tma && !b.persist && !b.pair &&
               march_eligible(mode, b, p.periodic, p.need_1der, a.nlines, a.inner)) ? 1 : 0;
    return true;
}

__global__ void scale_array_kernel(double* __restrict__ a, double k, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = k * a[i];
}

void scale_array(double* a, double k, long long n, cudaStream_t st) {
    const long long blocks = std::min<long long>((n + 255) / 256, 148LL * 16);
    scale_array_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, k, n);
}

static cudaError_t launch_any(int mode, const LineArgs& a_in, const DevPlan& p, const Sys2& s1, const Sys2& s2, bool periodic,
                              bool need1, bool contig, cudaStream_t st) {
    LineArgs a = a_in;
    Line2Args b;
    const bool fast = build_line2(mode, a, p, s1, s2, contig, b);
    // a pending factor of the accumulation target is fused by the contiguous fast kernel only; otherwise it is applied first
    const bool fuse_scale = fast && contig && !b.tma && a.accumulate != 0 && (mode == MODE_BURGERS || mode == MODE_P1);
    if (a.acc_scale != 1.0 && a.accumulate != 0 && !fuse_scale) {
        long long n = 0;
        if (contig) n = a.nlines * a.n;
        else n = (a.nlines / a.inner - 1) * a.outer_stride + (long long)(a.n - 1) * a.stride + a.inner;   // last element + 1
        scale_array(a.out1, a.acc_scale, n, st);
        a.acc_scale = 1.0;
    }
    b.acc_scale = fuse_scale ? a.acc_scale : 1.0;
    if (!fast) { ctx().general_launches++; return launch_lines(mode, a, periodic, need1, contig, st); }
    ctx().fast_launches++;
    if (b.march) { ctx().march_launches++; return launch_march(mode, b, periodic, need1, a.nlines, a.inner, st); }
    return launch_lines2(mode, b, periodic, need1, contig, a.nlines, a.inner, st);
}

static int check_dims(int dir, int nx, int ny, int nz, const tlab_plan_s* g) {
    const int n = (dir == 1) ? nx : (dir == 2 ? ny : nz);
    if (!g) return fail(TLAB_ERR_OPTION, "null plan");
    if (dir < 1 || dir > 3) return fail(TLAB_ERR_OPTION, "direction must be 1, 2 or 3");
    if (g->p.n != n) return fail(TLAB_ERR_DIMGRID, "plan size does not match the field extent along this direction");
    if (nx < 1 || ny < 1 || nz < 1) return fail(TLAB_ERR_DIMGRID, "non-positive extent");
    if (n > 1 && g->p.T > 512) return fail(TLAB_ERR_DIMGRID, "line longer than 8192 points");
    return 0;
}

int run_partial(int dir, int type, int nx, int ny, int nz, int ibc, tlab_plan_s* g, const double* u, double* result,
                double* tmp1, const double* u2, double scale, int accumulate) {
    int rc = check_dims(dir, nx, ny, nz, g);
    if (rc) return rc;
    const size_t bytes = (size_t)nx * ny * nz * sizeof(double);
    cudaStream_t st = ctx().stream;
    if (type != TLAB_OPR_P1 && type != TLAB_OPR_P2 && type != TLAB_OPR_P2_P1)
        return fail(TLAB_ERR_UNDEVELOP, "OPR_Partial: only OPR_P1, OPR_P2, OPR_P2_P1 are implemented");
    if (type == TLAB_OPR_P2_P1 && !tmp1) return fail(TLAB_ERR_OPTION, "OPR_P2_P1 needs tmp1");
    if (g->p.n == 1) {       // 2-D case: derivative set to zero (opr_partial.f90:174-177, 287-289)
        if (accumulate == 0) cudaMemsetAsync(result, 0, bytes, st);
        if (type == TLAB_OPR_P2_P1) cudaMemsetAsync(tmp1, 0, bytes, st);
        return 0;
    }
    if (ibc < 0 || ibc > 3) return fail(TLAB_ERR_OPTION, "bcs codes must be 0 or 1");
    const DevPlan& p = g->p;
    LineArgs a;
    fill_common(a, p);
    bool contig;
    set_geometry(a, dir, nx, ny, nz, contig);
    a.u = u; a.out1 = result; a.out2 = tmp1;
    a.u2 = u2; a.scale = scale; a.accumulate = accumulate;
    a.rhs1 = p.rhs1[ibc];
    a.lu1 = p.lu1[ibc];
    a.lu2 = p.lu2[0];
    const int mode = (type == TLAB_OPR_P1) ? MODE_P1 : (type == TLAB_OPR_P2 ? MODE_P2 : MODE_P2_P1);
    ProfScope ps(PC_PARTIAL_X + dir - 1);
    return cuda_check(launch_any(mode, a, p, p.sys1[ibc], p.sys2[0], p.periodic, p.need_1der, contig, st), "line kernel");
}

int run_burgers(int dir, int is, int nx, int ny, int nz, int ibc, tlab_plan_s* g, const double* s, const double* vel,
                double* result, int accumulate, double acc_scale) {
    int rc = check_dims(dir, nx, ny, nz, g);
    if (rc) return rc;
    cudaStream_t st = ctx().stream;
    if (g->p.n == 1) {       // opr_burgers.f90:207-210
        if (!accumulate) cudaMemsetAsync(result, 0, (size_t)nx * ny * nz * sizeof(double), st);
        else if (acc_scale != 1.0) scale_array(result, acc_scale, (long long)nx * ny * nz, st);
        return 0;
    }
    if (g->burgers_first < 0) return fail(TLAB_ERR_OPTION, "tlab_opr_burgers_init has not been called for this plan");
    if (is < 0 || is >= g->burgers_count) return fail(TLAB_ERR_OPTION, "scalar index out of range");
    if (ibc < 0 || ibc > 3) return fail(TLAB_ERR_OPTION, "bcs codes must be 0 or 1");
    const DevPlan& p = g->p;
    LineArgs a;
    fill_common(a, p);
    bool contig;
    set_geometry(a, dir, nx, ny, nz, contig);
    a.u = s; a.vel = vel; a.out1 = result; a.accumulate = accumulate; a.acc_scale = acc_scale;
    a.rhs1 = p.rhs1[ibc];
    a.lu1 = p.lu1[ibc];
    a.lu2 = p.lu2[g->burgers_first + is];
    ProfScope ps(PC_BURGERS_X + dir - 1);
    return cuda_check(launch_any(MODE_BURGERS, a, p, p.sys1[ibc], p.sys2[g->burgers_first + is], p.periodic, p.need_1der, contig, st), "burgers kernel");
}

// OPR_Burgers along one direction for several fields advected by the same velocity, accumulated into out[f]
// (the twelve calls of RHS_GLOBAL_INCOMPRESSIBLE_1, rhs_global_incompressible_1.f90:98-162, grouped by direction):
// one fused launch when the fast kernels apply, one launch per field otherwise
int run_burgers_multi(int dir, int nf, const int* is, const double* const* sf, const double* vel, double* const* out,
                      int nx, int ny, int nz, tlab_plan_s* g, long long* launches) {
    int rc = check_dims(dir, nx, ny, nz, g);
    if (rc) return rc;
    auto fallback = [&]() {
        for (int f = 0; f < nf; f++) {
            if (int r = run_burgers(dir, is[f], nx, ny, nz, 0, g, sf[f], vel, out[f], +1)) return r;
            if (launches) (*launches)++;
        }
        return 0;
    };
    if (g->p.n == 1 || nf < 2 || nf > 4 || g->burgers_first < 0 || !ctx().tune_fuse) return fallback();
    int is_b = -1;
    for (int f = 0; f < nf; f++) {
        if (is[f] < 0 || is[f] >= g->burgers_count) return fail(TLAB_ERR_OPTION, "scalar index out of range");
        if (is[f] != is[0]) { if (is_b < 0) is_b = is[f]; else if (is[f] != is_b) return fallback(); }
    }
    const DevPlan& p = g->p;
    LineArgs a;
    fill_common(a, p);
    bool contig;
    set_geometry(a, dir, nx, ny, nz, contig);
    a.u = sf[0]; a.vel = vel; a.out1 = out[0]; a.accumulate = +1;
    a.rhs1 = p.rhs1[0];
    a.lu1 = p.lu1[0];
    a.lu2 = p.lu2[g->burgers_first + is[0]];
    Line2Args b;
    if (!build_line2(MODE_BURGERS, a, p, p.sys1[0], p.sys2[g->burgers_first + is[0]], contig, b)) return fallback();
    if (is_b >= 0 && !p.sys2[g->burgers_first + is_b].ok) return fallback();
    auto al = [](const void* q) { return (reinterpret_cast<size_t>(q) & 15) == 0; };
    for (int f = 0; f < nf; f++) if (contig && !(al(sf[f]) && al(out[f]))) return fallback();
    b.persist = 0;
    b.tma = 0;
    b.nf = nf;
    b.pf_next = ctx().tune_pf_next;
    for (int f = 0; f < nf; f++) { b.fu[f] = sf[f]; b.fo[f] = out[f]; b.fsys[f] = (is[f] == is[0]) ? 0 : 1; }
    if (is_b >= 0) b.s2b = p.sys2[g->burgers_first + is_b];
    ctx().fast_launches++;
    if (launches) (*launches)++;
    ProfScope ps(PC_BURGERS_X + dir - 1);
    return cuda_check(launch_lines2(MODE_BURGERS, b, p.periodic, p.need_1der, contig, a.nlines, a.inner, ctx().stream), "fused burgers kernel");
}

int run_neumann_y(int ibc, int nx, int ny, int nz, tlab_plan_s* g, const double* u, double* hb, double* ht) {
    int rc = check_dims(2, nx, ny, nz, g);
    if (rc) return rc;
    cudaStream_t st = ctx().stream;
    const size_t pb = (size_t)nx * nz * sizeof(double);
    if (g->p.n == 1) {       // boundary_bcs.f90:390-391
        cudaMemsetAsync(hb, 0, pb, st);
        cudaMemsetAsync(ht, 0, pb, st);
        return 0;
    }
    if (g->p.periodic) return fail(TLAB_ERR_OPTION, "Neumann boundary values need a non-periodic direction");
    if (ibc < 1 || ibc > 3) return fail(TLAB_ERR_OPTION, "ibc must be 1, 2 or 3");
    const DevPlan& p = g->p;
    LineArgs a;
    fill_common(a, p);
    bool contig;
    set_geometry(a, 2, nx, ny, nz, contig);
    a.u = u;
    a.rhs1 = p.rhs1[ibc];
    a.lu1 = p.lu1[ibc];
    a.lu2 = p.lu2[0];
    a.bcs_hb = (ibc == BCS_ND || ibc == BCS_NN) ? hb : nullptr;
    a.bcs_ht = (ibc == BCS_DN || ibc == BCS_NN) ? ht : nullptr;
    std::memcpy(a.neu_bot, p.neu_bot[ibc], sizeof(a.neu_bot));
    std::memcpy(a.neu_top, p.neu_top[ibc], sizeof(a.neu_top));
    a.neu_lu_bot = p.neu_lu_bot[ibc];
    a.neu_lu_top = p.neu_lu_top[ibc];
    ProfScope ps(PC_NEUMANN);
    return cuda_check(launch_any(MODE_NEUMANN, a, p, p.sys1[ibc], p.sys2[0], false, false, false, st), "neumann kernel");
}

}  // namespace tlab

using namespace tlab;

extern "C" {

int tlab_gpu_init(int device) {
    Context& c = ctx();
    if (c.ready) return 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TLAB_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (int rc = cuda_check(cudaSetDevice(device), "cudaSetDevice")) return rc;
    c.device = device;
    if (int rc = cuda_check(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return rc;
    c.ready = true;
    return 0;
}

int tlab_gpu_finalize(void) {
    Context& c = ctx();
    if (!c.ready) return 0;
    cudaStreamSynchronize(c.stream);
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false;
    return 0;
}

const char* tlab_gpu_last_error(void) { return ctx().last_error.c_str(); }

int tlab_gpu_set_async(int on) { ctx().async = (on != 0); return 0; }

int tlab_gpu_synchronize(void) {
    if (int rc = need_ready()) return rc;
    return cuda_check(cudaStreamSynchronize(ctx().stream), "synchronize");
}

int tlab_gpu_malloc(void** ptr, size_t bytes) {
    if (int rc = need_ready()) return rc;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TLAB_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    return 0;
}

int tlab_gpu_free(void* ptr) { return cuda_check(cudaFree(ptr), "cudaFree"); }

int tlab_gpu_malloc_managed(void** ptr, size_t bytes) {
    if (int rc = need_ready()) return rc;
    if (!ptr) return fail(TLAB_ERR_OPTION, "null argument");
    cudaError_t e = cudaMallocManaged(ptr, bytes, cudaMemAttachGlobal);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TLAB_ERR_ALLOC, std::string("cudaMallocManaged: ") + cudaGetErrorString(e)); }
    cudaMemAdvise(*ptr, bytes, cudaMemAdviseSetPreferredLocation, ctx().device);
    cudaGetLastError();          // advice is best effort
    return 0;
}

int tlab_gpu_prefetch(const void* ptr, size_t bytes, int to_device) {
    if (int rc = need_ready()) return rc;
    if (int rc = cuda_check(cudaMemPrefetchAsync(ptr, bytes, to_device ? ctx().device : cudaCpuDeviceId, ctx().stream), "prefetch")) return rc;
    return finish();
}

int tlab_gpu_upload(void* dst, const void* src, size_t bytes) {
    if (int rc = need_ready()) return rc;
    if (int rc = cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().stream), "upload")) return rc;
    return cuda_check(cudaStreamSynchronize(ctx().stream), "upload");
}

int tlab_gpu_copy(void* dst, const void* src, size_t bytes) {
    if (int rc = need_ready()) return rc;
    if (int rc = cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx().stream), "copy")) return rc;
    return finish();
}

int tlab_gpu_download(void* dst, const void* src, size_t bytes) {
    if (int rc = need_ready()) return rc;
    if (int rc = cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream), "download")) return rc;
    return cuda_check(cudaStreamSynchronize(ctx().stream), "download");
}

int tlab_gpu_stream(void** stream) {
    if (int rc = need_ready()) return rc;
    if (!stream) return fail(TLAB_ERR_OPTION, "null argument");
    *stream = (void*)ctx().stream;
    return 0;
}

int tlab_gpu_profile(int on) {
    Context& c = ctx();
    if (on && !c.profiling) {
        for (auto& r : c.prof) { c.event_pool.push_back(r.a); c.event_pool.push_back(r.b); }
        c.prof.clear();
    }
    c.profiling = (on != 0);
    return 0;
}

int tlab_gpu_profile_report(double* ms_per_class, int* count_per_class, int nclass) {
    Context& c = ctx();
    if (!ms_per_class || !count_per_class || nclass < PC_COUNT) return fail(TLAB_ERR_OPTION, "profile report: need PC_COUNT entries");
    if (int rc = cuda_check(cudaStreamSynchronize(c.stream), "profile report")) return rc;
    for (int i = 0; i < nclass; i++) { ms_per_class[i] = 0.0; count_per_class[i] = 0; }
    for (auto& r : c.prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { ms_per_class[r.cls] += ms; count_per_class[r.cls]++; }
        c.event_pool.push_back(r.a); c.event_pool.push_back(r.b);
    }
    c.prof.clear();
    return 0;
}

int tlab_gpu_set_tuning(const char* key, int value) {
    if (!key) return fail(TLAB_ERR_OPTION, "null key");
    if (!std::strcmp(key, "lines_x")) ctx().tune_lines_x = value;
    else if (!std::strcmp(key, "lines_yz")) ctx().tune_lines_yz = value;
    else if (!std::strcmp(key, "prefetch")) set_prefetch(value != 0);
    else if (!std::strcmp(key, "fast")) ctx().tune_fast = value;
    else if (!std::strcmp(key, "pf_dist")) ctx().tune_pf_dist = value;
    else if (!std::strcmp(key, "persist")) ctx().tune_persist = value;
    else if (!std::strcmp(key, "tma")) ctx().tune_tma = value;
    else if (!std::strcmp(key, "neu_compact")) ctx().tune_neu_compact = value;
    else if (!std::strcmp(key, "pair")) ctx().tune_pair = value;
    else if (!std::strcmp(key, "fuse_update")) ctx().tune_fuse_update = value;
    else if (!std::strcmp(key, "tma_l2")) ctx().tune_tma_l2 = value;
    else if (!std::strcmp(key, "splitz")) ctx().tune_splitz = value;
    else if (!std::strcmp(key, "split_emulate")) ctx().tune_split_emulate = value;
    else if (!std::strcmp(key, "pf_l1")) ctx().tune_pf_l1 = value;
    else if (!std::strcmp(key, "march")) ctx().tune_march = value;
    else if (!std::strcmp(key, "circ")) ctx().tune_circ = value;
    else if (!std::strcmp(key, "split_trim")) ctx().tune_split_trim = value;
    else if (!std::strcmp(key, "split_local")) ctx().tune_split_local = value;
    else if (!std::strcmp(key, "pull_overlap")) ctx().tune_pull_overlap = value;
    else if (!std::strcmp(key, "lazy_scale")) ctx().tune_lazy_scale = value;
    else if (!std::strcmp(key, "march_red")) ctx().tune_march_red = value;
    else if (!std::strcmp(key, "march_cfg")) ctx().tune_march_cfg = value;
    else if (!std::strcmp(key, "march_pf")) ctx().tune_march_pf = value;
    else if (!std::strcmp(key, "march_peel")) ctx().tune_march_peel = value;
    else if (!std::strcmp(key, "fuse")) ctx().tune_fuse = value;
    else if (!std::strcmp(key, "overlap")) ctx().tune_overlap = value;
    else if (!std::strcmp(key, "kxsplit")) ctx().tune_kxsplit = value;
    else if (!std::strcmp(key, "p2p")) trp().p2p_enabled = (value != 0);
    else if (!std::strcmp(key, "p2p_ctas")) trp().p2p_ctas = value > 0 ? value : 148;
    else if (!std::strcmp(key, "p2p_dma")) trp().p2p_dma = (value != 0);
    else if (!std::strcmp(key, "pf_next")) ctx().tune_pf_next = value;
    else if (!std::strcmp(key, "poisson_minb")) ctx().tune_poisson_minb = value;
    else if (!std::strcmp(key, "poisson_split")) ctx().tune_poisson_split = value;
    else if (!std::strcmp(key, "poisson_warp")) ctx().tune_poisson_warp = value;
    else if (!std::strcmp(key, "poisson_pf")) ctx().tune_poisson_pf = value;
    else if (!std::strcmp(key, "poisson_il")) ctx().tune_poisson_il = value;
    else if (!std::strcmp(key, "poisson_factors")) ctx().tune_poisson_factors = value;
    else return fail(TLAB_ERR_OPTION, std::string("unknown tuning key ") + key);
    return 0;
}

int tlab_gpu_get_counter(const char* key, long long* value) {
    if (!key || !value) return fail(TLAB_ERR_OPTION, "null argument");
    if (!std::strcmp(key, "fast_launches")) *value = ctx().fast_launches;
    else if (!std::strcmp(key, "general_launches")) *value = ctx().general_launches;
    else if (!std::strcmp(key, "tma_launches")) *value = lines2_tma_launches();
    else if (!std::strcmp(key, "march_launches")) *value = ctx().march_launches;
    else if (!std::strcmp(key, "splitz_ops")) *value = splitz().ops;
    else if (!std::strcmp(key, "splitz_march_ops")) *value = splitz().march_ops;
    else if (!std::strcmp(key, "p2p_exchanges")) *value = trp().p2p_exchanges;
    else if (!std::strcmp(key, "nccl_exchanges")) *value = trp().nccl_exchanges;
    else return fail(TLAB_ERR_OPTION, std::string("unknown counter ") + key);
    return 0;
}

static int plan_create(int dir, int n, const double* nodes, int periodic, int uniform, int mode1, int mode2,
                       tlab_plan_t* out, bool device) {
    if (device) { if (int rc = need_ready()) return rc; }
    if (!out || !nodes || n < 1) return fail(TLAB_ERR_OPTION, "tlab_fdm_plan_create: bad arguments");
    if (n > 1 && n < 16) return fail(TLAB_ERR_DIMGRID, "a direction needs 1 or at least 16 grid points");
    tlab_plan_s* g = new tlab_plan_s();
    g->dir = dir;
    int rc = create_plan(nodes, n, periodic != 0, uniform != 0, mode1, mode2, g->p.h);
    if (rc) {
        delete g;
        return fail(rc, rc == 85 ? "grid must be uniform in a periodic direction"
                                 : (rc == 104 ? "finite-difference scheme not implemented on the GPU path" : "plan creation failed"));
    }
    g->p.n = n;
    if (device) {
        rc = devplan_build(g->p);
        if (rc) { devplan_free(g->p); delete g; return fail(rc, "device plan upload failed"); }
    }
    *out = g;
    return 0;
}

int tlab_fdm_plan_create(int dir, int n, const double* nodes, int periodic, int uniform, int mode1, int mode2,
                         tlab_plan_t* out) {
    return plan_create(dir, n, nodes, periodic, uniform, mode1, mode2, out, true);
}

int tlab_fdm_plan_create_host(int dir, int n, const double* nodes, int periodic, int uniform, int mode1, int mode2,
                              tlab_plan_t* out) {
    return plan_create(dir, n, nodes, periodic, uniform, mode1, mode2, out, false);
}

int tlab_fdm_plan_destroy(tlab_plan_t g) {
    if (!g) return 0;
    for (int i = 0; i < 3; i++) if (ctx().burgers_plans[i] == g) ctx().burgers_plans[i] = nullptr;
    devplan_free(g->p);
    delete g;
    return 0;
}

static int copy_mat(const Mat& m, int c0, int c1, double* out, int cap, int* count) {
    const int nr = m.nrow(), nc = c1 - c0 + 1;
    if (count) *count = nr * nc;
    if (cap < nr * nc) return fail(TLAB_ERR_ALLOC, "output buffer too small");
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) out[(size_t)j * nr + i] = m(m.r0 + i, c0 + j);
    return 0;
}

static int copy_vec(const std::vector<double>& v, double* out, int cap, int* count) {
    if (count) *count = (int)v.size();
    if (cap < (int)v.size()) return fail(TLAB_ERR_ALLOC, "output buffer too small");
    std::copy(v.begin(), v.end(), out);
    return 0;
}

int tlab_fdm_plan_get(tlab_plan_t g, const char* what, double* out, int cap, int* count) {
    if (!g || !what || !out) return fail(TLAB_ERR_OPTION, "tlab_fdm_plan_get: bad arguments");
    const HostPlan& h = g->p.h;
    const std::string w(what);
    if (w == "nodes") return copy_vec(h.nodes, out, cap, count);
    if (w == "mwn1") return copy_vec(h.der1.mwn, out, cap, count);
    if (w == "mwn2") return copy_vec(h.der2.mwn, out, cap, count);
    if (w == "jac1") return copy_mat(h.jac, 1, 1, out, cap, count);
    if (w == "jac2") return copy_mat(h.jac, 2, 2, out, cap, count);
    if (w == "jac3") return copy_mat(h.jac, 3, 3, out, cap, count);
    if (h.size <= 1) return fail(TLAB_ERR_OPTION, "plan of size 1 has no tables");
    if (w == "lhs1") return copy_mat(h.der1.lhs, 1, h.der1.ndl, out, cap, count);
    if (w == "rhs1") return copy_mat(h.der1.rhs, 1, h.der1.ndr, out, cap, count);
    if (w == "lu1") return copy_mat(h.der1.lu, h.der1.lu.c0, h.der1.lu.c1, out, cap, count);
    if (w == "rhs1_b") return copy_mat(h.der1.rhs_b, 0, 7, out, cap, count);
    if (w == "rhs1_t") return copy_mat(h.der1.rhs_t, 1, 7, out, cap, count);
    if (w == "lhs2") return copy_mat(h.der2.lhs, 1, h.der2.ndl, out, cap, count);
    if (w == "rhs2") return copy_mat(h.der2.rhs, 1, h.der2.ndr + h.der2.ndl, out, cap, count);
    if (w == "lu2") return copy_mat(h.der2.lu, h.der2.lu.c0, h.der2.lu.c1, out, cap, count);
    return fail(TLAB_ERR_OPTION, "tlab_fdm_plan_get: unknown table " + w);
}

// FDM_Int1_CreateSystem for one eigenvalue, assembled on the host exactly as the device threads do it
// (L0 + lambda*L1, then the reduction at the opposite end); used to validate the split tables.
int tlab_fdm_int1_system_host(tlab_plan_t g, int ibc, double lambda, double* lhs, double* rhs, double* rhs_b, double* rhs_t) {
    if (!g || !lhs || !rhs || !rhs_b || !rhs_t) return fail(TLAB_ERR_OPTION, "tlab_fdm_int1_system_host: null argument");
    if (ibc != BCS_MIN && ibc != BCS_MAX) return fail(TLAB_ERR_OPTION, "ibc must be BCS_MIN (1) or BCS_MAX (2)");
    HostInt1 H;
    if (int rc = int1_create_base(g->p.h.der1, ibc, H)) return fail(rc, "integral operator not available for this scheme");
    const int n = H.n;
    Mat L(1, n, 1, 5);
    for (int r = 1; r <= n; r++) for (int k = 1; k <= 5; k++) L(r, k) = H.L0(r, k) + lambda * H.L1(r, k);
    if (ibc == BCS_MIN) fdm_bcs_reduce(BCS_MAX, L, H.rhs, nullptr, &H.rhs_t0);
    else fdm_bcs_reduce(BCS_MIN, L, H.rhs, &H.rhs_b0, nullptr);
    for (int k = 1; k <= 5; k++) for (int r = 1; r <= n; r++) lhs[(size_t)(k - 1) * n + r - 1] = L(r, k);
    for (int k = 1; k <= 3; k++) for (int r = 1; r <= n; r++) rhs[(size_t)(k - 1) * n + r - 1] = H.rhs(r, k);
    for (int c = 0; c <= 7; c++) for (int r = 1; r <= 5; r++) rhs_b[(size_t)c * 5 + r - 1] = H.rhs_b0(r, c);
    for (int c = 1; c <= 8; c++) for (int r = 0; r <= 4; r++) rhs_t[(size_t)(c - 1) * 5 + r] = H.rhs_t0(r, c);
    return 0;
}

int tlab_opr_partial(int dir, int type, int nx, int ny, int nz, const int bcs[4], tlab_plan_t g, const double* u,
                     double* result, double* tmp1) {
    if (int rc = need_ready()) return rc;
    if (!bcs || !u || !result) return fail(TLAB_ERR_OPTION, "OPR_Partial: null argument");
    if (u == result || (tmp1 && (tmp1 == u || tmp1 == result)))
        return fail(TLAB_ERR_OPTION, "OPR_Partial: u, result and tmp1 must not alias");
    const int ibc = bcs[0] + bcs[1] * 2;       // opr_partial.f90:91
    if (int rc = run_partial(dir, type, nx, ny, nz, ibc, g, u, result, tmp1)) return rc;
    return finish();
}

int tlab_opr_burgers_init(tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz, double visc, int nscal, const double* schmidt) {
    if (int rc = need_ready()) return rc;
    if (nscal < 0 || (nscal > 0 && !schmidt)) return fail(TLAB_ERR_OPTION, "OPR_Burgers_Initialize: bad scalar list");
    tlab_plan_s* gs[3] = {gx, gy, gz};
    for (int ig = 0; ig < 3; ig++) {
        tlab_plan_s* g = gs[ig];
        ctx().burgers_plans[ig] = g;
        if (!g) continue;
        if (g->p.n <= 1) { g->burgers_first = 0; g->burgers_count = nscal + 1; continue; }
        if (g->p.h.der2.ndl != 3) return fail(TLAB_ERR_OPTION, "Burgers: more than 3 LHS diagonals in the second derivative");
        g->burgers_first = (int)g->p.lu2.size();
        g->burgers_count = nscal + 1;
        for (int is = 0; is <= nscal; is++) {
            const double d = (is == 0) ? visc : visc / schmidt[is - 1];
            devplan_add_diffusion(g->p, d);
        }
    }
    return cuda_check(cudaGetLastError(), "burgers init");
}

int tlab_opr_burgers(int dir, int ivel, int is, int nx, int ny, int nz, const int bcs[4], const double* s, const double* u,
                     double* result, double* tmp1, const double* u_t) {
    (void)tmp1; (void)u_t;
    if (int rc = need_ready()) return rc;
    if (!bcs || !s || !result) return fail(TLAB_ERR_OPTION, "OPR_Burgers: null argument");
    if (bcs[2] + bcs[3] > 0) return fail(TLAB_ERR_UNDEVELOP, "OPR_Burgers: only developed for biased BCs");  // opr_burgers.f90:460-463
    if (dir < 1 || dir > 3) return fail(TLAB_ERR_OPTION, "direction must be 1, 2 or 3");
    tlab_plan_s* g = ctx().burgers_plans[dir - 1];
    if (!g) return fail(TLAB_ERR_OPTION, "tlab_opr_burgers_init has not been called");
    const double* vel = (ivel == TLAB_OPR_B_SELF) ? s : u;
    if (!vel) return fail(TLAB_ERR_OPTION, "OPR_Burgers: velocity missing");
    if (result == s || result == vel) return fail(TLAB_ERR_OPTION, "OPR_Burgers: result must not alias the inputs");
    const int ibc = bcs[0] + bcs[1] * 2;
    if (int rc = run_burgers(dir, is, nx, ny, nz, ibc, g, s, vel, result, 0)) return rc;
    return finish();
}

int tlab_fdm_der1_solve(tlab_plan_t g, int nlines, int ibc, const double* u, double* result) {
    if (int rc = need_ready()) return rc;
    if (!g || !u || !result || nlines < 1) return fail(TLAB_ERR_OPTION, "FDM_Der1_Solve: bad arguments");
    // u(nlines, n): the derivative direction is the slow index -> same access pattern as a z-derivative
    if (int rc = run_partial(3, TLAB_OPR_P1, nlines, 1, g->p.n, ibc, g, u, result, nullptr)) return rc;
    return finish();
}

int tlab_fdm_der2_solve(tlab_plan_t g, int nlines, int is, const double* u, const double* du, double* result) {
    (void)du;   // the first derivative needed on non-uniform grids is recomputed in registers
    if (int rc = need_ready()) return rc;
    if (!g || !u || !result || nlines < 1) return fail(TLAB_ERR_OPTION, "FDM_Der2_Solve: bad arguments");
    if (is < 0) {
        if (int rc = run_partial(3, TLAB_OPR_P2, nlines, 1, g->p.n, 0, g, u, result, nullptr)) return rc;
        return finish();
    }
    return fail(TLAB_ERR_UNDEVELOP, "FDM_Der2_Solve with a diffusivity-scaled LU is only reachable through OPR_Burgers");
}

int tlab_boundary_bcs_neumann_y(int ibc, int nx, int ny, int nz, tlab_plan_t gy, const double* u, double* hb, double* ht) {
    if (int rc = need_ready()) return rc;
    if (!u || !hb || !ht) return fail(TLAB_ERR_OPTION, "BOUNDARY_BCS_NEUMANN_Y: null argument");
    if (int rc = run_neumann_y(ibc, nx, ny, nz, gy, u, hb, ht)) return rc;
    return finish();
}

}  // extern "C"
