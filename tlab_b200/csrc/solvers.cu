// Stand-alone batched Thomas solvers and the out-of-place 2-D transpose, with the reference's argument lists.
//
// Replaces the substitution stages TRIDSS / TRIDPSS (src/utils/linear3.f90:56-150, 321-442), PENTADSS /
// PENTADSS2 (src/utils/linear5.f90:76-131, 209-244) and TLab_Transpose / TLab_Transpose_COMPLEX
// (src/utils/tlab_transpose.f90:14-82, 148-210) for callers that bring their own factored diagonals.
// Data are in the reference's lines-first layout f(len, nmax): one thread owns one line and marches along
// nmax, consecutive threads own consecutive lines (coalesced).  The fused operators (lines.cu) do not go
// through these entry points; they exist so that the whole thomas3/thomas5 surface has a device equivalent.
#include "../../include/tlab_gpu.h"
#include "context.h"

namespace tlab {
namespace {

__global__ void tridss_kernel(int nmax, long long len, const double* __restrict__ a, const double* __restrict__ b,
                              const double* __restrict__ c, double* __restrict__ f) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= len) return;
    double prev = f[l];
    for (int n = 1; n < nmax; n++) {
        const double v = f[l + n * len] + a[n] * prev;
        f[l + n * len] = v;
        prev = v;
    }
    prev = prev * b[nmax - 1];
    f[l + (long long)(nmax - 1) * len] = prev;
    for (int n = nmax - 2; n >= 0; n--) {
        const double v = (f[l + n * len] + c[n] * prev) * b[n];
        f[l + n * len] = v;
        prev = v;
    }
}

__global__ void tridpss_kernel(int nmax, long long len, const double* __restrict__ a, const double* __restrict__ b,
                               const double* __restrict__ c, const double* __restrict__ d, const double* __restrict__ e,
                               double* __restrict__ f, double* __restrict__ wrk) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= len) return;
    double prev = f[l] * b[0];
    f[l] = prev;
    double w = d[0] * prev;
    for (int n = 1; n < nmax - 1; n++) {
        const double v = f[l + n * len] * b[n] + a[n] * prev;
        f[l + n * len] = v;
        w = w + d[n] * v;
        prev = v;
    }
    if (wrk) wrk[l] = w;
    const double xn = (f[l + (long long)(nmax - 1) * len] - w) * b[nmax - 1];
    f[l + (long long)(nmax - 1) * len] = xn;
    prev = e[nmax - 2] * xn + f[l + (long long)(nmax - 2) * len];
    f[l + (long long)(nmax - 2) * len] = prev;
    for (int n = nmax - 3; n >= 0; n--) {
        const double v = f[l + n * len] + c[n] * prev + e[n] * xn;
        f[l + n * len] = v;
        prev = v;
    }
}

__global__ void pentadss_kernel(int nmax, long long len, const double* __restrict__ a, const double* __restrict__ b,
                                const double* __restrict__ c, const double* __restrict__ d, const double* __restrict__ e,
                                double* __restrict__ f) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= len) return;
    double p2 = f[l];
    double p1 = f[l + len] + p2 * b[1];
    f[l + len] = p1;
    for (int n = 2; n < nmax; n++) {
        const double v = f[l + n * len] + p1 * b[n] + p2 * a[n];
        f[l + n * len] = v;
        p2 = p1; p1 = v;
    }
    double x1 = p1 * c[nmax - 1];
    f[l + (long long)(nmax - 1) * len] = x1;
    double x0 = (f[l + (long long)(nmax - 2) * len] + x1 * d[nmax - 2]) * c[nmax - 2];
    f[l + (long long)(nmax - 2) * len] = x0;
    double q1 = x0, q2 = x1;      // x(n+1), x(n+2)
    for (int n = nmax - 3; n >= 0; n--) {
        const double v = (f[l + n * len] + q1 * d[n] + q2 * e[n]) * c[n];
        f[l + n * len] = v;
        q2 = q1; q1 = v;
    }
}

__global__ void pentadss2_kernel(int nmax, long long len, const double* __restrict__ a, const double* __restrict__ b,
                                 const double* __restrict__ c, const double* __restrict__ d, const double* __restrict__ e,
                                 double* __restrict__ f) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= len) return;
    double q2 = f[l + (long long)(nmax - 1) * len];
    double q1 = f[l + (long long)(nmax - 2) * len] - q2 * d[nmax - 2];
    f[l + (long long)(nmax - 2) * len] = q1;
    for (int n = nmax - 3; n >= 0; n--) {
        const double v = f[l + n * len] - q1 * d[n] - q2 * e[n];
        f[l + n * len] = v;
        q2 = q1; q1 = v;
    }
    double p2 = q1 / c[0];
    f[l] = p2;
    double p1 = (f[l + len] - p2 * b[1]) / c[1];
    f[l + len] = p1;
    for (int n = 2; n < nmax; n++) {
        const double v = (f[l + n * len] - p1 * b[n] - p2 * a[n]) / c[n];
        f[l + n * len] = v;
        p2 = p1; p1 = v;
    }
}

// b(k, j) = a(j, k): a(ma, nca) with nra rows used, b(mb, nra); elements of ESIZE doubles
template <int ESIZE>
__global__ void transpose_kernel(const double* __restrict__ a, int nra, int nca, int ma, double* __restrict__ b, int mb) {
    __shared__ double tile[32][33 * ESIZE];
    const int j0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + threadIdx.x, k = k0 + r;
        if (j < nra && k < nca)
            for (int c = 0; c < ESIZE; c++) tile[r][threadIdx.x * ESIZE + c] = a[((size_t)k * ma + j) * ESIZE + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + threadIdx.x, j = j0 + r;
        if (j < nra && k < nca)
            for (int c = 0; c < ESIZE; c++) b[((size_t)j * mb + k) * ESIZE + c] = tile[threadIdx.x][r * ESIZE + c];
    }
}

inline unsigned nblk(long long len) { return (unsigned)((len + 127) / 128); }

}  // namespace
}  // namespace tlab

using namespace tlab;

#define CHECK_SOLVER_ARGS(name)                                                                   \
    if (int rc = tlab_gpu_init(-1)) return rc;                                                    \
    if (nmax < 3 || len < 1 || !f) return fail(TLAB_ERR_OPTION, name ": bad arguments");

extern "C" {

int tlab_tridss(int nmax, int len, const double* a, const double* b, const double* c, double* f) {
    CHECK_SOLVER_ARGS("TRIDSS")
    tridss_kernel<<<nblk(len), 128, 0, ctx().stream>>>(nmax, len, a, b, c, f);
    return finish();
}

int tlab_tridpss(int nmax, int len, const double* a, const double* b, const double* c, const double* d, const double* e,
                 double* f, double* wrk) {
    CHECK_SOLVER_ARGS("TRIDPSS")
    tridpss_kernel<<<nblk(len), 128, 0, ctx().stream>>>(nmax, len, a, b, c, d, e, f, wrk);
    return finish();
}

int tlab_pentadss(int nmax, int len, const double* a, const double* b, const double* c, const double* d, const double* e,
                  double* f) {
    CHECK_SOLVER_ARGS("PENTADSS")
    pentadss_kernel<<<nblk(len), 128, 0, ctx().stream>>>(nmax, len, a, b, c, d, e, f);
    return finish();
}

int tlab_pentadss2(int nmax, int len, const double* a, const double* b, const double* c, const double* d, const double* e,
                   double* f) {
    CHECK_SOLVER_ARGS("PENTADSS2")
    pentadss2_kernel<<<nblk(len), 128, 0, ctx().stream>>>(nmax, len, a, b, c, d, e, f);
    return finish();
}

int tlab_transpose(const double* a, int nra, int nca, int ma, double* b, int mb) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    if (!a || !b || a == b || nra < 1 || nca < 1 || ma < nra || mb < nca) return fail(TLAB_ERR_OPTION, "TLab_Transpose: bad arguments");
    dim3 grid((nra + 31) / 32, (nca + 31) / 32), block(32, 8);
    transpose_kernel<1><<<grid, block, 0, ctx().stream>>>(a, nra, nca, ma, b, mb);
    return finish();
}

int tlab_transpose_complex(const double* a, int nra, int nca, int ma, double* b, int mb) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    if (!a || !b || a == b || nra < 1 || nca < 1 || ma < nra || mb < nca) return fail(TLAB_ERR_OPTION, "TLab_Transpose_COMPLEX: bad arguments");
    dim3 grid((nra + 31) / 32, (nca + 31) / 32), block(32, 8);
    transpose_kernel<2><<<grid, block, 0, ctx().stream>>>(a, nra, nca, ma, b, mb);
    return finish();
}

}  // extern "C"
