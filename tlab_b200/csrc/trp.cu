// Global z-slab <-> z-pencil transposes over NCCL (one process per GPU, NVLink/NVSwitch).
//
// Replaces TLabMPI_Trp_PlanK / TLabMPI_Trp_ExecK_Forward/Backward (real and complex),
// src/base/tlab_mpi_transpose.f90:301-339, 343-553, and the rank layout of src/base/tlab_mpi_procs.f90:40-58
// for a pure z decomposition (ims_npro_k = P, ims_npro_i = 1).
//
// Forward map (PlanK): with nlines = nxy / P, rank r sends to rank p the sub-block
// a[p*nlines : (p+1)*nlines, 0 : nzl) of its slab a(nxy, nzl) and stores what it receives from rank q at
// b[0 : nlines, q*nzl : (q+1)*nzl) of the pencil b(nlines, nz).  The receive side is contiguous per peer;
// the send side is gathered by a pack kernel (the reference lets MPI derived types do that).  Backward is
// the inverse, with the scatter (optionally accumulating, +/-) done by the unpack kernel.
#include "../../include/tlab_gpu.h"
#include "context.h"
#include "trp.h"
#include <cstring>
#include <vector>

namespace tlab {

namespace {

// sendbuf[p][k][i] = a[k*nxy + p*nl + i]   (optionally a + scale*a2)
__global__ void pack_kernel(const double* __restrict__ a, const double* __restrict__ a2, double scale,
                            double* __restrict__ sendbuf, long long nxy, int nzl, long long nl, int P) {
    const long long total = nxy * nzl;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / nxy, j = e - k * nxy;       // j = p*nl + i
        const long long p = j / nl, i = j - p * nl;
        double v = a[e];
        if (a2 != nullptr) v = v + a2[e] * scale;
        sendbuf[(p * nzl + k) * nl + i] = v;
    }
}

// a[k*nxy + q*nl + i] (op)= recvbuf[q][k][i]
__global__ void unpack_kernel(const double* __restrict__ recvbuf, double* __restrict__ a, long long nxy, int nzl,
                              long long nl, int P, int accumulate) {
    const long long total = nxy * nzl;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / nxy, j = e - k * nxy;
        const long long q = j / nl, i = j - q * nl;
        const double v = recvbuf[(q * nzl + k) * nl + i];
        if (accumulate == 0) a[e] = v;
        else if (accumulate > 0) a[e] = a[e] + v;
        else a[e] = a[e] - v;
    }
}

inline unsigned blocks_for(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b < 148LL * 16 ? b : 148LL * 16);
}

}  // namespace

Trp& trp() {
    static Trp t;
    return t;
}

#ifdef TLAB_HAVE_NCCL
static int nccl_check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return 0;
    return fail(TLAB_ERR_CUDA, std::string(what) + ": " + ncclGetErrorString(r));
}
#endif

int Trp::ensure(size_t doubles) {
    if (cap >= doubles) return 0;
    if (sendbuf) cudaFree(sendbuf);
    sendbuf = nullptr;
    if (cudaMalloc(&sendbuf, doubles * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        cap = 0;
        return fail(TLAB_ERR_ALLOC, "transpose buffer: out of device memory");
    }
    cap = doubles;
    return 0;
}

// all-to-all of P contiguous blocks of `count` doubles each
int Trp::alltoall(const double* src, double* dst, size_t count) {
    cudaStream_t st = ctx().stream;
    if (P == 1) {
        return cuda_check(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, st), "transpose copy");
    }
#ifdef TLAB_HAVE_NCCL
    if (int rc = nccl_check(ncclGroupStart(), "ncclGroupStart")) return rc;
    for (int p = 0; p < P; p++) {
        if (int rc = nccl_check(ncclSend(src + (size_t)p * count, count, ncclDouble, p, comm, st), "ncclSend")) return rc;
        if (int rc = nccl_check(ncclRecv(dst + (size_t)p * count, count, ncclDouble, p, comm, st), "ncclRecv")) return rc;
    }
    return nccl_check(ncclGroupEnd(), "ncclGroupEnd");
#else
    return fail(TLAB_ERR_UNDEVELOP, "library built without NCCL");
#endif
}

int Trp::forward(const double* a, const double* a2, double scale, double* b, long long nxy, int nzl) {
    if (nxy % P) return fail(TLAB_ERR_PARPARTITION, "transpose: number of lines is not a multiple of the number of ranks");
    const long long nl = nxy / P;
    const size_t total = (size_t)nxy * nzl;
    if (int rc = ensure(total)) return rc;
    ProfScope ps(PC_TRANSPOSE);
    pack_kernel<<<blocks_for((long long)total), 256, 0, ctx().stream>>>(a, a2, scale, sendbuf, nxy, nzl, nl, P);
    launches++;
    return alltoall(sendbuf, b, (size_t)nl * nzl);
}

int Trp::backward(const double* b, double* a, long long nxy, int nzl, int accumulate) {
    if (nxy % P) return fail(TLAB_ERR_PARPARTITION, "transpose: number of lines is not a multiple of the number of ranks");
    const long long nl = nxy / P;
    const size_t total = (size_t)nxy * nzl;
    if (int rc = ensure(total)) return rc;
    ProfScope ps(PC_TRANSPOSE);
    if (int rc = alltoall(b, sendbuf, (size_t)nl * nzl)) return rc;
    unpack_kernel<<<blocks_for((long long)total), 256, 0, ctx().stream>>>(sendbuf, a, nxy, nzl, nl, P, accumulate);
    launches++;
    return cuda_check(cudaGetLastError(), "unpack");
}

}  // namespace tlab

using namespace tlab;

extern "C" {

int tlab_mpi_get_unique_id(void* id_out_128) {
#ifdef TLAB_HAVE_NCCL
    if (!id_out_128) return fail(TLAB_ERR_OPTION, "null argument");
    ncclUniqueId id;
    if (int rc = nccl_check(ncclGetUniqueId(&id), "ncclGetUniqueId")) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "unexpected ncclUniqueId size");
    std::memcpy(id_out_128, &id, 128);
    return 0;
#else
    (void)id_out_128;
    return fail(TLAB_ERR_UNDEVELOP, "library built without NCCL");
#endif
}

int tlab_mpi_init(int rank, int nranks, const void* id_128) {
    if (int rc = tlab_gpu_init(-1)) return rc;
    Trp& t = trp();
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(TLAB_ERR_PARPARTITION, "bad rank / number of ranks");
    t.rank = rank; t.P = nranks;
    if (nranks == 1) return 0;
#ifdef TLAB_HAVE_NCCL
    if (!id_128) return fail(TLAB_ERR_OPTION, "null unique id");
    ncclUniqueId id;
    std::memcpy(&id, id_128, 128);
    return nccl_check(ncclCommInitRank(&t.comm, nranks, id, rank), "ncclCommInitRank");
#else
    return fail(TLAB_ERR_UNDEVELOP, "library built without NCCL");
#endif
}

int tlab_mpi_finalize(void) {
    Trp& t = trp();
#ifdef TLAB_HAVE_NCCL
    if (t.comm) { ncclCommDestroy(t.comm); t.comm = nullptr; }
#endif
    if (t.sendbuf) cudaFree(t.sendbuf);
    t.sendbuf = nullptr; t.cap = 0; t.P = 1; t.rank = 0;
    return 0;
}

int tlab_mpi_rank(int* rank, int* nranks) {
    if (rank) *rank = trp().rank;
    if (nranks) *nranks = trp().P;
    return 0;
}

int tlab_trp_exec_k_forward(const double* a, double* b, int nlines_total, int kmax, int is_complex) {
    if (!a || !b || a == b) return fail(TLAB_ERR_OPTION, "TLabMPI_Trp_ExecK_Forward: bad arguments");
    const long long nxy = (long long)nlines_total * (is_complex ? 2 : 1);
    if (nlines_total % trp().P) return fail(TLAB_ERR_PARPARTITION, "number of lines is not a multiple of the number of ranks");
    if (int rc = trp().forward(a, nullptr, 0.0, b, nxy, kmax)) return rc;
    return finish();
}

int tlab_trp_exec_k_backward(const double* b, double* a, int nlines_total, int kmax, int is_complex) {
    if (!a || !b || a == b) return fail(TLAB_ERR_OPTION, "TLabMPI_Trp_ExecK_Backward: bad arguments");
    const long long nxy = (long long)nlines_total * (is_complex ? 2 : 1);
    if (nlines_total % trp().P) return fail(TLAB_ERR_PARPARTITION, "number of lines is not a multiple of the number of ranks");
    if (int rc = trp().backward(b, a, nxy, kmax, 0)) return rc;
    return finish();
}

}  // extern "C"
