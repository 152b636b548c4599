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
#include <algorithm>

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

// a (op)= b   /   c = a + scale*a2
__global__ void accumulate_kernel(double* __restrict__ a, const double* __restrict__ b, long long n, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a[i] = (accumulate > 0) ? a[i] + b[i] : a[i] - b[i];
}
__global__ void combine_kernel(double* __restrict__ c, const double* __restrict__ a, const double* __restrict__ a2, double scale,
                               long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c[i] = a[i] + a2[i] * scale;
}

// forward transpose by peer stores: b_p[(rank*nzl + k)*nl + i] = a[k*nxy + p*nl + i] (+ scale*a2)
// grid = (chunks of a run, nzl, P); every run of nl doubles is contiguous on both sides
__global__ void push_forward_kernel(const double* __restrict__ a, const double* __restrict__ a2, double scale, Trp::PeerTab dst,
                                    long long nxy, int nzl, long long nl, int rank, int P) {
    // runs (k, p) of nl contiguous doubles are dealt to the CTAs round-robin, starting with the next peer so that the
    // ranks do not all store into the same GPU at the same time
    const int nruns = nzl * P;
    for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const int k = run / P, p = (run % P + rank + 1) % P;
        const double* __restrict__ src = a + (long long)k * nxy + (long long)p * nl;
        const double* __restrict__ src2 = a2 ? a2 + (long long)k * nxy + (long long)p * nl : nullptr;
        double* __restrict__ d = dst.p[p] + ((long long)rank * nzl + k) * nl;
        if ((nl & 1) == 0 && ((reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(d) | reinterpret_cast<size_t>(src2)) & 15) == 0) {
            const long long n2 = nl >> 1;
            constexpr int U = 8;                   // independent 16-byte loads in flight per thread
            long long i = threadIdx.x;
            for (; i + (long long)(U - 1) * blockDim.x < n2; i += (long long)U * blockDim.x) {
                double2 v[U];
#pragma unroll
                for (int k = 0; k < U; k++) v[k] = __ldcs(reinterpret_cast<const double2*>(src) + i + (long long)k * blockDim.x);
                if (src2) {
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        const double2 w = __ldcs(reinterpret_cast<const double2*>(src2) + i + (long long)k * blockDim.x);
                        v[k].x = v[k].x + w.x * scale; v[k].y = v[k].y + w.y * scale;
                    }
                }
#pragma unroll
                for (int k = 0; k < U; k++) reinterpret_cast<double2*>(d)[i + (long long)k * blockDim.x] = v[k];
            }
            for (; i < n2; i += blockDim.x) {
                double2 v = __ldcs(reinterpret_cast<const double2*>(src) + i);
                if (src2) { const double2 w = __ldcs(reinterpret_cast<const double2*>(src2) + i); v.x = v.x + w.x * scale; v.y = v.y + w.y * scale; }
                reinterpret_cast<double2*>(d)[i] = v;
            }
        } else {
            for (long long i = threadIdx.x; i < nl; i += blockDim.x) {
                double v = src[i];
                if (src2) v = v + src2[i] * scale;
                d[i] = v;
            }
        }
    }
}

// backward transpose by peer loads: a[k*nxy + q*nl + i] (op)= b_q[(rank*nzl + k)*nl + i]
__global__ void pull_backward_kernel(Trp::PeerTab srcs, double* __restrict__ a, long long nxy, int nzl, long long nl, int rank,
                                     int P, int accumulate) {
    const int nruns = nzl * P;
    for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const int k = run / P, q = (run % P + rank + 1) % P;
        const double* __restrict__ src = srcs.p[q] + ((long long)rank * nzl + k) * nl;
        double* __restrict__ d = a + (long long)k * nxy + (long long)q * nl;
        if ((nl & 1) == 0 && ((reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(d)) & 15) == 0) {
            const long long n2 = nl >> 1;
            constexpr int U = 8;
            long long i = threadIdx.x;
            for (; i + (long long)(U - 1) * blockDim.x < n2; i += (long long)U * blockDim.x) {
                double2 v[U];
#pragma unroll
                for (int k = 0; k < U; k++) v[k] = reinterpret_cast<const double2*>(src)[i + (long long)k * blockDim.x];
                if (accumulate != 0) {
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        const double2 o = reinterpret_cast<const double2*>(d)[i + (long long)k * blockDim.x];
                        if (accumulate > 0) { v[k].x = o.x + v[k].x; v[k].y = o.y + v[k].y; } else { v[k].x = o.x - v[k].x; v[k].y = o.y - v[k].y; }
                    }
                }
#pragma unroll
                for (int k = 0; k < U; k++) reinterpret_cast<double2*>(d)[i + (long long)k * blockDim.x] = v[k];
            }
            for (; i < n2; i += blockDim.x) {
                double2 v = reinterpret_cast<const double2*>(src)[i];
                if (accumulate != 0) {
                    const double2 o = reinterpret_cast<const double2*>(d)[i];
                    if (accumulate > 0) { v.x = o.x + v.x; v.y = o.y + v.y; } else { v.x = o.x - v.x; v.y = o.y - v.y; }
                }
                reinterpret_cast<double2*>(d)[i] = v;
            }
        } else {
            for (long long i = threadIdx.x; i < nl; i += blockDim.x) {
                double v = src[i];
                if (accumulate > 0) v = d[i] + v; else if (accumulate < 0) v = d[i] - v;
                d[i] = v;
            }
        }
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

const Trp::PeerTab* Trp::find(const double* base) const {
    for (const auto& e : registry) if (e.first == base) return &e.second;
    return nullptr;
}

void Trp::unregister_buffer(const double* base) {
    for (size_t e = 0; e < registry.size(); e++) {
        if (registry[e].first != base) continue;
        for (int q = 0; q < 8; q++) if (q != rank && registry[e].second.p[q]) cudaIpcCloseMemHandle(registry[e].second.p[q]);
        registry.erase(registry.begin() + e);
        return;
    }
}

int Trp::barrier() {
#ifdef TLAB_HAVE_NCCL
    if (P == 1) return 0;
    if (!barrier_buf) {
        if (cudaMalloc(&barrier_buf, 64) != cudaSuccess) return fail(TLAB_ERR_ALLOC, "barrier buffer");
        cudaMemsetAsync(barrier_buf, 0, 64, ctx().stream);
    }
    return nccl_check(ncclAllReduce(barrier_buf, barrier_buf, 1, ncclInt, ncclMax, comm, ctx().stream), "barrier");
#else
    return 0;
#endif
}

// Collective.  Publishes `base` (a cudaMalloc allocation) to the other ranks and maps theirs.
int Trp::register_buffer(double* base) {
    if (P == 1 || !p2p_enabled || P > 8) return 0;
#ifdef TLAB_HAVE_NCCL
    cudaStream_t st = ctx().stream;
    cudaIpcMemHandle_t mine;
    int ok = (cudaIpcGetMemHandle(&mine, base) == cudaSuccess) ? 1 : 0;
    if (!ok) { cudaGetLastError(); std::memset(&mine, 0, sizeof(mine)); }
    // handles (64 bytes) + ok flag, gathered through NCCL
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 64;
    unsigned char* dbuf = nullptr;
    if (cudaMalloc(&dbuf, rec * (P + 1)) != cudaSuccess) return fail(TLAB_ERR_ALLOC, "ipc exchange buffer");
    std::vector<unsigned char> h(rec * (P + 1), 0);
    std::memcpy(h.data(), &mine, sizeof(mine));
    h[sizeof(mine)] = (unsigned char)ok;
    cudaMemcpyAsync(dbuf, h.data(), rec, cudaMemcpyHostToDevice, st);
    int rc = nccl_check(ncclAllGather(dbuf, dbuf + rec, rec, ncclChar, comm, st), "ncclAllGather(ipc handles)");
    if (!rc) rc = cuda_check(cudaMemcpyAsync(h.data(), dbuf, rec * (P + 1), cudaMemcpyDeviceToHost, st), "ipc handles");
    if (!rc) rc = cuda_check(cudaStreamSynchronize(st), "ipc handles");
    cudaFree(dbuf);
    if (rc) return rc;
    bool all_ok = true;
    for (int q = 0; q < P; q++) if (!h[rec * (q + 1) + sizeof(mine)]) all_ok = false;
    PeerTab tab;
    for (int q = 0; q < 8; q++) tab.p[q] = nullptr;
    int opened = 1;
    if (all_ok) {
        for (int q = 0; q < P && opened; q++) {
            if (q == rank) { tab.p[q] = base; continue; }
            cudaIpcMemHandle_t hq;
            std::memcpy(&hq, h.data() + rec * (q + 1), sizeof(hq));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; }
            tab.p[q] = (double*)ptr;
        }
    } else opened = 0;
    // every rank must agree, otherwise the barriers of the two paths would not match
    int* flag = nullptr;
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return fail(TLAB_ERR_ALLOC, "ipc flag");
    cudaMemcpyAsync(flag, &opened, sizeof(int), cudaMemcpyHostToDevice, st);
    rc = nccl_check(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, comm, st), "ncclAllReduce(ipc)");
    int agreed = 0;
    if (!rc) rc = cuda_check(cudaMemcpyAsync(&agreed, flag, sizeof(int), cudaMemcpyDeviceToHost, st), "ipc flag");
    if (!rc) rc = cuda_check(cudaStreamSynchronize(st), "ipc flag");
    cudaFree(flag);
    if (rc) return rc;
    if (agreed) registry.emplace_back(base, tab);
    else {
        for (int q = 0; q < P; q++) if (q != rank && tab.p[q]) cudaIpcCloseMemHandle(tab.p[q]);
        p2p_enabled = false;            // no peer mapping in this environment: stay on the NCCL path
        registry.clear();
    }
    return 0;
#else
    (void)base;
    return 0;
#endif
}

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
    const PeerTab* tabf = (P > 1 && p2p_enabled) ? find(b) : nullptr;
    if (tabf && p2p_dma) {
        // copy engines: one strided 2-D copy per peer, no SM involved (overlaps with the x/y kernels of the other stream)
        ProfScope ps(PC_TRANSPOSE);
        cudaStream_t st = ctx().stream;
        const double* src = a;
        if (a2 != nullptr) {
            if (int rc = ensure(total)) return rc;
            combine_kernel<<<blocks_for((long long)total), 256, 0, st>>>(sendbuf, a, a2, scale, (long long)total);
            launches++;
            src = sendbuf;
        }
        for (int d = 1; d <= P; d++) {
            const int q = (rank + d) % P;
            if (int rc = cuda_check(cudaMemcpy2DAsync(tabf->p[q] + (long long)rank * nzl * nl, (size_t)nl * sizeof(double), src + (long long)q * nl,
                                                      (size_t)nxy * sizeof(double), (size_t)nl * sizeof(double), (size_t)nzl,
                                                      cudaMemcpyDefault, st), "peer copy (forward)")) return rc;
        }
        p2p_exchanges++;
        return barrier();
    }
    if (const PeerTab* tab = tabf) {
        ProfScope ps(PC_TRANSPOSE);
        push_forward_kernel<<<(unsigned)std::min<long long>(p2p_ctas, (long long)nzl * P), 512, 0, ctx().stream>>>(a, a2, scale, *tab, nxy, nzl, nl, rank, P);
        launches++;
        p2p_exchanges++;
        return barrier();               // every pencil is complete when the consumers start
    }
    nccl_exchanges++;
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
    const PeerTab* tabb = (P > 1 && p2p_enabled) ? find(b) : nullptr;
    if (tabb && p2p_dma) {
        ProfScope ps(PC_TRANSPOSE);
        cudaStream_t st = ctx().stream;
        if (accumulate != 0) { if (int rc = ensure(total)) return rc; }
        if (int rc = barrier()) return rc;          // every rank has finished producing its pencil
        double* dst = (accumulate != 0) ? sendbuf : a;
        for (int d = 1; d <= P; d++) {
            const int q = (rank + d) % P;
            if (int rc = cuda_check(cudaMemcpy2DAsync(dst + (long long)q * nl, (size_t)nxy * sizeof(double), tabb->p[q] + (long long)rank * nzl * nl,
                                                      (size_t)nl * sizeof(double), (size_t)nl * sizeof(double), (size_t)nzl,
                                                      cudaMemcpyDefault, st), "peer copy (backward)")) return rc;
        }
        if (int rc = barrier()) return rc;          // the pencils may be overwritten again
        if (accumulate != 0) {
            accumulate_kernel<<<blocks_for((long long)total), 256, 0, st>>>(a, sendbuf, (long long)total, accumulate);
            launches++;
        }
        p2p_exchanges++;
        return cuda_check(cudaGetLastError(), "accumulate");
    }
    if (const PeerTab* tab = tabb) {
        ProfScope ps(PC_TRANSPOSE);
        if (int rc = barrier()) return rc;          // every rank has finished producing its pencil
        pull_backward_kernel<<<(unsigned)std::min<long long>(p2p_ctas, (long long)nzl * P), 512, 0, ctx().stream>>>(*tab, a, nxy, nzl, nl, rank, P, accumulate);
        launches++;
        p2p_exchanges++;
        return barrier();                            // the pencils may be overwritten again
    }
    nccl_exchanges++;
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
    if (!t.zstream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&t.zstream, cudaStreamNonBlocking, hi) != cudaSuccess) { cudaGetLastError(); t.zstream = nullptr; }
    }
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
    for (auto& e : t.registry)
        for (int q = 0; q < 8; q++) if (q != t.rank && e.second.p[q]) cudaIpcCloseMemHandle(e.second.p[q]);
    t.registry.clear();
    if (t.zstream) { cudaStreamSynchronize(t.zstream); cudaStreamDestroy(t.zstream); t.zstream = nullptr; }
    if (t.barrier_buf) cudaFree(t.barrier_buf);
    t.barrier_buf = nullptr;
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
