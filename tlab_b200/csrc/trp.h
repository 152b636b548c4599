#pragma once
#include "context.h"
#include <vector>
#include <utility>
#ifdef TLAB_HAVE_NCCL
#include <nccl.h>
#endif

namespace tlab {

// z-slab <-> z-pencil transposes among the P ranks of the z communicator
struct Trp {
    int rank = 0, P = 1;
#ifdef TLAB_HAVE_NCCL
    ncclComm_t comm = nullptr;
#endif
    // peer-memory path: pencil buffers registered through CUDA IPC; the transposes then are one kernel that gathers
    // from the local slab and stores straight into every peer's pencil over NVLink (forward), or loads from every
    // peer's pencil and scatters/accumulates into the local slab (backward), bracketed by stream-ordered barriers
    struct PeerTab { double* p[8]; };
    std::vector<std::pair<const double*, PeerTab>> registry;
    int* barrier_buf = nullptr;
    cudaStream_t zstream = nullptr;    // high-priority stream of the z operators of a split domain (overlap with x/y work)
    bool p2p_enabled = true;
    bool p2p_dma = false;               // peer copies by the copy engines (cudaMemcpy2DAsync) instead of ld/st kernels
    int p2p_ctas = 148;                 // CTAs of the peer-memory kernels: enough to fill NVLink, few enough to leave SMs to overlapped work
    long long p2p_exchanges = 0, nccl_exchanges = 0;
    int register_buffer(double* base);                  // collective: every rank calls it in the same order
    const PeerTab* find(const double* base) const;
    void unregister_buffer(const double* base);          // local: close the peer mappings of one buffer
    int barrier();
    double* sendbuf = nullptr;   // pack / unpack staging
    size_t cap = 0;
    long long launches = 0;
    int ensure(size_t doubles);
    int alltoall(const double* src, double* dst, size_t count);
    // a(nxy, nzl) [+ scale*a2] -> b(nxy/P, nzl*P)
    int forward(const double* a, const double* a2, double scale, double* b, long long nxy, int nzl);
    // b(nxy/P, nzl*P) -> a(nxy, nzl), a = / += / -= according to accumulate (0, +1, -1)
    int backward(const double* b, double* a, long long nxy, int nzl, int accumulate);
};

Trp& trp();

}  // namespace tlab
