#pragma once
#include "context.h"
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
