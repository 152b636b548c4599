// z operators of a z-split domain without transposes (replaces TLabMPI_Trp_ExecK_Forward -> OPR_*_1D ->
// TLabMPI_Trp_ExecK_Backward of src/operators/opr_partial.f90:186-249 and src/physics/opr_burgers.f90:387-424 when the
// z direction is periodic and uniform).  See splitz.cu.
#pragma once
#include "context.h"
#include <vector>

namespace tlab {

struct SplitZ {
    long long nxy = 0;       // lines = points of an xy plane
    int kmax = 0, nzg = 0;   // local and global extent in z
    int P = 1, rank = 0;
    int emulate = 0;         // > 1: that many virtual ranks on this device over one full field (tests, P = 1)
    bool ready = false;
    long long calls = 0;     // operator calls so far (parity of the exchange buffers)
    long long ops = 0;
    long long march_ops = 0; // ... of which the finishing phase ran as a march (splitz_march_kernel)
    std::vector<double*> block[2];   // exchange buffers per parity: [0] of this rank (P > 1) or one per virtual rank
    int init(long long nxy, int kmax, int nzg, int P, int rank, int emulate);
    bool eligible(const tlab_plan_s* g, int is) const;      // is < 0: first derivative only
    // out (+)= nu d2s/dz2 - vel ds/dz ;  out (+|-)= d/dz (a + scale*a2)
    int burgers(tlab_plan_s* g, int is, const double* s, const double* vel, double* out, int accumulate);
    int partial(tlab_plan_s* g, const double* a, const double* a2, double scale, double* out, int accumulate);
    void destroy();
};

SplitZ& splitz();

}  // namespace tlab
