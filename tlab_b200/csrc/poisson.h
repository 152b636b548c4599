#pragma once
#include "context.h"
#include <cufft.h>
#include <vector>

namespace tlab {

// one side of the factorised second-order problem: shared banded tables of  u' + lambda u = f
struct Int1Dev {
    const double* L0 = nullptr;   // [n][5] lambda-independent part of the pentadiagonal system
    const double* L1 = nullptr;   // [n][5] part proportional to lambda
    const double* rhs = nullptr;  // [n][3] tridiagonal right-hand side operator
    double rb[4][4];              // rhs_b(1:3, 0:3) boundary rows at row 1 (final for BCS_MIN)
    double rt[3][5];              // rhs_t(0:2, 1:4) boundary rows at row n (final for BCS_MAX)
    int n = 0, bc = 0;
};

struct PoissonDev {
    int nxh = 0, ny = 0, nz = 0;
    long long nmodes = 0;
    long long plane_sz = 0;           // doubles per scratch / fundamental plane (modes padded to 32)
    int il = 1;                       // planes that are read together are interleaved row by row (see plane() in poisson.cu)
    double norm = 1.0;
    int i_sing0 = 0, i_sing1 = 0, k_sing0 = 0, k_sing1 = 0;
    const double* lambda = nullptr;   // [nmodes]
    Int1Dev smin, smax;
    double* fund = nullptr;           // 5 planes [ny][nmodes]: v1, e-, u1, s+, e+
    double* scr = nullptr;            // 6 planes [ny][nmodes] of per-mode scratch
    double* amat = nullptr;           // 9 x [nmodes]: LU-decomposed 3x3 boundary system
    double* fac = nullptr;            // optional, 8 planes [ny][nmodes]: LU factors (la, lb, 1/c, -d) of the BCS_MIN(+lam) and
                                      // BCS_MAX(-lam) systems of every regular mode, computed once (they depend on lambda only)
};

struct Poisson {
    bool ready = false;
    int nx = 0, ny = 0, nz = 0, nxh = 0;   // nz: local slab thickness
    int nzg = 0, P = 1;                     // global extent in z, ranks in z
    double* c3 = nullptr;                   // pencil work array (P > 1, NCCL path)
    // kx-split spectral stage (P > 1, peer memory): rank p owns the wavenumbers kx0[p] .. kx0[p+1]-1 for all y and z, so
    // that the z transform AND the y solves run in the pencil layout: 3 complex exchanges per call instead of 6
    bool kxsplit = false;
    double *cpa = nullptr, *cpb = nullptr;  // complex pencils (kxl, ny, nzg)
    int kx0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    PoissonDev D;
    cufftHandle plan_fx = 0, plan_bx = 0, plan_z = 0;
    std::vector<void*> allocs;
    int init(tlab_plan_s* gx, tlab_plan_s* gy, tlab_plan_s* gz, int nz_local);
    int solve(double* p, double* c1, double* c2, const double* hb, const double* ht, double* dpdy);
    void release();
};

Poisson& poisson();

}  // namespace tlab
