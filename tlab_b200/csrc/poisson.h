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

// Tables of one side in "lane order" for the team-per-mode kernel: the entry of row r = 8 j + q + 1 (chunk j, q-th row of
// the chunk) sits at [q * T + j], so that the lanes of a team (lane = chunk) read consecutive addresses.
struct WarpSide {
    const double2* rh = nullptr;     // {rhs(r, 1), rhs(r, 2)}: interior rows of the tridiagonal right-hand side operator
    const double2* ab0 = nullptr;    // second sub-diagonal {L0(r, 1), L1(r, 1)}
    const double2* ab1 = nullptr;    // first sub-diagonal {L0(r, 2), L1(r, 2)}
    const double2* e = nullptr;      // second super-diagonal {L0(r, 5), L1(r, 5)}
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
    // team-per-mode kernel (ny a multiple of 8, at most 1024): per mode and side the upper LU factors (1/c, -d) of every
    // row as double2 in lane order [mode][q * T + j], the five fundamental lines as [mode][line][q * T + j]
    int T = 0;                        // chunks of 8 rows per line (0: kernel not available for this geometry)
    WarpSide wmin, wmax;
    const double2* wfac_min = nullptr;
    const double2* wfac_max = nullptr;
    const double* wfund = nullptr;
    int pf_dist = 0;                  // team kernel: L2 prefetch of the forcing tile of the CTA this many places later
    double* sing = nullptr;           // small planes (fund, scr layout, 32 slots) for the singular modes
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
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // two-stream way back of the kx-split stage
    std::vector<void*> allocs;
    int init(tlab_plan_s* gx, tlab_plan_s* gy, tlab_plan_s* gz, int nz_local);
    int solve(double* p, double* c1, double* c2, const double* hb, const double* ht, double* dpdy);
    void release();
};

Poisson& poisson();

}  // namespace tlab
