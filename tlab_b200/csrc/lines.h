#pragma once
#include "plan.h"

namespace tlab {

enum { MODE_P1 = 1, MODE_P2 = 2, MODE_P2_P1 = 3, MODE_BURGERS = 4, MODE_NEUMANN = 5 };

// kernel argument block (passed by value; ~1.5 KB)
struct LineArgs {
    int n = 0, T = 1, cbase = 0, crem = 0;
    int L = 1;                        // lines per CTA
    int xstride = 0;                  // shared-memory line stride (contiguous-line kernel)
    int accumulate = 0;               // +1: out1 += result, -1: out1 -= result, 0: out1 = result
    double scale = 0.0;               // input is u + scale * u2 when u2 != nullptr
    double acc_scale = 1.0;           // accumulate != 0: out1 = acc_scale * out1 +|- result (the pending `hq = hq*kco` of the RK scheme)
    long long nlines = 0;
    long long stride = 1;             // distance between consecutive points of a line
    long long inner = 1;              // line index -> offset: (line / inner) * outer_stride + line % inner
    long long outer_stride = 0;
    const double* u = nullptr;        // field to differentiate (s in the Burgers operator)
    const double* u2 = nullptr;       // optional second input, see scale
    const double* vel = nullptr;      // advecting velocity (Burgers)
    double* out1 = nullptr;           // P1: du, P2: d2u, P2_P1: d2u, BURGERS: nu d2s - vel ds
    double* out2 = nullptr;           // P2_P1: du
    double* bcs_hb = nullptr;         // NEUMANN: boundary planes
    double* bcs_ht = nullptr;
    const double* rhs_d1 = nullptr;
    const double* rhs2_rows = nullptr; // CompactDirect6 second derivative: per-row coefficients [n][5] of the pentadiagonal rhs (MatMul_5d)
    RhsTab rhs1, rhs2;
    SolveTab lu1, lu2;
    double neu_bot[BROW_W], neu_top[BROW_W];
    double neu_lu_bot = 0.0, neu_lu_top = 0.0;
};

int pick_lines_per_cta(int T, bool contig, int override_L);
void set_prefetch(bool on);
int xtile_stride(int n, int L);
cudaError_t launch_lines(int mode, const LineArgs& a, bool periodic, bool need1, bool contig, cudaStream_t s);

}  // namespace tlab
