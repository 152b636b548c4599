// Fast line kernels (lines2.cu): argument block and launcher.
#pragma once
#include "lines.h"

namespace tlab {

struct Line2Args {
    int n = 0, T = 1, L = 4, lshift = 2;
    int accumulate = 0;               // +1: out1 += result, -1: out1 -= result, 0: out1 = result
    int pf_dist = 0;                  // L2 prefetch distance in tiles (0: off)
    int xls = 0;                      // shared-memory line stride of the x tile (doubles)
    int pf_l1 = 0;                    // strided kernel: L1 prefetch of the velocity / accumulation target at kernel start
    int persist = 0;                  // strided kernel: persistent CTAs with cp.async staging
    unsigned ntiles = 0, tiles_x = 0; // persistent kernel: tile count and tiles per outer block
    int neu_nb = 0, neu_nt = 0;       // BOUNDARY_BCS_NEUMANN_Y: CTAs hold only these many chunks next to the bottom / top wall (0, 0: whole lines)
    int pair = 0;                     // strided kernel: a line is shared by a cluster of 2 CTAs (L lines x T/2 chunks each)
    int tma = 0;                      // strided kernel: persistent CTAs fed and drained by the TMA unit (lines2_strided_tma)
    int tma_rb = 0;                   // rows per TMA box
    int march = 0;                    // strided kernel: marching panels of 32 lines (march.cu)
    int march_cfg = 4;                // marching kernel: resident CTAs per SM it is compiled for (+10: velocity requested before the barriers)
    int march_pf = 0;                 // marching kernel: L2 prefetch of the finishing stage's operands at the start of a step
    int march_peel = 0;               // marching kernel, non-periodic direction: rounds 1 .. R-2 hold constant chunks only (peeled steps)
    int march_red = 0;                // marching kernel: accumulate with red.global.add.f64 instead of load + store
    int tma_l2 = 0;                   // L2 promotion of the tensor maps (0 none, 1/2/3: 64/128/256 bytes)
    double scale = 0.0;               // input is u + scale * u2 when u2 != nullptr
    double acc_scale = 1.0;           // contiguous kernel: out1 = acc_scale * out1 +|- result (explicitly rounded product first)
    long long stride = 1;             // distance between consecutive points of a line (strided kernel)
    long long inner = 1;              // tile (bx, by) starts at by * outer_stride + bx * L; lines by * inner + bx * L + l
    long long outer_stride = 0;
    const double* u = nullptr;
    const double* u2 = nullptr;
    const double* vel = nullptr;
    double* out1 = nullptr;
    double* out2 = nullptr;
    double* bcs_hb = nullptr;
    double* bcs_ht = nullptr;
    const double* cjac = nullptr;     // Jacobian correction of the second derivative, d2 -= jscale * cjac * d1 (DevPlan::cjac2)
    RhsTab rhs1, rhs2;
    Sys2 s1, s2;
    // fused Burgers launch (several fields advected by the same velocity): fields, results and which of s2 / s2b each uses
    int nf = 0;
    int pf_next = 0;                  // fused launch: prefetch the next field of the tile into L2
    int fsys[4] = {0, 0, 0, 0};
    const double* fu[4] = {nullptr, nullptr, nullptr, nullptr};
    double* fo[4] = {nullptr, nullptr, nullptr, nullptr};
    Sys2 s2b;
    double neu_bot[BROW_W], neu_top[BROW_W];
    double neu_lu_bot = 0.0, neu_lu_top = 0.0;
};

// can the fast kernels run this geometry?  Returns the number of lines per CTA in *L_out.
bool lines2_eligible(const DevPlan& p, const Sys2& s1, const Sys2* s2, int n, long long nlines, long long inner, bool contig,
                     int L_override, int* L_out);
int lines2_xstride(int T, int L);
size_t lines2_persist_smem(int T, int L);
long long lines2_tma_launches();
bool lines2_tma_eligible(int mode, const Line2Args& a);   // geometry, alignment and shared-memory budget of the TMA kernel
cudaError_t launch_lines2(int mode, const Line2Args& a, bool periodic, bool need1, bool contig, long long nlines,
                          long long inner, cudaStream_t s);

// marching-panel kernels (march.cu)
bool march_sys_ok(const std::vector<double>& crec, int T, int K0, int K1, bool periodic);
bool march_eligible(int mode, const Line2Args& a, bool periodic, bool need1, long long nlines, long long inner);
bool march_peelable(int mode, const Line2Args& a, bool periodic);     // non-periodic: constant-only rounds away from the walls
cudaError_t launch_march(int mode, const Line2Args& a, bool periodic, bool need1, long long nlines, long long inner, cudaStream_t s);

}  // namespace tlab
