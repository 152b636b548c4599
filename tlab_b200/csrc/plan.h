// Device-side plan: the tables the line kernels read (uploaded once per plan).
#pragma once
#include "fdm_host.h"
#include <cuda_runtime.h>
#include <vector>
#include <map>

namespace tlab {

constexpr int CHUNK = 16;        // points of one line owned by one thread (register-resident)
constexpr int MAX_BROWS = 4;     // special rows at each end of a banded right-hand side
constexpr int BROW_W = 8;        // columns a special row may touch (counted from the wall)

// One LU-factored tridiagonal system, rewritten as two first-order linear recurrences
//   forward : y_i = f_i * beta_i + alpha_i * y_{i-1}
//   backward: x_i = (y_i + gamma_i * x_{i+1}) * delta_i            (non-periodic)
//             x_i = (y_i + gamma_i * x_{i+1}) + pe_i * x_N         (circulant)
// plus, for circulant systems, the rank-one closure x_N = (y_N - sum_i pd_i y_i) * bN.
// Rows excluded by a homogeneous Neumann condition have alpha = gamma = 0 and a zero rhs row.
// Per chunk, a descriptor says whether the coefficients are constant over the chunk (interior of a
// uniform grid: the LU factors have converged), in which case the threads use scalars instead of tables.
struct ChunkDesc {
    double a, b, g, d;      // constant-chunk coefficients: alpha, beta, gamma, delta
    double Af, Ab;          // product of the forward / backward multipliers over the chunk
    int flags;              // CD_* bits
    int pad;
};
enum { CD_FWD_CONST = 1, CD_BWD_CONST = 2, CD_PD_ZERO = 4 };

// Coefficient records {alpha, beta, gamma, delta|pe} are stored per (chunk, register slot) and interleaved in
// groups of REC_GROUP chunks: record (t, j) sits at index ((t / G) * CHUNK + j) * G + t % G.  The chunks of a warp
// (tid = l + L*t) then read one 128-byte line per load, at an address that is a per-thread base plus a
// compile-time offset.
constexpr int REC_GROUP = 4;
struct SolveTab {
    const double2* rec = nullptr;    // [T padded][CHUNK] x {alpha, beta}, {gamma, delta|pe}, interleaved
    const double* pd = nullptr;      // same indexing (one double per record), circulant only
    const ChunkDesc* cd = nullptr;   // [T]
    double bN = 0.0;
    int Wf = 0, Wb = 0;              // look-back windows (in chunks), see DESIGN.md "chunked substitution"
};

// The same system in the form used by the fast kernels (lines2.cu): per point
//   forward : y_j = f_j + a_j y_{j-1}
//   backward: x_j = d_j y_j + g_j x_{j+1}  [+ e_j x_N, circulant]
// and, because both sweeps are linear, per chunk (zero-inflow sweeps x^ of the chunk alone)
//   x_j = x^_j + Q_j A + R_j B + S_j x_N
// with A the true y at the end of the previous chunk, B the x_N-free part of the true x at the start of the next chunk,
// x_N = sum_i p_i y_i (circulant closure).  A and B follow from the published chunk ends through look-back /
// look-ahead sums over at most LB2 chunks with precomputed weights (the dropped tail is < 2^-80).
// Chunks whose coefficients are constant (interior of a uniform grid) use the scalars below instead of the tables.
constexpr int LB2 = 6;
struct Sys2 {
    const double2* tab = nullptr;    // point records, 4 double2 per point: {a,p} {d,g} {Q,R} {S,0}; item (t, j, q) at
                                     // (((t>>3)*CHUNK + j)*4 + q)*8 + (t&7)  (the 8 chunks of a warp read one 128-byte line)
    const double* crec = nullptr;    // chunk records, 16 doubles: wf[6], wb[6], Q0, PP, isconst, 0
    double ca = 0.0, cd = 0.0, cg = 0.0;   // constant-chunk coefficients
    double cQ[CHUNK], cR[CHUNK];           // constant-chunk correction vectors
    int K0 = 0, K1 = 0;              // circulant: chunks [0,K0) and [T-K1,T) contribute to x_N
    int K0m = 0, K1m = 0;            // the same with a threshold of 2^-56 instead of 2^-80 (march.cu)
    // circulant form (periodic, hence uniform, direction): the reference's matrix is A0 diag(s) with A0 the constant-coefficient
    // circulant matrix and s_j the Jacobian factor of column j, and A0 factorises EXACTLY into two cyclic first-order recurrences
    // with the converged LU constants (ca, cd, cg).  So every chunk is a constant chunk, the windows wrap around the line, there
    // is no rank-one closure (no tables, no x_N), and the solution is multiplied point by point by rho_j = (1/s_j) / mean(1/s)
    // (the round-off noise of the reference's Jacobian, see plan.cu): wf[k] = (ca^16)^k, wb[k] = (cg^16)^k
    int circ = 0;
    double cwf[LB2], cwb[LB2];
    const double* rho = nullptr;     // item (t, j) at ((t>>3)*CHUNK + j)*8 + (t&7)
    int c_lo = 0, c_hi = -1;         // chunks c_lo .. c_hi are all constant chunks (non-periodic: everything but the chunks at the walls)
    double jscale = 1.0;             // factor of the solution (the diffusivity of a Burgers system): scales the Jacobian correction
    int ok = 0;                      // 0: look-back window too long for the fast kernels
    int march_ok = 0;                // 1: a window of 3 chunks suffices (dropped weights < 2^-64) and the closure chunks fit one round
                                     // of the marching kernels (march.cu)
};

// banded right-hand side B u: constant interior stencil + dense special rows at the ends
struct RhsTab {
    double rc = 0.0;                 // centre coefficient (symmetric stencils; 0 for antisymmetric)
    double r2 = 0.0, r3 = 0.0;       // 2nd and 3rd off-diagonals (1st off-diagonal is 1 by normalisation)
    int nb = 0;                      // # of special rows at each end (0: periodic)
    double bot[MAX_BROWS][BROW_W];   // f_i     = sum_k bot[i][k] * u_k            (i, k 0-based from the bottom)
    double top[MAX_BROWS][BROW_W];   // f_{n-1-q} = sum_k top[q][k] * u_{n-1-k}     (q, k 0-based from the top)
};

struct DevPlan {
    HostPlan h;
    int n = 0;
    int T = 1;                       // chunks per line
    int cbase = 0, crem = 0;         // chunk t starts at t*cbase + min(t, crem), has cbase + (t < crem) points
    bool periodic = false;
    bool need_1der = false;
    RhsTab rhs1[4];                  // first derivative, per ibc (only [0] for periodic)
    RhsTab rhs2;                     // second derivative
    SolveTab lu1[4];                 // first derivative LU per ibc
    std::vector<SolveTab> lu2;       // second derivative LU: [0] plain, [1 + is] scaled by diffusivity is (Burgers)
    Sys2 sys1[4];                    // the same systems in the form of lines2.cu
    std::vector<Sys2> sys2;
    const double* rhs_d1 = nullptr;  // [n][3] Jacobian correction of the second derivative (need_1der)
    const double* rhs2_rows = nullptr;  // [n][5] per-row rhs of a CompactDirect6 second derivative (general kernels only)
    const double* cjac2 = nullptr;   // lines2.cu: c_j = dx2_j / dx1_j^2, item (t, j) at ((t>>3)*CHUNK + j)*8 + (t&7).  The Jacobian term of
                                     // the second derivative is a diagonal correction of the solution: with lhs = A0 diag(dx1^2) and
                                     // rhs_d1 = -A0 diag(dx2) (fdm_com2_jacobian.f90:263-274; the extended-stencil entry of a tridiagonal
                                     // lhs is zero), A (d2u) = B u + rhs_d1 du  <=>  d2u = A^-1 B u - c du
    const double* d_mwn1 = nullptr;  // [n] modified wavenumbers of the first derivative (periodic)
    const double* d_jac = nullptr;   // [n] dx/ds
    // Neumann boundary-value closure (BOUNDARY_BCS_NEUMANN_Y): value = sum_k bcsrow[k] u_k + lu_coef * du_1
    double neu_bot[4][BROW_W], neu_top[4][BROW_W];
    double neu_lu_bot[4], neu_lu_top[4];
    std::vector<void*> allocs;       // device allocations owned by the plan
};

int devplan_build(DevPlan& p);                       // upload tables of p.h
int devplan_add_diffusion(DevPlan& p, double diff);  // append a diffusivity-scaled second-derivative LU; returns index or <0
void devplan_free(DevPlan& p);

inline int chunk_start(const DevPlan& p, int t) { return t * p.cbase + (t < p.crem ? t : p.crem); }

}  // namespace tlab
