// Library-wide state (one GPU per process, like one MPI rank of the reference).
#pragma once
#include "plan.h"
#include "lines.h"
#include "lines2.h"
#include <string>
#include <vector>
#include <cuda_runtime.h>

struct tlab_plan_s {
    tlab::DevPlan p;
    int dir = 0;
    int burgers_first = -1;   // index in p.lu2 of the is = 0 diffusion-scaled LU (-1: tlab_opr_burgers_init not called)
    int burgers_count = 0;
};

namespace tlab {

struct Context {
    bool ready = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool async = false;
    std::string last_error;
    int tune_lines_x = 0, tune_lines_yz = 0;
    int tune_fast = 1;        // 1: use the fast line kernels (lines2.cu) whenever the geometry allows
    int tune_pf_dist = -1;
    int tune_poisson_factors = 1;  // keep the per-mode LU factors of the Poisson y systems (8 more planes) instead of refactorising
    int tune_poisson_il = 0;      // Poisson: per-mode planes that are read together interleaved row by row (set before OPR_Elliptic_Initialize)
    int tune_poisson_split = -1;  // Poisson y solves: one thread per component instead of per mode (-1: when there are few modes)
    int tune_poisson_warp = 4;    // Poisson y solves: team of warps per mode with the lines in registers (ny = 8 T <= 1024); value = kx per CTA (2, 4, 8), 0: off (set before OPR_Elliptic_Initialize)
    int tune_poisson_pf = -1;     // team kernel: prefetch distance in CTAs (-1: two per SM, 0: off)
    int tune_poisson_minb = 3;  // resident CTAs per SM the Poisson y kernel is compiled for (register budget)
    int tune_pf_next = 0;     // fused Burgers launch: L2 prefetch of the tile's next field
    int tune_fuse = 0;        // RHS: one fused Burgers launch per direction (fields sharing the advecting velocity)
    int tune_kxsplit = 1;     // split domain + peer memory: kx-split spectral stage of the Poisson solver
    int tune_overlap = 1;     // split domain: z operators on a second stream, overlapped with the x/y operators
    int tune_pf_l1 = 0;       // strided fast kernels: early L1 prefetch of the operands needed after the solve
    int tune_persist = 0;     // strided fast kernels: persistent CTAs with asynchronous staging
    int tune_fuse_update = 1; // RK substep: `hq2 -= dpdy` and the wall planes of hq2 folded into the update of q2 (2 sweeps less)
    int tune_pair = 0;        // long periodic strided lines (> 32 chunks): one line shared by a cluster of 2 CTAs (128-byte rows).
                              // Off: measured slower at C3 (burgers_z 23.1 vs 21.4 ms), the cluster barriers cost more than the rows gain
    int tune_neu_compact = 1; // BOUNDARY_BCS_NEUMANN_Y: CTAs made of the wall chunks only
    int tune_splitz = 1;      // split domain: z operators on the slabs with halo / chunk-end exchange (splitz.cu) instead of transposes
    int tune_split_emulate = 0;  // P = 1: run the split-z kernels over this many virtual slabs of the field (tests)
    int tune_tma_l2 = 0;      // L2 promotion of the tensor maps of the TMA kernels
    int tune_tma = 0;         // strided fast kernels: persistent CTAs fed and drained by the TMA unit (tensor maps, reduce-add
                              // stores).  Off: measured slower than the LSU kernels, the TMA unit sustains ~20 GB/s per SM on
                              // rows of 32-128 bytes (profiles/ncu_full_tma_r01.json)
    int tune_lazy_scale = 1;  // RK substep: `hq = hq*kco` is not written by the update but folded into the first accumulation of the next
                              // substep (buoyancy source, OPR_Burgers_X): 8 B/pt less per field and substep, same bits
    int tune_pull_overlap = 1; // kx-split Poisson stage: pulls of p^ and dp^/dy on a second stream, beside the inverse transforms
    int tune_march_peel = 1;  // marching kernels of non-periodic directions: constant-only steps for the rounds away from the walls
    int tune_circ = 1;        // periodic directions: circulant form of the fast kernels (constant chunks everywhere, wrapping windows)
    int tune_split_trim = 1;  // split-z marching kernels: phase 1 publishes only the chunk ends phase 2 reads
    int tune_split_local = 0; // timing experiment: phase 1 writes its ends into the own block instead of the peers' (results are wrong)
    int tune_march = 1;       // strided fast kernels: marching panels of 32 lines (march.cu) for OPR_Partial P1 and OPR_Burgers
    int tune_march_cfg = 4;   // marching kernels: CTAs per SM they are compiled for (3 / 4; +10: velocity requested before the barriers)
    int tune_march_pf = 0;
    int tune_march_red = 0;   // marching kernels: accumulate with red.global.add.f64
    long long march_launches = 0;
    long long fast_launches = 0, general_launches = 0;    // L2 prefetch distance of the fast kernels in tiles (-1: automatic, 0: off)
    tlab_plan_s* burgers_plans[3] = {nullptr, nullptr, nullptr};
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
};

// optional per-class timing of the library's launches with CUDA events on the library stream
enum ProfClass { PC_BURGERS_X = 0, PC_BURGERS_Y, PC_BURGERS_Z, PC_PARTIAL_X, PC_PARTIAL_Y, PC_PARTIAL_Z, PC_NEUMANN,
                 PC_FFT, PC_POISSON_Y, PC_ELEMENTWISE, PC_TRANSPOSE, PC_COUNT };
struct ProfScope {
    int idx = -1;
    explicit ProfScope(int cls);
    ~ProfScope();
};

Context& ctx();
int fail(int code, const std::string& msg);
int cuda_check(cudaError_t e, const char* what);
int finish();          // synchronise unless async; maps errors

// launch helpers shared by the operator entry points and the RHS driver (all device pointers)
int run_partial(int dir, int type, int nx, int ny, int nz, int ibc, tlab_plan_s* g, const double* u, double* result,
                double* tmp1, const double* u2 = nullptr, double scale = 0.0, int accumulate = 0);
int run_burgers(int dir, int is, int nx, int ny, int nz, int ibc, tlab_plan_s* g, const double* s, const double* vel,
                double* result, int accumulate, double acc_scale = 1.0);   // accumulate != 0: result = acc_scale*result +|- operator
void scale_array(double* a, double k, long long n, cudaStream_t st);
int run_burgers_multi(int dir, int nf, const int* is, const double* const* sf, const double* vel, double* const* out,
                      int nx, int ny, int nz, tlab_plan_s* g, long long* launches);
int run_neumann_y(int ibc, int nx, int ny, int nz, tlab_plan_s* g, const double* u, double* hb, double* ht);

}  // namespace tlab
