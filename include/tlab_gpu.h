/* tlab_gpu.h -- C ABI of the B200-native tlab hot path (libtlab_gpu.so).
 *
 * Drop-in boundary for tlab's incompressible/Boussinesq right-hand side and Runge-Kutta substep.
 * Every entry point names the reference (turbulencia/tlab, Fortran) interface it replaces as
 * file:line under src/.  The Fortran host binds these through iso_c_binding (fortran/tlab_gpu_mod.f90,
 * INTEGRATION.md).  Conventions:
 *   - all functions return 0 on success or a tlab error code (src/include/dns_error.h numbering;
 *     TLAB_ERR_CUDA = 200 for CUDA/cuFFT/NCCL failures, text in tlab_gpu_last_error());
 *   - unless a name ends in _host, array arguments are DEVICE pointers (tlab_gpu_malloc) holding
 *     double precision data in the reference's layout a(nx,ny,nz), x fastest;
 *   - bcs is the Fortran bcs(2,2) in column-major order: {bcs(1,1), bcs(2,1), bcs(1,2), bcs(2,2)};
 *   - one caller thread per process/GPU (the reference is not re-entrant either, opr_elliptic.f90:64-81);
 *     calls are stream-ordered on the library stream and synchronous at return unless
 *     tlab_gpu_set_async(1) was called.
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef TLAB_GPU_H
#define TLAB_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLAB_ERR_DIMGRID 48      /* DNS_ERROR_DIMGRID   */
#define TLAB_ERR_PARPARTITION 45 /* DNS_ERROR_PARPARTITION */
#define TLAB_ERR_ALLOC 80        /* DNS_ERROR_ALLOC     */
#define TLAB_ERR_OPTION 85       /* DNS_ERROR_OPTION    */
#define TLAB_ERR_UNDEVELOP 104   /* DNS_ERROR_UNDEVELOP */
#define TLAB_ERR_CUDA 200

/* operator codes, src/operators/opr_partial.f90:19-21 and src/physics/opr_burgers.f90 (OPR_B_SELF/U_IN) */
#define TLAB_OPR_P1 1
#define TLAB_OPR_P2 2
#define TLAB_OPR_P2_P1 3
#define TLAB_OPR_B_SELF 0
#define TLAB_OPR_B_U_IN 1
/* boundary-condition codes, src/base/tlab_constants.f90:62-71 */
#define TLAB_BCS_DD 0
#define TLAB_BCS_ND 1
#define TLAB_BCS_DN 2
#define TLAB_BCS_NN 3
/* scheme codes, src/fdm/fdm_derivative.f90:51-58 */
#define TLAB_FDM_COM4_JACOBIAN 4
#define TLAB_FDM_COM6_JACOBIAN 6
#define TLAB_FDM_COM6_JACOBIAN_HYPER 7
#define TLAB_FDM_COM6_DIRECT 16      /* second derivative only (SpaceOrder2 = CompactDirect6, src/fdm/fdm_comx_direct.f90:305-412) */
/* DNS_BCS_* of src/tools/dns/boundary_bcs.f90:42-46 */
#define TLAB_DNS_BCS_DIRICHLET 3
#define TLAB_DNS_BCS_NEUMANN 4
/* time.f90 RKM_EXP3 / RKM_EXP4 */
#define TLAB_RKM_EXP3 3
#define TLAB_RKM_EXP4 4

typedef struct tlab_plan_s* tlab_plan_t; /* one direction: type(fdm_dt), src/fdm/fdm.f90:14-29 */
typedef struct tlab_dns_s* tlab_dns_t;   /* module state of tools/dns: q, s, hq, hs, txc, parameters */

#define TLAB_MAX_SCAL 8
/* parameters of the incompressible/Boussinesq path that the reference keeps in module variables
 * (TLab_Memory, NavierStokes, Gravity, BOUNDARY_BCS, DNS_LOCAL, TIME) */
typedef struct {
    int nx, ny, nz;                   /* imax, jmax, kmax */
    int nscal;                        /* inb_scal */
    int rkm_mode;                     /* TLAB_RKM_EXP3 | TLAB_RKM_EXP4 */
    int buoyancy_type;                /* 0 none, 1 EQNS_BOD_HOMOGENEOUS, 2 EQNS_BOD_LINEAR (first scalar) */
    int scal_limit;                   /* [Control] ScalLimit */
    int bcs_flow_jmin[3], bcs_flow_jmax[3];                         /* TLAB_DNS_BCS_DIRICHLET | _NEUMANN */
    int bcs_scal_jmin[TLAB_MAX_SCAL], bcs_scal_jmax[TLAB_MAX_SCAL];
    double visc;                      /* 1/Reynolds */
    double schmidt[TLAB_MAX_SCAL];
    double buoyancy_params[2];        /* linear: {c1, c0}; homogeneous: {b, -} */
    double buoyancy_vector[3];        /* [Gravity] Vector / Froude */
    double scal_min[TLAB_MAX_SCAL], scal_max[TLAB_MAX_SCAL];
} tlab_dns_params;

/* ---- runtime ------------------------------------------------------------------------------- */
int tlab_gpu_init(int device);                 /* TLab_Start, src/base/tlab_workflow.f90:36-101 (device part) */
int tlab_gpu_finalize(void);                   /* TLab_Stop */
const char* tlab_gpu_last_error(void);         /* text of the last failure (TLab_Write_ASCII(efile, ...)) */
int tlab_gpu_set_async(int on);
int tlab_gpu_synchronize(void);
int tlab_gpu_malloc(void** ptr, size_t bytes); /* TLab_Allocate_Real, src/base/tlab_memory.f90:164-216 */
int tlab_gpu_free(void* ptr);
/* Unified (managed) memory for a Fortran host that keeps q, s, txc as ordinary arrays (TLab_Allocate_Real with
 * c_f_pointer on this pointer): the same address is valid on the host and on the device, so the signature-exact Fortran
 * wrappers (fortran/tlab_gpu_mod.f90) pass c_loc(array) straight through.  tlab_gpu_prefetch moves the pages to the device
 * (to_device != 0) or back to the host ahead of use; without it they migrate on first touch. */
int tlab_gpu_malloc_managed(void** ptr, size_t bytes);
int tlab_gpu_prefetch(const void* ptr, size_t bytes, int to_device);
int tlab_gpu_upload(void* dst_device, const void* src_host, size_t bytes);
int tlab_gpu_download(void* dst_host, const void* src_device, size_t bytes);
int tlab_gpu_copy(void* dst_device, const void* src_device, size_t bytes);
/* Tuning keys (all optional; the defaults are what bench.py measures; every variant is an alternative schedule of the same
 * arithmetic and is tested against the oracle, tests/test_dns_gpu.py):
 *   "lines_x", "lines_yz"   lines per CTA of the line kernels (0 = automatic)
 *   "fast"                  1/0: the fast line kernels (lines2.cu, march.cu) where the geometry allows (default 1)
 *   "pf_dist"               their L2 prefetch distance in tiles (-1 automatic, 0 off)
 *   "circ"                  1/0: periodic directions in circulant form (default 1) / with the rank-one closure of TRIDPFS
 *   "march"                 marching-panel kernels: 1 = periodic directions, lines longer than 32 chunks and OPR_Burgers along
 *                           non-periodic directions whose interior chunks are constant (default), 2 = wherever eligible, 0 = off
 *   "march_peel"            1/0: constant-only steps for the interior rounds of a non-periodic march (default 1)
 *   "march_cfg", "march_red", "march_pf"   register budget / red.global.add accumulation / L2 prefetch of the marching kernels
 *   "fuse", "pf_next", "persist", "pair", "tma", "tma_l2", "neu_compact", "fuse_update", "lazy_scale"   see DESIGN.md section 4
 *   "splitz"                1/0: z operators of a z-split domain on the slabs (default 1) / through K-transposes
 *   "split_trim", "split_emulate", "split_local"   phase 1 publishes only what phase 2 reads (default 1); virtual slabs on one
 *                           GPU (tests); timing experiment (wrong results)
 *   "poisson_split", "poisson_il", "poisson_minb", "pull_overlap"   Poisson stage: kx-split exchange, table interleave, CTAs per
 *                           SM of the y solves, two-stream way back
 * Unknown keys return TLAB_ERR_OPTION. */
int tlab_gpu_set_tuning(const char* key, int value);
/* "tma": 1/0 run the y/z line operators as persistent CTAs fed and drained by the TMA unit (default 0: measured slower on rows
 * of 32-128 bytes; falls back to the LSU kernels when a geometry is not eligible).
 * launch counters: "fast_launches", "general_launches", "tma_launches" (line kernels since start-up) */
int tlab_gpu_get_counter(const char* key, long long* value);

/* the CUDA stream (cudaStream_t) every call is ordered on, for event timing by the host */
int tlab_gpu_stream(void** stream);
/* per-class device timing of this library's launches (USE_PROFILE of the reference, time.f90:195-198,311-329):
 * classes 0-2 Burgers x/y/z, 3-5 Partial x/y/z, 6 Neumann BCs, 7 FFTs, 8 Poisson y solves, 9 element-wise,
 * 10 transposes.  The report synchronises, fills 11 entries and clears the records. */
#define TLAB_PROF_CLASSES 11
int tlab_gpu_profile(int on);
int tlab_gpu_profile_report(double* ms_per_class, int* count_per_class, int nclass);

/* ---- plans ---------------------------------------------------------------------------------- */
/* FDM_CreatePlan(x, g), src/fdm/fdm.f90:143-252: builds Jacobians, scheme tables, Neumann reductions
 * and LU factors from the node positions and uploads them.  dir = 1,2,3 (x,y,z) is informative. */
int tlab_fdm_plan_create(int dir, int n, const double* nodes_host, int periodic, int uniform,
                         int mode_der1, int mode_der2, tlab_plan_t* out);
/* same tables on the host only (no device needed): the handle is valid for tlab_fdm_plan_get/destroy only */
int tlab_fdm_plan_create_host(int dir, int n, const double* nodes_host, int periodic, int uniform,
                              int mode_der1, int mode_der2, tlab_plan_t* out);
int tlab_fdm_plan_destroy(tlab_plan_t plan);
/* read back host tables of a plan (for hosts that want g%jac, g%der1%mwn, ... from this library):
 * what = "nodes" | "jac1" | "jac2" | "jac3" | "mwn1" | "mwn2" | "lhs1" | "rhs1" | "lu1" | "lhs2" | "rhs2" | "lu2"
 * | "rhs1_b" | "rhs1_t"; matrices are returned column-major (Fortran order); *count receives the length. */
int tlab_fdm_plan_get(tlab_plan_t plan, const char* what, double* out_host, int capacity, int* count);

/* FDM_Int1_CreateSystem(x, g, lambda, ibc, fdmi), src/fdm/fdm_integral.f90:91-214, on the host for one
 * eigenvalue (ibc = 1 BCS_MIN, 2 BCS_MAX).  Outputs, column-major: lhs(n,5) before LU, rhs(n,3),
 * rhs_b(1:5,0:7), rhs_t(0:4,1:8). */
int tlab_fdm_int1_system_host(tlab_plan_t plan, int ibc, double lambda, double* lhs, double* rhs, double* rhs_b,
                              double* rhs_t);

/* ---- operators ------------------------------------------------------------------------------ */
/* OPR_Partial_X/Y/Z(type, nx, ny, nz, bcs, g, u, result, tmp1), src/operators/opr_partial.f90:31-377 */
int tlab_opr_partial(int dir, int type, int nx, int ny, int nz, const int bcs[4], tlab_plan_t g,
                     const double* u, double* result, double* tmp1_or_null);

/* OPR_Burgers_Initialize, src/physics/opr_burgers.f90:52-115: diffusivity-scaled LU of the second
 * derivative for is = 0 (visc) and is = 1..nscal (visc/schmidt(is)) on the three plans */
int tlab_opr_burgers_init(tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz, double visc, int nscal,
                          const double* schmidt_host);
/* OPR_Burgers_X/Y/Z(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t), opr_burgers.f90:190-431.
 * result = diffusivity(is) * d2s - u * ds along dir.  tmp1 and u_t only carry the reference's transposed
 * copy of the velocity; no transposed copies exist here, they are accepted and ignored. */
int tlab_opr_burgers(int dir, int ivel, int is, int nx, int ny, int nz, const int bcs[4], const double* s,
                     const double* u, double* result, double* tmp1_ignored, const double* u_t_ignored);

/* FDM_Der1_Solve / FDM_Der2_Solve on the reference's lines-first view u(nlines, n),
 * src/fdm/fdm_derivative.f90:218-278, 413-459 */
int tlab_fdm_der1_solve(tlab_plan_t g, int nlines, int ibc, const double* u, double* result);
int tlab_fdm_der2_solve(tlab_plan_t g, int nlines, int is_or_minus1, const double* u, const double* du_ignored,
                        double* result);

/* thomas3 / thomas5 substitution stages on device arrays in the reference's lines-first layout f(len, nmax);
 * a..e are the factored diagonals produced by TRIDFS / TRIDPFS / PENTADFS / PENTADFS2:
 * TRIDSS(nmax, len, a, b, c, f), src/utils/linear3.f90:56-150; TRIDPSS(nmax, len, a, b, c, d, e, f, wrk), :321-442;
 * PENTADSS(nmax, len, a, b, c, d, e, f), src/utils/linear5.f90:76-131; PENTADSS2, :209-244 */
int tlab_tridss(int nmax, int len, const double* a, const double* b, const double* c, double* f);
int tlab_tridpss(int nmax, int len, const double* a, const double* b, const double* c, const double* d,
                 const double* e, double* f, double* wrk_or_null);
int tlab_pentadss(int nmax, int len, const double* a, const double* b, const double* c, const double* d,
                  const double* e, double* f);
int tlab_pentadss2(int nmax, int len, const double* a, const double* b, const double* c, const double* d,
                   const double* e, double* f);
/* TLab_Transpose(a, nra, nca, ma, b, mb) and _COMPLEX, src/utils/tlab_transpose.f90:14-82,148-210: b(k,j) = a(j,k) */
int tlab_transpose(const double* a, int nra, int nca, int ma, double* b, int mb);
int tlab_transpose_complex(const double* a, int nra, int nca, int ma, double* b, int mb);

/* BOUNDARY_BCS_NEUMANN_Y(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht, tmp1), src/tools/dns/boundary_bcs.f90:368-473 */
int tlab_boundary_bcs_neumann_y(int ibc, int nx, int ny, int nz, tlab_plan_t gy, const double* u, double* bcs_hb,
                                double* bcs_ht);

/* OPR_Elliptic_Initialize, src/operators/opr_elliptic.f90:86-250 (FourierXZ_Factorize): eigenvalues from the
 * x/z modified wavenumbers, integral-operator tables in y, cuFFT plans, fundamental solutions per mode.
 * kmax_local: thickness of this rank's z slab (0 or gz's size when the domain is not split) */
int tlab_opr_elliptic_init(tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz, int kmax_local);
/* OPR_Poisson(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy), opr_elliptic.f90:263-364.
 * p: forcing in, solution out; tmp1, tmp2: work arrays of (nx+2)*ny*nz doubles (the reference's
 * isize_txc_field); dpdy may be NULL.  Only ibc = BCS_NN. */
int tlab_opr_poisson(int nx, int ny, int nz, int ibc, double* p, double* tmp1, double* tmp2,
                     const double* bcs_hb, const double* bcs_ht, double* dpdy_or_null);

/* OPR_Fourier_X_Forward(nx, ny, nz, in, out) / _Backward, src/operators/opr_fourier.f90:219-329: real-to-complex transform
 * along x of ny*nz lines; the half spectrum is c(nx/2+1, ny, nz) (interleaved re, im; element nx/2+1 = Nyquist), i.e.
 * (nx+2)*ny*nz doubles = the reference's isize_txc_field.  Unnormalised like FFTW (backward(forward(a)) = nx * a); the
 * backward transform may overwrite `in`.  OPR_Fourier_Z_Forward(in, out) / _Backward, opr_fourier.f90:333-433: complex transform
 * along z of the (nx/2+1)*ny lines of that array (stride (nx/2+1)*ny), in-place allowed; the reference takes the sizes from
 * module variables, here they are arguments.  cuFFT, as north_star specifies.  Single domain only. */
int tlab_opr_fourier_x_forward(int nx, int ny, int nz, const double* in, double* out);
int tlab_opr_fourier_x_backward(int nx, int ny, int nz, double* in, double* out);
int tlab_opr_fourier_z_forward(int nx, int ny, int nz, double* in, double* out);
int tlab_opr_fourier_z_backward(int nx, int ny, int nz, double* in, double* out);

/* ---- domain decomposition (z slabs, one process per GPU) -------------------------------------- */
/* TLabMPI_Initialize + TLabMPI_Trp_Initialize, src/base/tlab_mpi_procs.f90:17-116, tlab_mpi_transpose.f90:68-200,
 * for ims_npro_k = nranks, ims_npro_i = 1.  Rank 0 obtains a 128-byte NCCL id, the host broadcasts it
 * (MPI_Bcast / torch.distributed), every rank calls tlab_mpi_init before creating plans that use it. */
int tlab_mpi_get_unique_id(void* id_out_128);
int tlab_mpi_init(int rank, int nranks, const void* id_128);
int tlab_mpi_finalize(void);
int tlab_mpi_rank(int* rank, int* nranks);
/* TLabMPI_Trp_ExecK_Forward / _Backward (real or complex), tlab_mpi_transpose.f90:343-553: slab a(nlines_total, kmax)
 * <-> pencil b(nlines_total/P, kmax*P); a and b are distinct device arrays of nlines_total*kmax elements */
int tlab_trp_exec_k_forward(const double* a, double* b, int nlines_total, int kmax, int is_complex);
int tlab_trp_exec_k_backward(const double* b, double* a, int nlines_total, int kmax, int is_complex);

/* ---- time advance --------------------------------------------------------------------------- */
/* Device-resident state: allocates q(3), s(nscal), hq, hs and work arrays (TLab_Initialize_Memory,
 * src/base/tlab_memory.f90:164-216; dns_main.f90:103-104), calls OPR_Burgers_Initialize and
 * OPR_Elliptic_Initialize.  bbackground_host: reference buoyancy profile (ny) or NULL for zero. */
int tlab_dns_create(const tlab_dns_params* prm, tlab_plan_t gx, tlab_plan_t gy, tlab_plan_t gz,
                    const double* bbackground_host, tlab_dns_t* out);
int tlab_dns_destroy(tlab_dns_t dns);
/* device pointer of "q1".."q3", "s1"..,"hq1".., "hs1".., "p" (pressure of the last RHS), "dpdy" */
int tlab_dns_field(tlab_dns_t dns, const char* name, double** dev_ptr);
int tlab_dns_upload_host(tlab_dns_t dns, const char* name, const double* src_host);
int tlab_dns_download_host(tlab_dns_t dns, const char* name, double* dst_host);
int tlab_dns_launch_count(tlab_dns_t dns, long long* count); /* kernels of this library launched so far */
/* TIME_INITIALIZE coefficients, src/tools/dns/time.f90:86-112 */
int tlab_time_rk_coefficients(int rkm_mode, double* kdt, double* ktime, double* kco, int* nsub);
/* RHS_GLOBAL_INCOMPRESSIBLE_1, src/tools/dns/rhs_global_incompressible_1.f90:15-405: accumulates into hq, hs */
int tlab_rhs_global_incompressible_1(tlab_dns_t dns, double dte);
/* one Runge-Kutta stage: TLab_Sources_Flow + RHS + q += dte*hq (TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT,
 * time.f90:559-670), DNS_BOUNDS_LIMIT, and hq *= kco when scale_h != 0 (time.f90:261-298) */
int tlab_time_substep(tlab_dns_t dns, double dte, double kco, int scale_h);
/* stage `stage` (0-based) of TIME_RUNGEKUTTA's loop: zeroes hq, hs when stage == 0 (time.f90:212-216), then
 * tlab_time_substep with dte = dtime*kdt(stage) and kco(stage) */
int tlab_time_rungekutta_stage(tlab_dns_t dns, double dtime, int stage);
/* TIME_RUNGEKUTTA, time.f90:185-333: hq = hs = 0, then all stages with dte = dtime*kdt(s) */
int tlab_time_rungekutta(tlab_dns_t dns, double dtime);
/* TIME_COURANT, time.f90:365-548 (incompressible branch): dt = min(cfla / max(|u|/dx+|v|/dy+|w|/dz),
 * cfld / (max(1, 1/Pr, 1/min Sc) * visc * max sum 1/d^2)) with d = g%jac(:,1); *dtime is left untouched when cfla <= 0.
 * Also returns the reference's logged CFL and diffusion numbers (dns.out columns). */
int tlab_time_courant(tlab_dns_t dns, double cfla, double cfld, double prandtl, double* dtime, double* cfl_number,
                      double* diffusion_number);
/* dilatation bounds of DNS_BOUNDS_CONTROL, src/tools/dns/dns_local.f90:94-234, via FI_INVARIANT_P
 * (src/mappings/fi_vectorcalculus.f90:111-141): minimum and maximum of div u (dns.out DilMin, DilMax) */
int tlab_dns_bounds_control(tlab_dns_t dns, double* dil_min, double* dil_max);
/* same, starting from and returning to HOST arrays q(N,3), s(N,nscal) (pinned memory recommended) */
int tlab_time_rungekutta_host(tlab_dns_t dns, double dtime, double* q_host, double* s_host);

#ifdef __cplusplus
}
#endif
#endif
