/* abi_smoke.c -- one Runge-Kutta step through the C ABI alone (include/tlab_gpu.h, no Python, no C++).
 *
 * Catches drift between the header a Fortran/C host binds and the library: this file is compiled by a C compiler
 * against the header only (tests/test_abi_cpu.py: -Wall -Werror, then linked to libtlab_gpu.so) and, on a GPU box,
 * run (tests/test_abi_gpu.py).  The sequence is the reference's own: FDM_CreatePlan x3, OPR_Burgers_Initialize +
 * OPR_Elliptic_Initialize (inside tlab_dns_create, dns_main.f90:103-135), TIME_RUNGEKUTTA (time.f90:185-333),
 * DNS_BOUNDS_CONTROL (dns_local.f90:94-234): a solenoidal Taylor-Green-like field must stay finite and solenoidal.
 * Exit code 0 = ok, 77 = no CUDA device (the library has no CPU fallback), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "tlab_gpu.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != 0) {                                                                    \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, tlab_gpu_last_error());          \
            return (rc_ == TLAB_ERR_CUDA && !device_seen) ? 77 : 1;                        \
        }                                                                                  \
    } while (0)

int main(void) {
    int device_seen = 0;
    const int nx = 32, ny = 33, nz = 32;
    const double pi = 3.14159265358979323846;
    const size_t N = (size_t)nx * ny * nz;
    double *x = malloc(nx * sizeof(double)), *y = malloc(ny * sizeof(double)), *z = malloc(nz * sizeof(double));
    double *q = malloc(3 * N * sizeof(double)), *s = malloc(N * sizeof(double));
    if (!x || !y || !z || !q || !s) return 2;
    for (int i = 0; i < nx; i++) x[i] = 2.0 * pi * i / nx;
    for (int k = 0; k < nz; k++) z[k] = 2.0 * pi * k / nz;
    for (int j = 0; j < ny; j++) {                       /* mildly stretched wall-normal grid on [0, 1] */
        const double t = (double)j / (ny - 1);
        y[j] = t + 0.1 * sin(pi * t) / pi;
    }
    CHECK(tlab_gpu_init(0));
    device_seen = 1;
    tlab_plan_t gx, gy, gz;
    CHECK(tlab_fdm_plan_create(1, nx, x, 1, 1, TLAB_FDM_COM6_JACOBIAN, TLAB_FDM_COM6_JACOBIAN_HYPER, &gx));
    CHECK(tlab_fdm_plan_create(2, ny, y, 0, 0, TLAB_FDM_COM6_JACOBIAN, TLAB_FDM_COM6_JACOBIAN_HYPER, &gy));
    CHECK(tlab_fdm_plan_create(3, nz, z, 1, 1, TLAB_FDM_COM6_JACOBIAN, TLAB_FDM_COM6_JACOBIAN_HYPER, &gz));

    tlab_dns_params prm;
    for (size_t i = 0; i < sizeof(prm); i++) ((char*)&prm)[i] = 0;
    prm.nx = nx; prm.ny = ny; prm.nz = nz;
    prm.nscal = 1;
    prm.rkm_mode = TLAB_RKM_EXP4;
    prm.buoyancy_type = 2;                                /* linear in the first scalar */
    prm.buoyancy_params[0] = 1.0; prm.buoyancy_params[1] = 0.0;
    prm.buoyancy_vector[1] = 1.0;
    prm.visc = 1.0 / 1000.0;
    prm.schmidt[0] = 1.0;
    for (int i = 0; i < 3; i++) { prm.bcs_flow_jmin[i] = TLAB_DNS_BCS_DIRICHLET; prm.bcs_flow_jmax[i] = TLAB_DNS_BCS_DIRICHLET; }
    prm.bcs_scal_jmin[0] = TLAB_DNS_BCS_DIRICHLET; prm.bcs_scal_jmax[0] = TLAB_DNS_BCS_NEUMANN;
    tlab_dns_t dns;
    CHECK(tlab_dns_create(&prm, gx, gy, gz, NULL, &dns));

    /* u = sin x cos z g(y), w = -cos x sin z g(y), v = 0 with g = sin^2(pi y): solenoidal, no-slip walls */
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const size_t p = ((size_t)k * ny + j) * nx + i;
                const double g = sin(pi * y[j]) * sin(pi * y[j]);
                q[p] = 0.1 * sin(x[i]) * cos(z[k]) * g;
                q[N + p] = 0.0;
                q[2 * N + p] = -0.1 * cos(x[i]) * sin(z[k]) * g;
                s[p] = y[j] + 0.01 * sin(x[i]) * g;
            }
    double dil0[2], dil1[2], dt = 1.0, cfl, dif;
    CHECK(tlab_dns_upload_host(dns, "q1", q));
    CHECK(tlab_dns_upload_host(dns, "q2", q + N));
    CHECK(tlab_dns_upload_host(dns, "q3", q + 2 * N));
    CHECK(tlab_dns_upload_host(dns, "s1", s));
    CHECK(tlab_dns_bounds_control(dns, &dil0[0], &dil0[1]));
    CHECK(tlab_time_courant(dns, 1.2, 0.25, 1.0, &dt, &cfl, &dif));
    CHECK(tlab_time_rungekutta_host(dns, dt, q, s));
    CHECK(tlab_dns_bounds_control(dns, &dil1[0], &dil1[1]));
    double sum = 0.0;
    for (size_t p = 0; p < 3 * N; p++) {
        if (!isfinite(q[p])) { fprintf(stderr, "non-finite velocity at %zu\n", p); return 3; }
        sum += q[p] * q[p];
    }
    for (size_t p = 0; p < N; p++)
        if (!isfinite(s[p])) { fprintf(stderr, "non-finite scalar at %zu\n", p); return 3; }
    long long launches = 0;
    CHECK(tlab_dns_launch_count(dns, &launches));
    printf("ABI_SMOKE dt=%.6e cfl=%.4f dif=%.4f dil_before=[%.3e,%.3e] dil_after=[%.3e,%.3e] ke=%.6e launches=%lld\n", dt, cfl, dif,
           dil0[0], dil0[1], dil1[0], dil1[1], 0.5 * sum / N, launches);
    if (!(dt > 0.0 && dt < 1.0) || launches <= 0) return 4;
    if (!(isfinite(dil1[0]) && isfinite(dil1[1]) && fabs(dil1[0]) < 1.0 && fabs(dil1[1]) < 1.0)) { fprintf(stderr, "dilatation out of bounds after the step\n"); return 5; }
    CHECK(tlab_dns_destroy(dns));
    CHECK(tlab_fdm_plan_destroy(gx));
    CHECK(tlab_fdm_plan_destroy(gy));
    CHECK(tlab_fdm_plan_destroy(gz));
    CHECK(tlab_gpu_finalize());
    free(x); free(y); free(z); free(q); free(s);
    printf("ABI_SMOKE_OK\n");
    return 0;
}
