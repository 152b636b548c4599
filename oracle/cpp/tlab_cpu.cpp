// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing under tlab_b200/ may link, load or call this file.
//
// CPU baseline of the incompressible/Boussinesq RK substep: a C++17 + OpenMP restatement of the reference's algorithm
// (what `--impl reference` and the `cpu_baseline` leg of bench.py time on the host cores; the image has no Fortran compiler,
// so the reference itself cannot be built: DESIGN.md section 2).  It follows the numpy oracle routine by routine and is checked
// against it (tests/test_cpu_baseline.py, <= 1e-12); the plan tables (scheme coefficients, LU factors, integral-operator
// systems per eigenvalue) are produced by the oracle in Python and handed over as plain arrays.
//
//   banded right-hand sides      src/fdm/fdm_matmul.f90:70-642   (MatMul_5d_antisym, MatMul_7d_sym, MatMul_3d, MatMul_3d_add)
//   Thomas substitution stages   src/utils/linear3.f90:56-150 (TRIDSS), :321-442 (TRIDPSS); src/utils/linear5.f90:76-131 (PENTADSS)
//   OPR_Partial / OPR_Burgers    src/operators/opr_partial.f90:31-377, src/physics/opr_burgers.f90:190-521
//   FDM_Int1_Solve               src/fdm/fdm_integral.f90:219-314
//   OPR_ODE2_Factorize_NN[_Sing] src/operators/opr_odes.f90:37-96,165-183,265-386
//
// Schedule.  The reference transposes a field so that the lines of a direction are the slow index, then sweeps row by row over
// all lines (vector loops over `len`, OpenMP over slices of `len`).  Here a thread takes a block of LB adjacent lines, gathers
// it into a cache-resident buffer [n][LB] (which is the transposition, done in cache), runs the same row-by-row sweeps on it and
// scatters the result: the same arithmetic per line, two passes over memory less per operator.  That favours the CPU, which is
// the conservative side for a baseline.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

constexpr int LB = 32;       // lines per block
constexpr int BW = 8;        // columns a dense boundary row may touch

}  // namespace

extern "C" {

// banded right-hand side: interior stencil with constants, dense rows at the ends (non-periodic)
struct Band {
    int n, periodic, sym, nb;
    double rc, r2, r3;
    double bot[4][BW];       // f[i]       = sum_k bot[i][k] u[k]
    double top[4][BW];       // f[n-1-q]   = sum_k top[q][k] u[n-1-k]
};
// factored tridiagonal system: TRIDFS (a, b, c) on rows nmin..nmax (1-based), or TRIDPFS (a, b, c, d, e)
struct Tri {
    int n, periodic, nmin, nmax;
    const double *a, *b, *c, *d, *e;
};

}  // extern "C"

namespace {

inline int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

void band_apply(const Band& B, const double* __restrict__ u, double* __restrict__ f, int W) {
    const int n = B.n;
    auto row = [&](int i, int im1, int ip1, int im2, int ip2, int im3, int ip3) {
        const double *u0 = u + (size_t)i * W, *a1 = u + (size_t)ip1 * W, *b1 = u + (size_t)im1 * W, *a2 = u + (size_t)ip2 * W,
                     *b2 = u + (size_t)im2 * W, *a3 = u + (size_t)ip3 * W, *b3 = u + (size_t)im3 * W;
        double* o = f + (size_t)i * W;
        if (B.sym) {
            if (B.r3 != 0.0) for (int w = 0; w < W; w++) o[w] = B.rc * u0[w] + a1[w] + b1[w] + B.r2 * (a2[w] + b2[w]) + B.r3 * (a3[w] + b3[w]);
            else for (int w = 0; w < W; w++) o[w] = B.rc * u0[w] + a1[w] + b1[w] + B.r2 * (a2[w] + b2[w]);
        } else {
            if (B.r3 != 0.0) for (int w = 0; w < W; w++) o[w] = a1[w] - b1[w] + B.r2 * (a2[w] - b2[w]) + B.r3 * (a3[w] - b3[w]);
            else if (B.r2 != 0.0) for (int w = 0; w < W; w++) o[w] = a1[w] - b1[w] + B.r2 * (a2[w] - b2[w]);
            else for (int w = 0; w < W; w++) o[w] = a1[w] - b1[w];
        }
    };
    if (B.periodic) {
        for (int i = 0; i < n; i++) row(i, wrap(i - 1, n), wrap(i + 1, n), wrap(i - 2, n), wrap(i + 2, n), wrap(i - 3, n), wrap(i + 3, n));
        return;
    }
    for (int i = B.nb; i < n - B.nb; i++) row(i, i - 1, i + 1, i - 2, i + 2, std::max(i - 3, 0), std::min(i + 3, n - 1));
    for (int i = 0; i < B.nb; i++) {
        double* o = f + (size_t)i * W;
        for (int w = 0; w < W; w++) o[w] = 0.0;
        for (int k = 0; k < BW && k < n; k++) {
            const double c = B.bot[i][k];
            if (c == 0.0) continue;
            const double* s = u + (size_t)k * W;
            for (int w = 0; w < W; w++) o[w] += c * s[w];
        }
        double* ot = f + (size_t)(n - 1 - i) * W;
        for (int w = 0; w < W; w++) ot[w] = 0.0;
        for (int k = 0; k < BW && k < n; k++) {
            const double c = B.top[i][k];
            if (c == 0.0) continue;
            const double* s = u + (size_t)(n - 1 - k) * W;
            for (int w = 0; w < W; w++) ot[w] += c * s[w];
        }
    }
}

// MatMul_3d_add (fdm_matmul.f90:126-153): f += rhs_d1 * du, rows r[i] = {r1, r2, r3}; extended stencils in the first and last row
void jac_add(const double* __restrict__ r, int n, const double* __restrict__ du, double* __restrict__ f, int W) {
    for (int w = 0; w < W; w++) f[w] += du[w] * r[1] + du[W + w] * r[2] + du[2 * W + w] * r[0];
    for (int i = 1; i < n - 1; i++) {
        const double r1 = r[3 * i], r2 = r[3 * i + 1], r3 = r[3 * i + 2];
        const double *m = du + (size_t)(i - 1) * W, *c = du + (size_t)i * W, *p = du + (size_t)(i + 1) * W;
        double* o = f + (size_t)i * W;
        for (int w = 0; w < W; w++) o[w] += m[w] * r1 + c[w] * r2 + p[w] * r3;
    }
    const int i = n - 1;
    double* o = f + (size_t)i * W;
    for (int w = 0; w < W; w++)
        o[w] += du[(size_t)(i - 2) * W + w] * r[3 * i + 2] + du[(size_t)(i - 1) * W + w] * r[3 * i] + du[(size_t)i * W + w] * r[3 * i + 1];
}

// TRIDSS (linear3.f90:56-150) on rows nmin..nmax; TRIDPSS (:321-442)
void tri_solve(const Tri& T, double* __restrict__ f, int W, double* __restrict__ wrk) {
    if (!T.periodic) {
        const int i0 = T.nmin - 1, i1 = T.nmax - 1;           // 0-based first and last active row
        const double *a = T.a, *b = T.b, *c = T.c;            // indexed by active row number 0..m-1
        const int m = i1 - i0 + 1;
        double* g = f + (size_t)i0 * W;
        for (int k = 1; k < m; k++) {
            const double ak = a[k];
            double* o = g + (size_t)k * W;
            const double* p = o - W;
            for (int w = 0; w < W; w++) o[w] = o[w] + ak * p[w];
        }
        {
            double* o = g + (size_t)(m - 1) * W;
            const double bk = b[m - 1];
            for (int w = 0; w < W; w++) o[w] = o[w] * bk;
        }
        for (int k = m - 2; k >= 0; k--) {
            const double bk = b[k], ck = c[k];
            double* o = g + (size_t)k * W;
            const double* p = o + W;
            for (int w = 0; w < W; w++) o[w] = (o[w] + ck * p[w]) * bk;
        }
        return;
    }
    const int n = T.n;
    const double *a = T.a, *b = T.b, *c = T.c, *d = T.d, *e = T.e;
    for (int w = 0; w < W; w++) f[w] = f[w] * b[0];
    for (int k = 1; k < n - 1; k++) {
        double* o = f + (size_t)k * W;
        const double* p = o - W;
        const double bk = b[k], ak = a[k];
        for (int w = 0; w < W; w++) o[w] = o[w] * bk + ak * p[w];
    }
    for (int w = 0; w < W; w++) wrk[w] = 0.0;
    for (int k = 0; k < n - 1; k++) {
        const double dk = d[k];
        const double* p = f + (size_t)k * W;
        for (int w = 0; w < W; w++) wrk[w] = wrk[w] + dk * p[w];
    }
    double* last = f + (size_t)(n - 1) * W;
    for (int w = 0; w < W; w++) last[w] = (last[w] - wrk[w]) * b[n - 1];
    {
        double* o = f + (size_t)(n - 2) * W;
        const double ek = e[n - 2];
        for (int w = 0; w < W; w++) o[w] = ek * last[w] + o[w];
    }
    for (int k = n - 3; k >= 0; k--) {
        double* o = f + (size_t)k * W;
        const double* p = o + W;
        const double ck = c[k], ek = e[k];
        for (int w = 0; w < W; w++) o[w] = o[w] + ck * p[w] + ek * last[w];
    }
}

struct Geo { long long nblocks; int n; };

// block b of LB lines along dir: element (i, w) of the block lives at base(b) + i * pstride + w * lstride
struct Blocker {
    int dir, nx, ny, nz, n;
    long long nlines, pstride, lstride;
    Blocker(int dir_, int nx_, int ny_, int nz_) : dir(dir_), nx(nx_), ny(ny_), nz(nz_) {
        if (dir == 0) { n = nx; nlines = (long long)ny * nz; pstride = 1; lstride = nx; }
        else if (dir == 1) { n = ny; nlines = (long long)nx * nz; pstride = nx; lstride = 1; }
        else { n = nz; nlines = (long long)nx * ny; pstride = (long long)nx * ny; lstride = 1; }
    }
    long long nblocks() const {
        if (dir == 1) return (long long)nz * ((nx + LB - 1) / LB);
        return (nlines + LB - 1) / LB;
    }
    // first element and width of block b
    void locate(long long b, long long& base, int& W) const {
        if (dir == 0) { const long long l0 = b * LB; W = (int)std::min<long long>(LB, nlines - l0); base = l0 * nx; }
        else if (dir == 1) {
            const int per = (nx + LB - 1) / LB;
            const long long k = b / per; const int i0 = (int)(b % per) * LB;
            W = std::min(LB, nx - i0); base = k * (long long)nx * ny + i0;
        } else { const long long l0 = b * LB; W = (int)std::min<long long>(LB, nlines - l0); base = l0; }
    }
    void gather(const double* __restrict__ src, long long base, int W, double* __restrict__ buf) const {
        if (dir == 0) {
            for (int w = 0; w < W; w++) { const double* s = src + base + (long long)w * nx; for (int i = 0; i < n; i++) buf[(size_t)i * W + w] = s[i]; }
        } else {
            for (int i = 0; i < n; i++) std::memcpy(buf + (size_t)i * W, src + base + (long long)i * pstride, (size_t)W * sizeof(double));
        }
    }
    template <class F>
    void scatter(double* __restrict__ dst, long long base, int W, const double* __restrict__ buf, F combine) const {
        if (dir == 0) {
            for (int w = 0; w < W; w++) { double* d = dst + base + (long long)w * nx; for (int i = 0; i < n; i++) d[i] = combine(d[i], buf[(size_t)i * W + w]); }
        } else {
            for (int i = 0; i < n; i++) { double* d = dst + base + (long long)i * pstride; const double* s = buf + (size_t)i * W; for (int w = 0; w < W; w++) d[w] = combine(d[w], s[w]); }
        }
    }
};

}  // namespace

extern "C" {

int cpu_threads(void) { return omp_get_max_threads(); }

// mode 1: out (op)= d/ds (u + scale*u2)                       OPR_Partial OPR_P1 (opr_partial.f90), FDM_Der1_Solve
// mode 4: out (op)= [d2 (diffusivity-scaled LU)] - vel * d1     OPR_Burgers_1D (opr_burgers.f90:439-521)
// accumulate: 0 out = r, +1 out += r, -1 out -= r
void cpu_line_op(int mode, int dir, int nx, int ny, int nz, const Band* b1, const Tri* t1, const Band* b2, const Tri* t2,
                 const double* rhs_d1, const double* u, const double* u2, double scale, const double* vel, double* out,
                 int accumulate) {
    const Blocker G(dir, nx, ny, nz);
    const long long nb = G.nblocks();
    const int n = G.n;
#pragma omp parallel
    {
        std::vector<double> bu((size_t)n * LB), bf1((size_t)n * LB), bf2((size_t)n * LB), bv((size_t)n * LB), wrk(LB);
#pragma omp for schedule(static)
        for (long long b = 0; b < nb; b++) {
            long long base; int W;
            G.locate(b, base, W);
            G.gather(u, base, W, bu.data());
            if (u2) {
                G.gather(u2, base, W, bv.data());
                for (size_t i = 0; i < (size_t)n * W; i++) bu[i] = bu[i] + bv[i] * scale;
            }
            band_apply(*b1, bu.data(), bf1.data(), W);
            tri_solve(*t1, bf1.data(), W, wrk.data());
            const double* res = bf1.data();
            if (mode == 4) {
                band_apply(*b2, bu.data(), bf2.data(), W);
                if (rhs_d1) jac_add(rhs_d1, n, bf1.data(), bf2.data(), W);
                tri_solve(*t2, bf2.data(), W, wrk.data());
                G.gather(vel, base, W, bv.data());
                for (size_t i = 0; i < (size_t)n * W; i++) bf2[i] = bf2[i] - bv[i] * bf1[i];
                res = bf2.data();
            }
            if (accumulate > 0) G.scatter(out, base, W, res, [](double o, double r) { return o + r; });
            else if (accumulate < 0) G.scatter(out, base, W, res, [](double o, double r) { return o - r; });
            else G.scatter(out, base, W, res, [](double, double r) { return r; });
        }
    }
}

// ---- element-wise sweeps (each is one pass of the reference over a field) -------------------------------------------
void cpu_axpy(long long n, double a, const double* x, double* y) {          // y = y + a x
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = y[i] + a * x[i];
}
void cpu_scale(long long n, double a, double* y) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = a * y[i];
}
void cpu_zero(long long n, double* y) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = 0.0;
}
void cpu_sub(long long n, const double* x, double* y) {                     // y = y - x
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = y[i] - x[i];
}
void cpu_clip(long long n, double lo, double hi, double* y) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = std::min(std::max(y[i], lo), hi);
}
// hq += g * (c1*s - (ref(j) - c0))   (Gravity_Buoyancy EQNS_BOD_LINEAR + TLab_Sources_Flow)
void cpu_buoyancy_linear(int nx, int ny, int nz, double g, double c1, double c0, const double* ref, const double* s, double* hq) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++) {
            const double dummy = ref[j] - c0;
            const size_t o = ((size_t)k * ny + j) * nx;
            for (int i = 0; i < nx; i++) hq[o + i] = hq[o + i] + g * (c1 * s[o + i] - dummy);
        }
}
void cpu_get_planes(int nx, int ny, int nz, const double* f, double* hb, double* ht) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; k++) {
        std::memcpy(hb + (size_t)k * nx, f + (size_t)k * ny * nx, (size_t)nx * sizeof(double));
        std::memcpy(ht + (size_t)k * nx, f + ((size_t)k * ny + ny - 1) * nx, (size_t)nx * sizeof(double));
    }
}
void cpu_set_planes(int nx, int ny, int nz, double* f, const double* hb, const double* ht) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; k++) {
        std::memcpy(f + (size_t)k * ny * nx, hb + (size_t)k * nx, (size_t)nx * sizeof(double));
        std::memcpy(f + ((size_t)k * ny + ny - 1) * nx, ht + (size_t)k * nx, (size_t)nx * sizeof(double));
    }
}

// BOUNDARY_BCS_NEUMANN_Y (boundary_bcs.f90:368-473): wall values such that the normal derivative vanishes.  The reference
// differentiates the whole field for it (banded product with the Neumann rows, reduced solve) and keeps two planes; so does this.
//   b1, t1: first-derivative band (dense rows of the Neumann variant ibc) and the LU of the reduced system;
//   fb[k], ft[k]: the functionals bcs_b, bcs_t of the product on the first / last BW points; lu_b, lu_t: closure coefficients
void cpu_neumann_y(int ibc, int nx, int ny, int nz, const Band* b1, const Tri* t1, const double* fb, const double* ft, double lu_b,
                   double lu_t, const double* u, double* hb, double* ht) {
    const Blocker G(1, nx, ny, nz);
    const long long nb = G.nblocks();
    const int n = ny;
#pragma omp parallel
    {
        std::vector<double> bu((size_t)n * LB), bf((size_t)n * LB), wrk(LB);
#pragma omp for schedule(static)
        for (long long b = 0; b < nb; b++) {
            long long base; int W;
            G.locate(b, base, W);
            G.gather(u, base, W, bu.data());
            band_apply(*b1, bu.data(), bf.data(), W);
            tri_solve(*t1, bf.data(), W, wrk.data());
            const long long k = base / ((long long)nx * ny);
            const int i0 = (int)(base - k * (long long)nx * ny);
            for (int w = 0; w < W; w++) {
                double sb = 0.0, st = 0.0;
                for (int c = 0; c < BW; c++) { sb += fb[c] * bu[(size_t)c * W + w]; st += ft[c] * bu[(size_t)(n - 1 - c) * W + w]; }
                if (ibc & 1) hb[k * nx + i0 + w] = sb + lu_b * bf[(size_t)1 * W + w];
                if (ibc & 2) ht[k * nx + i0 + w] = st + lu_t * bf[(size_t)(n - 2) * W + w];
            }
        }
    }
}

// ---- Poisson: the y problems of the (kx, kz) modes ----------------------------------------------------------------------
// Per side (BCS_MIN with +sqrt(lambda), BCS_MAX with -sqrt(lambda)) and mode m the oracle provides, from FDM_Int1_Initialize:
//   L[m][n][5]   pentadiagonal system; rows 2..n-1 LU-factored by PENTADFS, rows 1 and n as reduced by FDM_Bcs_Reduce
//   rb[m][10]    rhs_b(1,1:3), rhs_b(2,1:3), rhs_b(3,0:3);   rt[m][10]  rhs_t(0,1:4), rhs_t(1,1:3), rhs_t(2,1:3)
// and, shared by all modes, the tridiagonal right-hand side operator rhs[n][3] of each side.
struct Int1 {
    int n, bc;                 // bc: 1 BCS_MIN, 2 BCS_MAX
    const double* L;           // [n][5] of this mode
    const double* rb;          // [10]
    const double* rt;          // [10]
    const double* rhs;         // [n][3]
};

}  // extern "C"

namespace {

// FDM_Int1_Solve (fdm_integral.f90:219-314) for nl lines stored as f[n][nl]; result holds the boundary value on entry.
// Returns the derivative at the boundary in du[nl] when du != nullptr.
void int1_solve(const Int1& S, const double* __restrict__ f, double* __restrict__ r, int nl, double* __restrict__ du,
                double* __restrict__ bcsb, double* __restrict__ bcst) {
    const int n = S.n;
    const double* L = S.L;
    auto Lr = [&](int row1, int col1) { return L[(size_t)(row1 - 1) * 5 + (col1 - 1)]; };
    auto R = [&](int row1) { return r + (size_t)(row1 - 1) * nl; };
    auto F = [&](int row1) { return f + (size_t)(row1 - 1) * nl; };
    if (S.bc == 1) for (int l = 0; l < nl; l++) R(n)[l] = F(n)[l];
    else for (int l = 0; l < nl; l++) R(1)[l] = F(1)[l];
    // MatMul_3d with BCS_BOTH (fdm_matmul.f90:70-121)
    const double* rb = S.rb;
    const double* rt = S.rt;
    for (int l = 0; l < nl; l++) {
        bcsb[l] = R(1)[l] * rb[1] + F(2)[l] * rb[2] + F(3)[l] * rb[0];
        const double r2v = R(1)[l] * rb[3] + F(2)[l] * rb[4] + F(3)[l] * rb[5];
        const double r3v = R(1)[l] * rb[6] + F(2)[l] * rb[7] + F(3)[l] * rb[8] + F(4)[l] * rb[9];
        R(2)[l] = r2v; R(3)[l] = r3v;
    }
    for (int i = 4; i <= n - 3; i++) {
        const double c1 = S.rhs[(size_t)(i - 1) * 3], c2 = S.rhs[(size_t)(i - 1) * 3 + 1];
        const double *fm = F(i - 1), *fc = F(i), *fp = F(i + 1);
        double* o = R(i);
        for (int l = 0; l < nl; l++) o[l] = fm[l] * c1 + fc[l] * c2 + fp[l];
    }
    for (int l = 0; l < nl; l++) {
        const double rn = R(n)[l];
        const double a = F(n - 3)[l] * rt[0] + F(n - 2)[l] * rt[1] + F(n - 1)[l] * rt[2] + rn * rt[3];
        const double b = F(n - 2)[l] * rt[4] + F(n - 1)[l] * rt[5] + rn * rt[6];
        bcst[l] = F(n - 2)[l] * rt[9] + F(n - 1)[l] * rt[7] + rn * rt[8];
        R(n - 2)[l] = a; R(n - 1)[l] = b;
    }
    // PENTADSS (linear5.f90:76-131) on rows 2..n-1
    {
        const int m = n - 2;
        auto A = [&](int k, int col1) { return L[(size_t)(k + 1) * 5 + (col1 - 1)]; };      // k = 0..m-1 -> row k+2
        double* g = R(2);
        for (int l = 0; l < nl; l++) g[nl + l] = g[nl + l] + g[l] * A(1, 2);
        for (int k = 2; k < m; k++) {
            const double bk = A(k, 2), ak = A(k, 1);
            double* o = g + (size_t)k * nl;
            for (int l = 0; l < nl; l++) o[l] = o[l] + o[l - nl] * bk + o[l - 2 * nl] * ak;
        }
        {
            double* o = g + (size_t)(m - 1) * nl;
            const double ck = A(m - 1, 3);
            for (int l = 0; l < nl; l++) o[l] = o[l] * ck;
        }
        {
            double* o = g + (size_t)(m - 2) * nl;
            const double ck = A(m - 2, 3), dk = A(m - 2, 4);
            for (int l = 0; l < nl; l++) o[l] = (o[l] + o[l + nl] * dk) * ck;
        }
        for (int k = m - 3; k >= 0; k--) {
            const double ck = A(k, 3), dk = A(k, 4), ek = A(k, 5);
            double* o = g + (size_t)k * nl;
            for (int l = 0; l < nl; l++) o[l] = (o[l] + o[l + nl] * dk + o[l + 2 * nl] * ek) * ck;
        }
    }
    if (S.bc == 2) {
        for (int l = 0; l < nl; l++) {
            double r1 = bcsb[l];
            r1 = r1 + Lr(1, 4) * R(2)[l];
            r1 = r1 + Lr(1, 5) * R(3)[l];
            r1 = r1 + Lr(1, 1) * R(4)[l];
            R(1)[l] = r1;
        }
        if (du) for (int l = 0; l < nl; l++) {
            double d = Lr(n, 3) * R(n)[l];
            d = d + Lr(n, 2) * R(n - 1)[l];
            d = d + Lr(n, 1) * R(n - 2)[l];
            d = d + Lr(n, 5) * R(n - 3)[l];
            d = d + S.rhs[(size_t)(n - 1) * 3 + 0] * F(n - 1)[l];
            du[l] = d;
        }
    } else {
        for (int l = 0; l < nl; l++) {
            double rn = bcst[l];
            rn = rn + Lr(n, 2) * R(n - 1)[l];
            rn = rn + Lr(n, 1) * R(n - 2)[l];
            rn = rn + Lr(n, 5) * R(n - 3)[l];
            R(n)[l] = rn;
        }
        if (du) for (int l = 0; l < nl; l++) {
            double d = Lr(1, 3) * R(1)[l];
            d = d + Lr(1, 4) * R(2)[l];
            d = d + Lr(1, 5) * R(3)[l];
            d = d + Lr(1, 1) * R(4)[l];
            d = d + S.rhs[2] * F(2)[l];
            du[l] = d;
        }
    }
}

}  // namespace

extern "C" {

// OPR_ODE2_Factorize_NN (opr_odes.f90:265-386) / _NN_Sing (:165-183 -> _DN_Sing :37-96) for every mode of the half spectrum
// c[nz][ny][nxh] (complex, interleaved); on return c holds p^ and cv dp^/dy.  lam[m] = lambda of mode m = k*nxh + i (the
// reference passes sqrt(lambda) to the integral operators), sing[m] != 0 marks the singular modes.
void cpu_poisson_modes(int nxh, int ny, int nz, double* c, double* cv, const double* lam, const unsigned char* sing,
                       const double* Lmin, const double* Lmax, const double* rbmin, const double* rtmin, const double* rbmax,
                       const double* rtmax, const double* rhsmin, const double* rhsmax) {
    const long long nm = (long long)nxh * nz;
    const int n = ny;
#pragma omp parallel
    {
        // lines: 2 (re, im) for the data, 3 for the fundamental solutions
        std::vector<double> f((size_t)n * 2), v((size_t)n * 2), u((size_t)n * 2), w1((size_t)n * 3), w2((size_t)n * 3);
        double bb[3], bt[3], du0[2], der[3], d1[1];
#pragma omp for schedule(static)
        for (long long m = 0; m < nm; m++) {
            const long long k = m / nxh; const int i = (int)(m - k * nxh);
            double* cm = c + ((size_t)k * n * nxh + i) * 2;                  // row j at + j*nxh*2
            double* cvm = cv + ((size_t)k * n * nxh + i) * 2;
            for (int j = 0; j < n; j++) { f[2 * j] = cm[(size_t)j * nxh * 2]; f[2 * j + 1] = cm[(size_t)j * nxh * 2 + 1]; }
            const double bcs0[2] = {f[0], f[1]}, bcs1[2] = {f[2 * (n - 1)], f[2 * (n - 1) + 1]};
            Int1 Smin{n, 1, Lmin + (size_t)m * n * 5, rbmin + (size_t)m * 10, rtmin + (size_t)m * 10, rhsmin};
            Int1 Smax{n, 2, Lmax + (size_t)m * n * 5, rbmax + (size_t)m * 10, rtmax + (size_t)m * 10, rhsmax};
            const double lm = std::sqrt(lam[m]);
            std::fill(u.begin(), u.end(), 0.0);
            std::fill(v.begin(), v.end(), 0.0);
            if (sing[m]) {
                // OPR_ODE2_Factorize_DN_Sing with bcs(1) = 0
                f[0] = 0.0; f[1] = 0.0;
                v[2 * (n - 1)] = bcs1[0]; v[2 * (n - 1) + 1] = bcs1[1];
                int1_solve(Smax, f.data(), v.data(), 2, nullptr, bb, bt);
                std::vector<double> f1(n, 0.0), v1(n, 0.0), u1(n, 0.0);
                f1[0] = 1.0;
                int1_solve(Smax, f1.data(), v1.data(), 1, nullptr, bb, bt);
                u[0] = 0.0; u[1] = 0.0;
                int1_solve(Smin, v.data(), u.data(), 2, du0, bb, bt);
                int1_solve(Smin, v1.data(), u1.data(), 1, d1, bb, bt);
                const double ff = 1.0 / (d1[0] - v1[0]);
                const double a0 = (v[0] - du0[0]) * ff, a1 = (v[1] - du0[1]) * ff;
                for (int j = 0; j < n; j++) {
                    u[2 * j] = u[2 * j] + a0 * u1[j]; u[2 * j + 1] = u[2 * j + 1] + a1 * u1[j];
                    v[2 * j] = v[2 * j] + a0 * v1[j]; v[2 * j + 1] = v[2 * j + 1] + a1 * v1[j];
                }
            } else {
                f[2 * (n - 1)] = 0.0; f[2 * (n - 1) + 1] = 0.0;
                int1_solve(Smin, f.data(), v.data(), 2, nullptr, bb, bt);
                std::fill(w1.begin(), w1.end(), 0.0);
                std::fill(w2.begin(), w2.end(), 0.0);
                w1[3 * (n - 1) + 0] = 1.0;          // f1(n)
                w2[1] = 1.0;                        // em(1)
                int1_solve(Smin, w1.data(), w2.data(), 3, nullptr, bb, bt);
                // v1 = w2[:,0], em = w2[:,1]
                int1_solve(Smax, v.data(), u.data(), 2, du0, bb, bt);
                w1[3 * (n - 1) + 0] = 0.0; w1[3 * (n - 1) + 1] = 0.0;
                for (int j = 0; j < n; j++) w2[3 * j + 2] = 0.0;
                w1[3 * (n - 1) + 2] = 1.0;          // ep(n)
                int1_solve(Smax, w2.data(), w1.data(), 3, der, bb, bt);
                auto V1 = [&](int j) { return w2[3 * j]; };
                auto EM = [&](int j) { return w2[3 * j + 1]; };
                auto U1 = [&](int j) { return w1[3 * j]; };
                auto SP = [&](int j) { return w1[3 * j + 1]; };
                auto EP = [&](int j) { return w1[3 * j + 2]; };
                double a11 = 1.0 + lm * SP(0), a21 = EM(n - 1), a31 = der[1];
                double a12 = lm * EP(0), a22 = lm, a32 = der[2];
                double a13 = lm * U1(0), a23 = V1(n - 1), a33 = der[0];
                a12 = a12 / a11;
                a22 = a22 - a21 * a12;
                a32 = a32 - a31 * a12;
                a13 = a13 / a11;
                a23 = (a23 - a21 * a13) / a22;
                a33 = a33 - a31 * a13 - a32 * a23;
                for (int q = 0; q < 2; q++) {
                    const double b0 = q ? bcs0[1] : bcs0[0], b1 = q ? bcs1[1] : bcs1[0];
                    double v0 = (b0 - lm * u[q]) / a11;
                    double un = (b1 - v[2 * (n - 1) + q] - a21 * v0) / a22;
                    const double fn = (b1 - du0[q] - a31 * v0 - a32 * un) / a33;
                    un = un - a23 * fn;
                    v0 = v0 - a12 * un - a13 * fn;
                    v[q] = v0; u[2 * (n - 1) + q] = un;
                    int j = n - 1;
                    v[2 * j + q] = v[2 * j + q] + fn * V1(j) + v0 * EM(j) + lm * u[2 * j + q];
                    for (j = n - 2; j >= 1; j--) {
                        u[2 * j + q] = u[2 * j + q] + fn * U1(j) + v0 * SP(j) + un * EP(j);
                        v[2 * j + q] = v[2 * j + q] + fn * V1(j) + v0 * EM(j) + lm * u[2 * j + q];
                    }
                    j = 0;
                    u[q] = u[q] + fn * U1(j) + v0 * SP(j) + un * EP(j);
                    v[q] = v[q] + lm * u[q];
                }
            }
            for (int j = 0; j < n; j++) {
                cm[(size_t)j * nxh * 2] = u[2 * j]; cm[(size_t)j * nxh * 2 + 1] = u[2 * j + 1];
                cvm[(size_t)j * nxh * 2] = v[2 * j]; cvm[(size_t)j * nxh * 2 + 1] = v[2 * j + 1];
            }
        }
    }
}

}  // extern "C"
