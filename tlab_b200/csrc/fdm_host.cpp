// Host-side plan builder: compact-scheme coefficient tables, Neumann reductions and LU factors.
//
// What is computed here is fixed by the numerical method tlab uses (Lele 1992 compact schemes in
// Jacobian form, Carpenter et al. 1993 boundary closures, Lamballais et al. 2011 hyperviscous second
// derivative) and by the table layout its operators expect; the reference restates it in
//   src/fdm/fdm_com1_jacobian.f90:38-291, src/fdm/fdm_com2_jacobian.f90:39-282,
//   src/fdm/fdm_base.f90:194-391, src/utils/linear3.f90:29-51,269-316,
//   src/fdm/fdm_derivative.f90:63-459, src/fdm/fdm.f90:143-252, src/fdm/fdm_integral.f90:91-214.
// The arithmetic order matters for parity at the 1e-12 level only through round-off; the table
// *layout* (which slot holds the extended boundary stencil, which LU block belongs to which BC)
// is part of the drop-in contract and is kept.
#include "fdm_host.h"
#include <cmath>
#include <algorithm>

namespace tlab {

static const double PI = 3.14159265358979323846;

namespace {

struct SchemeRows {
    int ndl, ndr;
    double interior[5];                 // a1, a2, b1, b2, b3
    int nbc;                            // number of special boundary rows given (0 for periodic)
    std::vector<std::vector<double>> bc; // bc[k] = {a1, a2, b1, b2, ...} of row k+1
};

// circular shift of a 1-based vector: out(i) = v(i + s)
std::vector<double> cshift(const std::vector<double>& v, int s) {
    int n = (int)v.size() - 1;
    std::vector<double> o(n + 1, 0.0);
    for (int i = 1; i <= n; i++) {
        int j = ((i - 1 + s) % n + n) % n + 1;
        o[i] = v[j];
    }
    return o;
}

// Fill lhs/rhs with the interior stencil and the biased rows at both ends.
// sign = -1: antisymmetric closure mirror (first derivative); +1: symmetric (second derivative).
void fill_rows(const SchemeRows& S, int n, bool second, Mat& lhs, Mat& rhs) {
    const int ndl = S.ndl, ndr = S.ndr, idl = ndl / 2 + 1, idr = ndr / 2 + 1;
    const double* c = S.interior - 1;   // c[1..5]
    for (int i = 1; i <= n; i++) {
        lhs(i, idl) = 1.0;
        for (int ic = 1; ic < idl; ic++) { lhs(i, idl - ic) = c[ic]; lhs(i, idl + ic) = c[ic]; }
        double centre = 0.0;
        for (int ic = 1; ic < idr; ic++) {
            if (second) {
                centre = centre - 2.0 * c[ic + 2];
                rhs(i, idr - ic) = c[ic + 2];
            } else {
                rhs(i, idr - ic) = -c[ic + 2];
            }
            rhs(i, idr + ic) = c[ic + 2];
        }
        rhs(i, idr) = centre;
    }
    const double mirror = second ? 1.0 : -1.0;
    for (int k = 0; k < S.nbc; k++) {
        const std::vector<double>& bv = S.bc[k];
        auto b = [&](int i) { return (i >= 1 && i <= (int)bv.size()) ? bv[i - 1] : 0.0; };
        const int row = k + 1;
        for (int j = 1; j <= ndl; j++) lhs(row, j) = 0.0;
        for (int j = 1; j <= ndr; j++) rhs(row, j) = 0.0;
        if (row == 1) {
            lhs(1, idl) = 1.0;
            int icmax = std::min(idl - 1, 2);
            for (int ic = 1; ic <= icmax; ic++) lhs(1, idl + ic) = b(ic);
            icmax = std::min(idr, 4);
            for (int ic = 0; ic < icmax; ic++) rhs(1, idr + ic) = b(3 + ic);
            // slot (1,1) carries one more stencil point when the diagonals run out; it is 0 when the
            // closure already fits (7-diagonal rhs) -- the reference indexes past its 6-element
            // closure array there, which is undefined; the intended value is 0.
            rhs(1, 1) = b(3 + icmax);
        } else if (row == 2) {
            lhs(2, idl - 1) = b(1); lhs(2, idl) = 1.0; lhs(2, idl + 1) = b(2);
            int icmax = std::min(idr + 1, 4);
            for (int ic = 0; ic < icmax; ic++) rhs(2, idr - 1 + ic) = b(3 + ic);
        } else {
            lhs(3, idl - 1) = b(1); lhs(3, idl) = 1.0; lhs(3, idl + 1) = b(2);
            int icmax = std::min(idr + 2, 6);
            for (int ic = 0; ic < icmax; ic++) rhs(3, idr - 2 + ic) = b(3 + ic);
        }
        const int mrow = n - k;
        for (int j = 1; j <= ndl; j++) lhs(mrow, j) = lhs(row, ndl + 1 - j);
        for (int j = 1; j <= ndr; j++) rhs(mrow, j) = mirror * rhs(row, ndr + 1 - j);
    }
}

SchemeRows scheme_der1(int mode, bool periodic) {
    SchemeRows S;
    if (mode == FDM_COM4_JACOBIAN) {
        S.ndl = 3; S.ndr = 3;
        double c[5] = {0.25, 0.0, 0.75, 0.0, 0.0};
        std::copy(c, c + 5, S.interior);
        if (!periodic) S.bc = {{2.0, 0.0, -2.5, 2.0, 0.5, 0.0}};
    } else {   // 6th-order tridiagonal
        S.ndl = 3; S.ndr = 5;
        double c[5] = {1.0 / 3.0, 0.0, 7.0 / 9.0, 1.0 / 36.0, 0.0};
        std::copy(c, c + 5, S.interior);
        if (!periodic) S.bc = {{2.0, 0.0, -2.5, 2.0, 0.5, 0.0},
                               {1.0 / 6.0, 0.5, -5.0 / 9.0, -0.5, 1.0, 1.0 / 18.0}};
    }
    S.nbc = (int)S.bc.size();
    return S;
}

SchemeRows scheme_der2(int mode, bool periodic) {
    SchemeRows S;
    const std::vector<double> b1 = {11.0, 0.0, 13.0, -27.0, 15.0, -1.0};
    const std::vector<double> b2 = {0.1, 0.1, 1.2, -2.4, 1.2, 0.0};
    if (mode == FDM_COM4_JACOBIAN) {
        S.ndl = 3; S.ndr = 5;
        double c[5] = {0.1, 0.0, 1.2, 0.0, 0.0};
        std::copy(c, c + 5, S.interior);
        if (!periodic) S.bc = {b1};
    } else if (mode == FDM_COM6_JACOBIAN || mode == FDM_COM6_JACOBIAN_PENTA) {
        S.ndl = 3; S.ndr = 5;
        double c[5] = {2.0 / 11.0, 0.0, 12.0 / 11.0, 3.0 / 44.0, 0.0};
        std::copy(c, c + 5, S.interior);
        if (!periodic) S.bc = {b1, b2};
    } else {   // hyperviscous 6th order, Lamballais et al. 2011
        S.ndl = 3; S.ndr = 7;
        const double kc = std::pow(PI, 2.0);
        double c[5] = {(272.0 - 45.0 * kc) / (416.0 - 90.0 * kc), 0.0,
                       (48.0 - 135.0 * kc) / (1664.0 - 360.0 * kc),
                       (528.0 - 81.0 * kc) / (208.0 - 45.0 * kc) / 4.0,
                       -(432.0 - 63.0 * kc) / (1664.0 - 360.0 * kc) / 9.0};
        std::copy(c, c + 5, S.interior);
        if (!periodic) S.bc = {b1, b2,
                               {2.0 / 11.0, 2.0 / 11.0, 3.0 / 44.0, 12.0 / 11.0, -51.0 / 22.0, 12.0 / 11.0, 3.0 / 44.0, 0.0}};
    }
    S.nbc = (int)S.bc.size();
    return S;
}

std::vector<double> wavenumbers(int n) {
    std::vector<double> wn(n);
    for (int i = 1; i <= n; i++)
        wn[i - 1] = 2.0 * PI * double(i <= n / 2 + 1 ? i - 1 : i - 1 - n) / double(n);
    return wn;
}

// ---- Thomas factorizations on column views of a Mat -------------------------------------------
struct Col {
    Mat* m; int c; int r0;   // element k (1-based) is (*m)(r0 + k - 1, c)
    double& operator[](int k) { return (*m)(r0 + k - 1, c); }
};

void tridfs(int nmax, Col a, Col b, Col c) {
    for (int n = 2; n <= nmax; n++) {
        a[n] = a[n] / b[n - 1];
        b[n] = b[n] - a[n] * c[n - 1];
    }
    for (int n = 1; n <= nmax; n++) { a[n] = -a[n]; b[n] = 1.0 / b[n]; c[n] = -c[n]; }
}

void tridpfs(int nmax, Col a, Col b, Col c, Col d, Col e) {
    c[1] = c[1] / b[1];
    e[1] = a[1] / b[1];
    d[1] = c[nmax];
    for (int n = 2; n <= nmax - 2; n++) {
        b[n] = b[n] - a[n] * c[n - 1];
        c[n] = c[n] / b[n];
        e[n] = -a[n] * e[n - 1] / b[n];
        d[n] = -d[n - 1] * c[n - 1];
    }
    b[nmax - 1] = b[nmax - 1] - a[nmax - 1] * c[nmax - 2];
    e[nmax - 1] = (c[nmax - 1] - a[nmax - 1] * e[nmax - 2]) / b[nmax - 1];
    d[nmax - 1] = a[nmax] - d[nmax - 2] * c[nmax - 2];
    double sum = 0.0;
    for (int n = 1; n <= nmax - 1; n++) sum = sum + d[n] * e[n];
    b[nmax] = b[nmax] - sum;
    for (int n = 1; n <= nmax; n++) {
        b[n] = 1.0 / b[n];
        a[n] = -a[n] * b[n];
        c[n] = -c[n];
        e[n] = -e[n];
    }
}

// Neumann reduction of the first/last rows (rhs_b rows 1.., cols 0..; rhs_t rows 0.., cols 1..)
void bcs_neumann(int ibc, Mat& lhs, int lc0, int ndl, const Mat& rhs, int ndr, Mat& rhs_b, Mat& rhs_t) {
    // lhs columns are lc0+1 .. lc0+ndl of the LU table
    auto L = [&](int i, int j) -> double& { return lhs(i, lc0 + j); };
    const int idl = ndl / 2 + 1, idr = ndr / 2 + 1, nx = rhs.r1;
    if (ibc == BCS_ND || ibc == BCS_NN) {
        for (int i = 1; i <= idr; i++) for (int j = 1; j <= ndr; j++) rhs_b(i, j) = rhs(i, j);
        const double dummy = 1.0 / rhs(1, idr);
        for (int j = 1; j <= ndr; j++) rhs_b(1, j) = -rhs_b(1, j) * dummy;
        for (int ir = 1; ir <= idr - 1; ir++) {
            for (int ic = idr + 1; ic <= ndr; ic++)
                rhs_b(1 + ir, ic - ir) = rhs_b(1 + ir, ic - ir) + rhs_b(1 + ir, idr - ir) * rhs_b(1, ic);
            int ic = ndr + 1;
            rhs_b(1 + ir, ic - ir) = rhs_b(1 + ir, ic - ir) + rhs_b(1 + ir, idr - ir) * rhs_b(1, 1);
        }
        for (int j = 1; j <= ndl; j++) L(1, j) = L(1, j) * dummy;
        for (int ir = 1; ir <= idr - 1; ir++) {
            for (int ic = idl + 1; ic <= ndl; ic++)
                L(1 + ir, ic - ir) = L(1 + ir, ic - ir) - rhs_b(1 + ir, idr - ir) * L(1, ic);
            rhs_b(1 + ir, idr - ir) = rhs_b(1 + ir, idr - ir) * L(1, idl);
        }
        for (int ir = 1; ir <= idl - 1; ir++)
            rhs_b(1 + ir, idr - ir) = rhs_b(1 + ir, idr - ir) - L(1 + ir, idl - ir);
        rhs_b(1, idr) = L(1, idl);
    }
    if (ibc == BCS_DN || ibc == BCS_NN) {
        for (int i = 1; i <= idr; i++) for (int j = 1; j <= ndr; j++) rhs_t(i, j) = rhs(nx - idr + i, j);
        const double dummy = 1.0 / rhs(nx, idr);
        for (int j = 1; j <= ndr; j++) rhs_t(idr, j) = -rhs_t(idr, j) * dummy;
        for (int ir = 1; ir <= idr - 1; ir++) {
            for (int ic = 1; ic <= idr - 1; ic++)
                rhs_t(idr - ir, ic + ir) = rhs(nx - ir, ic + ir) + rhs(nx - ir, idr + ir) * rhs_t(idr, ic);
            int ic = 0;
            rhs_t(idr - ir, ic + ir) = rhs_t(idr - ir, ic + ir) + rhs(nx - ir, idr + ir) * rhs_t(idr, ndr);
        }
        for (int j = 1; j <= ndl; j++) L(nx, j) = L(nx, j) * dummy;
        for (int ir = 1; ir <= idr - 1; ir++) {
            for (int ic = 1; ic <= idl - 1; ic++)
                L(nx - ir, ic + ir) = L(nx - ir, ic + ir) - rhs(nx - ir, idr + ir) * L(nx, ic);
            rhs_t(idr - ir, idr + ir) = rhs_t(idr - ir, idr + ir) * L(nx, idl);
        }
        for (int ir = 1; ir <= idl - 1; ir++)
            rhs_t(idr - ir, idr + ir) = rhs_t(idr - ir, idr + ir) - L(nx - ir, idl + ir);
        rhs_t(idr, idr) = L(nx, idl);
    }
}

// ---- banded mat-vec for one line (host side; only needed to compute the Jacobians) ------------
// antisymmetric (first derivative) or symmetric (second derivative) interior, generic boundary rows
void matmul_line(const HostDer& g, bool second, int ibc, const double* u, double* f) {
    const int n = g.size, ndr = g.ndr, idr = ndr / 2 + 1, h = idr - 1;
    const Mat& r = g.rhs;
    auto U = [&](int i) { return u[i - 1]; };
    auto Up = [&](int i) { return u[((i - 1) % n + n) % n]; };
    const bool per = (ibc == BCS_PERIODIC);
    if (second && g.mode_fdm == FDM_COM6_DIRECT) {
        // MatMul_5d with per-row coefficients (fdm_matmul.f90:266-320)
        for (int i = 1; i <= n; i++) {
            double s = 0.0;
            if (i == 1) s = U(1) * r(1, 3) + U(2) * r(1, 4) + U(3) * r(1, 5) + U(4) * r(1, 1);
            else if (i == 2) s = U(1) * r(2, 2) + U(2) * r(2, 3) + U(3) * r(2, 4) + U(4) * r(2, 5);
            else if (i == n - 1) s = U(n - 3) * r(i, 1) + U(n - 2) * r(i, 2) + U(n - 1) * r(i, 3) + U(n) * r(i, 4);
            else if (i == n) s = U(n - 3) * r(n, 5) + U(n - 2) * r(n, 1) + U(n - 1) * r(n, 2) + U(n) * r(n, 3);
            else if (i <= 4 || i >= n - 3) s = U(i - 2) * r(i, 1) + U(i - 1) * r(i, 2) + U(i) * r(i, 3) + U(i + 1) * r(i, 4) + U(i + 2) * r(i, 5);
            else s = U(i - 2) * r(i, 1) + U(i - 1) * r(i, 2) + U(i) * r(i, 3) + U(i + 1) + U(i + 2) * r(i, 5);
            f[i - 1] = s;
        }
        return;
    }
    // rows with constant interior stencil: nb+1 .. n-nb (nb = idr-1 special rows for antisym/sym kernels)
    const int nb = per ? 0 : h;
    const int ref_row = h + 2;      // a row guaranteed to hold interior coefficients
    for (int i = 1; i <= n; i++) {
        if (!per && (i <= nb || i > n - nb)) continue;
        double s;
        if (second) {
            s = r(ref_row, idr) * Up(i) + Up(i + 1) + Up(i - 1);
            for (int k = 2; k <= h; k++) s = s + r(ref_row, idr + k) * (Up(i + k) + Up(i - k));
        } else {
            s = Up(i + 1) - Up(i - 1);
            for (int k = 2; k <= h; k++) s = s + r(ref_row, idr + k) * (Up(i + k) - Up(i - k));
        }
        f[i - 1] = s;
    }
    if (per) return;
    const bool nb_min = (!second) && (ibc == BCS_ND || ibc == BCS_NN);
    const bool nb_max = (!second) && (ibc == BCS_DN || ibc == BCS_NN);
    for (int i = 1; i <= nb; i++) {     // bottom rows
        double s = 0.0;
        if (nb_min) {
            if (i == 1) continue;       // Neumann row: value imposed (0), not part of the system
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; if (col >= 2 && col <= n) s += g.rhs_b(i, j) * U(col); }
        } else {
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; if (col >= 1 && col <= n) s += r(i, j) * U(col); }
            if (i == 1) s += r(1, 1) * U(idr + 1);          // extended stencil
        }
        f[i - 1] = s;
    }
    for (int q = 1; q <= nb; q++) {     // top rows, i = n - q + 1
        const int i = n - q + 1;
        double s = 0.0;
        if (nb_max) {
            if (q == 1) continue;
            const int tr = idr - q + 1;   // row of rhs_t: idr <-> n
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; if (col >= 1 && col <= n - 1) s += g.rhs_t(tr, j) * U(col); }
        } else {
            if (q == 1) s = r(n, ndr) * U(n - idr);         // extended stencil comes first in the last row
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; if (col >= 1 && col <= n) s += r(i, j) * U(col); }
        }
        f[i - 1] = s;
    }
}

void tridss_line(int nmax, const Mat& lu, int r0, int c0, double* f) {
    auto a = [&](int k) { return lu(r0 + k - 1, c0 + 1); };
    auto b = [&](int k) { return lu(r0 + k - 1, c0 + 2); };
    auto c = [&](int k) { return lu(r0 + k - 1, c0 + 3); };
    for (int n = 2; n <= nmax; n++) f[n - 1] = f[n - 1] + a(n) * f[n - 2];
    f[nmax - 1] = f[nmax - 1] * b(nmax);
    for (int n = nmax - 1; n >= 1; n--) f[n - 1] = (f[n - 1] + c(n) * f[n]) * b(n);
}

void tridpss_line(int nmax, const Mat& lu, double* f) {
    auto a = [&](int k) { return lu(k, 1); };
    auto b = [&](int k) { return lu(k, 2); };
    auto c = [&](int k) { return lu(k, 3); };
    auto d = [&](int k) { return lu(k, 4); };
    auto e = [&](int k) { return lu(k, 5); };
    f[0] = f[0] * b(1);
    for (int n = 2; n <= nmax - 1; n++) f[n - 1] = f[n - 1] * b(n) + a(n) * f[n - 2];
    double wrk = 0.0;
    for (int n = 1; n <= nmax - 1; n++) wrk = wrk + d(n) * f[n - 1];
    f[nmax - 1] = (f[nmax - 1] - wrk) * b(nmax);
    f[nmax - 2] = e(nmax - 1) * f[nmax - 1] + f[nmax - 2];
    for (int n = nmax - 2; n >= 1; n--) f[n - 1] = f[n - 1] + c(n) * f[n] + e(n) * f[nmax - 1];
}

void der1_create(const std::vector<double>& dx /*1-based*/, int n, HostDer& g, bool periodic) {
    SchemeRows S = scheme_der1(g.mode_fdm, periodic);
    g.size = n; g.periodic = periodic; g.ndl = S.ndl; g.ndr = S.ndr;
    std::copy(S.interior, S.interior + 5, g.coef);
    g.lhs = Mat(1, n, 1, 5);
    g.rhs = Mat(1, n, 1, 7);
    fill_rows(S, n, false, g.lhs, g.rhs);
    const int idl = S.ndl / 2 + 1;
    for (int i = 1; i <= n; i++) g.lhs(i, idl) = g.lhs(i, idl) * dx[i];
    for (int ic = 1; ic < idl; ic++) {
        std::vector<double> dm = cshift(dx, -ic), dp = cshift(dx, +ic);
        for (int i = 1; i <= n; i++) {
            g.lhs(i, idl - ic) = g.lhs(i, idl - ic) * dm[i];
            g.lhs(i, idl + ic) = g.lhs(i, idl + ic) * dp[i];
        }
    }
    for (int i = 1; i <= n; i++) {
        for (int j = 1; j <= S.ndl; j++) g.lhs(i, j) = g.lhs(i, j) / S.interior[2];
        for (int j = 1; j <= S.ndr; j++) g.rhs(i, j) = g.rhs(i, j) / S.interior[2];
    }
    g.mwn.assign(n, 0.0);
    if (periodic) {
        std::vector<double> wn = wavenumbers(n);
        const double* c = S.interior - 1;
        for (int i = 0; i < n; i++)   // cos(wn) multiplies a2 as well, as in the reference (fdm_derivative.f90:207)
            g.mwn[i] = 2.0 * (c[3] * std::sin(wn[i]) + c[4] * std::sin(2.0 * wn[i]) + c[5] * std::sin(3.0 * wn[i])) /
                       (1.0 + 2.0 * c[1] * std::cos(wn[i]) + 2.0 * c[2] * std::cos(wn[i]));
    }
}

void der1_initialize(const std::vector<double>& dx, int n, HostDer& g, bool periodic, const std::vector<int>& cases) {
    der1_create(dx, n, g, periodic);
    g.rhs_b = Mat(1, 4, 0, 7);
    g.rhs_t = Mat(0, 4, 1, 7);
    const int ndl = g.ndl;
    if (periodic) {
        g.lu = Mat(1, n, 1, ndl + 2);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= ndl; j++) g.lu(i, j) = g.lhs(i, j);
        tridpfs(n, {&g.lu, 1, 1}, {&g.lu, 2, 1}, {&g.lu, 3, 1}, {&g.lu, 4, 1}, {&g.lu, 5, 1});
    } else {
        g.lu = Mat(1, n, 1, 20);
        Mat rhs_view(1, n, 1, g.ndr);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= g.ndr; j++) rhs_view(i, j) = g.rhs(i, j);
        for (size_t ib = 0; ib < cases.size(); ib++) {
            const int ip = (int)ib * 5, bc = cases[ib];
            for (int i = 1; i <= n; i++) for (int j = 1; j <= ndl; j++) g.lu(i, ip + j) = g.lhs(i, j);
            bcs_neumann(bc, g.lu, ip, ndl, rhs_view, g.ndr, g.rhs_b, g.rhs_t);
            int nmin = 1, nmax = n;
            if (bc == BCS_ND || bc == BCS_NN) nmin++;
            if (bc == BCS_DN || bc == BCS_NN) nmax--;
            tridfs(nmax - nmin + 1, {&g.lu, ip + 1, nmin}, {&g.lu, ip + 2, nmin}, {&g.lu, ip + 3, nmin});
        }
    }
}

void der2_initialize(const std::vector<double>& dx1, const std::vector<double>& dx2, int n, HostDer& g,
                     bool periodic, bool uniform) {
    SchemeRows S = scheme_der2(g.mode_fdm, periodic);
    g.size = n; g.periodic = periodic; g.ndl = S.ndl; g.ndr = S.ndr;
    std::copy(S.interior, S.interior + 5, g.coef);
    g.lhs = Mat(1, n, 1, 5);
    g.rhs = Mat(1, n, 1, 12);
    Mat rhs(1, n, 1, S.ndr);
    fill_rows(S, n, true, g.lhs, rhs);
    const int idl = S.ndl / 2 + 1, ndr = S.ndr;
    // first-derivative correction block -A2 * d2x/ds2, then fold dx/ds into A2
    for (int i = 1; i <= n; i++) g.rhs(i, ndr + idl) = -g.lhs(i, idl) * dx2[i];
    for (int ic = 1; ic < idl; ic++) {
        std::vector<double> dm = cshift(dx2, -ic), dp = cshift(dx2, +ic);
        for (int i = 1; i <= n; i++) {
            g.rhs(i, ndr + idl - ic) = -g.lhs(i, idl - ic) * dm[i];
            g.rhs(i, ndr + idl + ic) = -g.lhs(i, idl + ic) * dp[i];
        }
    }
    for (int i = 1; i <= n; i++) g.lhs(i, idl) = g.lhs(i, idl) * dx1[i] * dx1[i];
    for (int ic = 1; ic < idl; ic++) {
        std::vector<double> dm = cshift(dx1, -ic), dp = cshift(dx1, +ic);
        for (int i = 1; i <= n; i++) {
            g.lhs(i, idl - ic) = g.lhs(i, idl - ic) * dm[i] * dm[i];
            g.lhs(i, idl + ic) = g.lhs(i, idl + ic) * dp[i] * dp[i];
        }
    }
    for (int i = 1; i <= n; i++) {
        for (int j = 1; j <= S.ndl; j++) g.lhs(i, j) = g.lhs(i, j) / S.interior[2];
        for (int j = 1; j <= ndr; j++) g.rhs(i, j) = rhs(i, j) / S.interior[2];
        for (int j = 1; j <= S.ndl; j++) g.rhs(i, ndr + j) = g.rhs(i, ndr + j) / S.interior[2];
    }
    if (!uniform) g.need_1der = true;
    g.mwn.assign(n, 0.0);
    if (periodic) {
        std::vector<double> wn = wavenumbers(n);
        const double* c = S.interior - 1;
        for (int i = 0; i < n; i++)
            g.mwn[i] = 2.0 * (c[3] * (1.0 - std::cos(wn[i])) + c[4] * (1.0 - std::cos(2.0 * wn[i])) +
                              c[5] * (1.0 - std::cos(3.0 * wn[i]))) /
                       (1.0 + 2.0 * c[1] * std::cos(wn[i]) + 2.0 * c[2] * std::cos(2.0 * wn[i]));
    }
    if (periodic) {
        g.lu = Mat(1, n, 1, 5);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= 3; j++) g.lu(i, j) = g.lhs(i, j);
        tridpfs(n, {&g.lu, 1, 1}, {&g.lu, 2, 1}, {&g.lu, 3, 1}, {&g.lu, 4, 1}, {&g.lu, 5, 1});
    } else {
        g.lu = Mat(1, n, 1, 3);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= 3; j++) g.lu(i, j) = g.lhs(i, j);
        tridfs(n, {&g.lu, 1, 1}, {&g.lu, 2, 1}, {&g.lu, 3, 1});
    }
}


// ------------------------------------------------------------------------------------------------
// CompactDirect6 second derivative: the compact scheme derived on the actual (non-uniform) nodes instead of on a uniform
// computational grid (reference src/fdm/fdm_comx_direct.f90:305-412 FDM_C2N6_Direct with its coefficient functions :468-783 and
// the Lagrange helpers of src/fdm/fdm_base.f90:31-143).  Tridiagonal lhs, pentadiagonal rhs with per-row coefficients, no
// Jacobian term.  x is 1-based here, as in the formulas.
struct DirectC2N6 {
    const double* x;      // x[i], i = 1..n
    double Pi(int j, std::initializer_list<int> idx) const { double f = 1.0; for (int k : idx) f *= (x[j] - x[k]); return f; }
    double Pi_p(int j, std::initializer_list<int> idx) const {
        double f = 0.0;
        int kk = 0;
        for (int k : idx) {
            (void)k;
            double d = 1.0;
            int mm = 0;
            for (int m : idx) { if (mm != kk) d *= (x[j] - x[m]); mm++; }
            f += d;
            kk++;
        }
        return f;
    }
    double Pi_pp_3(int j, int a, int b, int c) const { return 2.0 * (x[j] - x[a] + x[j] - x[b] + x[j] - x[c]); }
    double Lag(int j, int i, std::initializer_list<int> idx) const {
        double f = 1.0;
        for (int k : idx) if (k != i) f = f * (x[j] - x[k]) / (x[i] - x[k]);
        return f;
    }
    double Lag_p(int j, int i, std::initializer_list<int> idx) const {
        double den = 1.0, f = 0.0;
        int kk = 0;
        for (int k : idx) {
            if (k != i) {
                double d = 1.0;
                int mm = 0;
                for (int m : idx) { if (m != i && mm != kk) d *= (x[j] - x[m]); mm++; }
                f += d;
                den *= (x[i] - x[k]);
            }
            kk++;
        }
        return f / den;
    }
    double Lag_pp_3(int j, int i, std::initializer_list<int> idx) const {
        (void)j;
        double f = 2.0;
        for (int k : idx) if (k != i) f = f / (x[i] - x[k]);
        return f;
    }
    double PIp_o_PI(int j, int i) const {
        double f = (x[j] - x[i + 2]) * (x[j] - x[i - 2]) + (x[j] - x[i]) * (x[j] - x[i - 2]) + (x[j] - x[i]) * (x[j] - x[i + 2]);
        return f / Pi(j, {i - 2, i, i + 2});
    }
    double PIpp_o_PI(int j, int i) const {
        const double f = x[j] - x[i + 2] + x[j] - x[i - 2] + x[j] - x[i];
        return 2.0 * f / Pi(j, {i - 2, i, i + 2});
    }
    double D_coef(int i) const {
        const double dx = x[i + 1] - x[i - 1];
        return 6.0 + 4.0 * dx * (PIp_o_PI(i + 1, i) - PIp_o_PI(i - 1, i)) - 2.0 * dx * dx * PIp_o_PI(i + 1, i) * PIp_o_PI(i - 1, i);
    }
    double A1D(int im, int ip, int i) const {
        const double dx = x[ip] - x[im];
        return -4.0 * PIp_o_PI(ip, i) - 2.0 * PIp_o_PI(im, i) + 2.0 * dx * (PIp_o_PI(ip, i) * PIp_o_PI(im, i) - PIpp_o_PI(ip, i)) +
               dx * dx * PIpp_o_PI(ip, i) * PIp_o_PI(im, i);
    }
    double A2D(int im, int ip, int i) const {
        const double dx = x[ip] - x[im];
        return 4.0 * PIp_o_PI(ip, i) * PIp_o_PI(im, i) - PIpp_o_PI(ip, i) - 2.0 / dx * (PIp_o_PI(ip, i) - PIp_o_PI(im, i)) +
               dx * PIpp_o_PI(ip, i) * PIp_o_PI(im, i);
    }
    double B1D(int im, int ip, int i) const { const double dx = x[ip] - x[im]; return -(2.0 / dx + PIp_o_PI(ip, i)) * dx * dx; }
    double B2D(int im, int ip, int i) const { const double dx = x[ip] - x[im]; return 1.0 - dx * PIp_o_PI(im, i); }
    double C1D(int j, int i) const {
        const double dx = x[i + 1] - x[i - 1], dxp = x[i + 1] - x[j], dxm = x[j] - x[i - 1];
        return (dxp - dxm) / (dxp * dxm) * (6.0 - 4.0 * dx * dx / (dxp * dxm)) +
               2.0 * dx * (dxm / dxp - dxp / dxm) * PIp_o_PI(i - 1, i) * PIp_o_PI(i + 1, i) +
               PIp_o_PI(i - 1, i) * (4.0 * dx / dxp - 4.0 * dx / dxm - 2.0 * dx * dx / (dxp * dxp)) -
               PIp_o_PI(i + 1, i) * (4.0 * dx / dxp - 4.0 * dx / dxm + 2.0 * dx * dx / (dxm * dxm));
    }
    double C2D(int j, int i) const {
        const double dx = x[i + 1] - x[i - 1], dxp = x[i + 1] - x[j], dxm = x[j] - x[i - 1];
        return 2.0 * (1.0 / (dxp * dxp) + 1.0 / (dxm * dxm) - 1.0 / (dxp * dxm)) +
               2.0 * dx * dx / (dxp * dxm) * PIp_o_PI(i + 1, i) * PIp_o_PI(i - 1, i) -
               2.0 * PIp_o_PI(i + 1, i) * dx / dxm * (1.0 / dxp - 1.0 / dxm) - 2.0 * PIp_o_PI(i - 1, i) * dx / dxp * (1.0 / dxp - 1.0 / dxm);
    }
    double a2n6(int im, int ip, int i) const {
        const double dx = x[ip] - x[im], dxp = x[i] - x[ip], dxm = x[i] - x[im];
        double f1 = B1D(ip, im, i) * (dxm + dxp) + B2D(im, ip, i) * dxp * (dxp + 2.0 * dxm);
        f1 = f1 * 2.0 * Pi_p(i, {i - 2, i, i + 2});
        double f2 = B1D(ip, im, i) + B2D(im, ip, i) * dxp;
        f2 = f2 * Pi_pp_3(i, i - 2, i, i + 2) * dxp * dxm;
        return -(f1 + f2) / dx / Pi(ip, {i - 2, i, i + 2});
    }
    double b2n6(int im, int ip, int i) const {
        const double dx = x[ip] - x[im], dxp = x[i] - x[ip], dxm = x[i] - x[im];
        const double D = D_coef(i);
        double f1 = 1.0 + A1D(im, ip, i) / D * (dxm + dxp) + A2D(im, ip, i) / D * dxp * (dxp + 2.0 * dxm);
        f1 = f1 * 2.0 * Pi_p(i, {i - 2, i, i + 2});
        double f2 = 1.0 + A1D(im, ip, i) / D * dxp + A2D(im, ip, i) / D * dxp * dxp;
        f2 = f2 * Pi_pp_3(i, i - 2, i, i + 2) * dxm;
        return (f1 + f2) / dx / Pi(ip, {i - 2, i, i + 2});
    }
    double c2n6(int j, int i) const {
        const double dx = x[i] - x[j], dxp = x[i] - x[i + 1], dxm = x[i] - x[i - 1], dxp2 = x[j] - x[i + 1], dxm2 = x[j] - x[i - 1];
        const double D = D_coef(i);
        double f1 = C1D(j, i) / D * (1.0 + dx / dxp + dx / dxm) + C2D(j, i) / D * (2.0 + dx / dxp + dx / dxm) * dx + 1.0 / dxp + 1.0 / dxm;
        f1 = f1 * 2.0 * Lag_p(i, j, {i - 2, i, i + 2}) * dxp * dxm / (dxp2 * dxm2);
        double f2 = 1.0 + C1D(j, i) / D * dx + C2D(j, i) / D * dx * dx;
        f2 = f2 * Lag_pp_3(j, j, {i - 2, i, i + 2}) * dxp * dxm / (dxp2 * dxm2);
        return f1 + f2;
    }
    void c2n4(int i, double (&c)[6]) const {
        const double dx = x[i + 1] - x[i - 1], dxp = x[i + 1] - x[i], dxm = x[i] - x[i - 1];
        const double D = dxp * dxm + dx * dx;
        c[0] = (dxm * dxm - dxp * dxp + dxp * dxm) * dxp / dx / D;
        c[1] = 1.0;
        c[2] = (dxp * dxp - dxm * dxm + dxp * dxm) * dxm / dx / D;
        c[3] = dxp / dx * 12.0 / D;
        c[4] = -12.0 / D;
        c[5] = dxm / dx * 12.0 / D;
    }
    void c2n3_biased(int i, bool backwards, double (&c)[6]) const {
        const int i1 = i, i2 = backwards ? i - 1 : i + 1, i3 = backwards ? i - 2 : i + 2, i4 = backwards ? i - 3 : i + 3;
        const double dx1 = x[i2] - x[i1], dx3 = x[i2] - x[i3], dx4 = x[i2] - x[i4];
        auto M = std::initializer_list<int>{i1, i3, i4};
        const double a1 = 1.0;
        const double a2 = (0.5 * dx1 * Pi_pp_3(i1, i1, i3, i4) - Pi_p(i1, M)) / Pi_p(i2, M);
        double b2 = Pi_pp_3(i1, i1, i3, i4) + 0.5 * dx1 * Pi_pp_3(i1, i1, i3, i4) * Pi_pp_3(i2, i1, i3, i4) / Pi_p(i2, M) -
                    Pi_p(i1, M) / Pi_p(i2, M) * Pi_pp_3(i2, i1, i3, i4);
        b2 = b2 / Pi(i2, M);
        double D = Lag(i2, i1, M) + dx1 * Lag_p(i2, i1, M);
        double b1 = -2.0 * Lag_p(i1, i1, M) * (Lag(i2, i1, M) + 2.0 * dx1 * Lag_p(i2, i1, M)) + 2.0 * Lag_p(i2, i1, M);
        b1 = b1 / D / dx1 + Lag_pp_3(i1, i1, M);
        D = Lag(i2, i3, M) + dx3 * Lag_p(i2, i3, M);
        double b3 = (Lag(i2, i3, M) + dx1 * Lag_p(i2, i3, M)) * dx1 * Lag_pp_3(i1, i3, M) -
                    2.0 * (Lag(i2, i3, M) + 2.0 * dx1 * Lag_p(i2, i3, M)) * Lag_p(i1, i3, M);
        b3 = b3 / D / dx3;
        D = Lag(i2, i4, M) + dx4 * Lag_p(i2, i4, M);
        double b4 = (Lag(i2, i4, M) + dx1 * Lag_p(i2, i4, M)) * dx1 * Lag_pp_3(i1, i4, M) -
                    2.0 * (Lag(i2, i4, M) + 2.0 * dx1 * Lag_p(i2, i4, M)) * Lag_p(i1, i4, M);
        b4 = b4 / D / dx4;
        c[0] = a1; c[1] = a2; c[2] = b1; c[3] = b2; c[4] = b3; c[5] = b4;
    }
};

// lhs(1..n, 1..3), rhs(1..n, 1..5); the fourth coefficient of the first / last row sits in rhs(1, 1) / rhs(n, 5)
void c2n6_direct(const std::vector<double>& nodes0, int n, Mat& lhs, Mat& rhs) {
    std::vector<double> xp(n + 1, 0.0);
    for (int i = 1; i <= n; i++) xp[i] = nodes0[i - 1];
    DirectC2N6 d{xp.data()};
    double c[6];
    d.c2n3_biased(1, false, c);
    double dummy = 1.0 / c[2];
    lhs(1, 2) = c[0] * dummy; lhs(1, 3) = c[1] * dummy;
    rhs(1, 3) = c[2] * dummy; rhs(1, 4) = c[3] * dummy; rhs(1, 5) = c[4] * dummy; rhs(1, 1) = c[5] * dummy;
    d.c2n3_biased(n, true, c);
    dummy = 1.0 / c[2];
    lhs(n, 2) = c[0] * dummy; lhs(n, 1) = c[1] * dummy;
    rhs(n, 3) = c[2] * dummy; rhs(n, 2) = c[3] * dummy; rhs(n, 1) = c[4] * dummy; rhs(n, 5) = c[5] * dummy;
    for (int i : {2, n - 1}) {
        d.c2n4(i, c);
        dummy = 1.0 / c[4];
        for (int j = 0; j < 3; j++) { lhs(i, 1 + j) = c[j] * dummy; rhs(i, 2 + j) = c[3 + j] * dummy; }
    }
    for (int i = 3; i <= n - 2; i++) {
        const double D = d.D_coef(i);
        const double a = 1.0, ap1 = d.a2n6(i - 1, i + 1, i) / D, am1 = d.a2n6(i + 1, i - 1, i) / D;
        const double bp1 = d.b2n6(i - 1, i + 1, i), bm1 = d.b2n6(i + 1, i - 1, i);
        const double dxp = xp[i] - xp[i + 1], dxm = xp[i] - xp[i - 1];
        const double lp = d.Lag_p(i, i, {i - 2, i, i + 2});
        const double b = 2.0 * d.C2D(i, i) / D + 2.0 * d.C1D(i, i) / D * ((dxm + dxp) / (dxp * dxm) + lp) +
                         (2.0 + 2.0 * lp * (dxm + dxp)) / (dxp * dxm) + d.Lag_pp_3(i, i, {i - 2, i, i + 2});
        const double bp2 = d.c2n6(i + 2, i), bm2 = d.c2n6(i - 2, i);
        dummy = 1.0 / bp1;
        lhs(i, 1) = am1 * dummy; lhs(i, 2) = a * dummy; lhs(i, 3) = ap1 * dummy;
        rhs(i, 1) = bm2 * dummy; rhs(i, 2) = bm1 * dummy; rhs(i, 3) = b * dummy; rhs(i, 4) = bp1 * dummy; rhs(i, 5) = bp2 * dummy;
    }
}

void der2_initialize_direct(const std::vector<double>& nodes0, int n, HostDer& g) {
    g.size = n; g.periodic = false; g.ndl = 3; g.ndr = 5;
    for (double& c : g.coef) c = 0.0;
    g.lhs = Mat(1, n, 1, 5);
    g.rhs = Mat(1, n, 1, 12);
    c2n6_direct(nodes0, n, g.lhs, g.rhs);
    g.need_1der = false;
    g.mwn.assign(n, 0.0);
    g.lu = Mat(1, n, 1, 3);
    for (int i = 1; i <= n; i++) for (int j = 1; j <= 3; j++) g.lu(i, j) = g.lhs(i, j);
    tridfs(n, {&g.lu, 1, 1}, {&g.lu, 2, 1}, {&g.lu, 3, 1});
}

}  // namespace

void der1_solve_line(const HostDer& g, int ibc, const double* u, double* result) {
    const int n = g.size;
    int ibc_loc = g.periodic ? BCS_PERIODIC : ibc;
    int nmin = 1, nmax = n;
    if (ibc_loc == BCS_ND || ibc_loc == BCS_NN) { result[0] = 0.0; nmin++; }
    if (ibc_loc == BCS_DN || ibc_loc == BCS_NN) { result[n - 1] = 0.0; nmax--; }
    matmul_line(g, false, ibc_loc, u, result);
    if (g.periodic) tridpss_line(n, g.lu, result);
    else tridss_line(nmax - nmin + 1, g.lu, nmin, ibc * 5, result + (nmin - 1));
}

void der2_solve_line(const HostDer& g, const double* u, const double* du, double* result) {
    const int n = g.size;
    matmul_line(g, true, g.periodic ? BCS_PERIODIC : BCS_DD, u, result);
    if (g.need_1der) {
        const int ip = g.ndr;
        auto R = [&](int i, int j) { return g.rhs(i, ip + j); };
        result[0] += du[0] * R(1, 2) + du[1] * R(1, 3) + du[2] * R(1, 1);
        for (int i = 2; i <= n - 1; i++) result[i - 1] += du[i - 2] * R(i, 1) + du[i - 1] * R(i, 2) + du[i] * R(i, 3);
        result[n - 1] += du[n - 3] * R(n, 3) + du[n - 2] * R(n, 1) + du[n - 1] * R(n, 2);
    }
    if (g.periodic) tridpss_line(n, g.lu, result);
    else tridss_line(n, g.lu, 1, 0, result);
}

int create_plan(const double* nodes, int n, bool periodic, bool uniform, int mode1, int mode2, HostPlan& g) {
    if (periodic && !uniform) return 85;   // DNS_ERROR_OPTION: grid must be uniform in a periodic direction
    if (mode1 != FDM_COM4_JACOBIAN && mode1 != FDM_COM6_JACOBIAN) return 104;   // DNS_ERROR_UNDEVELOP
    if (periodic && mode2 == FDM_COM6_DIRECT) mode2 = FDM_COM6_JACOBIAN_HYPER;      // the same on uniform grids (fdm.f90:158)
    if (mode2 != FDM_COM4_JACOBIAN && mode2 != FDM_COM6_JACOBIAN && mode2 != FDM_COM6_JACOBIAN_HYPER && mode2 != FDM_COM6_DIRECT) return 104;
    g.size = n; g.periodic = periodic; g.uniform = uniform;
    g.der1 = HostDer(); g.der2 = HostDer();
    g.der1.mode_fdm = mode1; g.der2.mode_fdm = mode2;
    g.nodes.assign(nodes, nodes + n);
    g.jac = Mat(1, std::max(n, 1), 1, 3);
    if (n > 1) {
        g.scale = nodes[n - 1] - nodes[0];
        if (periodic) g.scale = g.scale * (1.0 + 1.0 / double(n - 1));
    } else {
        g.scale = 1.0;
        for (int j = 1; j <= 3; j++) g.jac(1, j) = 1.0;
        g.der1.size = g.der2.size = 1;
        return 0;
    }
    if (n < 8) return 48;   // DNS_ERROR_DIMGRID: boundary closures need at least 8 points
    std::vector<double> one(n + 1, 1.0), zero(n + 1, 0.0), tmp(n), j1(n + 1), j2(n + 1);
    // dx/ds from the first-derivative scheme applied to the nodes on a unit computational grid
    der1_initialize(one, n, g.der1, false, {BCS_DD});
    der1_solve_line(g.der1, BCS_DD, nodes, tmp.data());
    for (int i = 1; i <= n; i++) { g.jac(i, 1) = tmp[i - 1]; j1[i] = tmp[i - 1]; }
    der1_initialize(j1, n, g.der1, periodic, {BCS_DD, BCS_ND, BCS_DN, BCS_NN});
    if (periodic) for (double& w : g.der1.mwn) w = w / g.jac(1, 1);
    // d2x/ds2 from the second-derivative scheme on the unit grid
    if (mode2 == FDM_COM6_DIRECT) {
        std::vector<double> unit(n);
        for (int i = 0; i < n; i++) unit[i] = double(i);
        der2_initialize_direct(unit, n, g.der2);
    } else {
        der2_initialize(one, zero, n, g.der2, false, true);
    }
    der2_solve_line(g.der2, nodes, nodes, tmp.data());
    for (int i = 1; i <= n; i++) { g.jac(i, 3) = tmp[i - 1]; j2[i] = tmp[i - 1]; g.jac(i, 2) = g.jac(i, 1); }
    g.der2.need_1der = false;
    if (mode2 == FDM_COM6_DIRECT) der2_initialize_direct(g.nodes, n, g.der2);
    else der2_initialize(j1, j2, n, g.der2, periodic, uniform);
    if (periodic) for (double& w : g.der2.mwn) w = w / (g.jac(1, 1) * g.jac(1, 1));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Reduction of the first/last row of a banded system into its neighbours (used by the integral
// operators).  lhs(1..n, 1..ndl) in place; rhs(1..n, 1..ndr); rhs_b rows 1.., cols 0..;
// rhs_t rows 0.., cols 1..
void fdm_bcs_reduce(int ibc, Mat& lhs, const Mat& rhs, Mat* rhs_b, Mat* rhs_t) {
    const int ndl = lhs.c1, idl = ndl / 2 + 1, ndr = rhs.c1, idr = ndr / 2 + 1, nx = lhs.r1;
    const int nx_t = idr, m = std::max(idl, idr + 1);
    if (ibc == BCS_MIN || ibc == BCS_BOTH) {
        const double dummy = 1.0 / lhs(1, idl);
        for (int j = 1; j <= ndl; j++) lhs(1, j) = -lhs(1, j) * dummy;
        lhs(1, idl) = 1.0;
        for (int ir = 1; ir <= idl - 1; ir++) {
            for (int ic = idl + 1; ic <= ndl; ic++)
                lhs(1 + ir, ic - ir) = lhs(1 + ir, ic - ir) + lhs(1 + ir, idl - ir) * lhs(1, ic);
            int ic = ndl + 1;
            lhs(1 + ir, ic - ir) = lhs(1 + ir, ic - ir) + lhs(1 + ir, idl - ir) * lhs(1, 1);
        }
        if (rhs_b) {
            Mat& B = *rhs_b;
            for (int i = 1; i <= m; i++) for (int j = 1; j <= ndr; j++) B(i, j) = rhs(i, j);
            for (int j = 1; j <= ndr; j++) B(1, j) = B(1, j) * dummy;
            for (int ir = 1; ir <= idl - 1; ir++) {
                for (int ic = idr; ic <= ndr; ic++)
                    B(1 + ir, ic - ir) = B(1 + ir, ic - ir) - lhs(1 + ir, idl - ir) * B(1, ic);
                int ic = ndr + 1;
                B(1 + ir, ic - ir) = B(1 + ir, ic - ir) - lhs(1 + ir, idl - ir) * B(1, 1);
            }
        }
    }
    if (ibc == BCS_MAX || ibc == BCS_BOTH) {
        const double dummy = 1.0 / lhs(nx, idl);
        for (int j = 1; j <= ndl; j++) lhs(nx, j) = -lhs(nx, j) * dummy;
        lhs(nx, idl) = 1.0;
        for (int ir = 1; ir <= idl - 1; ir++) {
            int ic = 0;
            lhs(nx - ir, ic + ir) = lhs(nx - ir, ic + ir) + lhs(nx - ir, idl + ir) * lhs(nx, ndl);
            for (ic = 1; ic <= idl - 1; ic++)
                lhs(nx - ir, ic + ir) = lhs(nx - ir, ic + ir) + lhs(nx - ir, idl + ir) * lhs(nx, ic);
        }
        if (rhs_t) {
            Mat& T = *rhs_t;
            for (int i = 1; i <= m; i++) for (int j = 1; j <= ndr; j++) T(nx_t - m + i, j) = rhs(nx - m + i, j);
            for (int j = 1; j <= ndr; j++) T(nx_t, j) = T(nx_t, j) * dummy;
            for (int ir = 1; ir <= idl - 1; ir++) {
                int ic = 0;
                T(nx_t - ir, ic + ir) = T(nx_t - ir, ic + ir) - lhs(nx - ir, idl + ir) * T(nx_t, ndr);
                for (ic = 1; ic <= idr; ic++)
                    T(nx_t - ir, ic + ir) = T(nx_t - ir, ic + ir) - lhs(nx - ir, idl + ir) * T(nx_t, ic);
            }
        }
    }
}

// lambda-affine split of the integral system (before the reduction at the opposite end)
int int1_create_base(const HostDer& g, int ibc, HostInt1& o) {
    const int ndl = g.ndl, idl = ndl / 2 + 1, ndr = g.ndr, idr = ndr / 2 + 1, nx = g.size;
    if (ndl != 3 || ndr != 5) return 104;     // only the tridiagonal 6th-order first derivative
    o.n = nx; o.bc = ibc;
    o.L0 = Mat(1, nx, 1, ndr); o.L1 = Mat(1, nx, 1, ndr);
    o.rhs = Mat(1, nx, 1, ndl);
    o.rhs_b0 = Mat(1, 5, 0, 7); o.rhs_t0 = Mat(0, 4, 1, 8);
    Mat grhs(1, nx, 1, ndr);
    for (int i = 1; i <= nx; i++) {
        for (int j = 1; j <= ndl; j++) o.rhs(i, j) = g.lhs(i, j);
        for (int j = 1; j <= ndr; j++) grhs(i, j) = g.rhs(i, j);
    }
    Mat rhsr_b(1, 5, 0, 7), rhsr_t(0, 4, 1, 8);
    fdm_bcs_reduce(ibc, o.rhs, grhs, &rhsr_b, &rhsr_t);
    Mat& rb = o.rhs_b0; Mat& rt = o.rhs_t0;
    if (ibc == BCS_MIN) {
        for (int i = 1; i <= idl + 1; i++) for (int j = 1; j <= ndl; j++) rb(i, j) = o.rhs(i, j);
        for (int ir = 1; ir <= idr - 1; ir++) rb(1 + ir, idl - ir) = -rhsr_b(1 + ir, idr - ir);
    } else {
        for (int i = 0; i <= idl; i++) for (int j = 1; j <= ndl; j++) rt(i, j) = o.rhs(nx - idl + i, j);
        for (int ir = 1; ir <= idr - 1; ir++) rt(idl - ir, idl + ir) = -rhsr_t(idr - ir, idr + ir);
    }
    // C = B + lambda A, split as L0 + lambda L1
    for (int i = 1; i <= nx; i++) {
        for (int j = 1; j <= ndr; j++) o.L0(i, j) = g.rhs(i, j);
        o.L1(i, idr) = g.lhs(i, idl);
    }
    for (int k = 1; k <= idl - 1; k++) {
        for (int i = 1 + k; i <= nx; i++) o.L1(i, idr - k) = g.lhs(i, idl - k);
        for (int i = 1; i <= nx - k; i++) o.L1(i, idr + k) = g.lhs(i, idl + k);
    }
    if (ibc == BCS_MIN) {
        for (int i = 1; i <= idr; i++) for (int j = 1; j <= ndr; j++) { o.L0(i, j) = rhsr_b(i, j); o.L1(i, j) = 0.0; }
        for (int j = 1; j <= idl - 1; j++) o.L1(1, idr + j) = -rb(1, idl + j);
        for (int ir = 1; ir <= idr - 1; ir++)
            for (int j = 1; j <= ndl; j++) o.L1(1 + ir, idr - idl + j) = rb(1 + ir, j);
    } else {
        for (int i = 1; i <= idr; i++) for (int j = 1; j <= ndr; j++) { o.L0(nx - idr + i, j) = rhsr_t(i, j); o.L1(nx - idr + i, j) = 0.0; }
        for (int j = 1; j <= idl - 1; j++) o.L1(nx, idr - idl + j) = -rt(idl, j);
        for (int ir = 1; ir <= idr - 1; ir++)
            for (int j = 1; j <= ndl; j++) o.L1(nx - ir, idr - idl + j) = rt(idl - ir, j);
    }
    // normalisation (lambda-independent factors)
    const int m = std::max(idr, idl + 1);
    for (int ir = 1; ir <= m; ir++) {
        double dummy = 1.0 / o.rhs(ir, idl);
        for (int j = 0; j <= ndl; j++) rb(ir, j) = rb(ir, j) * dummy;
        dummy = 1.0 / o.rhs(nx - ir + 1, idl);
        for (int j = 1; j <= ndl + 1; j++) rt(idl - ir + 1, j) = rt(idl - ir + 1, j) * dummy;
        dummy = 1.0 / o.rhs(ir, idl);
        for (int j = 1; j <= ndl; j++) o.rhs(ir, j) = o.rhs(ir, j) * dummy;
        for (int j = 1; j <= ndr; j++) { o.L0(ir, j) = o.L0(ir, j) * dummy; o.L1(ir, j) = o.L1(ir, j) * dummy; }
        dummy = 1.0 / o.rhs(nx - ir + 1, idl);
        for (int j = 1; j <= ndl; j++) o.rhs(nx - ir + 1, j) = o.rhs(nx - ir + 1, j) * dummy;
        for (int j = 1; j <= ndr; j++) { o.L0(nx - ir + 1, j) = o.L0(nx - ir + 1, j) * dummy; o.L1(nx - ir + 1, j) = o.L1(nx - ir + 1, j) * dummy; }
    }
    for (int ir = m + 1; ir <= nx - m; ir++) {
        const double dummy = 1.0 / o.rhs(ir, idl + 1);
        for (int j = 1; j <= ndl; j++) o.rhs(ir, j) = o.rhs(ir, j) * dummy;
        for (int j = 1; j <= ndr; j++) { o.L0(ir, j) = o.L0(ir, j) * dummy; o.L1(ir, j) = o.L1(ir, j) * dummy; }
    }
    return 0;
}

}  // namespace tlab
