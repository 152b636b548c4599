// Host-side construction of the compact finite-difference plans (C++17).
//
// This is the product's own plan builder: what the Fortran host does in
// FDM_CreatePlan (reference src/fdm/fdm.f90:143-252) before any operator is
// called.  A Fortran caller can instead hand its already-built arrays to
// tlab_fdm_plan_create_from_arrays(); a caller without the Fortran host (the
// Python mirror, the tests, bench.py) builds them here from the grid nodes.
// The kernels only ever see the uploaded tables (plan.h).
#pragma once
#include <vector>
#include <cassert>
#include <cstddef>

namespace tlab {

// Dense 2-D table with arbitrary inclusive index ranges (rows r0..r1, cols c0..c1),
// so that the index conventions of the scheme definitions can be used directly.
struct Mat {
    int r0 = 1, r1 = 0, c0 = 1, c1 = 0;
    std::vector<double> v;
    Mat() {}
    Mat(int r0_, int r1_, int c0_, int c1_) : r0(r0_), r1(r1_), c0(c0_), c1(c1_),
        v((size_t)(r1_ - r0_ + 1) * (c1_ - c0_ + 1), 0.0) {}
    int ncol() const { return c1 - c0 + 1; }
    int nrow() const { return r1 - r0 + 1; }
    double& operator()(int i, int j) {
        assert(i >= r0 && i <= r1 && j >= c0 && j <= c1);
        return v[(size_t)(i - r0) * ncol() + (j - c0)];
    }
    double operator()(int i, int j) const {
        assert(i >= r0 && i <= r1 && j >= c0 && j <= c1);
        return v[(size_t)(i - r0) * ncol() + (j - c0)];
    }
};

enum { BCS_PERIODIC = -1, BCS_DD = 0, BCS_ND = 1, BCS_DN = 2, BCS_NN = 3 };
enum { BCS_NONE = 0, BCS_MIN = 1, BCS_MAX = 2, BCS_BOTH = 3 };
// scheme codes (reference fdm_derivative.f90:51-58)
enum { FDM_COM4_JACOBIAN = 4, FDM_COM6_JACOBIAN_PENTA = 5, FDM_COM6_JACOBIAN = 6, FDM_COM6_JACOBIAN_HYPER = 7, FDM_COM6_DIRECT = 16 };

struct HostDer {
    int mode_fdm = 0;
    int size = 0;
    bool periodic = false;
    bool need_1der = false;
    int ndl = 0, ndr = 0;          // # of lhs / rhs diagonals
    Mat lhs, rhs;                  // (1..n, 1..ndl), (1..n, 1..ndr[+ndl])
    Mat rhs_b, rhs_t;              // (1..4, 0..7), (0..4, 1..7): Neumann-reduced boundary rows
    std::vector<double> mwn;       // modified wavenumbers (periodic)
    Mat lu;                        // der1: (1..n, 1..20) biased or (1..n, 1..ndl+2) periodic; der2: (1..n, 1..3|5)
    double coef[5] = {0, 0, 0, 0, 0};
};

struct HostPlan {
    int size = 0;
    bool periodic = false, uniform = false;
    double scale = 1.0;
    std::vector<double> nodes;     // 0-based
    Mat jac;                       // (1..n, 1..3)
    HostDer der1, der2;
};

// returns 0 or a DNS_ERROR_* code
int create_plan(const double* nodes, int n, bool periodic, bool uniform, int mode1, int mode2, HostPlan& g);

// line solves on the host (used to obtain the Jacobians; n-major single line)
void der1_solve_line(const HostDer& g, int ibc, const double* u, double* result);
void der2_solve_line(const HostDer& g, const double* u, const double* du, double* result);

// First-order integral operator u' + lambda u = f (reference fdm_integral.f90:91-214), split into the
// lambda-independent and the lambda-proportional part of the system *before* the reduction at the
// opposite end, so that a device thread can assemble lhs = L0 + lambda*L1 for its own eigenvalue.
struct HostInt1 {
    int n = 0, bc = 0;
    Mat L0, L1;        // (1..n, 1..5)
    Mat rhs;           // (1..n, 1..3)   lambda-independent, normalised
    Mat rhs_b0;        // (1..5, 0..7)   for BCS_MIN: final; for BCS_MAX: before the opposite-end reduction
    Mat rhs_t0;        // (0..4, 1..8)
};
int int1_create_base(const HostDer& g, int ibc, HostInt1& out);

void fdm_bcs_reduce(int ibc, Mat& lhs, const Mat& rhs, Mat* rhs_b, Mat* rhs_t);

}  // namespace tlab
