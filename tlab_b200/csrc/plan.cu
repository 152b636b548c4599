// Build the device tables of a plan from the host plan (see plan.h).
#include "plan.h"
#include "lines2.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace tlab {

namespace {

const double* upload(DevPlan& p, const std::vector<double>& v) {
    double* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(double)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
    p.allocs.push_back(d);
    return d;
}

int chunk_cnt(const DevPlan& p, int t) { return p.cbase + (t < p.crem ? 1 : 0); }

// smallest look-back window W such that any product of W consecutive chunk multipliers is < 2^-80
int window(const std::vector<double>& A, bool forward) {
    const int T = (int)A.size();
    const double tiny = std::ldexp(1.0, -80);
    for (int W = 1; W < T; W++) {
        bool ok = true;
        for (int t = 0; t < T && ok; t++) {
            // chunks feeding chunk t: forward t-W..t-1, backward t+1..t+W; only complete windows matter
            int lo = forward ? t - W : t + 1, hi = forward ? t - 1 : t + W;
            if (lo < 0 || hi > T - 1) continue;     // window reaches the end of the line: nothing is dropped
            double prod = 1.0;
            for (int k = lo; k <= hi; k++) prod *= std::fabs(A[k]);
            if (!(prod < tiny)) ok = false;
        }
        if (ok) return W;
    }
    return std::max(T - 1, 0);
}

template <class T>
const T* upload_t(DevPlan& p, const std::vector<T>& v) {
    T* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    p.allocs.push_back(d);
    return d;
}

bool nearly_const(const std::vector<double>& v, int s0, int c) {
    double lo = v[s0], hi = v[s0];
    for (int j = 1; j < c; j++) { lo = std::min(lo, v[s0 + j]); hi = std::max(hi, v[s0 + j]); }
    return (hi - lo) <= std::ldexp(std::max(std::fabs(lo), std::fabs(hi)), -50);
}

void finish_solve(DevPlan& p, SolveTab& s, std::vector<double>& alpha, std::vector<double>& beta,
                  std::vector<double>& gamma, std::vector<double>& delta, std::vector<double>& pd,
                  std::vector<double>& pe, bool periodic) {
    const int T = p.T;
    std::vector<double> Af(T, 1.0), Ab(T, 1.0);
    const int Tp = (T + REC_GROUP - 1) / REC_GROUP * REC_GROUP;
    std::vector<double2> rec((size_t)Tp * CHUNK * 2, make_double2(0.0, 1.0));
    std::vector<double> pdv((size_t)Tp * CHUNK, 0.0);
    std::vector<ChunkDesc> cd(T);
    double pdmax = 0.0, pemax = 0.0;
    for (size_t i = 0; i < pd.size(); i++) { pdmax = std::max(pdmax, std::fabs(pd[i])); pemax = std::max(pemax, std::fabs(pe[i])); }
    for (int t = 0; t < T; t++) {
        const int s0 = chunk_start(p, t), c = chunk_cnt(p, t);
        const int j0 = (t == T - 1) ? CHUNK - c : 0;
        for (int j = 0; j < c; j++) {
            const size_t r = ((size_t)(t / REC_GROUP) * CHUNK + (j0 + j)) * REC_GROUP + t % REC_GROUP;
            rec[2 * r] = make_double2(alpha[s0 + j], beta[s0 + j]);
            rec[2 * r + 1] = make_double2(gamma[s0 + j], periodic ? pe[s0 + j] : delta[s0 + j]);
            pdv[r] = pd[s0 + j];
        }
        double w = 1.0;
        for (int j = 0; j < c; j++) w *= alpha[s0 + j];
        Af[t] = w;
        w = 1.0;
        for (int j = c - 1; j >= 0; j--) w *= periodic ? gamma[s0 + j] : gamma[s0 + j] * delta[s0 + j];
        Ab[t] = w;
        ChunkDesc& d = cd[t];
        d.a = alpha[s0]; d.b = beta[s0]; d.g = gamma[s0]; d.d = periodic ? 1.0 : delta[s0];
        d.Af = Af[t]; d.Ab = Ab[t]; d.flags = 0; d.pad = 0;
        // terms below 2^-80 of the largest coefficient are dropped, like the look-back tails
        bool pd0 = true, pe0 = true;
        for (int j = 0; j < c; j++) {
            if (std::fabs(pd[s0 + j]) > std::ldexp(pdmax, -80)) pd0 = false;
            if (std::fabs(pe[s0 + j]) > std::ldexp(pemax, -80)) pe0 = false;
        }
        if (!periodic) { pd0 = true; pe0 = true; }
        if (c == CHUNK && nearly_const(alpha, s0, c) && nearly_const(beta, s0, c)) d.flags |= CD_FWD_CONST;
        if (c == CHUNK && nearly_const(gamma, s0, c) && (periodic ? pe0 : nearly_const(delta, s0, c))) d.flags |= CD_BWD_CONST;
        if (pd0) d.flags |= CD_PD_ZERO;
    }
    s.Wf = window(Af, true);
    s.Wb = window(Ab, false);
    s.rec = upload_t(p, rec);
    s.cd = upload_t(p, cd);
    s.pd = periodic ? upload(p, pdv) : nullptr;
}


// ---- the same system for lines2.cu (see plan.h, Sys2) ------------------------------------------
void build_sys2(DevPlan& p, Sys2& s2, const std::vector<double>& alpha, const std::vector<double>& beta,
                const std::vector<double>& gamma, const std::vector<double>& delta, const std::vector<double>& pd,
                const std::vector<double>& pe, double bN, bool periodic, const std::vector<double>* sdiag = nullptr) {
    s2 = Sys2();
    const int n = p.n, T = p.T;
    if (p.crem != 0 || p.cbase != CHUNK) return;            // fast kernels need full chunks
    std::vector<double> a(n), d(n), g(n), e(n, 0.0), pp(n, 0.0);
    for (int i = 0; i < n; i++) {
        if (periodic) {
            a[i] = (i > 0) ? alpha[i] * beta[i - 1] / beta[i] : 0.0;
            d[i] = beta[i]; g[i] = gamma[i]; e[i] = pe[i];
            pp[i] = -bN * pd[i] * beta[i];
        } else {
            a[i] = alpha[i]; d[i] = delta[i]; g[i] = gamma[i] * delta[i];
        }
    }
    if (periodic) { d[n - 1] = 0.0; g[n - 1] = 0.0; e[n - 1] = 1.0; pp[n - 1] = bN * beta[n - 1]; a[n - 1] = alpha[n - 1] * beta[n - 2] / beta[n - 1]; }
    // global backward sweep of the e vector: S_i = e_i + g_i S_{i+1}
    std::vector<double> S(n + 1, 0.0);
    for (int i = n - 1; i >= 0; i--) S[i] = e[i] + g[i] * S[i + 1];
    std::vector<double> P(n), Q(n), R(n), Af(T), Rb(T), Q0(T), PP(T);
    for (int t = 0; t < T; t++) {
        const int s0 = t * CHUNK;
        double w = 1.0, acc = 0.0;
        for (int j = 0; j < CHUNK; j++) { w *= a[s0 + j]; P[s0 + j] = w; acc += pp[s0 + j] * w; }
        Af[t] = w; PP[t] = acc;
        w = 1.0;
        double q = 0.0;
        for (int j = CHUNK - 1; j >= 0; j--) { w *= g[s0 + j]; R[s0 + j] = w; q = d[s0 + j] * P[s0 + j] + g[s0 + j] * q; Q[s0 + j] = q; }
        Rb[t] = w; Q0[t] = q;
    }
    // look-back windows must fit LB2 chunks
    if (window(Af, true) > LB2 || window(Rb, false) > LB2) return;
    // constant chunks: coefficients equal to the converged LU factors (interior of a uniform direction).  The reference
    // builds the Jacobian by differentiating the node positions, which leaves round-off noise of relative size
    // ~eps*n in its LU factors; the constants are the mean over the middle half of the line and a chunk counts as
    // constant if it deviates by less than 2^-41 (4.5e-13) from them -- far inside the 1e-12 parity tolerance.
    // Non-periodic directions, any grid: the reference's matrix is A0 diag(s) with A0 the constant-coefficient matrix of the scheme
    // (boundary closures in its first and last rows) and s_j the Jacobian factor of column j (dx_j or dx_j^2: stretched grids,
    // or the round-off noise of a uniform one).  Its LU therefore has the CLEAN forward multipliers of A0, d_j = d0_j / s_j and
    // g_j = g0_j s_{j+1} / s_j: away from the walls the factors of A0 have converged, and a chunk is a constant chunk in the
    // variable w_j = x_j / rho_j, rho_j = s_m / s_j (m = n/2 the reference row).  Only the chunks next to the walls keep tables.
    const bool unscaled = !periodic && sdiag != nullptr && n >= 4 * CHUNK;
    std::vector<double> rho(n, 1.0);
    if (unscaled) {
        const std::vector<double>& sd = *sdiag;
        const int m = n / 2;
        s2.ca = a[m]; s2.cd = d[m]; s2.cg = g[m] * sd[m] / sd[m + 1];
        for (int i = 0; i < n; i++) rho[i] = sd[m] / sd[i];
    } else {
        // means as reference value + mean deviation: a plain running sum of 512 equal numbers is already off by ~1e-14
        const double ra = a[n / 2], rd = d[n / 2], rg = g[n / 2];
        double sa = 0.0, sd = 0.0, sg = 0.0;
        int cnt = 0;
        for (int i = n / 4; i < n - n / 4; i++) { sa += a[i] - ra; sd += d[i] - rd; sg += g[i] - rg; cnt++; }
        s2.ca = ra + sa / cnt; s2.cd = rd + sd / cnt; s2.cg = rg + sg / cnt;
    }
    {
        double w = 1.0, q = 0.0;
        std::vector<double> Pc(CHUNK);
        for (int j = 0; j < CHUNK; j++) { w *= s2.ca; Pc[j] = w; }
        w = 1.0;
        for (int j = CHUNK - 1; j >= 0; j--) { w *= s2.cg; s2.cR[j] = w; q = s2.cd * Pc[j] + s2.cg * q; s2.cQ[j] = q; }
    }
    double pmax = 0.0, smax = 0.0;
    for (int i = 0; i < n; i++) { pmax = std::max(pmax, std::fabs(pp[i])); smax = std::max(smax, std::fabs(S[i])); }
    auto close = [](double v, double ref) { return std::fabs(v - ref) <= std::ldexp(std::fabs(ref), -41); };
    auto close48 = [](double v, double ref) { return std::fabs(v - ref) <= std::ldexp(std::fabs(ref), -48); };
    std::vector<int> isc(T, 0);
    for (int t = 0; t < T; t++) {
        bool c = true;
        for (int j = 0; j < CHUNK && c; j++) {
            const int i = t * CHUNK + j;
            if (unscaled) {
                const std::vector<double>& sd = *sdiag;
                c = i + 1 < n && close48(a[i], s2.ca) && close48(d[i] / rho[i], s2.cd) && close48(g[i] * sd[i] / sd[i + 1], s2.cg);
                continue;
            }
            c = close(a[i], s2.ca) && close(d[i], s2.cd) && close(g[i], s2.cg);
            if (periodic && c) c = std::fabs(pp[i]) <= std::ldexp(pmax, -80) && std::fabs(S[i]) <= std::ldexp(smax, -80);
        }
        isc[t] = c ? 1 : 0;
    }
    // circulant closure: chunks [0,K0) and [T-K1,T) carry it; everything in between must be constant (p, S negligible)
    if (periodic) {
        int k0 = 0, k1 = 0;
        auto carries = [&](int t) {
            for (int j = 0; j < CHUNK; j++) {
                const int i = t * CHUNK + j;
                if (std::fabs(pp[i]) > std::ldexp(pmax, -80) || std::fabs(S[i]) > std::ldexp(smax, -80)) return true;
            }
            return false;
        };
        while (k0 < T && carries(k0)) k0++;
        while (k1 < T - k0 && carries(T - 1 - k1)) k1++;
        for (int t = k0; t < T - k1; t++) if (carries(t)) { k0 = T; k1 = 0; break; }
        s2.K0 = k0; s2.K1 = k1;
        for (int t = 0; t < T; t++) if (t < k0 || t >= T - k1) isc[t] = 0;
        // the same at 2^-56 (eps/8) for the marching kernels, whose closure chunks must fit one round
        auto carries56 = [&](int t) {
            for (int j = 0; j < CHUNK; j++) {
                const int i = t * CHUNK + j;
                if (std::fabs(pp[i]) > std::ldexp(pmax, -56) || std::fabs(S[i]) > std::ldexp(smax, -56)) return true;
            }
            return false;
        };
        int m0 = 0, m1 = 0;
        while (m0 < T && carries56(m0)) m0++;
        while (m1 < T - m0 && carries56(T - 1 - m1)) m1++;
        for (int t = m0; t < T - m1; t++) if (carries56(t)) { m0 = T; m1 = 0; break; }
        s2.K0m = m0; s2.K1m = m1;
    }
    // circulant form.  The reference scales the COLUMNS of the constant-coefficient matrix A0 by the Jacobian s_j (dx_j or
    // dx_j^2, computed from the node positions and therefore carrying round-off noise of relative size eps*j: 5e-13 at the
    // end of a line of 1024 points, 3.5e-12 at 2048), i.e. x = diag(1/s) A0^-1 f: the forward multipliers of its LU are the
    // clean constant, d_j = d0 / s_j and g_j = g0 s_{j+1} / s_j.  So the constant-coefficient solution is multiplied point by
    // point by rho_j = (1/s_j) / mean(1/s) and equals the reference's to round-off for any line length.  Checked here: with
    // the noise divided out the factors must be flat to 2^-48 over the middle half of the line.
    if (periodic && T >= 8 && sdiag != nullptr) {
        const std::vector<double>& sd = *sdiag;
        const double rm = 1.0 / sd[n / 2];
        double m = 0.0;
        int cnt = 0;
        for (int i = n / 4; i < n - n / 4; i++) { m += 1.0 / sd[i] - rm; cnt++; }
        m = rm + m / cnt;
        for (int i = 0; i < n; i++) rho[i] = (1.0 / sd[i]) / m;
        bool flat = true;
        for (int i = n / 4; i < n - n / 4 && flat; i++)
            flat = close48(a[i], s2.ca) && close48(d[i] / rho[i], s2.cd) && close48(g[i] * sd[i] / sd[i + 1], s2.cg);
        if (getenv("TLAB_DEBUG_PLAN")) {
            double ea = 0, ed = 0, eg = 0;
            for (int i = n / 4; i < n - n / 4; i++) {
                ea = std::max(ea, std::fabs(a[i] / s2.ca - 1)); ed = std::max(ed, std::fabs(d[i] / rho[i] / s2.cd - 1));
                eg = std::max(eg, std::fabs(g[i] * sd[i] / sd[i + 1] / s2.cg - 1));
            }
            fprintf(stderr, "[circ] n=%d flat=%d dev a=%g d=%g g=%g (2^-48=%g) sd[mid]=%g\n", n, (int)flat, ea, ed, eg, std::ldexp(1.0, -48), sd[n / 2]);
        }
        if (flat) {
            const double af = std::pow(s2.ca, CHUNK), rb = std::pow(s2.cg, CHUNK);
            double wa = 1.0, wb = 1.0;
            for (int k = 0; k < LB2; k++) { s2.cwf[k] = wa; s2.cwb[k] = wb; wa *= af; wb *= rb; }
            if (std::fabs(wa) <= std::ldexp(1.0, -80) && std::fabs(wb) <= std::ldexp(1.0, -80)) {
                const int Tq = (T + 7) / 8 * 8;
                std::vector<double> rt((size_t)Tq * CHUNK, 1.0);
                for (int t = 0; t < T; t++)
                    for (int j = 0; j < CHUNK; j++) rt[((size_t)(t >> 3) * CHUNK + j) * 8 + (t & 7)] = rho[t * CHUNK + j];
                s2.rho = upload(p, rt);
                s2.circ = s2.rho ? 1 : 0;
            }
        }
    }
    const int Tp = (T + 7) / 8 * 8;
    std::vector<double2> tab((size_t)Tp * CHUNK * 4, make_double2(0.0, 0.0));
    std::vector<double> crec((size_t)T * 16, 0.0);
    for (int t = 0; t < T; t++) {
        for (int j = 0; j < CHUNK; j++) {
            const int i = t * CHUNK + j;
            const size_t base = (((size_t)(t >> 3) * CHUNK + j) * 4) * 8 + (t & 7);
            tab[base + 0 * 8] = make_double2(a[i], pp[i]);
            tab[base + 1 * 8] = make_double2(d[i], g[i]);
            tab[base + 2 * 8] = make_double2(Q[i], R[i]);
            tab[base + 3 * 8] = make_double2(S[i], 0.0);
        }
        double* c = &crec[(size_t)t * 16];
        double w = 1.0;
        for (int k = 1; k <= LB2; k++) {            // A(t) = sum_k wf[k] yend(t-k)
            c[k - 1] = (t - k >= 0) ? w : 0.0;
            if (t - k >= 0) w *= Af[t - k];
        }
        w = 1.0;
        for (int k = 1; k <= LB2; k++) {            // B(t) = sum_k wb[k] z(t+k)
            c[LB2 + k - 1] = (t + k <= T - 1) ? w : 0.0;
            if (t + k <= T - 1) w *= Rb[t + k];
        }
        c[12] = Q0[t]; c[13] = PP[t]; c[14] = isc[t] ? 1.0 : 0.0;
        c[15] = (unscaled && t < T - 1) ? 1.0 / rho[(t + 1) * CHUNK] : 1.0;      // B in the variable w of a constant chunk
    }
    if (getenv("TLAB_DEBUG_PLAN")) {
        for (int i : {0, 1, 2, 16, 40, 80, n / 2, n / 2 + 1, n - 80, n - 40, n - 3, n - 2, n - 1})
            fprintf(stderr, "   i=%d a=%.17g d=%.17g g=%.17g e=%g pp=%g S=%g\n", i, a[i], d[i], g[i], e[i], pp[i], S[i]);
        int nc = 0;
        for (int t = 0; t < T; t++) nc += isc[t];
        fprintf(stderr, "[sys2] n=%d T=%d periodic=%d Wf=%d Wb=%d const_chunks=%d K0=%d K1=%d ca=%g cd=%g cg=%g Af=%g Rb=%g\n", n, T,
                (int)periodic, window(Af, true), window(Rb, false), nc, s2.K0, s2.K1, s2.ca, s2.cd, s2.cg, Af[T / 2], Rb[T / 2]);
    }
    {
        // longest run of constant chunks
        int best = 0, lo = 0;
        for (int t = 0; t < T; t++) {
            if (!isc[t]) { lo = t + 1; continue; }
            if (t - lo + 1 > best) { best = t - lo + 1; s2.c_lo = lo; s2.c_hi = t; }
        }
    }
    if (unscaled) {
        int nc = 0;
        for (int t = 0; t < T; t++) nc += isc[t];
        if (nc > 0) {
            const int Tq = (T + 7) / 8 * 8;
            std::vector<double> rt((size_t)Tq * CHUNK, 1.0);
            for (int t = 0; t < T; t++)
                for (int j = 0; j < CHUNK; j++) rt[((size_t)(t >> 3) * CHUNK + j) * 8 + (t & 7)] = rho[t * CHUNK + j];
            s2.rho = upload(p, rt);
            if (!s2.rho) return;
        } else {
            for (int t = 0; t < T; t++) crec[(size_t)t * 16 + 15] = 1.0;
        }
    }
    s2.march_ok = march_sys_ok(crec, T, s2.K0m, s2.K1m, periodic) ? 1 : 0;
    s2.tab = upload_t(p, tab);
    s2.crec = upload(p, crec);
    s2.ok = (s2.tab && s2.crec) ? 1 : 0;
}

// lu columns c0+1..c0+3 (c0+1..c0+5 periodic); rows nmin..nmax active; scale = diffusivity (1 = none)
void make_solve(DevPlan& p, SolveTab& s, Sys2& s2, const Mat& lu, int c0, int nmin, int nmax, bool periodic, double diff,
                bool scaled, const Mat* lhs = nullptr) {
    const int n = p.n;
    std::vector<double> alpha(n, 0.0), beta(n, 1.0), gamma(n, 0.0), delta(n, 1.0), pd(n, 0.0), pe(n, 0.0);
    if (periodic) {
        for (int r = 1; r <= n; r++) {
            const int i = r - 1;
            double a = lu(r, 1), b = lu(r, 2), c = lu(r, 3), d = lu(r, 4), e = lu(r, 5);
            if (scaled) { b = b * diff; d = d / diff; }
            if (r >= 2 && r <= n - 1) alpha[i] = a;
            if (r <= n - 1) { beta[i] = b; pd[i] = d; pe[i] = e; }
            if (r <= n - 2) gamma[i] = c;
            if (r == n) s.bN = b;
        }
    } else {
        for (int r = 1; r <= n; r++) {
            const int i = r - 1;
            if (r < nmin || r > nmax) continue;
            double a = lu(r, c0 + 1), b = lu(r, c0 + 2), c = lu(r, c0 + 3);
            if (scaled) { b = b * diff; c = c / diff; }
            if (r > nmin) alpha[i] = a;
            if (r < nmax) gamma[i] = c;
            delta[i] = b;
        }
    }
    std::vector<double> sdiag;
    if (lhs != nullptr) {
        sdiag.resize(n);
        for (int r = 1; r <= n; r++) sdiag[r - 1] = (*lhs)(r, 2);      // centre column of the tridiagonal lhs: 1 * s_j
    }
    build_sys2(p, s2, alpha, beta, gamma, delta, pd, pe, s.bN, periodic, sdiag.empty() ? nullptr : &sdiag);
    s2.jscale = scaled ? diff : 1.0;
    finish_solve(p, s, alpha, beta, gamma, delta, pd, pe, periodic);
}

// special rows of the banded right-hand side, densified
void make_rhs(const HostDer& g, bool second, int ibc, RhsTab& R) {
    std::memset(&R, 0, sizeof(R));
    const int n = g.size, ndr = g.ndr, idr = ndr / 2 + 1;
    const int nb = second ? idr - 1 : idr;
    const int ref = nb + 1;
    R.rc = second ? g.rhs(ref, idr) : 0.0;
    R.r2 = (ndr >= 5) ? g.rhs(ref, idr + 2) : 0.0;
    R.r3 = (ndr >= 7) ? g.rhs(ref, idr + 3) : 0.0;
    if (g.periodic) { R.nb = 0; return; }
    R.nb = nb;
    const bool neu_min = !second && (ibc == BCS_ND || ibc == BCS_NN);
    const bool neu_max = !second && (ibc == BCS_DN || ibc == BCS_NN);
    for (int i = 1; i <= nb; i++) {
        if (neu_min) {
            if (i == 1) continue;
            for (int j = 0; j <= ndr; j++) { int col = i + j - idr; if (col >= 2 && col <= BROW_W) R.bot[i - 1][col - 1] += g.rhs_b(i, j); }
        } else {
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; if (col >= 1 && col <= BROW_W) R.bot[i - 1][col - 1] += g.rhs(i, j); }
            if (i == 1) R.bot[0][idr] += g.rhs(1, 1);
        }
    }
    for (int q = 0; q < nb; q++) {
        const int i = n - q;
        if (neu_max) {
            if (q == 0) continue;
            const int tr = idr - q;
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; int k = n - col; if (k >= 1 && k < BROW_W) R.top[q][k] += g.rhs_t(tr, j); }
        } else {
            for (int j = 1; j <= ndr; j++) { int col = i + j - idr; int k = n - col; if (k >= 0 && k < BROW_W) R.top[q][k] += g.rhs(i, j); }
            if (q == 0) R.top[0][idr] += g.rhs(n, ndr);
        }
    }
}

}  // namespace

int devplan_build(DevPlan& p) {
    const HostPlan& h = p.h;
    p.n = h.size;
    p.periodic = h.periodic;
    if (p.n <= 1) { p.T = 1; p.cbase = 1; p.crem = 0; return 0; }
    p.need_1der = h.der2.need_1der;
    p.T = (p.n + CHUNK - 1) / CHUNK;
    p.cbase = p.n / p.T;
    p.crem = p.n % p.T;
    const int n = p.n;
    // first derivative
    if (h.periodic) {
        make_rhs(h.der1, false, BCS_PERIODIC, p.rhs1[0]);
        make_solve(p, p.lu1[0], p.sys1[0], h.der1.lu, 0, 1, n, true, 1.0, false, &h.der1.lhs);
        for (int b = 1; b < 4; b++) { p.rhs1[b] = p.rhs1[0]; p.lu1[b] = p.lu1[0]; p.sys1[b] = p.sys1[0]; }
    } else {
        for (int ibc = 0; ibc < 4; ibc++) {
            make_rhs(h.der1, false, ibc, p.rhs1[ibc]);
            int nmin = 1, nmax = n;
            if (ibc == BCS_ND || ibc == BCS_NN) nmin++;
            if (ibc == BCS_DN || ibc == BCS_NN) nmax--;
            make_solve(p, p.lu1[ibc], p.sys1[ibc], h.der1.lu, ibc * 5, nmin, nmax, false, 1.0, false, &h.der1.lhs);
        }
    }
    // second derivative
    make_rhs(h.der2, true, BCS_DD, p.rhs2);
    if (h.der2.mode_fdm == FDM_COM6_DIRECT) {
        // per-row coefficients: no constant interior stencil, no dense special rows (the kernels branch on rhs2_rows)
        std::memset(&p.rhs2, 0, sizeof(p.rhs2));
        std::vector<double> rows(5 * (size_t)n);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= 5; j++) rows[5 * (size_t)(i - 1) + (j - 1)] = h.der2.rhs(i, j);
        p.rhs2_rows = upload(p, rows);
    }
    p.lu2.clear();
    p.lu2.emplace_back();
    p.sys2.clear();
    p.sys2.emplace_back();
    make_solve(p, p.lu2[0], p.sys2[0], h.der2.lu, 0, 1, n, h.periodic, 1.0, false, &h.der2.lhs);
    if (p.need_1der) {
        std::vector<double> r(3 * (size_t)n);
        for (int i = 1; i <= n; i++) for (int j = 1; j <= 3; j++) r[3 * (size_t)(i - 1) + (j - 1)] = h.der2.rhs(i, h.der2.ndr + j);
        p.rhs_d1 = upload(p, r);
        // the fast kernels need the tridiagonal lhs without an extended-stencil entry (rhs_d1(1,1) = rhs_d1(n,3) = 0)
        if (p.crem == 0 && p.cbase == CHUNK && r[0] == 0.0 && r[3 * (size_t)(n - 1) + 2] == 0.0) {
            const int Tp = (p.T + 7) / 8 * 8;
            std::vector<double> cj((size_t)Tp * CHUNK, 0.0);
            for (int t = 0; t < p.T; t++)
                for (int j = 0; j < CHUNK; j++) {
                    const int i = t * CHUNK + j + 1;
                    cj[((size_t)(t >> 3) * CHUNK + j) * 8 + (t & 7)] = h.jac(i, 3) / (h.jac(i, 2) * h.jac(i, 2));
                }
            p.cjac2 = upload(p, cj);
        }
    }
    {
        std::vector<double> j1(n);
        for (int i = 1; i <= n; i++) j1[i - 1] = h.jac(i, 1);
        p.d_jac = upload(p, j1);
        p.d_mwn1 = upload(p, h.der1.mwn);
    }
    // Neumann boundary-value closures
    std::memset(p.neu_bot, 0, sizeof(p.neu_bot));
    std::memset(p.neu_top, 0, sizeof(p.neu_top));
    std::memset(p.neu_lu_bot, 0, sizeof(p.neu_lu_bot));
    std::memset(p.neu_lu_top, 0, sizeof(p.neu_lu_top));
    if (!h.periodic) {
        const HostDer& g = h.der1;
        const int ndr = g.ndr, idr = ndr / 2 + 1;
        for (int ibc = 1; ibc < 4; ibc++) {
            const int ip = ibc * 5;
            if (ibc == BCS_ND || ibc == BCS_NN) {
                for (int j = idr + 1; j <= ndr; j++) p.neu_bot[ibc][j - idr] += g.rhs_b(1, j);   // col 1+j-idr, 0-based k = j-idr
                p.neu_bot[ibc][idr] += g.rhs_b(1, 1);                                            // extended stencil, col idr+1
                p.neu_lu_bot[ibc] = g.lu(1, ip + 3);
            }
            if (ibc == BCS_DN || ibc == BCS_NN) {
                for (int j = 1; j <= idr - 1; j++) p.neu_top[ibc][idr - j] += g.rhs_t(idr, j);  // col n+j-idr, k = idr-j
                p.neu_top[ibc][idr] += g.rhs_t(idr, ndr);                                        // extended stencil, col n-idr
                p.neu_lu_top[ibc] = g.lu(n, ip + 1);
            }
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 80;
}

int devplan_add_diffusion(DevPlan& p, double diff) {
    if (p.n <= 1) return 0;
    p.lu2.emplace_back();
    p.sys2.emplace_back();
    make_solve(p, p.lu2.back(), p.sys2.back(), p.h.der2.lu, 0, 1, p.n, p.h.periodic, diff, true, &p.h.der2.lhs);
    return (int)p.lu2.size() - 1;
}

void devplan_free(DevPlan& p) {
    for (void* a : p.allocs) cudaFree(a);
    p.allocs.clear();
}

}  // namespace tlab
