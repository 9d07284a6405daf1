// ORACLE (test infrastructure, NOT product code) -- C++ port of the reference's *literal* CPU
// algorithm, used as the timed CPU baseline (bench.py cpu_baseline / --impl reference) and as a
// second checker.  PARITY UNPINNED: the reference is Julia and cannot run in this image.
//
// Follows, statement by statement:
//   calculate_unitary_and_derivatives      reference src/UnitaryCalculations.jl:20-155
//   calculate_fidelity_and_derivatives     reference src/FidelityCalculations.jl:19-119
//   calculate_common! (cost/grad assembly) reference src/FidelityCalculations.jl:174-184
// with concrete complex<double> storage (the reference allocates abstract `Complex`/`Real`
// arrays, src/UnitaryCalculations.jl:34-42, so this port is *faster* than the Julia original).
// The arithmetic the reference delegates to Julia's stdlib LinearAlgebra (unpinned version, no
// Manifest) is restated from the published algorithms: `exp` = Higham 2005 Pade scaling-and-squaring
// with the degree-3/5/7/9/13 thresholds 0.015/0.25/0.95/2.1/5.4 (without the gebal balancing step,
// which only permutes for these Hamiltonians), `inv` = LU with partial pivoting.
//
// Hamiltonians come in through the same term-list descriptor as the CUDA library
// (include/robustgrape_b200.h), evaluated *by value* at the perturbed inputs exactly as the
// reference's closures would be -- none of the CUDA path's difference algebra is used here.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>
#include "../include/robustgrape_b200.h"

typedef std::complex<double> cd;

#define ORACLE_MAXD 10
struct Mat {
    int n;
    cd a[ORACLE_MAXD * ORACLE_MAXD];   // column-major, fixed capacity: no heap traffic in the matrix operators
    Mat() : n(0) {}
    explicit Mat(int n_) : n(n_) { for (int i = 0; i < n_ * n_; ++i) a[i] = cd(0, 0); }
    cd& operator()(int i, int j) { return a[i + (size_t)n * j]; }
    const cd& operator()(int i, int j) const { return a[i + (size_t)n * j]; }
    static Mat eye(int n) { Mat m(n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
};
static Mat operator*(const Mat& A, const Mat& B) {
    const int n = A.n; Mat C(n);
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k) { const cd b = B(k, j); if (b == cd(0, 0)) continue; for (int i = 0; i < n; ++i) C(i, j) += A(i, k) * b; }
    return C;
}
static Mat operator+(const Mat& A, const Mat& B) { Mat C(A.n); for (int i = 0; i < C.n * C.n; ++i) C.a[i] = A.a[i] + B.a[i]; return C; }
static Mat operator-(const Mat& A, const Mat& B) { Mat C(A.n); for (int i = 0; i < C.n * C.n; ++i) C.a[i] = A.a[i] - B.a[i]; return C; }
static Mat operator*(cd s, const Mat& A) { Mat C(A.n); for (int i = 0; i < C.n * C.n; ++i) C.a[i] = s * A.a[i]; return C; }
static Mat adj(const Mat& A) { Mat C(A.n); for (int i = 0; i < A.n; ++i) for (int j = 0; j < A.n; ++j) C(j, i) = std::conj(A(i, j)); return C; }
static cd trace(const Mat& A) { cd s = 0; for (int i = 0; i < A.n; ++i) s += A(i, i); return s; }
static double norm1(const Mat& A) { double m = 0; for (int j = 0; j < A.n; ++j) { double s = 0; for (int i = 0; i < A.n; ++i) s += std::abs(A(i, j)); m = std::max(m, s); } return m; }

// Solve A X = B in place (LU, partial pivoting); B is overwritten with X.
static void gesv(Mat A, Mat& B) {
    const int n = A.n;
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::abs(A(k, k));
        for (int i = k + 1; i < n; ++i) if (std::abs(A(i, k)) > best) { best = std::abs(A(i, k)); p = i; }
        if (p != k) for (int j = 0; j < n; ++j) { std::swap(A(k, j), A(p, j)); std::swap(B(k, j), B(p, j)); }
        const cd piv = A(k, k);
        for (int i = k + 1; i < n; ++i) {
            const cd f = A(i, k) / piv;
            if (f == cd(0, 0)) continue;
            for (int j = k + 1; j < n; ++j) A(i, j) -= f * A(k, j);
            for (int j = 0; j < n; ++j) B(i, j) -= f * B(k, j);
        }
    }
    for (int j = 0; j < n; ++j)
        for (int i = n - 1; i >= 0; --i) {
            cd s = B(i, j);
            for (int k = i + 1; k < n; ++k) s -= A(i, k) * B(k, j);
            B(i, j) = s / A(i, i);
        }
}
static Mat inv(const Mat& A) { Mat X = Mat::eye(A.n); gesv(A, X); return X; }

// exp(A): Higham (2005) Pade approximants with scaling and squaring.
static Mat expm(Mat A) {
    const int n = A.n;
    const double nA = norm1(A);
    const Mat I = Mat::eye(n);
    if (nA <= 2.1) {
        static const double C9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
        static const double C7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
        static const double C5[] = {30240., 15120., 3360., 420., 30., 1.};
        static const double C3[] = {120., 60., 12., 1.};
        const double* C; int nc;
        if (nA > 0.95) { C = C9; nc = 10; } else if (nA > 0.25) { C = C7; nc = 8; } else if (nA > 0.015) { C = C5; nc = 6; } else { C = C3; nc = 4; }
        const Mat A2 = A * A;
        Mat P = I, U = cd(C[1]) * I, V = cd(C[0]) * I;
        for (int k = 1; k <= nc / 2 - 1; ++k) { P = P * A2; U = U + cd(C[2 * k + 1]) * P; V = V + cd(C[2 * k]) * P; }
        U = A * U;
        Mat X = V + U;
        gesv(V - U, X);
        return X;
    }
    const double s = std::log2(nA / 5.4);
    int si = 0;
    if (s > 0) { si = (int)std::ceil(s); A = cd(std::ldexp(1.0, -si)) * A; }
    static const double CC[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                                10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
    const Mat A2 = A * A, A4 = A2 * A2, A6 = A2 * A4;
    Mat U = A * (A6 * (cd(CC[13]) * A6 + cd(CC[11]) * A4 + cd(CC[9]) * A2) + cd(CC[7]) * A6 + cd(CC[5]) * A4 + cd(CC[3]) * A2 + cd(CC[1]) * I);
    Mat V = A6 * (cd(CC[12]) * A6 + cd(CC[10]) * A4 + cd(CC[8]) * A2) + cd(CC[6]) * A6 + cd(CC[4]) * A4 + cd(CC[2]) * A2 + cd(CC[0]) * I;
    Mat X = V + U;
    gesv(V - U, X);
    for (int t = 0; t < si; ++t) X = X * X;
    return X;
}

// ---- descriptor evaluation by value ---------------------------------------------------------
static cd factor_value(const rg_factor& f, int k, const double* xk, const double* xadd, double err, const rg_problem_desc* d) {
    switch (f.kind) {
    case RG_F_ERR: return err;
    case RG_F_ERR1P_M1: { volatile double t = 1.0 + err; return t - 1.0; }
    case RG_F_TABLE: return d->table[(size_t)f.index * d->ntimes + k];
    default: break;
    }
    const double v = (f.space == RG_S_MAIN) ? xk[f.index] : xadd[f.index];
    const double u = (f.scale == 1.0 && f.offset == 0.0) ? v : f.scale * v + f.offset;
    if (f.kind == RG_F_VAR) return u;
    if (f.kind == RG_F_COS) return std::cos(u);
    if (f.kind == RG_F_SIN) return std::sin(u);
    return cd(std::cos(u), std::sin(u));
}
// owner: RG_OWNER_H0, error index, or RG_OWNER_TARGET
static Mat op_eval(const rg_problem_desc* d, const rg_term* terms, int nterms, int owner, int k, const double* xk,
                   const double* xadd, double err) {
    Mat M(d->ndim);
    for (int t = 0; t < nterms; ++t) {
        const rg_term& tm = terms[t];
        if (tm.owner != owner) continue;
        cd c(tm.coef_re, tm.coef_im);
        for (int f = 0; f < tm.nfactors; ++f) c *= factor_value(tm.factors[f], k, xk, xadd, err, d);
        for (int z = 0; z < tm.nnz; ++z) M(tm.rows[z], tm.cols[z]) += c * cd(tm.vals[2 * z], tm.vals[2 * z + 1]);
    }
    return M;
}

// ---- the literal algorithm --------------------------------------------------------------------
static void one_pulse(const rg_problem_desc* d, const double* x, double* Fo, double* Fdx_tot, double* F2o, double* F2dx_tot) {
    const int n = d->ndim, N = d->ntimes, p = d->nparam, na = d->nb_additional_param, ne = d->nerr;
    const int nx = p * N + na;
    const double eps = d->eps, eps2 = d->eps2, dt = d->t0 / d->ntimes;
    const cd mI(0, -1);
    const double* x_add = x + (size_t)p * N;
    std::vector<double> xac(x_add, x_add + na), xmc(p);
    auto H0 = [&](int k, const double* xk, const double* xa) { return op_eval(d, d->terms, d->nterms, RG_OWNER_H0, k, xk, xa, 0.0); };
    auto He = [&](int e, int k, const double* xk, const double* xa, double err) { return op_eval(d, d->terms, d->nterms, e, k, xk, xa, err); };
    auto prop = [&](const Mat& H) { return expm((mI * dt) * H); };

    Mat cum = Mat::eye(n), old = cum;
    std::vector<Mat> A_dx((size_t)p * N), A_dxa((size_t)na * N), A_derr((size_t)ne * N), A_derr_dx((size_t)p * ne * N), A_derr_dxa((size_t)na * ne * N);
    std::vector<Mat> arr_err(ne), arr_dx(p), arr_dxa(na);
    for (int nt = 0; nt < N; ++nt) {                                            // :44
        const double* xk = x + (size_t)nt * p;
        const Mat U = prop(H0(nt, xk, x_add));                                  // :45
        cum = U * cum;                                                          // :46
        const Mat cinv = inv(cum);                                              // :47
        std::copy(xk, xk + p, xmc.begin());
        for (int i = 0; i < p; ++i) {                                           // :49-56
            xmc[i] += eps;
            const Mat Ud = prop(H0(nt, xmc.data(), x_add));
            A_dx[i + (size_t)p * nt] = cinv * (cd(1 / eps) * (Ud - U)) * old;
            xmc[i] = xk[i] + eps2;
            arr_dx[i] = prop(H0(nt, xmc.data(), x_add));
            xmc[i] = xk[i];
        }
        for (int j = 0; j < na; ++j) {                                          // :57-64
            xac[j] += eps;
            const Mat Ud = prop(H0(nt, xk, xac.data()));
            A_dxa[j + (size_t)na * nt] = cinv * (cd(1 / eps) * (Ud - U)) * old;
            xac[j] = x_add[j] + eps2;
            arr_dxa[j] = prop(H0(nt, xk, xac.data()));
            xac[j] = x_add[j];
        }
        for (int e = 0; e < ne; ++e) {                                          // :66-98
            const Mat Ue = prop(He(e, nt, xk, x_add, eps) + H0(nt, xk, x_add));
            A_derr[e + (size_t)ne * nt] = cinv * (cd(1 / eps) * (Ue - U)) * old;
            arr_err[e] = prop(He(e, nt, xk, x_add, eps2) + H0(nt, xk, x_add));
            for (int i = 0; i < p; ++i) {
                xmc[i] += eps2;
                const Mat Um = prop(He(e, nt, xmc.data(), x_add, eps2) + H0(nt, xmc.data(), x_add));
                A_derr_dx[i + (size_t)p * (e + (size_t)ne * nt)] = cinv * (cd(1 / (eps2 * eps2)) * (Um + U - arr_err[e] - arr_dx[i])) * old;
                xmc[i] = xk[i];
            }
            for (int j = 0; j < na; ++j) {
                xac[j] += eps2;
                const Mat Um = prop(He(e, nt, xk, xac.data(), eps2) + H0(nt, xk, xac.data()));
                A_derr_dxa[j + (size_t)na * (e + (size_t)ne * nt)] = cinv * (cd(1 / (eps2 * eps2)) * (Um + U - arr_err[e] - arr_dxa[j])) * old;
                xac[j] = x_add[j];
            }
        }
        old = cum;                                                              // :99
    }
    // contractions :106-152
    std::vector<Mat> U_dx((size_t)p * N), U_dxa(na), U_derr(ne), U_derr_dx((size_t)p * N * ne), U_derr_dxa((size_t)na * ne);
    for (int nt = 0; nt < N; ++nt) for (int i = 0; i < p; ++i) U_dx[i + (size_t)p * nt] = cum * A_dx[i + (size_t)p * nt];
    for (int j = 0; j < na; ++j) { Mat S(n); for (int nt = 0; nt < N; ++nt) S = S + A_dxa[j + (size_t)na * nt]; U_dxa[j] = cum * S; }
    for (int e = 0; e < ne; ++e) {
        std::vector<Mat> cs(N), rcs(N);
        Mat acc(n);
        for (int nt = 0; nt < N; ++nt) { acc = acc + A_derr[e + (size_t)ne * nt]; cs[nt] = acc; }
        acc = Mat(n);
        for (int nt = N - 1; nt >= 0; --nt) { acc = acc + A_derr[e + (size_t)ne * nt]; rcs[nt] = acc; }
        U_derr[e] = cum * cs[N - 1];
        for (int nt = 0; nt < N; ++nt)
            for (int i = 0; i < p; ++i) {
                Mat T(n);
                if (nt >= 1) T = T + A_dx[i + (size_t)p * nt] * cs[nt - 1];
                if (nt < N - 1) T = T + rcs[nt + 1] * A_dx[i + (size_t)p * nt];
                T = T + A_derr_dx[i + (size_t)p * (e + (size_t)ne * nt)];
                U_derr_dx[i + (size_t)p * (nt + (size_t)N * e)] = cum * T;
            }
        for (int j = 0; j < na; ++j) {
            Mat T(n);
            for (int nt = 1; nt < N; ++nt) T = T + A_dxa[j + (size_t)na * nt] * cs[nt - 1];
            for (int nt = 0; nt < N - 1; ++nt) T = T + rcs[nt + 1] * A_dxa[j + (size_t)na * nt];
            for (int nt = 0; nt < N; ++nt) T = T + A_derr_dxa[j + (size_t)na * (e + (size_t)ne * nt)];
            U_derr_dxa[j + (size_t)na * e] = cum * T;
        }
    }
    // fidelity part, src/FidelityCalculations.jl:32-118
    const Mat& U = cum;
    const Mat U0 = op_eval(d, d->target_terms, d->ntarget_terms, RG_OWNER_TARGET, 0, nullptr, x_add, 0.0);
    std::vector<Mat> V(na);
    for (int j = 0; j < na; ++j) {
        xac[j] += eps;
        V[j] = cd(1 / eps) * (op_eval(d, d->target_terms, d->ntarget_terms, RG_OWNER_TARGET, 0, nullptr, xac.data(), 0.0) - U0);
        xac[j] = x_add[j];
    }
    Mat P0(n), P(n);
    for (int i = 0; i < n * n; ++i) { P0.a[i] = d->projector[i]; P.a[i] = (d->projector[i] != 0.0) ? 1.0 : 0.0; }
    const double D = std::real(trace(P0)), DD = D * (D + 1);
    auto trm = [&](const Mat& A) { return trace(P0 * A); };
    const Mat U0h = adj(U0), Uh = adj(U);
    const cd tau = trm(P * U0h * U);
    *Fo = (std::real(trm(P * U0h * U * P * Uh * U0)) + std::norm(tau)) / DD;                                    // :54
    for (int nt = 0; nt < N; ++nt)
        for (int i = 0; i < p; ++i) {                                                                             // :56-65
            const Mat& X = U_dx[i + (size_t)p * nt];
            Fdx_tot[i + (size_t)p * nt] = (std::real(trm(P * U0h * X * P * Uh * U0 + P * U0h * U * P * adj(X) * U0))
                                           + 2 * std::real(std::conj(trm(P * U0h * U)) * trm(P * U0h * X))) / DD;
        }
    for (int j = 0; j < na; ++j) {                                                                                // :67-76
        const Mat& X = U_dxa[j];
        Fdx_tot[(size_t)p * N + j] = (std::real(trm(P * U0h * X * P * Uh * U0 + P * U0h * U * P * adj(X) * U0 + P * adj(V[j]) * U * P * Uh * U0 + P * U0h * U * P * Uh * V[j]))
                                      + 2 * std::real(std::conj(trm(P * U0h * U)) * trm(P * U0h * X + P * adj(V[j]) * U))) / DD;
    }
    for (int e = 0; e < ne; ++e) {                                                                                // :78-114
        const Mat& E = U_derr[e];
        const Mat Eh = adj(E);
        F2o[e] = 2 * (std::real(trm(P * U0h * E * P * Eh * U0 - P * Eh * E)) + std::norm(trm(P * U0h * E)) - D * std::real(trm(P * Eh * E))) / DD;
        for (int nt = 0; nt < N; ++nt)
            for (int i = 0; i < p; ++i) {
                const Mat& Z = U_derr_dx[i + (size_t)p * (nt + (size_t)N * e)];
                const Mat Zh = adj(Z);
                F2dx_tot[(size_t)e * nx + i + (size_t)p * nt] = 2 * (
                    std::real(trm(P * U0h * Z * P * Eh * U0 + P * U0h * E * P * Zh * U0 - P * Zh * E - P * Eh * Z))
                    + 2 * std::real(std::conj(trm(P * U0h * E)) * trm(P * U0h * Z))
                    - D * std::real(trm(P * Zh * E + P * Eh * Z))) / DD;
            }
        for (int j = 0; j < na; ++j) {
            const Mat& Z = U_derr_dxa[j + (size_t)na * e];
            const Mat Zh = adj(Z);
            F2dx_tot[(size_t)e * nx + (size_t)p * N + j] = 2 * (
                std::real(trm(P * adj(V[j]) * E * P * Eh * U0 + P * U0h * Z * P * Eh * U0 + P * U0h * E * P * Zh * U0
                              + P * U0h * E * P * Eh * V[j] - P * Zh * E - P * Eh * Z))
                + 2 * std::real(std::conj(trm(P * U0h * E)) * trm(P * adj(V[j]) * E + P * U0h * Z))
                - D * std::real(trm(P * Zh * E + P * Eh * Z))) / DD;
        }
    }
}

// Dynamic scheduling of pulses over host threads (std::thread: no OpenMP runtime dependency).
template <class Fn>
static void parallel_for(int B, int nthreads, Fn fn) {
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, B));
    if (nt == 1) { for (int b = 0; b < B; ++b) fn(b); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() { for (int b = next++; b < B; b = next++) fn(b); });
    for (auto& t : th) t.join();
}

extern "C" {
// Batched calculate_fidelity_and_derivatives; layouts as in rg_fidelity_and_derivatives_batch.
int oracle_fidelity_and_derivatives_batch(const rg_problem_desc* d, int B, const double* X, double* F, double* Fdx,
                                          double* F2, double* F2dx, int nthreads) {
    if (d->ndim > ORACLE_MAXD) return -1;
    const int nx = d->nparam * d->ntimes + d->nb_additional_param, ne = d->nerr;
    parallel_for(B, nthreads, [&](int b) {
        std::vector<double> f2(std::max(1, ne)), f2dx((size_t)std::max(1, ne) * nx), fdx(nx);
        double f;
        one_pulse(d, X + (size_t)b * nx, &f, fdx.data(), f2.data(), f2dx.data());
        if (F) F[b] = f;
        if (Fdx) std::copy(fdx.begin(), fdx.end(), Fdx + (size_t)b * nx);
        if (F2) for (int e = 0; e < ne; ++e) F2[(size_t)b * ne + e] = f2[e];
        if (F2dx) for (size_t i = 0; i < (size_t)ne * nx; ++i) F2dx[(size_t)b * ne * nx + i] = f2dx[i];
    });
    return 0;
}
// calculate_common! without regularisation, src/FidelityCalculations.jl:177-184
int oracle_cost_and_grad_batch(const rg_problem_desc* d, int B, const double* X, const double* coeff, double* cost,
                               double* grad, int nthreads) {
    if (d->ndim > ORACLE_MAXD) return -1;
    const int nx = d->nparam * d->ntimes + d->nb_additional_param, ne = d->nerr;
    parallel_for(B, nthreads, [&](int b) {
        std::vector<double> f2(std::max(1, ne)), f2dx((size_t)std::max(1, ne) * nx), fdx(nx);
        double f;
        one_pulse(d, X + (size_t)b * nx, &f, fdx.data(), f2.data(), f2dx.data());
        double c = 1 - f;
        for (int i = 0; i < nx; ++i) grad[(size_t)b * nx + i] = -fdx[i];
        for (int e = 0; e < ne; ++e) {
            c += coeff[e] * f2[e] * f2[e];
            for (int i = 0; i < nx; ++i) grad[(size_t)b * nx + i] += 2 * coeff[e] * f2[e] * f2dx[(size_t)e * nx + i];
        }
        cost[b] = c;
    });
    return 0;
}
int oracle_max_threads(void) {
    return std::max(1u, std::thread::hardware_concurrency());
}
}
