// rg_block2.cuh -- workspace-free fused path for Hamiltonians whose coupling graph decomposes into blocks of at most
// two levels (the 5-level symmetric-blockaded and 7-level full-blockaded Rydberg models, src/RydbergTools.jl:31-39,71-81:
// 1 + 2 + 2 (+ 2)).
//
// Nothing of size N is stored: the step propagators U_k = exp(-i dt H_k) and their finite differences are *recomputed*
// inside the two sweeps (chunk aggregate, backward gradient) from x_k.  That is affordable because a 2 x 2 skew-Hermitian
// block has a closed-form exponential,
//      A = i mu I + A0,  A0 = [[i nu, w], [-conj(w), -i nu]],  A0^2 = -z I,  z = nu^2 + |w|^2
//      exp(A) = e^{i mu} (C(z) I + S(z) A0),   C(z) = cos sqrt z = sum (-z)^k/(2k)!,  S(z) = sinc sqrt z = sum (-z)^k/(2k+1)!
// and the reference's finite differences (src/UnitaryCalculations.jl:50-52,66-70,76-83) are evaluated *exactly*, without
// subtracting rounded exponentials, by running the same formula in "difference arithmetic": every scalar is carried as a
// jet (value, difference in a, difference in b, mixed second difference) and products use the exact rule
//      d(pq) = dp (q + dq) + p dq ,   d_ab(pq) = p_ab (q+q_a+q_b+q_ab) + p_a (q_b+q_ab) + p_b (q_a+q_ab) + p q_ab .
// This is the scalar counterpart of the differenced Horner recurrences of rg_smalld.cuh, so the semantics (perturbed input
// fl(x+eps), nominal eps in the quotient, O(eps) truncation bias of the reference) are unchanged.
// Range: z <= 4.5 per block (||dt H||_1 up to ~2.1); beyond that the kernels raise the status flag and the host falls back
// to the general scaling-and-squaring kernels.
#pragma once
#include "rg_steps_t.cuh"

// (-1)^k/(2k)!  and  (-1)^k/(2k+1)!
__constant__ double c_cosq[14] = {1.0, -1.0 / 2, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600, -1.0 / 87178291200.0,
                                  1.0 / 20922789888000.0, -1.0 / 6402373705728000.0, 1.0 / 2432902008176640000.0,
                                  -1.0 / 1124000727777607680000.0, 1.0 / 620448401733239439360000.0, -1.0 / 403291461126605635584000000.0};
__constant__ double c_sincq[14] = {1.0, -1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800.0,
                                   -1.0 / 1307674368000.0, 1.0 / 355687428096000.0, -1.0 / 121645100408832000.0,
                                   1.0 / 51090942171709440000.0, -1.0 / 25852016738884976640000.0, 1.0 / 15511210043330985984000000.0,
                                   -1.0 / 10888869450418352160768000000.0};

// series degree K with remainder (K+1) z^{K+1}/(2K+2)! below 1e-17 (differences converge like the derivative series)
__device__ __forceinline__ int b2_degree(double z) {
    // branch-free: the number of thresholds below z
    const int k = 2 + (z > 1.3e-5) + (z > 5.6e-4) + (z > 5.9e-3) + (z > 3.0e-2) + (z > 0.10) + (z > 0.26) + (z > 0.57) + (z > 1.05) +
                  (z > 1.85) + (z > 2.95);
    return (z > 4.5) ? 99 : k;
}

// ---- difference arithmetic ------------------------------------------------------------------------------------------
__device__ __forceinline__ double jm(double a, double b) { return a * b; }
__device__ __forceinline__ cplx jm(double a, cplx b) { return cscale(b, a); }
__device__ __forceinline__ cplx jm(cplx a, double b) { return cscale(a, b); }
__device__ __forceinline__ cplx jm(cplx a, cplx b) { return cmul(a, b); }
__device__ __forceinline__ double ja(double a, double b) { return a + b; }
__device__ __forceinline__ cplx ja(cplx a, cplx b) { return cadd(a, b); }

// r = p * q for jets of order O (1 << O slots: value, d_a, d_b, d_ab); r must not alias p or q
template <int O, class A, class B, class R>
__device__ __forceinline__ void jprod(const A* p, const B* q, R* r) {
    r[0] = jm(p[0], q[0]);
    if constexpr (O >= 1) r[1] = ja(jm(p[1], ja(q[0], q[1])), jm(p[0], q[1]));
    if constexpr (O >= 2) {
        r[2] = ja(jm(p[2], ja(q[0], q[2])), jm(p[0], q[2]));
        r[3] = ja(ja(jm(p[3], ja(ja(q[0], q[1]), ja(q[2], q[3]))), jm(p[1], ja(q[2], q[3]))),
                  ja(jm(p[2], ja(q[1], q[3])), jm(p[0], q[3])));
    }
}
// e^{ih} - 1 = -2 sin^2(h/2) + i sin h, accurate for tiny h
static __device__ __noinline__ void expm1i_large(double h, cplx* out) {      // |h| >= 1e-3: never taken with the reference's eps = 1e-8, eps2 = 1e-4
    const double s2 = sin(0.5 * h);
    *out = cmk(-2.0 * s2 * s2, sin(h));
}
__device__ __forceinline__ cplx expm1i(double h) {
    if (fabs(h) < 1e-3) {
        const double h2 = h * h;
        return cmk(-0.5 * h2 * (1.0 - h2 * (1.0 / 12.0) * (1.0 - h2 * (1.0 / 30.0))),
                   h * (1.0 - h2 * (1.0 / 6.0) * (1.0 - h2 * (1.0 / 20.0))));
    }
    cplx r;
    expm1i_large(h, &r);
    return r;
}
// E = e^{i mu} as a jet
template <int O>
__device__ __forceinline__ void jexpi(const double* mu, cplx* E) {
    double s, c;
    rg_sincos(mu[0], s, c);
    E[0] = cmk(c, s);
    if constexpr (O >= 1) {
        const cplx ea = expm1i(mu[1]);
        E[1] = cmul(E[0], ea);
        if constexpr (O >= 2) {
            const cplx eb = expm1i(mu[2]), eab = expm1i(mu[3]);
            E[2] = cmul(E[0], eb);
            // (1+ea)(1+eb)(1+eab) - (1+ea) - (1+eb) + 1 = ea eb + eab (1+ea)(1+eb)
            const cplx pa = cmk(1.0 + ea.x, ea.y), pb = cmk(1.0 + eb.x, eb.y);
            E[3] = cmul(E[0], cadd(cmul(ea, eb), cmul(eab, cmul(pa, pb))));
        }
    }
}

// One 2 x 2 skew-Hermitian block [[i a1, w], [-conj(w), i a2]] given as jets -> jets of the four entries of its exponential.
// DIAG = false: a1 = a2 = 0 at compile time (pure coupling, e.g. the resonant Rydberg drive).  Returns the series degree used
// (99 = out of range).
template <int O, bool DIAG>
__device__ __forceinline__ int block2_exp(const double* a1, const double* a2, const cplx* w, cplx* u11, cplx* u12, cplx* u21, cplx* u22) {
    constexpr int n = 1 << O;
    double nu[n], mu[n], z[n], t1[n], t2[n], wr[n], wi[n];
#pragma unroll
    for (int s = 0; s < n; ++s) { wr[s] = w[s].x; wi[s] = w[s].y; }
    jprod<O>(wr, wr, t1); jprod<O>(wi, wi, t2);
#pragma unroll
    for (int s = 0; s < n; ++s) z[s] = t1[s] + t2[s];
    if constexpr (DIAG) {
#pragma unroll
        for (int s = 0; s < n; ++s) { mu[s] = 0.5 * (a1[s] + a2[s]); nu[s] = 0.5 * (a1[s] - a2[s]); }
        jprod<O>(nu, nu, t1);
#pragma unroll
        for (int s = 0; s < n; ++s) z[s] += t1[s];
    }
    double zb = fabs(z[0]);
#pragma unroll
    for (int s = 1; s < n; ++s) zb += fabs(z[s]);
    int K = b2_degree(zb * 1.0001);
    const int Kret = K;
    if (K == 99) K = 12;
    double C[n], S[n];
    C[0] = c_cosq[K]; S[0] = c_sincq[K];
#pragma unroll
    for (int s = 1; s < n; ++s) { C[s] = 0.0; S[s] = 0.0; }
    // Horner in difference arithmetic, v <- c_j + z v
#ifndef RG_B2_SERIES_SWITCH
#define RG_B2_SERIES_SWITCH 1      // 1: fall-through switch instead of a loop (measured on B200, C4: 0.647 vs 0.683 ms)
#endif
#define RG_B2_SERIES_STEP(j)                                                              \
    {                                                                                     \
        jprod<O>(z, C, t1); jprod<O>(z, S, t2);                                           \
        _Pragma("unroll") for (int s = 0; s < n; ++s) { C[s] = t1[s]; S[s] = t2[s]; }     \
        C[0] += c_cosq[j]; S[0] += c_sincq[j];                                            \
    }
#if RG_B2_SERIES_SWITCH
    switch (K) {
    case 12: RG_B2_SERIES_STEP(11)
    case 11: RG_B2_SERIES_STEP(10)
    case 10: RG_B2_SERIES_STEP(9)
    case 9: RG_B2_SERIES_STEP(8)
    case 8: RG_B2_SERIES_STEP(7)
    case 7: RG_B2_SERIES_STEP(6)
    case 6: RG_B2_SERIES_STEP(5)
    case 5: RG_B2_SERIES_STEP(4)
    case 4: RG_B2_SERIES_STEP(3)
    case 3: RG_B2_SERIES_STEP(2)
    default: RG_B2_SERIES_STEP(1) RG_B2_SERIES_STEP(0)
    }
#else
    for (int j = K - 1; j >= 0; --j) RG_B2_SERIES_STEP(j)
#endif
#undef RG_B2_SERIES_STEP
    cplx o12[n];
    jprod<O>(S, w, o12);
    if constexpr (DIAG) {
        cplx g1[n], g2[n], o21[n], E[n];
        jprod<O>(nu, S, t1);
#pragma unroll
        for (int s = 0; s < n; ++s) { g1[s] = cmk(C[s], t1[s]); g2[s] = cmk(C[s], -t1[s]); o21[s] = cmk(-o12[s].x, o12[s].y); }
        jexpi<O>(mu, E);
        jprod<O>(E, g1, u11); jprod<O>(E, g2, u22); jprod<O>(E, o12, u12); jprod<O>(E, o21, u21);
    } else {
#pragma unroll
        for (int s = 0; s < n; ++s) { u11[s] = cmk(C[s], 0.0); u22[s] = cmk(C[s], 0.0); u12[s] = o12[s]; u21[s] = cmk(-o12[s].x, o12[s].y); }
    }
    return Kret;
}

// ---- block structure of a pattern (compile time) ----------------------------------------------------------------------
// partner of level l inside the closure of UMASK: the other member of its two-level block, -1 for a one-level block,
// -2 if the block has more than two members (pattern not eligible)
__host__ __device__ constexpr int b2_partner(int d, unsigned tri, int l) {
    const u64 cm = closure_from_tri(d, tri);
    int p = -1;
    for (int j = 0; j < d; ++j)
        if (j != l && ((cm >> (l + d * j)) & 1ull)) { if (p >= 0) return -2; p = j; }
    return p;
}
// The same table with its entries forced into constant expressions: inside an unrolled device loop a plain call of the
// constexpr function above is *not* a constant-expression context, and for d = 7 the optimiser no longer folds the closure
// loops -- it emitted them as run-time code (40 copies of a 343-iteration loop per thread; measured 4.7 ms instead of
// 0.4 ms for the 7-level model at 8192 x 1000).
template <int D, unsigned UMASK> struct B2P {
    static constexpr int p0 = b2_partner(D, UMASK, 0), p1 = b2_partner(D, UMASK, 1), p2 = b2_partner(D, UMASK, 2), p3 = b2_partner(D, UMASK, 3),
                         p4 = b2_partner(D, UMASK, 4), p5 = b2_partner(D, UMASK, 5), p6 = b2_partner(D, UMASK, 6), p7 = b2_partner(D, UMASK, 7);
    __host__ __device__ static constexpr int partner(int l) {
        return l == 0 ? p0 : l == 1 ? p1 : l == 2 ? p2 : l == 3 ? p3 : l == 4 ? p4 : l == 5 ? p5 : l == 6 ? p6 : p7;
    }
};
__host__ __device__ constexpr bool b2_eligible(int d, unsigned tri) {
    for (int l = 0; l < d; ++l)
        if (b2_partner(d, tri, l) == -2) return false;
    return true;
}

enum { B2_VALUE = 0, B2_VAR = 1, B2_ERR = 2, B2_MIXED = 3 };

// Jets of the upper triangle of A = -i dt H at one time step.
//   O = 0            : slot 0 = H0
//   O = 1, B2_VAR    : slot 1 = difference of H0 in variable v with the step actually taken at eps (:50)
//   O = 1, B2_ERR    : slot 1 = error Hamiltonian es at err = eps (:66-68)
//   O = 2, B2_MIXED  : a = difference of H0 in v at eps2, b = error Hamiltonian at eps2, ab = its difference in v (:76-79)
template <int D, unsigned UMASK, int O>
__device__ __forceinline__ void b2_assemble(const DevProblem& P, const StagedPlan& sp, const double* xk, const double* xadd, int k,
                                            int kind, int v, int es, cplx (&tj)[Tri<D>::n][1 << O], const TrigSlots& tr) {
    constexpr int NP = Tri<D>::n;
    constexpr int n = 1 << O;
    int sp_ = RG_S_NONE, ix = 0;
    double h = 0.0, errv = 0.0;
    if (kind == B2_VAR || kind == B2_MIXED) {
        sp_ = P.var_space[v]; ix = P.var_index[v];
        const double val = (sp_ == RG_S_MAIN) ? xk[ix] : xadd[ix];
        const double e = (kind == B2_VAR) ? P.eps : P.eps2;
        h = __dsub_rn(__dadd_rn(val, e), val);
    }
    if (kind == B2_ERR) errv = P.eps;
    if (kind == B2_MIXED) errv = P.eps2;
    EvalCtx ec{xk, xadd, errv, P.table, P.N, k, tr.n, tr.s0, tr.c0, tr.s1, tr.c1};
#pragma unroll
    for (int pos = 0; pos < NP; ++pos)
#pragma unroll
        for (int s = 0; s < n; ++s) tj[pos][s] = cmk(0.0, 0.0);
    for (int t = 0; t < P.nterms; ++t) {
        const DevTerm& tm = sp.terms[t];
        const bool isH0 = tm.owner == RG_OWNER_H0;
        if (!sp.used[t] || !(isH0 || ((kind == B2_ERR || kind == B2_MIXED) && tm.owner == es))) continue;
        cplx base, del;
        term_coef(tm, ec, sp_, ix, h, base, del);
        const cplx cb = cmk(base.y * P.dt, -base.x * P.dt), cd = cmk(del.y * P.dt, -del.x * P.dt);   // (-i dt) * coefficient
        const cplx* dv = sp.dense + t * NP;
#pragma unroll
        for (int pos = 0; pos < NP; ++pos) {
            if (!((UMASK >> pos) & 1u)) continue;
            const cplx m = dv[pos];
            if (m.x == 0.0 && m.y == 0.0) continue;            // warp-uniform
            if (isH0) {
                cfma(tj[pos][0], cb, m);
                if constexpr (O >= 1) { if (kind != B2_ERR) cfma(tj[pos][1], cd, m); }
            } else {
                if constexpr (O == 1) cfma(tj[pos][1], cb, m);
                if constexpr (O == 2) { cfma(tj[pos][2], cb, m); cfma(tj[pos][3], cd, m); }
            }
        }
    }
}

// Step propagator (and its differences) as pattern matrices: out[s] = slot s of exp(A-jet).
template <int D, unsigned UMASK, int O>
struct B2Blocks {
    typedef Pat<D, stored_from_tri(D, UMASK)> PT;
    typedef PMat<D, stored_from_tri(D, UMASK)> M;
    template <int l>
    static __device__ __forceinline__ void run(const cplx (&tj)[Tri<D>::n][1 << O], M (&out)[1 << O], int& Kmax) {
        constexpr int n = 1 << O;
        if constexpr (l < D) {
            constexpr int pr = b2_partner(D, UMASK, l);
            if constexpr (pr > l) {
                constexpr bool DIAG = ((UMASK >> Tri<D>::idx(l, l)) & 1u) || ((UMASK >> Tri<D>::idx(pr, pr)) & 1u);
                double a1[n], a2[n];
                cplx w[n], u11[n], u12[n], u21[n], u22[n];
#pragma unroll
                for (int s = 0; s < n; ++s) {
                    a1[s] = ((UMASK >> Tri<D>::idx(l, l)) & 1u) ? tj[Tri<D>::idx(l, l)][s].y : 0.0;
                    a2[s] = ((UMASK >> Tri<D>::idx(pr, pr)) & 1u) ? tj[Tri<D>::idx(pr, pr)][s].y : 0.0;
                    w[s] = tj[Tri<D>::idx(l, pr)][s];
                }
                const int K = block2_exp<O, DIAG>(a1, a2, w, u11, u12, u21, u22);
                Kmax = max(Kmax, K);
#pragma unroll
                for (int s = 0; s < n; ++s) {
                    out[s].v[PT::idx(l, l)] = u11[s]; out[s].v[PT::idx(l, pr)] = u12[s];
                    out[s].v[PT::idx(pr, l)] = u21[s]; out[s].v[PT::idx(pr, pr)] = u22[s];
                }
            } else if constexpr (pr == -1 && PT::has(l, l)) {
                // one-level block with a diagonal term: U = e^{i a}
                double a[n];
                cplx E[n];
#pragma unroll
                for (int s = 0; s < n; ++s) a[s] = tj[Tri<D>::idx(l, l)][s].y;
                jexpi<O>(a, E);
#pragma unroll
                for (int s = 0; s < n; ++s) out[s].v[PT::idx(l, l)] = E[s];
            }
            run<l + 1>(tj, out, Kmax);
        }
    }
};

template <int D, unsigned UMASK, int O>
__device__ __forceinline__ int b2_step(const DevProblem& P, const StagedPlan& sp, const double* xk, const double* xadd, int k,
                                       int kind, int v, int es, PMat<D, stored_from_tri(D, UMASK)> (&out)[1 << O], const TrigSlots& tr) {
    cplx tj[Tri<D>::n][1 << O];
    b2_assemble<D, UMASK, O>(P, sp, xk, xadd, k, kind, v, es, tj, tr);
    int Kmax = 0;
    B2Blocks<D, UMASK, O>::template run<0>(tj, out, Kmax);
    return Kmax;
}

#ifndef RG_B2_AGG_CTAS
#define RG_B2_AGG_CTAS 3
#endif
#ifndef RG_B2_GRAD_CTAS
#define RG_B2_GRAD_CTAS 2
#endif

// ---- chunk aggregates, one thread per (pulse, chunk): Q <- U_k Q ; Wl_e <- U_k Wl_e + D_k^e Q_old (recomputed U_k, D_k)
template <int D, unsigned UMASK>
__global__ void __launch_bounds__(128, RG_B2_AGG_CTAS)
k_agg_b2(const DevProblem P, const TriPlanDev tp, const double* __restrict__ X, int B, int L, int nc, cplx* __restrict__ Qb,
         cplx* __restrict__ Wlb, int* __restrict__ status) {
    constexpr u64 CMS = stored_from_tri(D, UMASK);
    typedef PMat<D, CMS> M;
    constexpr int DD = D * D;
    extern __shared__ cplx smem[];
    const StagedPlan sp = stage_plan(P, tp, reinterpret_cast<unsigned char*>(smem));
    const long long total = (long long)B * nc;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = item < total;
    if (!live) item = total - 1;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    int Kmax = 0;
    M q; q.identity();
    if (ne == 0) {
        for (int k = k0; k < k1; ++k) {
            for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
            const TrigSlots tr = trig_eval(P, xk, xadd);
            M u[1];
            Kmax = max(Kmax, b2_step<D, UMASK, 0>(P, sp, xk, xadd, k, B2_VALUE, 0, 0, u, tr));
            M qn; pmat_mul<D, CMS, false, false>(qn, u[0], q);
            q = qn;
        }
    } else {
        for (int e = 0; e < ne; ++e) {
            q.identity();
            M wl; wl.zero();
            for (int k = k0; k < k1; ++k) {
                for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
                const TrigSlots tr = trig_eval(P, xk, xadd);
                M u[2];
                Kmax = max(Kmax, b2_step<D, UMASK, 1>(P, sp, xk, xadd, k, B2_ERR, 0, e, u, tr));
                M wn; pmat_mul<D, CMS, false, false>(wn, u[0], wl); pmat_mul<D, CMS, false, true>(wn, u[1], q);
                M qn; pmat_mul<D, CMS, false, false>(qn, u[0], q);
                wl = wn; q = qn;
            }
            if (live) pmat_to_dense<D, CMS>(wl, Wlb + (((size_t)b * nc + ch) * ne + e) * DD, false);
        }
    }
    if (live) pmat_to_dense<D, CMS>(q, Qb + ((size_t)b * nc + ch) * DD, true);
    if (Kmax == 99) atomicOr(status, 2);
}

// ---- backward gradient sweep, fidelity role: out0[b*nx + p*k + v] = scale0 * Re tr(G_k dU_k^v C_{k-1})
template <int D, unsigned UMASK>
__global__ void __launch_bounds__(128, RG_B2_GRAD_CTAS)
k_grad_b2(const DevProblem P, const TriPlanDev tp, const double* __restrict__ X, int B, int L, int nc, const cplx* __restrict__ Cb,
          const cplx* __restrict__ Gb, double* __restrict__ out0, double scale0, double* __restrict__ addS) {
    constexpr u64 CMS = stored_from_tri(D, UMASK);
    typedef Pat<D, CMS> PT;
    typedef PMat<D, CMS> M;
    constexpr int DD = D * D;
    extern __shared__ cplx smem[];
    const StagedPlan sp = stage_plan(P, tp, reinterpret_cast<unsigned char*>(smem));
    const long long total = (long long)B * nc;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int nv = P.nvar, ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    M c, g;
    pmat_from_dense<D, CMS>(c, Cb + ((size_t)b * nc + ch) * DD);
    {   // k_scan stores the co-state by rows: G(i,j) = Gb[i*D + j]
        const cplx* gp = Gb + ((size_t)b * nc + ch) * DD;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (PT::has(i, j)) g.v[PT::idx(i, j)] = gp[i * D + j];
    }
    for (int k = k1 - 1; k >= k0; --k) {
        for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
        const TrigSlots tr = trig_eval(P, xk, xadd);
        M u;
        M cp;
        for (int v = 0; v < max(nv, 1); ++v) {
            M ud[2];
            b2_step<D, UMASK, 1>(P, sp, xk, xadd, k, nv ? B2_VAR : B2_VALUE, v, 0, ud, tr);
            if (v == 0) { u = ud[0]; pmat_mul<D, CMS, true, false>(cp, u, c); }      // C_{k-1} = U_k^dag C_k
            if (nv == 0) break;
            M t; pmat_mul<D, CMS, false, false>(t, ud[1], cp);                        // dU C_{k-1}
            const double s = pmat_retrace<D, CMS>(g, t) * scale0;
            if (P.var_space[v] == RG_S_MAIN) out0[(size_t)b * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
            else addS[(((size_t)b * (1 + ne)) * P.a + P.var_index[v]) * P.N + k] = s;
        }
        M gn; pmat_mul<D, CMS, false, false>(gn, g, u);                               // G_{k-1} = G_k U_k
        g = gn; c = cp;
    }
}

// ---- backward sweep of the sensitivity gradient, error source e = blockIdx.y (see k_grad_err_t)
template <int D, unsigned UMASK>
__global__ void __launch_bounds__(128)
k_grad_err_b2(const DevProblem P, const TriPlanDev tp, const double* __restrict__ X, int B, int L, int nc,
              const cplx* __restrict__ Cb, const cplx* __restrict__ Wb, const cplx* __restrict__ G1b, const cplx* __restrict__ H1b,
              double* __restrict__ out1, double* __restrict__ addS) {
    constexpr u64 CMS = stored_from_tri(D, UMASK);
    typedef Pat<D, CMS> PT;
    typedef PMat<D, CMS> M;
    constexpr int DD = D * D;
    extern __shared__ cplx smem[];
    const StagedPlan sp = stage_plan(P, tp, reinterpret_cast<unsigned char*>(smem));
    const long long total = (long long)B * nc;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int es = blockIdx.y;
    const int nv = P.nvar, ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    const double f1 = 2.0 / DD1 * P.inv_eps * P.inv_eps, f2 = 2.0 / DD1 * P.inv_eps2sq;
    M c, w, g, h;
    pmat_from_dense<D, CMS>(c, Cb + ((size_t)b * nc + ch) * DD);
    {
        const size_t off = (((size_t)b * ne + es) * nc + ch) * DD;
        pmat_from_dense<D, CMS>(w, Wb + off);
        const cplx* gp = G1b + off;          // co-states are stored by rows
        const cplx* hp = H1b + off;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (PT::has(i, j)) { g.v[PT::idx(i, j)] = gp[i * D + j]; h.v[PT::idx(i, j)] = hp[i * D + j]; }
    }
    for (int k = k1 - 1; k >= k0; --k) {
        for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
        const TrigSlots tr = trig_eval(P, xk, xadd);
        M ue[2];                               // U_k and D_k^e at err = eps
        b2_step<D, UMASK, 1>(P, sp, xk, xadd, k, B2_ERR, 0, es, ue, tr);
        {   // rewind: C_{k-1} = U^dag C_k ;  W_{k-1} = U^dag (W_k - D_k C_{k-1})
            M cp; pmat_mul<D, CMS, true, false>(cp, ue[0], c);
            c = cp;
            M t; pmat_mul<D, CMS, false, false>(t, ue[1], c);
#pragma unroll
            for (int i = 0; i < PT::nnz; ++i) t.v[i] = csub(w.v[i], t.v[i]);
            pmat_mul<D, CMS, true, false>(w, ue[0], t);
        }
        for (int v = 0; v < nv; ++v) {
            double s1, s2;
            {
                M ud[2];
                b2_step<D, UMASK, 1>(P, sp, xk, xadd, k, B2_VAR, v, 0, ud, tr);
                M t; pmat_mul<D, CMS, false, false>(t, ud[1], c);
                s1 = pmat_retrace<D, CMS>(h, t);
                pmat_mul<D, CMS, false, false>(t, ud[1], w);
                s1 += pmat_retrace<D, CMS>(g, t);
            }
            {
                M u4[4];
                b2_step<D, UMASK, 2>(P, sp, xk, xadd, k, B2_MIXED, v, es, u4, tr);
                M t; pmat_mul<D, CMS, false, false>(t, u4[3], c);
                s2 = pmat_retrace<D, CMS>(g, t);
            }
            const double s = f1 * s1 + f2 * s2;
            if (P.var_space[v] == RG_S_MAIN) out1[((size_t)b * ne + es) * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
            else addS[(((size_t)b * (1 + ne) + 1 + es) * P.a + P.var_index[v]) * P.N + k] = s;
        }
        {   // advance: H' <- H' U + G' D ;  G' <- G' U
            M hn; pmat_mul<D, CMS, false, false>(hn, h, ue[0]); pmat_mul<D, CMS, false, true>(hn, g, ue[1]);
            M gn; pmat_mul<D, CMS, false, false>(gn, g, ue[0]);
            h = hn; g = gn;
        }
    }
}
