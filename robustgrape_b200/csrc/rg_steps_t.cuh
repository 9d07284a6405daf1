// rg_steps_t.cuh -- thread-per-time-step propagator kernel for Hermitian problems with d <= 5, and the
// chunk-aggregate kernel that follows it.
//
// k_steps_t: one thread owns one (pulse, time step).  A = -i dt H and every perturbation matrix dA are
//   skew-Hermitian, so their upper triangles (d(d+1)/2 complex numbers each) live in registers; the thread
//   runs the (value, difference) Horner recurrence of rg_smalld.cuh column by column.  No shared-memory
//   operand traffic, no warp synchronisation, all 32 lanes busy, and the coefficient evaluation (sincos)
//   is done once per step instead of once per lane of a group.  Structural zeros of H (positions of the
//   upper triangle no term touches) are skipped with warp-uniform branches when MASKED.
// k_chunk_agg: per (pulse, chunk) group: Q_c = U_last ... U_first and Wl_c = dQ_c/derr from the stored
//   step matrices (src/UnitaryCalculations.jl:46 restricted to a chunk).
#pragma once
#include "rg_smalld.cuh"

#define RG_T_MAX_TERMS 16
#ifndef RG_STEPS_T_CTAS
#define RG_STEPS_T_CTAS 6      // resident CTAs per SM requested for the sparse-pattern k_steps_t (register cap 65536 / (128 n))
#endif
#ifndef RG_SO_T_CTAS
#define RG_SO_T_CTAS 4
#endif
#ifndef RG_AGG_T_CTAS
#define RG_AGG_T_CTAS 3        // 168 registers, no spills; the gradient sweeps spill (and slow down) under any cap
#endif

struct TriPlanDev {
    int nent;                 // plan entries
    const int* ptr;           // [npos+1] entry range of each upper-triangle position
    const int* term;          // [nent] term index
    const double* val;        // [nent*2] matrix value (re, im)
    const double* colw;       // [nterms*d] column 1-norms of each term's matrix (for the ||A||_1 bound)
    const int* used;          // [nterms] term has upper-triangle entries
    unsigned maskA;           // positions touched by H0 terms
    unsigned maskVar[RG_MAX_VARS];   // ... by H0 terms depending on variable v
    unsigned maskErr[RG_MAX_ERR];    // ... by terms of error source e
};

struct StagedPlan {
    const int* ptr; const int* term; const cplx* val; const double* colw; const int* used; const DevTerm* terms;
    const cplx* dense;        // [nterms][npos] matrix value of term t at upper-triangle position pos (0 where absent)
};
__host__ __device__ inline size_t staged_plan_bytes(int nterms, int nent, int d) {
    const int npos = d * (d + 1) / 2;
    return rg_align16((size_t)nterms * sizeof(DevTerm)) + rg_align16((size_t)(npos + 1) * 4) + rg_align16((size_t)nent * 4) +
           rg_align16((size_t)nent * 16) + rg_align16((size_t)nterms * d * 8) + rg_align16((size_t)nterms * 4) +
           rg_align16((size_t)nterms * npos * 16);
}
__device__ inline StagedPlan stage_plan(const DevProblem& P, const TriPlanDev& tp, unsigned char* sm) {
    const int npos = P.d * (P.d + 1) / 2;
    unsigned char* p = sm;
    int* t32 = reinterpret_cast<int*>(p); p += rg_align16((size_t)P.nterms * sizeof(DevTerm));
    int* ptr = reinterpret_cast<int*>(p); p += rg_align16((size_t)(npos + 1) * 4);
    int* term = reinterpret_cast<int*>(p); p += rg_align16((size_t)tp.nent * 4);
    double* val = reinterpret_cast<double*>(p); p += rg_align16((size_t)tp.nent * 16);
    double* colw = reinterpret_cast<double*>(p); p += rg_align16((size_t)P.nterms * P.d * 8);
    int* used = reinterpret_cast<int*>(p); p += rg_align16((size_t)P.nterms * 4);
    double* dense = reinterpret_cast<double*>(p);
    for (int i = threadIdx.x; i < 2 * P.nterms * npos; i += blockDim.x) dense[i] = 0.0;
    const int nt4 = P.nterms * (int)(sizeof(DevTerm) / 4);
    for (int i = threadIdx.x; i < nt4; i += blockDim.x) t32[i] = reinterpret_cast<const int*>(P.terms)[i];
    for (int i = threadIdx.x; i <= npos; i += blockDim.x) ptr[i] = tp.ptr[i];
    for (int i = threadIdx.x; i < tp.nent; i += blockDim.x) { term[i] = tp.term[i]; val[2 * i] = tp.val[2 * i]; val[2 * i + 1] = tp.val[2 * i + 1]; }
    for (int i = threadIdx.x; i < P.nterms * P.d; i += blockDim.x) colw[i] = tp.colw[i];
    for (int i = threadIdx.x; i < P.nterms; i += blockDim.x) used[i] = tp.used[i];
    __syncthreads();
    for (int pos = threadIdx.x; pos < npos; pos += blockDim.x)          // one thread per position: no write conflicts
        for (int e = ptr[pos]; e < ptr[pos + 1]; ++e) {
            dense[2 * (term[e] * npos + pos)] += val[2 * e];
            dense[2 * (term[e] * npos + pos) + 1] += val[2 * e + 1];
        }
    __syncthreads();
    StagedPlan s{ptr, term, reinterpret_cast<const cplx*>(val), colw, used, reinterpret_cast<const DevTerm*>(t32),
                 reinterpret_cast<const cplx*>(dense)};
    return s;
}

template <int D, unsigned UMASK, int l, bool CHECK>
__device__ __forceinline__ void horner_loop_tri(const cplx (&ta)[Tri<D>::n], const cplx (&td)[Tri<D>::n], unsigned mA, unsigned mD,
                                                int m, cplx (&y)[D], cplx (&dl)[D]) {
    typedef Pat<D, closure_from_tri(D, UMASK)> PT;
    for (int j = m - 1; j >= 1; --j) {
        const double inv = c_inv_j[j];
        cplx t[D], u[D];
#pragma unroll
        for (int i = 0; i < D; ++i) { t[i] = cmk(0.0, 0.0); u[i] = cmk(0.0, 0.0); }
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int i = 0; i <= k; ++i) {
                const int pos = Tri<D>::idx(i, k);
                if (!((UMASK >> pos) & 1u)) continue;
                if (!PT::has(k, l) && !PT::has(i, l)) continue;      // both operands structurally zero in this column
                if (!CHECK || ((mA >> pos) & 1u)) {
                    const cplx a = ta[pos];
                    if (PT::has(k, l)) { cfma(t[i], a, y[k]); cfma(u[i], a, dl[k]); }
                    if (i != k && PT::has(i, l)) { cfma_nconj(t[k], a, y[i]); cfma_nconj(u[k], a, dl[i]); }
                }
                if (!CHECK || ((mD >> pos) & 1u)) {
                    const cplx d = td[pos];
                    if (PT::has(k, l)) cfma(u[i], d, cadd(y[k], dl[k]));
                    if (i != k && PT::has(i, l)) cfma_nconj(u[k], d, cadd(y[i], dl[i]));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            y[i] = cscale(t[i], inv);
            if (i == l) y[i].x += 1.0;
            dl[i] = cscale(u[i], inv);
        }
    }
}

// (value, difference) Horner recurrence for one column l with register-resident triangles.
// UMASK: compile-time superset of every position any matrix of this problem touches (pruned code and
// registers for the rest); mA / mD: run-time (warp-uniform) masks of A and of the current dA, tested only
// when UMASK is not the full triangle.
template <int D, unsigned UMASK, int l>
__device__ __forceinline__ void horner_col_tri(const cplx (&ta)[Tri<D>::n], const cplx (&td)[Tri<D>::n], unsigned mA, unsigned mD,
                                               int m, cplx (&y)[D], cplx (&dl)[D]) {
    constexpr bool MASKED = (UMASK != ((1u << Tri<D>::n) - 1u));
    // column l of exp(A) is confined to the rows the closure pattern allows: y[k] == 0 elsewhere
    typedef Pat<D, closure_from_tri(D, UMASK)> PT;
    {
        const double inv = c_inv_j[m];
#pragma unroll
        for (int i = 0; i < D; ++i) { y[i] = cmk(0.0, 0.0); dl[i] = cmk(0.0, 0.0); }
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int i = 0; i <= k; ++i) {
                if (!((UMASK >> Tri<D>::idx(i, k)) & 1u)) continue;
                const cplx a = ta[Tri<D>::idx(i, k)], d = td[Tri<D>::idx(i, k)];
                if (k == l) { y[i] = cscale(a, inv); dl[i] = cscale(d, inv); }
                if (i == l && i != k) { y[k] = cscale(cmk(-a.x, a.y), inv); dl[k] = cscale(cmk(-d.x, d.y), inv); }
            }
#pragma unroll
        for (int i = 0; i < D; ++i) if (i == l) y[i].x += 1.0;
    }
    // run-time masks cost an ISETP per position and iteration; when both matrices fill the compile-time pattern (the usual
    // case for the pre-instantiated Rydberg patterns) take the test-free loop
    if (!MASKED || ((mA & mD & UMASK) == UMASK)) horner_loop_tri<D, UMASK, l, false>(ta, td, mA, mD, m, y, dl);
    else horner_loop_tri<D, UMASK, l, true>(ta, td, mA, mD, m, y, dl);
}

// Column l of a compact (pattern) matrix to global memory. With an even number of stored elements every matrix starts on
// a 32-byte boundary, so two consecutive stored elements of the column whose first compact index is even go out as one
// STG.256.
template <int D, u64 CMS>
__host__ __device__ constexpr int pat_next_row(int i, int l) {          // next stored row after i in column l (D if none)
    for (int r = i + 1; r < D; ++r)
        if (Pat<D, CMS>::has(r, l)) return r;
    return D;
}
template <int D, u64 CMS>
__host__ __device__ constexpr int pat_prev_row(int i, int l) {          // previous stored row before i in column l (-1 if none)
    for (int r = i - 1; r >= 0; --r)
        if (Pat<D, CMS>::has(r, l)) return r;
    return -1;
}
// (row index as a template parameter: every pattern query below is a constant expression, nothing is evaluated at run time)
template <int D, u64 CMS, int l, int i>
__device__ __forceinline__ void store_col_row(cplx* __restrict__ dst, const cplx (&v)[D]) {
    typedef Pat<D, CMS> PT;
    if constexpr (i < D) {
        if constexpr (PT::has(i, l)) {
            constexpr bool PAIRS = (PT::nnz & 1) == 0;
            constexpr int prev = pat_prev_row<D, CMS>(i, l), next = pat_next_row<D, CMS>(i, l);
            constexpr int me = PT::idx(i, l);
            // pairs are formed greedily from the top of the column: element i is the second of a pair iff the element before
            // it has an even compact index (the compact indices of a column are consecutive)
            constexpr bool second = PAIRS && prev >= 0 && ((me - 1) & 1) == 0;
            constexpr bool first = PAIRS && next < D && (me & 1) == 0;
            if constexpr (!second) {
                if constexpr (first) st256(dst + me, v[i], v[next < D ? next : i]);
                else dst[me] = v[i];
            }
        }
        store_col_row<D, CMS, l, i + 1>(dst, v);
    }
}
template <int D, u64 CMS, int l>
__device__ __forceinline__ void store_col(cplx* __restrict__ dst, const cplx (&v)[D]) {
    store_col_row<D, CMS, l, 0>(dst, v);
}

template <int D, unsigned UMASK, int l>
__device__ __forceinline__ void columns(const cplx (&ta)[Tri<D>::n], const cplx (&td)[Tri<D>::n], unsigned mA, unsigned mD, int m,
                                        bool live, cplx* __restrict__ dstD, cplx* __restrict__ dstU, cplx* smU = nullptr) {
    typedef Pat<D, stored_from_tri(D, UMASK)> PT;
    if constexpr (l < D) {
        if constexpr (PT::has(l, l)) {              // inert levels: U(l,l) = 1, dU = 0, nothing computed or stored
            cplx y[D], dl[D];
            horner_col_tri<D, UMASK, l>(ta, td, mA, mD, m, y, dl);
            if (smU) {                              // fused chunk aggregate: U of this step, element-major in shared memory
#pragma unroll
                for (int i = 0; i < D; ++i)
                    if (PT::has(i, l)) smU[PT::idx(i, l) * 128] = live ? y[i] : cmk(i == l ? 1.0 : 0.0, 0.0);
            }
            if (live) {
                if (dstD) store_col<D, stored_from_tri(D, UMASK), l>(dstD, dl);
                if (dstU) store_col<D, stored_from_tri(D, UMASK), l>(dstU, y);
            }
        }
        columns<D, UMASK, l + 1>(ta, td, mA, mD, m, live, dstD, dstU, smU);
    }
}

// ---- matrices restricted to a compile-time pattern (used by the thread-per-chunk sweeps and the fused aggregate)
template <int D, u64 CM>
struct PMat {                       // matrix restricted to the pattern, stored compactly
    cplx v[Pat<D, CM>::nnz];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < Pat<D, CM>::nnz; ++i) v[i] = cmk(0.0, 0.0);
    }
    __device__ __forceinline__ void identity() {
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (Pat<D, CM>::has(i, j)) v[Pat<D, CM>::idx(i, j)] = cmk(i == j ? 1.0 : 0.0, 0.0);
    }
    static __device__ __forceinline__ void prefetch(const cplx* __restrict__ p) {
#pragma unroll
        for (int i = 0; i < Pat<D, CM>::nnz; i += 8) prefetch_line(p + i);           // 8 complex numbers per 128-byte line
    }
    __device__ __forceinline__ void load(const cplx* __restrict__ p) {
        if constexpr ((Pat<D, CM>::nnz & 1) == 0) {        // matrices start on 32-byte boundaries: 256-bit loads
#pragma unroll
            for (int i = 0; i < Pat<D, CM>::nnz; i += 2) ld256(p + i, v[i], v[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < Pat<D, CM>::nnz; ++i) v[i] = p[i];
        }
    }
};
// dense d x d (column-major) <-> pattern
template <int D, u64 CM>
__device__ __forceinline__ void pmat_from_dense(PMat<D, CM>& m, const cplx* __restrict__ p) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (Pat<D, CM>::has(i, j)) m.v[Pat<D, CM>::idx(i, j)] = p[i + D * j];
}
// unit_inert: entries (l,l) of inert levels (diagonal bit absent) are written as 1 (products of propagators), else 0
template <int D, u64 CM>
__device__ __forceinline__ void pmat_to_dense(const PMat<D, CM>& m, cplx* __restrict__ p, bool unit_inert) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
            p[i + D * j] = Pat<D, CM>::has(i, j) ? m.v[Pat<D, CM>::idx(i, j)] : cmk((i == j && unit_inert) ? 1.0 : 0.0, 0.0);
}
// C = op(A) B with op = none (ADJ = false) or conjugate transpose (ADJ = true); the closure pattern is closed under
// products and adjoints, so the result stays inside it.
template <int D, u64 CM, bool ADJ, bool ACC>
__device__ __forceinline__ void pmat_mul(PMat<D, CM>& c, const PMat<D, CM>& a, const PMat<D, CM>& b) {
    typedef Pat<D, CM> PT;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (!PT::has(i, j)) continue;
            cplx acc = ACC ? c.v[PT::idx(i, j)] : cmk(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (!PT::has(k, j)) continue;
                if (ADJ) { if (PT::has(k, i)) cfma_conj(acc, a.v[PT::idx(k, i)], b.v[PT::idx(k, j)]); }
                else { if (PT::has(i, k)) cfma(acc, a.v[PT::idx(i, k)], b.v[PT::idx(k, j)]); }
            }
            c.v[PT::idx(i, j)] = acc;
        }
}
// Re tr(A B)
template <int D, u64 CM>
__device__ __forceinline__ double pmat_retrace(const PMat<D, CM>& a, const PMat<D, CM>& b) {
    typedef Pat<D, CM> PT;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (PT::has(i, j) && PT::has(j, i)) {
                const cplx x = a.v[PT::idx(i, j)], y = b.v[PT::idx(j, i)];
                s = fma(x.x, y.x, s); s = fma(-x.y, y.y, s);
            }
    return s;
}

template <int D, unsigned UMASK>
__global__ void __launch_bounds__(128, (UMASK == ((1u << (D * (D + 1) / 2)) - 1u)) ? 2 : RG_STEPS_T_CTAS)
k_steps_t(const DevProblem P, const TriPlanDev tp, const double* __restrict__ X, int B, cplx* __restrict__ ws,
          int* __restrict__ status, int aggL = 0, int nc = 0, cplx* __restrict__ Qb = nullptr, int agg_off = 0) {
    // aggL > 0 (a power of two <= 32, no error sources): the time axis is padded to nc*aggL so that every aligned group of
    // aggL lanes is one chunk of one pulse, and the chunk aggregate Q = U_last ... U_first is formed here by a tree product in
    // shared memory instead of re-reading the step matrices in k_chunk_agg_t.
    constexpr int NP = Tri<D>::n;
    typedef Pat<D, stored_from_tri(D, UMASK)> PT;
    extern __shared__ cplx smem[];
    const int Np = aggL ? nc * aggL : P.N;                       // padded steps per pulse
    const long long total = (long long)B * Np;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_grid = item < total;
    if (!in_grid) item = total - 1;
    int b, kp;
    if (total < (1ll << 31)) { b = (int)((unsigned)item / (unsigned)Np); kp = (int)((unsigned)item - (unsigned)b * (unsigned)Np); }
    else { b = (int)(item / Np); kp = (int)(item % Np); }
    const bool live = in_grid && kp < P.N;
    const int k = min(kp, P.N - 1);
    // the control values are requested before the plan is staged so that their latency hides behind it
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
    const StagedPlan sp = stage_plan(P, tp, reinterpret_cast<unsigned char*>(smem));
    cplx* smU = aggL ? reinterpret_cast<cplx*>(reinterpret_cast<unsigned char*>(smem) + agg_off) + threadIdx.x : nullptr;
    cplx* wsk = ws + ((size_t)b * P.N + k) * (size_t)PT::nnz;            // object 0 of this step
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;                    // stride between objects
    const int nt = P.nterms, nv = P.nvar, ne = P.e, nfo = nv + ne;

    cplx ta[NP], td[NP];
    int m = 0;
    for (int o = 0; o < max(nfo, 1); ++o) {
        int sp_ = RG_S_NONE, ix = 0, es = -3;
        double h = 0.0, errv = 0.0;
        if (nfo > 0 && o < nv) {
            sp_ = P.var_space[o]; ix = P.var_index[o];
            const double v = (sp_ == RG_S_MAIN) ? xk[ix] : xadd[ix];
            h = __dsub_rn(__dadd_rn(v, P.eps), v);                 // the step actually taken (:50)
        } else if (nfo > 0) { es = o - nv; errv = P.eps; }
        EvalCtx ec{xk, xadd, errv, P.table, P.N, k};
        // ---- triangles of A = -i dt H (first pass) and of this pass's dA, one term at a time: the coefficient is consumed
        // as soon as it is known (no per-thread coefficient arrays), and the column-sum bound of ||A||_1 rides along
        double colsum[D];
#pragma unroll
        for (int kk = 0; kk < D; ++kk) colsum[kk] = 0.0;
#pragma unroll
        for (int pos = 0; pos < NP; ++pos) { if (o == 0) ta[pos] = cmk(0, 0); td[pos] = cmk(0, 0); }
        for (int t = 0; t < nt; ++t) {
            const DevTerm& tm = sp.terms[t];
            const bool isH0 = tm.owner == RG_OWNER_H0;
            // H0 terms: value on the first pass, difference on variable passes; error terms: only on their own pass
            if (!sp.used[t] || !((isH0 && (o == 0 || es < 0)) || tm.owner == es)) continue;
            cplx base, del;
            term_coef(tm, ec, sp_, ix, h, base, del);
            const cplx ca = isH0 ? cmk(base.y * P.dt, -base.x * P.dt) : cmk(0, 0);     // (-i dt) * coefficient
            const cplx dsel = isH0 ? del : base;                   // error terms enter with their value at err = eps
            const cplx cd = (isH0 && es >= 0) ? cmk(0, 0) : cmk(dsel.y * P.dt, -dsel.x * P.dt);
            const cplx* dv = sp.dense + t * NP;
#pragma unroll
            for (int pos = 0; pos < NP; ++pos) {
                if (!((UMASK >> pos) & 1u)) continue;
                const cplx v = dv[pos];
                if (v.x == 0.0 && v.y == 0.0) continue;            // warp-uniform
                if (o == 0) cfma(ta[pos], ca, v);
                cfma(td[pos], cd, v);
            }
            if (o == 0 && isH0) {
                // |c| <= max(|re|,|im|) + (sqrt2 - 1) min(|re|,|im|)  (strict upper bound, <= 8.3 % high; no square roots)
                const double ax = fabs(ca.x), ay = fabs(ca.y);
                const double w = fmax(ax, ay) + 0.41421356237309515 * fmin(ax, ay);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) colsum[kk] = fma(w, sp.colw[t * D + kk], colsum[kk]);
            }
        }
        if (o == 0) {
            // ||A||_1 <= max_k sum_t |c_t| * (column-k weight of M_t, upper triangle mirrored: A is skew-Hermitian)
            double nrm = 0.0;
#pragma unroll
            for (int kk = 0; kk < D; ++kk) nrm = fmax(nrm, colsum[kk]);
            m = taylor_degree(nrm * 1.001 + 2.0 * P.eps2 * P.dt);
            m = __reduce_max_sync(0xffffffffu, m);
            if (m == 99) { if ((threadIdx.x & 31) == 0) atomicOr(status, 2); m = 18; }
        }
        const unsigned mD = (nfo == 0) ? 0u : (o < nv ? tp.maskVar[o] : tp.maskErr[o - nv]);
        // ---- columns (unrolled: the pattern of each column is known at compile time)
        columns<D, UMASK, 0>(ta, td, tp.maskA, mD, m, live, nfo > 0 ? wsk + (size_t)(1 + o) * objS : nullptr,
                             o == 0 ? wsk : nullptr, o == 0 ? smU : nullptr);
    }
    if constexpr (PT::nnz <= 12) if (aggL) {
        // tree product over the aggL lanes of a chunk: after level s lane t (t % 2s == 0) holds U_{t+2s-1} ... U_t
        typedef PMat<D, stored_from_tri(D, UMASK)> M;
        const int lane = threadIdx.x & 31;
        __syncwarp();
        for (int s = 1; s < aggL; s <<= 1) {
            if ((lane & (2 * s - 1)) == 0) {
                M lo, hi, pr;
#pragma unroll
                for (int i = 0; i < PT::nnz; ++i) { lo.v[i] = smU[i * 128]; hi.v[i] = smU[i * 128 + s]; }
                pmat_mul<D, stored_from_tri(D, UMASK), false, false>(pr, hi, lo);
#pragma unroll
                for (int i = 0; i < PT::nnz; ++i) smU[i * 128] = pr.v[i];
            }
            __syncwarp();
        }
        if ((lane & (aggL - 1)) == 0 && in_grid) {
            M q;
#pragma unroll
            for (int i = 0; i < PT::nnz; ++i) q.v[i] = smU[i * 128];
            pmat_to_dense<D, stored_from_tri(D, UMASK)>(q, Qb + ((size_t)b * nc + kp / aggL) * (D * D), true);
        }
    }
}


// ------------------------------------------------------------------ mixed second differences, thread per step
// t += M v for a skew-Hermitian M given by its upper triangle; column l of the closure pattern decides which
// components of v can be non-zero.  mrt: run-time (warp-uniform) mask of M within the compile-time UMASK.
template <int D, unsigned UMASK, int l, bool CHECK = true>
__device__ __forceinline__ void tri_matvec_acc(cplx (&t)[D], const cplx (&tri)[Tri<D>::n], unsigned mrt, const cplx (&v)[D]) {
    typedef Pat<D, closure_from_tri(D, UMASK)> PT;
    constexpr bool MASKED = CHECK && (UMASK != ((1u << Tri<D>::n) - 1u));
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            const int pos = Tri<D>::idx(i, k);
            if (!((UMASK >> pos) & 1u)) continue;
            if (!PT::has(k, l) && !PT::has(i, l)) continue;
            if (MASKED && !((mrt >> pos) & 1u)) continue;
            const cplx a = tri[pos];
            if (PT::has(k, l)) cfma(t[i], a, v[k]);
            if (i != k && PT::has(i, l)) cfma_nconj(t[k], a, v[i]);
        }
}
// column l of a skew-Hermitian matrix given by its triangle, scaled
template <int D, unsigned UMASK, int l>
__device__ __forceinline__ void tri_column(cplx (&y)[D], const cplx (&tri)[Tri<D>::n], double sc) {
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            if (!((UMASK >> Tri<D>::idx(i, k)) & 1u)) continue;
            const cplx a = tri[Tri<D>::idx(i, k)];
            if (k == l) y[i] = cscale(a, sc);
            if (i == l && i != k) y[k] = cscale(cmk(-a.x, a.y), sc);
        }
}

template <int D, unsigned UMASK, int l, bool CHECK>
__device__ __forceinline__ void so_columns(const cplx (&ta)[Tri<D>::n], const cplx (&tal)[Tri<D>::n], const cplx (&tbe)[Tri<D>::n],
                                           const cplx (&tga)[Tri<D>::n], unsigned mA, unsigned mAl, unsigned mBe, int m,
                                           bool live, cplx* __restrict__ dst) {
    typedef Pat<D, stored_from_tri(D, UMASK)> PT;
    if constexpr (l < D) {
      if constexpr (PT::has(l, l)) {
        cplx y[D], da[D], db[D], dab[D];
        const double inv0 = c_inv_j[m];
        tri_column<D, UMASK, l>(y, ta, inv0);
        y[l].x += 1.0;
        tri_column<D, UMASK, l>(da, tal, inv0);
        tri_column<D, UMASK, l>(db, tbe, inv0);
        tri_column<D, UMASK, l>(dab, tga, inv0);
        for (int j = m - 1; j >= 1; --j) {
            const double inv = c_inv_j[j];
            cplx acc[D], s[D];
            // dab' = (A dab + al (db + dab) + be (da + dab) + ga (y + da + db + dab)) / j
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, ta, mA, dab);
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = cadd(db[i], dab[i]);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, tal, mAl, s);
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = cadd(da[i], dab[i]);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, tbe, mBe, s);
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = cadd(cadd(y[i], da[i]), cadd(db[i], dab[i]));
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, tga, mBe, s);
#pragma unroll
            for (int i = 0; i < D; ++i) dab[i] = cscale(acc[i], inv);
            // da' = (A da + al (y + da)) / j
#pragma unroll
            for (int i = 0; i < D; ++i) { acc[i] = cmk(0.0, 0.0); s[i] = cadd(y[i], da[i]); }
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, ta, mA, da);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, tal, mAl, s);
#pragma unroll
            for (int i = 0; i < D; ++i) da[i] = cscale(acc[i], inv);
            // db' = (A db + be (y + db)) / j
#pragma unroll
            for (int i = 0; i < D; ++i) { acc[i] = cmk(0.0, 0.0); s[i] = cadd(y[i], db[i]); }
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, ta, mA, db);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, tbe, mBe, s);
#pragma unroll
            for (int i = 0; i < D; ++i) db[i] = cscale(acc[i], inv);
            // y' = I + A y / j
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
            tri_matvec_acc<D, UMASK, l, CHECK>(acc, ta, mA, y);
#pragma unroll
            for (int i = 0; i < D; ++i) y[i] = cscale(acc[i], inv);
            y[l].x += 1.0;
        }
        if (live) {
            store_col<D, stored_from_tri(D, UMASK), l>(dst, dab);
        }
      }
        so_columns<D, UMASK, l + 1, CHECK>(ta, tal, tbe, tga, mA, mAl, mBe, m, live, dst);
    }
}

// One thread per time step: mixed second differences d2U^{v,e} at (eps2, eps2) for every (variable, error source).
// Only instantiated for structural masks small enough to keep four triangles in registers.
template <int D, unsigned UMASK>
__global__ void __launch_bounds__(128, RG_SO_T_CTAS)
k_steps_so_t(const DevProblem P, const TriPlanDev tp, const double* __restrict__ X, int B, cplx* __restrict__ ws,
             int* __restrict__ status) {
    constexpr int NP = Tri<D>::n;
    typedef Pat<D, stored_from_tri(D, UMASK)> PT;
    extern __shared__ cplx smem[];
    const long long total = (long long)B * P.N;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = item < total;
    if (!live) item = total - 1;
    int b, k;
    if (total < (1ll << 31)) { b = (int)((unsigned)item / (unsigned)P.N); k = (int)((unsigned)item - (unsigned)b * (unsigned)P.N); }
    else { b = (int)(item / P.N); k = (int)(item % P.N); }
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
    const StagedPlan sp = stage_plan(P, tp, reinterpret_cast<unsigned char*>(smem));
    cplx* wsk = ws + ((size_t)b * P.N + k) * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    const int nt = P.nterms, nv = P.nvar, ne = P.e;

    cplx ta[NP], tal[NP], tbe[NP], tga[NP];
    int m = 0;
    for (int e = 0; e < ne; ++e)
        for (int v = 0; v < nv; ++v) {
            const int sp_ = P.var_space[v], ix = P.var_index[v];
            const double val = (sp_ == RG_S_MAIN) ? xk[ix] : xadd[ix];
            const double h2 = __dsub_rn(__dadd_rn(val, P.eps2), val);
            EvalCtx ec{xk, xadd, P.eps2, P.table, P.N, k};
            // triangles, one term at a time (see k_steps_t):  A : H0 value | alpha : H0 difference in v |
            // beta : error value at eps2 | gamma : error difference in v at eps2
            double colsum[D];
#pragma unroll
            for (int kk = 0; kk < D; ++kk) colsum[kk] = 0.0;
#pragma unroll
            for (int pos = 0; pos < NP; ++pos) { ta[pos] = cmk(0, 0); tal[pos] = cmk(0, 0); tbe[pos] = cmk(0, 0); tga[pos] = cmk(0, 0); }
            for (int t = 0; t < nt; ++t) {
                const DevTerm& tm = sp.terms[t];
                const bool isH0 = tm.owner == RG_OWNER_H0;
                if (!sp.used[t] || !(isH0 || tm.owner == e)) continue;
                cplx base, del;
                term_coef(tm, ec, sp_, ix, h2, base, del);
                const cplx sb = cmk(base.y * P.dt, -base.x * P.dt), sd2 = cmk(del.y * P.dt, -del.x * P.dt);
                const cplx* dv = sp.dense + t * NP;
#pragma unroll
                for (int pos = 0; pos < NP; ++pos) {
                    if (!((UMASK >> pos) & 1u)) continue;
                    const cplx vv = dv[pos];
                    if (vv.x == 0.0 && vv.y == 0.0) continue;          // warp-uniform
                    if (isH0) { cfma(ta[pos], sb, vv); cfma(tal[pos], sd2, vv); }
                    else { cfma(tbe[pos], sb, vv); cfma(tga[pos], sd2, vv); }
                }
                if (e == 0 && v == 0 && isH0) {
                    const double ax = fabs(sb.x), ay = fabs(sb.y);
                    const double w = fmax(ax, ay) + 0.41421356237309515 * fmin(ax, ay);
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) colsum[kk] = fma(w, sp.colw[t * D + kk], colsum[kk]);
                }
            }
            if (e == 0 && v == 0) {
                double nrm = 0.0;
#pragma unroll
                for (int kk = 0; kk < D; ++kk) nrm = fmax(nrm, colsum[kk]);
                m = taylor_degree(nrm * 1.001 + 2.0 * P.eps2 * P.dt);
                m = __reduce_max_sync(0xffffffffu, m);
                if (m == 99) { if ((threadIdx.x & 31) == 0) atomicOr(status, 2); m = 18; }
            }
            cplx* dst = wsk + (size_t)(1 + nv + ne + e * nv + v) * objS;
            const unsigned mA = tp.maskA, mAl = tp.maskVar[v], mBe = tp.maskErr[e];
            if ((mA & mAl & mBe & UMASK) == UMASK) so_columns<D, UMASK, 0, false>(ta, tal, tbe, tga, mA, mAl, mBe, m, live, dst);
            else so_columns<D, UMASK, 0, true>(ta, tal, tbe, tga, mA, mAl, mBe, m, live, dst);
        }
}


// ------------------------------------------------------------------ thread-per-chunk kernels (sparse patterns)
// For structural patterns with few non-zeros the whole forward state / co-state fits in one thread's registers,
// so the chunk aggregate and the backward gradient sweep need no shared memory, no cp.async staging and no
// cross-lane reduction: one thread = one (pulse, chunk); step matrices stream from HBM in the compact layout.
// Chunk aggregates, one thread per (pulse, chunk): Q <- U_k Q ; Wl_e <- U_k Wl_e + D_k^e Q_old
template <int D, u64 CM>
__global__ void __launch_bounds__(128, RG_AGG_T_CTAS)
k_chunk_agg_t(const DevProblem P, int B, int L, int nc, const cplx* __restrict__ ws, cplx* __restrict__ Qb, cplx* __restrict__ Wlb) {
    typedef Pat<D, CM> PT;
    typedef PMat<D, CM> M;
    constexpr int DD = D * D;
    const long long total = (long long)B * nc;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int nv = P.nvar, ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const cplx* wsb = ws + (size_t)b * P.N * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    M q; q.identity();
    if (ne == 0) {
        M u; u.load(wsb + (size_t)k0 * PT::nnz);
        for (int k = k0; k < k1; ++k) {
            M un;
            if (k + 1 < k1) un.load(wsb + (size_t)(k + 1) * PT::nnz);      // prefetch
            M qn; pmat_mul<D, CM, false, false>(qn, u, q);
            q = qn; u = un;
        }
    } else {
        // error aggregates are processed one error source at a time to bound registers (Q is recomputed)
        for (int e = 0; e < ne; ++e) {
            q.identity();
            M wl; wl.zero();
            for (int k = k0; k < k1; ++k) {
                const cplx* wsk = wsb + (size_t)k * PT::nnz;
                if (k + 1 < k1) { M::prefetch(wsk + PT::nnz); M::prefetch(wsk + PT::nnz + (size_t)(1 + nv + e) * objS); }
                M u, de; u.load(wsk); de.load(wsk + (size_t)(1 + nv + e) * objS);
                M wn; pmat_mul<D, CM, false, false>(wn, u, wl); pmat_mul<D, CM, false, true>(wn, de, q);
                M qn; pmat_mul<D, CM, false, false>(qn, u, q);
                wl = wn; q = qn;
            }
            pmat_to_dense<D, CM>(wl, Wlb + (((size_t)b * nc + ch) * ne + e) * DD, false);
        }
    }
    pmat_to_dense<D, CM>(q, Qb + ((size_t)b * nc + ch) * DD, true);
}

// Backward gradient sweep, one thread per (pulse, chunk), fidelity role only (ERR roles use k_grad):
//   out0[b*nx + p*k + v] = scale0 * Re tr(G_k dU_k^v C_{k-1})
template <int D, u64 CM>
__global__ void __launch_bounds__(128)
k_grad_t(const DevProblem P, int B, int L, int nc, const cplx* __restrict__ ws, const cplx* __restrict__ Cb,
         const cplx* __restrict__ Gb, double* __restrict__ out0, double scale0, double* __restrict__ addS) {
    typedef Pat<D, CM> PT;
    typedef PMat<D, CM> M;
    constexpr int DD = D * D;
    const long long total = (long long)B * nc;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int nv = P.nvar, ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const cplx* wsb = ws + (size_t)b * P.N * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    M c, g;
    pmat_from_dense<D, CM>(c, Cb + ((size_t)b * nc + ch) * DD);
    {   // k_scan stores the co-state by rows (row l contiguous): G(i,j) = Gb[i*D + j]
        const cplx* gp = Gb + ((size_t)b * nc + ch) * DD;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (PT::has(i, j)) g.v[PT::idx(i, j)] = gp[i * D + j];
    }
    M u; u.load(wsb + (size_t)(k1 - 1) * PT::nnz);
    for (int k = k1 - 1; k >= k0; --k) {
        const cplx* wsk = wsb + (size_t)k * PT::nnz;
        M un;
        if (k > k0) un.load(wsb + (size_t)(k - 1) * PT::nnz);       // prefetch the next step's U (a cache prefetch of dU measured slower)
        M cp; pmat_mul<D, CM, true, false>(cp, u, c);                         // C_{k-1} = U_k^dag C_k
        for (int v = 0; v < nv; ++v) {
            M du; du.load(wsk + (size_t)(1 + v) * objS);
            M t; pmat_mul<D, CM, false, false>(t, du, cp);                    // dU C_{k-1}
            const double s = pmat_retrace<D, CM>(g, t) * scale0;
            if (P.var_space[v] == RG_S_MAIN) out0[(size_t)b * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
            else addS[(((size_t)b * (1 + ne)) * P.a + P.var_index[v]) * P.N + k] = s;
        }
        M gn; pmat_mul<D, CM, false, false>(gn, g, u);                        // G_{k-1} = G_k U_k
        g = gn; c = cp; u = un;
    }
}


// Backward sweep of the sensitivity gradient, one thread per (pulse, chunk), error source e = blockIdx.y:
//   out1[(b*ne+e)*nx + p*k + v] = (2/DD1) Re{ [tr(G' dU W_{k-1}) + tr(H' dU C_{k-1})]/eps^2 + tr(G' d2U C_{k-1})/eps2^2 }
// (W and H' carry an un-normalised eps, see k_grad).  Hermitian problems only (rewind with the adjoint).
template <int D, u64 CM>
__global__ void __launch_bounds__(128)
k_grad_err_t(const DevProblem P, int B, int L, int nc, const cplx* __restrict__ ws, const cplx* __restrict__ Cb,
             const cplx* __restrict__ Wb, const cplx* __restrict__ G1b, const cplx* __restrict__ H1b,
             double* __restrict__ out1, double* __restrict__ addS) {
    typedef Pat<D, CM> PT;
    typedef PMat<D, CM> M;
    constexpr int DD = D * D;
    const long long total = (long long)B * nc;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int es = blockIdx.y;
    const int nv = P.nvar, ne = P.e;
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const cplx* wsb = ws + (size_t)b * P.N * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    const double f1 = 2.0 / DD1 * P.inv_eps * P.inv_eps, f2 = 2.0 / DD1 * P.inv_eps2sq;
    M c, w, g, h;
    pmat_from_dense<D, CM>(c, Cb + ((size_t)b * nc + ch) * DD);
    {
        const size_t off = (((size_t)b * ne + es) * nc + ch) * DD;
        pmat_from_dense<D, CM>(w, Wb + off);
        const cplx* gp = G1b + off;          // co-states are stored by rows: G(i,j) = gp[i*D + j]
        const cplx* hp = H1b + off;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (PT::has(i, j)) { g.v[PT::idx(i, j)] = gp[i * D + j]; h.v[PT::idx(i, j)] = hp[i * D + j]; }
    }
    for (int k = k1 - 1; k >= k0; --k) {
        const cplx* wsk = wsb + (size_t)k * PT::nnz;
        if (k > k0) {                          // next step's operands on their way while this step is computed
            const cplx* nx = wsk - PT::nnz;
            M::prefetch(nx); M::prefetch(nx + (size_t)(1 + nv + es) * objS);
            for (int v = 0; v < nv; ++v) { M::prefetch(nx + (size_t)(1 + v) * objS); M::prefetch(nx + (size_t)(1 + nv + ne + es * nv + v) * objS); }
        }
        {   // rewind: C_{k-1} = U^dag C_k ;  W_{k-1} = U^dag (W_k - D_k C_{k-1})
            M u, de; u.load(wsk); de.load(wsk + (size_t)(1 + nv + es) * objS);
            M cp; pmat_mul<D, CM, true, false>(cp, u, c);
            c = cp;
            M t; pmat_mul<D, CM, false, false>(t, de, c);
#pragma unroll
            for (int i = 0; i < PT::nnz; ++i) t.v[i] = csub(w.v[i], t.v[i]);
            pmat_mul<D, CM, true, false>(w, u, t);
        }
        for (int v = 0; v < nv; ++v) {
            double s1, s2;
            {
                M du; du.load(wsk + (size_t)(1 + v) * objS);
                M t; pmat_mul<D, CM, false, false>(t, du, c);
                s1 = pmat_retrace<D, CM>(h, t);
                pmat_mul<D, CM, false, false>(t, du, w);
                s1 += pmat_retrace<D, CM>(g, t);
            }
            {
                M d2; d2.load(wsk + (size_t)(1 + nv + ne + es * nv + v) * objS);
                M t; pmat_mul<D, CM, false, false>(t, d2, c);
                s2 = pmat_retrace<D, CM>(g, t);
            }
            const double s = f1 * s1 + f2 * s2;
            if (P.var_space[v] == RG_S_MAIN) out1[((size_t)b * ne + es) * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
            else addS[(((size_t)b * (1 + ne) + 1 + es) * P.a + P.var_index[v]) * P.N + k] = s;
        }
        {   // advance: H' <- H' U + G' D ;  G' <- G' U
            M u, de; u.load(wsk); de.load(wsk + (size_t)(1 + nv + es) * objS);
            M hn; pmat_mul<D, CM, false, false>(hn, h, u); pmat_mul<D, CM, false, true>(hn, g, de);
            M gn; pmat_mul<D, CM, false, false>(gn, g, u);
            h = hn; g = gn;
        }
    }
}

// Chunk aggregates from the stored step matrices: q <- U_k q ; wl_e <- U_k wl_e + D_k^e q_old.
template <int D, u64 CM, u64 CMS>
__global__ void __launch_bounds__(128)
k_chunk_agg(const DevProblem P, int B, int L, int nc, const cplx* __restrict__ ws, cplx* __restrict__ Qb, cplx* __restrict__ Wlb) {
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    const long long total = (long long)B * nc;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < total;
    if (!live) item = total - 1;
    const int b = (int)(item / nc), ch = (int)(item % nc);
    const int nv = P.nvar, ne = P.e;
    const int nload = 1 + ne;

    extern __shared__ cplx smem[];
    cplx* base = smem + (size_t)(warp * G + g) * kagg_group_stride(D, ne);
    cplx* buf0 = base;
    cplx* buf1 = base + nload * DD;
    cplx* wl = base + 2 * nload * DD + l * D;      // + e*DD, private columns
    typedef Pat<D, CMS> PT;            // stored (compact) pattern; CM is the closure used by the products
    if (!PT::full) {
        for (int s = 0; s < 2 * nload; ++s)
#pragma unroll
            for (int i = 0; i < D; ++i) base[s * DD + l * D + i] = cmk((s % nload == 0 && i == l && !PT::has(l, l)) ? 1.0 : 0.0, 0.0);
        __syncwarp(amask);
    }
    const cplx* wsb = ws + (size_t)b * P.N * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    int coff = 0; unsigned rows = 0;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (PT::has(i, j)) { if (j < l) ++coff; if (j == l) rows |= 1u << i; }
    auto issue = [&](int k, cplx* dstbuf) {
        const cplx* wsk = wsb + (size_t)k * PT::nnz;
        for (int s = 0; s < nload; ++s) {
            const int obj = (s == 0) ? 0 : (1 + nv + (s - 1));
            const cplx* src = wsk + (size_t)obj * objS + coff;
            cplx* dst = dstbuf + s * DD + l * D;
            int r = 0;
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (PT::full || ((rows >> i) & 1u)) { cp_async16(dst + i, src + r); ++r; }
        }
        cp_async_commit();
    };
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    cplx q[D];
#pragma unroll
    for (int i = 0; i < D; ++i) q[i] = cmk(i == l ? 1.0 : 0.0, 0.0);
    for (int e = 0; e < ne; ++e)
#pragma unroll
        for (int i = 0; i < D; ++i) wl[e * DD + i] = cmk(0.0, 0.0);
    issue(k0, buf0);
    for (int kk = 0; kk < L; ++kk) {
        const bool ghost = (k0 + kk >= k1);
        cplx* cur = (kk & 1) ? buf1 : buf0;
        cplx* nxt = (kk & 1) ? buf0 : buf1;
        if (kk + 1 < L) { issue(min(k0 + kk + 1, k1 - 1), nxt); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp(amask);
        if (!ghost) {
            for (int e = 0; e < ne; ++e) {
                cplx w[D], wn[D];
#pragma unroll
                for (int i = 0; i < D; ++i) w[i] = wl[e * DD + i];
                matvec<D, CM>(cur, w, wn);
                matvec_acc<D, CM>(cur + (1 + e) * DD, q, wn);
#pragma unroll
                for (int i = 0; i < D; ++i) wl[e * DD + i] = wn[i];
            }
            cplx qn[D];
            matvec<D, CM>(cur, q, qn);
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = qn[i];
        }
        __syncwarp(amask);
    }
    if (live) {
        cplx* dst = Qb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) dst[i] = q[i];
        for (int e = 0; e < ne; ++e) {
            cplx* dw = Wlb + (((size_t)b * nc + ch) * ne + e) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) dw[i] = wl[e * DD + i];
        }
    }
}
