// rg_smalld.cuh -- fused GRAPE path for small Hilbert spaces (d <= 16), templated on D.
//
// Mapping: a *group* of D consecutive lanes owns one (pulse, time-chunk) work item; lane l of
// the group carries column l of every forward quantity and row l of every backward co-state in
// registers, and reads the d x d step matrices from shared memory (broadcast within the group).
// 32/D groups share a warp; only __syncwarp is ever needed.
//
//   k_steps  (K1): per time step, assemble A = -i dt H and the perturbation matrices from the term
//                  list, evaluate U = exp(A) and the *differenced* exponentials
//                  exp(A+dA)-exp(A) (first order) and the mixed second difference by Horner
//                  recurrences on (value, difference) pairs -- no cancellation, so the reference's
//                  1/eps and 1/eps2^2 quotients (src/UnitaryCalculations.jl:52,60,70,80-83,92-95)
//                  are exact to rounding.  Stores the step matrices to the HBM workspace and the
//                  chunk aggregates (product Q_c and its error derivatives Wl_c).
//   k_scan   (K2): per pulse, sequential scan over chunk aggregates: prefix C, W=dC/derr at chunk
//                  ends; F, F_d2err (src/FidelityCalculations.jl:54,79-83) and the co-state seeds
//                  K, K'; suffix co-states at chunk ends.
//   k_grad   (K3): per (pulse, chunk), backward sweep: rewinds the forward state with U_k^dagger,
//                  advances the co-states, and contracts them with the stored differences into
//                  F_dx / F_d2err_dx entries (src/FidelityCalculations.jl:56-65,85-97 in co-state form).
#pragma once
#include "rg_common.cuh"

// 1/j for the Horner recurrences (statically initialised: every translation unit has its own copy)
__constant__ double c_inv_j[32] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
                                   1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
                                   1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

enum { VK_DX = 0, VK_ERR = 1, VK_ERR_DX = 2, VK_BASE = 3, VK_TGT = 4, VK_TGT_DX = 5 };
enum { OPN = 0, OPC = 1, OPT = 2 };

// group stride (in cplx units) padded to an odd number of 16-byte words -> the 32/D groups of a
// warp hit distinct bank quads when they read the same matrix element.
__host__ __device__ inline int rg_odd(int n) { return (n & 1) ? n : n + 1; }
__host__ __device__ inline int k1_group_stride(int D, int nterms, int ne) {
    return rg_odd((4 + ne) * D * D + 2 * nterms);
}
__host__ __device__ inline int k1b_group_stride(int D, int nterms) { return rg_odd(5 * D * D + 4 * nterms); }
__host__ __device__ inline int kagg_group_stride(int D, int ne) { return rg_odd((2 * (1 + ne) + ne) * D * D + 1); }
__host__ __device__ inline int kmat_group_stride(int D, int nload) { return rg_odd((nload + 2) * D * D + 1); }
#ifndef RG_SCAN_RING
#define RG_SCAN_RING 2          // cp.async ring depth; 2 + one warp per CTA fits 5 CTAs/SM (2 waves at B = 8192), measured 0.134 vs 0.157 ms
#endif
#ifndef RG_SCAN_WPC
#define RG_SCAN_WPC 1          // warps per CTA of k_scan
#endif
__host__ __device__ inline int k2_group_stride(int D) { return rg_odd((14 + 2 * RG_SCAN_RING) * D * D + D + 1); }
__host__ __device__ inline int k3_group_stride(int D, int nload) { return rg_odd(2 * nload * D * D + D + 1); }

template <int D> struct GroupInfo {
    static constexpr int G = 32 / D;
    static constexpr unsigned amask = (G * D == 32) ? 0xffffffffu : ((1u << (G * D)) - 1u);
};


// ------------------------------------------------------------------ descriptor staging
// The term list, entry list and column pointers are copied to shared memory once per CTA so the
// per-step prologue never waits on global memory for them.
struct StagedDesc { const DevTerm* terms; const DevEntry* ents; const int* colptr; };
__host__ __device__ inline size_t rg_align16(size_t n) { return (n + 15) & ~(size_t)15; }
__host__ __device__ inline size_t staged_desc_bytes(int nterms, int nent, int d) {
    return rg_align16((size_t)nterms * sizeof(DevTerm)) + rg_align16((size_t)nent * sizeof(DevEntry)) + rg_align16((size_t)(d + 1) * sizeof(int));
}
// must be called by every thread of the CTA (contains __syncthreads)
__device__ inline StagedDesc stage_desc(const DevProblem& P, unsigned char* sm) {
    StagedDesc sd;
    DevTerm* t = reinterpret_cast<DevTerm*>(sm);
    DevEntry* e = reinterpret_cast<DevEntry*>(sm + rg_align16((size_t)P.nterms * sizeof(DevTerm)));
    int* cp = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(e) + rg_align16((size_t)P.nent * sizeof(DevEntry)));
    const int nt4 = P.nterms * (int)(sizeof(DevTerm) / 4), ne4 = P.nent * (int)(sizeof(DevEntry) / 4);
    for (int i = threadIdx.x; i < nt4; i += blockDim.x) reinterpret_cast<int*>(t)[i] = reinterpret_cast<const int*>(P.terms)[i];
    for (int i = threadIdx.x; i < ne4; i += blockDim.x) reinterpret_cast<int*>(e)[i] = reinterpret_cast<const int*>(P.ents)[i];
    for (int i = threadIdx.x; i <= P.d; i += blockDim.x) cp[i] = P.colptr[i];
    __syncthreads();
    sd.terms = t; sd.ents = e; sd.colptr = cp;
    return sd;
}

// ------------------------------------------------------------------ coefficient variants
// Fill coef[t] for all terms (distributed over the group's lanes) for one matrix variant.
//   VK_BASE   : H0 terms, value                         (scaled by -i dt)
//   VK_DX     : H0 terms, difference in variable v      (scaled by -i dt)
//   VK_ERR    : error-source `es` terms at err=errv      (scaled by -i dt)
//   VK_ERR_DX : error-source `es` terms at err=errv, difference in variable v
template <int D>
__device__ inline void fill_coefs(const DevProblem& P, const DevTerm* terms, cplx* coef, int kind, int es, double errv,
                                  int pspace, int pindex, double h, const double* xk, const double* xadd,
                                  int k, int l) {
    EvalCtx ec{xk, xadd, errv, P.table, P.N, k};
    for (int t = l; t < P.nterms; t += D) {
        const DevTerm& tm = terms[t];
        cplx out = cmk(0.0, 0.0);
        const bool isH0 = (tm.owner == RG_OWNER_H0);
        const bool use = (kind == VK_BASE || kind == VK_DX) ? isH0 : (tm.owner == es);
        if (use) {
            cplx b, dl;
            term_coef(tm, ec, pspace, pindex, h, b, dl);
            const cplx c = (kind == VK_BASE || kind == VK_ERR) ? b : dl;
            out = cmk(c.y * P.dt, -c.x * P.dt);          // (-i dt) * c
        }
        coef[t] = out;
    }
}

// Assemble column l of a d x d matrix in shared memory from the column-sorted entry list.
template <int D>
__device__ inline void assemble_col(const DevEntry* __restrict__ ents, const int* __restrict__ colptr,
                                    const cplx* coef, cplx* M, int l, bool zero = false) {
#pragma unroll
    for (int i = 0; i < D; ++i) M[i + D * l] = cmk(0.0, 0.0);
    if (zero) return;
    for (int idx = colptr[l]; idx < colptr[l + 1]; ++idx) {
        const DevEntry en = ents[idx];
        cplx acc = M[en.row + D * l];
        cfma(acc, coef[en.term], cmk(en.vr, en.vi));
        M[en.row + D * l] = acc;
    }
}

// ------------------------------------------------------------------ Horner recurrences
// Taylor polynomial T_m(A) = I + A(I + A/2(I + ... (I + A/m))) applied column-wise.
// First-order pair: (Y, Dl) with Y -> exp(A) column, Dl -> [exp(A+dA) - exp(A)] column:
//     Y' = I + A Y / j ,   Dl' = (A Dl + dA (Y + Dl)) / j .
template <int D>
__device__ __forceinline__ void horner_fo(const cplx* mA, const cplx* mD, int l, int m,
                                          cplx (&y)[D], cplx (&dl)[D]) {
    double inv = c_inv_j[m];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        y[i] = cscale(mA[i + D * l], inv);
        dl[i] = cscale(mD[i + D * l], inv);
        if (i == l) y[i].x += 1.0;
    }
    for (int j = m - 1; j >= 1; --j) {
        asm volatile("" ::: "memory");   // keep the (loop-invariant) matrix loads in the loop: stream from smem
        inv = c_inv_j[j];
        cplx t[D], u[D];
#pragma unroll
        for (int i = 0; i < D; ++i) { t[i] = cmk(0.0, 0.0); u[i] = cmk(0.0, 0.0); }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const cplx yk = y[k], dk = dl[k], sk = cadd(yk, dk);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const cplx a = mA[i + D * k];
                const cplx dm = mD[i + D * k];
                cfma(t[i], a, yk);
                cfma(u[i], a, dk);
                cfma(u[i], dm, sk);
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

// Second-order quadruple (Y, Da, Db, Dab): Dab -> exp(A+a+b+g) - exp(A+a) - exp(A+b) + exp(A):
//     Dab' = (A Dab + a (Db + Dab) + b (Da + Dab) + g (Y + Da + Db + Dab)) / j .
template <int D>
__device__ __forceinline__ void horner_so(const cplx* mA, const cplx* mAl, const cplx* mBe, const cplx* mGa, int l, int m,
                                          cplx (&y)[D], cplx (&da)[D], cplx (&db)[D], cplx (&dab)[D], unsigned amask) {
    double inv = c_inv_j[m];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        y[i] = cscale(mA[i + D * l], inv);
        if (i == l) y[i].x += 1.0;
        da[i] = cscale(mAl[i + D * l], inv);
        db[i] = cscale(mBe[i + D * l], inv);
        dab[i] = cscale(mGa[i + D * l], inv);
    }
    for (int j = m - 1; j >= 1; --j) {
        asm volatile("" ::: "memory");
        inv = c_inv_j[j];
        cplx acc[D];
        // mixed second difference first (needs the old value of everything)
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const cplx s1 = cadd(db[k], dab[k]);
            const cplx s2 = cadd(da[k], dab[k]);
            const cplx s3 = cadd(cadd(y[k], da[k]), s1);
            const cplx dk = dab[k];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                cfma(acc[i], mA[i + D * k], dk);
                cfma(acc[i], mAl[i + D * k], s1);
                cfma(acc[i], mBe[i + D * k], s2);
                cfma(acc[i], mGa[i + D * k], s3);
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) dab[i] = cscale(acc[i], inv);
        __syncwarp(amask);   // real barrier: stops ptxas from merging the LDS of the four sections
        // da
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const cplx dk = da[k], sk = cadd(y[k], dk);
#pragma unroll
            for (int i = 0; i < D; ++i) { cfma(acc[i], mA[i + D * k], dk); cfma(acc[i], mAl[i + D * k], sk); }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) da[i] = cscale(acc[i], inv);
        __syncwarp(amask);
        // db
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const cplx dk = db[k], sk = cadd(y[k], dk);
#pragma unroll
            for (int i = 0; i < D; ++i) { cfma(acc[i], mA[i + D * k], dk); cfma(acc[i], mBe[i + D * k], sk); }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) db[i] = cscale(acc[i], inv);
        __syncwarp(amask);
        // y
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const cplx yk = y[k];
#pragma unroll
            for (int i = 0; i < D; ++i) cfma(acc[i], mA[i + D * k], yk);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            y[i] = cscale(acc[i], inv);
            if (i == l) y[i].x += 1.0;
        }
    }
}


// ------------------------------------------------------------------ skew-Hermitian fast path
// For Hermitian H (and Hermitian perturbations) A = -i dt H and dA are skew-Hermitian:
// M[k][i] = -conj(M[i][k]).  The thread-per-step kernels (rg_steps_t.cuh) keep only the upper triangles, in
// registers, so their Horner loops run without any shared-memory traffic (the dense group loop above is
// LSU-wavefront bound: every 16-byte operand costs 4 wavefronts and feeds only 4-8 DFMA).
template <int D> struct Tri {
    static constexpr int n = D * (D + 1) / 2;
    __host__ __device__ static constexpr int idx(int i, int k) { return k * (k + 1) / 2 + i; }   // i <= k
};
// Fused pass over a unitary step matrix U (shared memory): cp = U^dagger c and gn = g U.
// Each loaded element feeds two complex FMAs (the two separate passes were LSU-wavefront bound).
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void rewind_advance(const cplx* __restrict__ M, const cplx (&c)[D], const cplx (&g)[D],
                                               cplx (&cp)[D], cplx (&gn)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) { cp[j] = cmk(0.0, 0.0); gn[j] = cmk(0.0, 0.0); }
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (!Pat<D, CM>::has(i, j)) continue;
            const cplx u = M[i + D * j];
            cfma_conj(cp[j], u, c[i]);
            cfma(gn[j], g[i], u);
        }
}

// out = M * v   (M in shared memory, column-major; elements outside the pattern CM are skipped)
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void matvec(const cplx* __restrict__ M, const cplx (&v)[D], cplx (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const cplx vk = v[k];
#pragma unroll
        for (int i = 0; i < D; ++i) if (Pat<D, CM>::has(i, k)) cfma(out[i], M[i + D * k], vk);
    }
}
// out += M * v
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void matvec_acc(const cplx* __restrict__ M, const cplx (&v)[D], cplx (&out)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const cplx vk = v[k];
#pragma unroll
        for (int i = 0; i < D; ++i) if (Pat<D, CM>::has(i, k)) cfma(out[i], M[i + D * k], vk);
    }
}
// out = M^dagger * v
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void matvec_adj(const cplx* __restrict__ M, const cplx (&v)[D], cplx (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        cplx a = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) if (Pat<D, CM>::has(k, i)) cfma_conj(a, M[k + D * i], v[k]);
        out[i] = a;
    }
}
// out = g * M   (row vector times matrix)
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void vecmat(const cplx (&g)[D], const cplx* __restrict__ M, cplx (&out)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        cplx a = cmk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < D; ++i) if (Pat<D, CM>::has(i, j)) cfma(a, g[i], M[i + D * j]);
        out[j] = a;
    }
}
template <int D, u64 CM = full_cmask<D>()>
__device__ __forceinline__ void vecmat_acc(const cplx (&g)[D], const cplx* __restrict__ M, cplx (&out)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int i = 0; i < D; ++i) if (Pat<D, CM>::has(i, j)) cfma(out[j], g[i], M[i + D * j]);
    }
}
// Re(g . t)
template <int D>
__device__ __forceinline__ double redot(const cplx (&g)[D], const cplx (&t)[D]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) { s = fma(g[i].x, t[i].x, s); s = fma(-g[i].y, t[i].y, s); }
    return s;
}

// ======================================================================================
// K1: per-step propagators and first-order differences, chunk aggregates
// ======================================================================================
// Work item = (pulse, chunk of L steps).  Per step: A, then one Horner pass per first-order object
// (variables, then error sources).  The chunk product q and the error aggregates wl live in shared
// memory (private columns) so the Horner loops own the register file.
template <int D>
__global__ void __launch_bounds__(128)
k_steps(const DevProblem P, const double* __restrict__ X, int B, int L, int nc,
        cplx* __restrict__ ws, cplx* __restrict__ Qb, cplx* __restrict__ Wlb, int* __restrict__ status) {
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ cplx smem[];
    const StagedDesc sd = stage_desc(P, reinterpret_cast<unsigned char*>(smem));
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    const long long total = (long long)B * nc;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < total;
    if (!live) item = total - 1;
    const int b = (int)(item / nc), ch = (int)(item % nc);

    const int nt = P.nterms, ne = P.e, nv = P.nvar;
    cplx* base = smem + staged_desc_bytes(P.nterms, P.nent, D) / sizeof(cplx) + (size_t)(warp * G + g) * k1_group_stride(D, nt, ne);
    cplx* mA = base;
    cplx* mD = base + DD;
    cplx* mX = base + 2 * DD;
    cplx* qS = base + 3 * DD + l * D;          // private column of the chunk product
    cplx* wl = base + 4 * DD + l * D;          // + e*DD: private column of Wl_e
    cplx* coef = base + (4 + ne) * DD;         // 2 * nt

    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];

#pragma unroll
    for (int i = 0; i < D; ++i) qS[i] = cmk(i == l ? 1.0 : 0.0, 0.0);
    for (int e = 0; e < ne; ++e)
#pragma unroll
        for (int i = 0; i < D; ++i) wl[e * DD + i] = cmk(0.0, 0.0);

    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    const int nfo = nv + ne;
    // Uniform trip count across the warp (groups sync with __syncwarp): steps past the end of the
    // pulse are ghost steps with A = 0, i.e. U = I exactly and all differences 0; nothing is stored.
    double xnext[RG_MAX_MAIN];
    for (int i = 0; i < P.p; ++i) xnext[i] = xp[(size_t)k0 * P.p + i];
    for (int kk = 0; kk < L; ++kk) {
        const bool ghost = (k0 + kk >= k1);
        const int k = ghost ? (k1 - 1) : (k0 + kk);
        const bool st = live && !ghost;
        double xk[RG_MAX_MAIN];
        {   // controls of this step were fetched one iteration ago; fetch the next step's now
            const int kn = min(k0 + kk + 1, k1 - 1);
            for (int i = 0; i < P.p; ++i) { xk[i] = xnext[i]; xnext[i] = xp[(size_t)kn * P.p + i]; }
        }
        cplx* wsk = ws + ((size_t)b * P.N + k) * (size_t)P.wsm;          // object 0 of this step
        const size_t objS = (size_t)P.wsB * P.N * P.wsm;                 // stride between objects

        // ---- base matrix A = -i dt H0(x_k)
        fill_coefs<D>(P, sd.terms, coef, VK_BASE, 0, 0.0, RG_S_NONE, 0, 0.0, xk, xadd, k, l);
        __syncwarp(amask);
        if (P.hstack) hstack_col<D>(P, b, k, 0, -1, mA, l, ghost);
        else assemble_col<D>(sd.ents, sd.colptr, coef, mA, l, ghost);
        double nrm = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { const cplx a = mA[i + D * l]; nrm += sqrt(a.x * a.x + a.y * a.y); }
        // Taylor degree and number of squarings, uniform across the warp
        int m, sq;
        {
            nrm = nrm * 1.001 + 2.0 * P.eps2 * P.dt;
            unsigned long long nb = __double_as_longlong(nrm);            // positive doubles order like integers
            nb = __reduce_max_sync(amask, (unsigned)(nb >> 32));
            expm_plan(__longlong_as_double((long long)((nb + 1ull) << 32)), m, sq);
            if (m == 99) { if (lane == 0) atomicOr(status, 1); m = 12; }
        }
        const double sc = sq ? scalbn(1.0, -sq) : 1.0;
        if (sq) {
            __syncwarp(amask);
#pragma unroll
            for (int i = 0; i < D; ++i) mA[i + D * l] = cscale(mA[i + D * l], sc);
        }

        // ---- first-order objects: variables then error sources (at least one pass, for U itself)
        for (int o = 0; o < max(nfo, 1); ++o) {
            cplx* cf = coef + nt;
            if (nfo == 0) {
                for (int t = l; t < nt; t += D) cf[t] = cmk(0.0, 0.0);
            } else if (o < nv) {
                const int sp = P.var_space[o], ix = P.var_index[o];
                const double v = (sp == RG_S_MAIN) ? xk[ix] : xadd[ix];
                const double h = __dsub_rn(__dadd_rn(v, P.eps), v);     // the step actually taken
                fill_coefs<D>(P, sd.terms, cf, VK_DX, 0, 0.0, sp, ix, h, xk, xadd, k, l);
            } else {
                fill_coefs<D>(P, sd.terms, cf, VK_ERR, o - nv, P.eps, RG_S_NONE, 0, 0.0, xk, xadd, k, l);
            }
            __syncwarp(amask);
            if (P.hstack) {
                // host-evaluated closures: difference of the stacked Hamiltonians (variables at eps, then H0 + Herr_e(eps))
                if (nfo == 0) hstack_col<D>(P, b, k, 0, -1, mD, l, true);
                else hstack_col<D>(P, b, k, o < nv ? 1 + o : 1 + 2 * nv + (o - nv), 0, mD, l, ghost);
            } else {
                assemble_col<D>(sd.ents, sd.colptr, cf, mD, l, ghost);
            }
            if (sq) {
#pragma unroll
                for (int i = 0; i < D; ++i) mD[i + D * l] = cscale(mD[i + D * l], sc);
            }
            __syncwarp(amask);
            cplx y[D], dl[D];
            horner_fo<D>(mA, mD, l, m, y, dl);
            // squarings: Y <- Y Y ;  Dl <- Dl (Y + Dl) + Y Dl   (mX holds Y, mD is free after the Horner pass)
            for (int q2 = 0; q2 < sq; ++q2) {
                __syncwarp(amask);
#pragma unroll
                for (int i = 0; i < D; ++i) { mX[i + D * l] = y[i]; mD[i + D * l] = dl[i]; }
                __syncwarp(amask);
                cplx yn[D], dn[D], sy[D];
#pragma unroll
                for (int i = 0; i < D; ++i) sy[i] = cadd(y[i], dl[i]);
                matvec<D>(mX, y, yn);
                matvec<D>(mD, sy, dn);
                matvec_acc<D>(mX, dl, dn);
#pragma unroll
                for (int i = 0; i < D; ++i) { y[i] = yn[i]; dl[i] = dn[i]; }
            }
            if (sq) __syncwarp(amask);
            if (st && nfo > 0) {
                cplx* dst = wsk + (size_t)(1 + o) * objS + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = dl[i];
            }
            if (o == 0) {
                if (st) {
                    cplx* dst = wsk + l * D;
#pragma unroll
                    for (int i = 0; i < D; ++i) dst[i] = y[i];
                }
                // wl_e <- U wl_e   (q is advanced at the end of the step; D_k q_old is still needed)
#pragma unroll
                for (int i = 0; i < D; ++i) mX[i + D * l] = y[i];
                __syncwarp(amask);
                for (int e = 0; e < ne; ++e) {
                    cplx w[D], wn[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) w[i] = wl[e * DD + i];
                    matvec<D>(mX, w, wn);
#pragma unroll
                    for (int i = 0; i < D; ++i) wl[e * DD + i] = wn[i];
                }
                if (ne == 0) {
                    cplx q[D], qn[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) q[i] = qS[i];
                    matvec<D>(mX, q, qn);
#pragma unroll
                    for (int i = 0; i < D; ++i) qS[i] = qn[i];
                }
                __syncwarp(amask);
            }
            if (nfo > 0 && o >= nv) {
                // error source e: wl_e += D_k q_old
                const int e = o - nv;
#pragma unroll
                for (int i = 0; i < D; ++i) mX[i + D * l] = dl[i];
                __syncwarp(amask);
                cplx w[D], q[D];
#pragma unroll
                for (int i = 0; i < D; ++i) { w[i] = wl[e * DD + i]; q[i] = qS[i]; }
                matvec_acc<D>(mX, q, w);
#pragma unroll
                for (int i = 0; i < D; ++i) wl[e * DD + i] = w[i];
                __syncwarp(amask);
                if (o == nfo - 1) {
                    // last pass of the step: advance q with U (recomputed identically in this pass)
#pragma unroll
                    for (int i = 0; i < D; ++i) mX[i + D * l] = y[i];
                    __syncwarp(amask);
                    cplx qn[D];
                    matvec<D>(mX, q, qn);
#pragma unroll
                    for (int i = 0; i < D; ++i) qS[i] = qn[i];
                    __syncwarp(amask);
                }
            }
        }
        if (!P.hermitian) {
            // U_k^{-1} = exp(-A) (the reference uses inv(), src/UnitaryCalculations.jl:47): one more Horner pass with
            // -A and a zero perturbation; stored as the last object of the step for the backward sweeps.
            __syncwarp(amask);
#pragma unroll
            for (int i = 0; i < D; ++i) { mX[i + D * l] = cmk(-mA[i + D * l].x, -mA[i + D * l].y); mD[i + D * l] = cmk(0.0, 0.0); }
            __syncwarp(amask);
            cplx y[D], dl[D];
            horner_fo<D>(mX, mD, l, m, y, dl);
            for (int q2 = 0; q2 < sq; ++q2) {
                __syncwarp(amask);
#pragma unroll
                for (int i = 0; i < D; ++i) mD[i + D * l] = y[i];
                __syncwarp(amask);
                cplx yn[D];
                matvec<D>(mD, y, yn);
#pragma unroll
                for (int i = 0; i < D; ++i) y[i] = yn[i];
            }
            if (st) {
                cplx* dst = wsk + (size_t)(P.nstore - 1) * objS + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = y[i];
            }
            __syncwarp(amask);
        }
    }
    if (live) {
        cplx* dst = Qb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) dst[i] = qS[i];
        for (int e = 0; e < ne; ++e) {
            cplx* dw = Wlb + (((size_t)b * nc + ch) * ne + e) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) dw[i] = wl[e * DD + i];
        }
    }
}

// ======================================================================================
// K1b: mixed second differences (variable v, error source e) at eps2 -- one work item per time step
// ======================================================================================
template <int D>
__global__ void __launch_bounds__(128)
k_steps_so(const DevProblem P, const double* __restrict__ X, int B, cplx* __restrict__ ws, int* __restrict__ status) {
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ cplx smem[];
    const StagedDesc sd = stage_desc(P, reinterpret_cast<unsigned char*>(smem));
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    const long long total = (long long)B * P.N;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < total;
    if (!live) item = total - 1;
    const int b = (int)(item / P.N), k = (int)(item % P.N);

    const int nt = P.nterms, ne = P.e, nv = P.nvar;
    cplx* base = smem + staged_desc_bytes(P.nterms, P.nent, D) / sizeof(cplx) + (size_t)(warp * G + g) * k1b_group_stride(D, nt);
    cplx* mA = base;
    cplx* mAl = base + DD;
    cplx* mBe = base + 2 * DD;
    cplx* mGa = base + 3 * DD;
    cplx* mY = base + 4 * DD;       // squaring scratch
    cplx* coef = base + 5 * DD;     // 4 * nt

    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
    cplx* wsk = ws + ((size_t)b * P.N + k) * (size_t)P.wsm;
    const size_t objS = (size_t)P.wsB * P.N * P.wsm;

    fill_coefs<D>(P, sd.terms, coef, VK_BASE, 0, 0.0, RG_S_NONE, 0, 0.0, xk, xadd, k, l);
    __syncwarp(amask);
    if (P.hstack) hstack_col<D>(P, b, k, 0, -1, mA, l);
    else assemble_col<D>(sd.ents, sd.colptr, coef, mA, l);
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) { const cplx a = mA[i + D * l]; nrm += sqrt(a.x * a.x + a.y * a.y); }
    int m, sq;
    {
        nrm = nrm * 1.001 + 2.0 * P.eps2 * P.dt;
        unsigned long long nb = __double_as_longlong(nrm);
        nb = __reduce_max_sync(amask, (unsigned)(nb >> 32));
        expm_plan(__longlong_as_double((long long)((nb + 1ull) << 32)), m, sq);
        if (m == 99) { if (lane == 0) atomicOr(status, 1); m = 12; }
    }
    const double sc = sq ? scalbn(1.0, -sq) : 1.0;
    if (sq) {
        __syncwarp(amask);
#pragma unroll
        for (int i = 0; i < D; ++i) mA[i + D * l] = cscale(mA[i + D * l], sc);
    }

    for (int e = 0; e < ne; ++e) {
        fill_coefs<D>(P, sd.terms, coef + 2 * nt, VK_ERR, e, P.eps2, RG_S_NONE, 0, 0.0, xk, xadd, k, l);   // beta
        __syncwarp(amask);
        if (P.hstack) hstack_col<D>(P, b, k, 1 + 2 * nv + ne + e, 0, mBe, l);                 // beta = H0 + Herr_e(eps2) - H0
        else assemble_col<D>(sd.ents, sd.colptr, coef + 2 * nt, mBe, l);
        if (sq) {
#pragma unroll
            for (int i = 0; i < D; ++i) mBe[i + D * l] = cscale(mBe[i + D * l], sc);
        }
        for (int v = 0; v < nv; ++v) {
            const int sp = P.var_space[v], ix = P.var_index[v];
            const double val = (sp == RG_S_MAIN) ? xk[ix] : xadd[ix];
            const double h2 = __dsub_rn(__dadd_rn(val, P.eps2), val);
            fill_coefs<D>(P, sd.terms, coef + nt, VK_DX, 0, 0.0, sp, ix, h2, xk, xadd, k, l);              // alpha
            fill_coefs<D>(P, sd.terms, coef + 3 * nt, VK_ERR_DX, e, P.eps2, sp, ix, h2, xk, xadd, k, l);   // gamma
            __syncwarp(amask);
            if (P.hstack) {
                // alpha = H0(x + eps2 e_v) - H0 ;  gamma = [H0(x+eps2 e_v) + Herr_e(x+eps2 e_v, eps2)] - H0(x+eps2 e_v) - beta
                hstack_col<D>(P, b, k, 1 + nv + v, 0, mAl, l);
                hstack_col<D>(P, b, k, 1 + 2 * nv + 2 * ne + e * nv + v, 1 + nv + v, mGa, l);
                const double sc0 = sq ? 1.0 / sc : 1.0;              // mBe is already scaled
#pragma unroll
                for (int i = 0; i < D; ++i) mGa[i + D * l] = csub(mGa[i + D * l], cscale(mBe[i + D * l], sc0));
            } else {
                assemble_col<D>(sd.ents, sd.colptr, coef + nt, mAl, l);
                assemble_col<D>(sd.ents, sd.colptr, coef + 3 * nt, mGa, l);
            }
            if (sq) {
#pragma unroll
                for (int i = 0; i < D; ++i) { mAl[i + D * l] = cscale(mAl[i + D * l], sc); mGa[i + D * l] = cscale(mGa[i + D * l], sc); }
            }
            __syncwarp(amask);
            cplx y[D], da[D], db[D], dab[D];
            horner_so<D>(mA, mAl, mBe, mGa, l, m, y, da, db, dab, amask);
            if (sq) {
                // squarings of the (value, da, db, dab) quadruple; beta (mBe) is saved in registers and restored
                cplx be_save[D];
#pragma unroll
                for (int i = 0; i < D; ++i) be_save[i] = mBe[i + D * l];
                for (int q2 = 0; q2 < sq; ++q2) {
                    __syncwarp(amask);
#pragma unroll
                    for (int i = 0; i < D; ++i) { mY[i + D * l] = y[i]; mAl[i + D * l] = da[i]; mBe[i + D * l] = db[i]; mGa[i + D * l] = dab[i]; }
                    __syncwarp(amask);
                    cplx yn[D], an[D], bn[D], abn[D], t[D];
                    // dab' = dab (Y+da+db+dab) + (Y+da+db) dab + da db + db da
#pragma unroll
                    for (int i = 0; i < D; ++i) t[i] = cadd(cadd(y[i], da[i]), cadd(db[i], dab[i]));
                    matvec<D>(mGa, t, abn);
                    matvec_acc<D>(mY, dab, abn); matvec_acc<D>(mAl, dab, abn); matvec_acc<D>(mBe, dab, abn);
                    matvec_acc<D>(mAl, db, abn); matvec_acc<D>(mBe, da, abn);
                    // da' = da (Y+da) + Y da ; db' likewise ; Y' = Y Y
#pragma unroll
                    for (int i = 0; i < D; ++i) t[i] = cadd(y[i], da[i]);
                    matvec<D>(mAl, t, an); matvec_acc<D>(mY, da, an);
#pragma unroll
                    for (int i = 0; i < D; ++i) t[i] = cadd(y[i], db[i]);
                    matvec<D>(mBe, t, bn); matvec_acc<D>(mY, db, bn);
                    matvec<D>(mY, y, yn);
#pragma unroll
                    for (int i = 0; i < D; ++i) { y[i] = yn[i]; da[i] = an[i]; db[i] = bn[i]; dab[i] = abn[i]; }
                }
                __syncwarp(amask);
#pragma unroll
                for (int i = 0; i < D; ++i) mBe[i + D * l] = be_save[i];
            }
            if (live) {
                cplx* dst = wsk + (size_t)(1 + nv + ne + e * nv + v) * objS;
#pragma unroll
                for (int i = 0; i < D; ++i) if (pat_has(P.cmask, D, i, l)) dst[pat_idx(P.cmask, D, i, l)] = dab[i];
            }
            __syncwarp(amask);
        }
        __syncwarp(amask);
    }
}

// ======================================================================================
// K2: per-pulse scan over chunk aggregates, fidelity algebra, co-state seeds
// ======================================================================================
template <int D, int OP>
__device__ __forceinline__ cplx melem(const cplx* A, int i, int k) {
    if (OP == OPN) return A[i + D * k];
    const cplx a = A[k + D * i];
    return (OP == OPC) ? cconj(a) : a;
}
// column l of dst (+)= alpha * op(A) op(B); dst must not alias A or B.
template <int D, int OPA, int OPB>
__device__ __forceinline__ void gmm(cplx* dst, const cplx* A, const cplx* B, int l, unsigned amask,
                                    bool accumulate = false, double alpha = 1.0) {
    __syncwarp(amask);
    cplx bcol[D];                       // column l of op(B), loaded once
#pragma unroll
    for (int k = 0; k < D; ++k) bcol[k] = melem<D, OPB>(B, k, l);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        cplx acc = cmk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < D; ++k) cfma(acc, melem<D, OPA>(A, i, k), bcol[k]);
        acc = cscale(acc, alpha);
        dst[i + D * l] = accumulate ? cadd(dst[i + D * l], acc) : acc;
    }
    __syncwarp(amask);
}
// sum over the group's lanes; every lane receives the result (fixed order -> deterministic)
template <int D>
__device__ inline cplx greduce(cplx v, cplx* scratch, int l, unsigned amask) {
    __syncwarp(amask);
    scratch[l] = v;
    __syncwarp(amask);
    cplx s = cmk(0.0, 0.0);
    for (int i = 0; i < D; ++i) s = cadd(s, scratch[i]);
    __syncwarp(amask);
    return s;
}
// tr(op(A) op(B))
template <int D, int OPA, int OPB>
__device__ __forceinline__ cplx gtrace2(const cplx* A, const cplx* B, cplx* scratch, int l, unsigned amask) {
    __syncwarp(amask);
    cplx acc = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < D; ++k) cfma(acc, melem<D, OPA>(A, l, k), melem<D, OPB>(B, k, l));
    return greduce<D>(acc, scratch, l, amask);
}

// Fidelity algebra of one pulse, carried out by the D lanes of a group on matrices in shared memory (src/FidelityCalculations.jl
// :54-65 for role 0, :79-97 for role 1+e, :35-40,72-74,102-109 for the target-derivative parts of the x_add gradient).
//   in : base + 2*DD (mU) = U (role 0) or E_e = W_N/eps (role 1+e), column-major
//   out: base + 9*DD (mK) = co-state seed K (role 0) or K' (role 1+e); returns F resp. F_d2err[e]; addT_dst[j] receives the
//        target-derivative part of the x_add[j] gradient entry (written by lane 0 when `live`).
// base needs (14*DD + D + 1) complex numbers.
template <int D>
__device__ __forceinline__ double fid_algebra(const DevProblem& P, cplx* base, const double* xadd, int role, int l, unsigned amask,
                                              double* __restrict__ addT_dst, bool live) {
    constexpr int DD = D * D;
    cplx* mX = base; cplx* mW = base + DD; cplx* mU = base + 2 * DD; cplx* mU0 = base + 3 * DD; cplx* mM = base + 4 * DD;
    cplx* T1 = base + 5 * DD; cplx* T2 = base + 6 * DD; cplx* T3 = base + 7 * DD; cplx* T4 = base + 8 * DD; cplx* mK = base + 9 * DD;
    cplx* cPP = base + 10 * DD; cplx* cPPt = base + 11 * DD; cplx* cP = base + 12 * DD; cplx* mV = base + 13 * DD;
    cplx* scratch = base + 14 * DD;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        cPP[i + D * l] = cmk(P.PP[i + D * l], 0.0);
        cPPt[i + D * l] = cmk(P.PPt[i + D * l], 0.0);
        cP[i + D * l] = cmk(P.Pm[i + D * l], 0.0);
    }
    // target U0(x_add)
    {
        EvalCtx ec{nullptr, xadd, 0.0, P.table, P.N, 0};
        cplx* coef = mX;    // reuse as coefficient table (ntt <= D*D enforced on host)
        __syncwarp(amask);
        for (int t = l; t < P.ntt; t += D) {
            cplx bb, dl;
            term_coef(P.tterms[t], ec, RG_S_NONE, 0, 0.0, bb, dl);
            coef[t] = bb;
        }
        __syncwarp(amask);
        assemble_col<D>(P.tents, P.tcolptr, coef, mU0, l);
        if (P.tstack) {
#pragma unroll
            for (int i = 0; i < D; ++i) mU0[i + D * l] = P.tstack[i + D * l];
        }
    }
    const double Dt = P.Dtr, DD1 = Dt * (Dt + 1.0);
    gmm<D, OPC, OPN>(mM, mU0, mU, l, amask);              // M = U0^dag U   (or U0^dag E)
    gmm<D, OPN, OPN>(T1, cPP, mM, l, amask);              // T1 = PP M
    cplx tau = cmk(0.0, 0.0);
    {
        cplx d = T1[l + D * l];
        tau = greduce<D>(d, scratch, l, amask);
    }
    gmm<D, OPN, OPC>(T2, cP, mM, l, amask);               // T2 = P M^dag
    const cplx tr12 = gtrace2<D, OPN, OPN>(T1, T2, scratch, l, amask);
    double Fval = (tr12.x + tau.x * tau.x + tau.y * tau.y) / DD1;
    gmm<D, OPN, OPN>(T3, T2, cPP, l, amask);              // P M^dag PP
    gmm<D, OPT, OPC>(T4, cP, mM, l, amask);               // P^T M^dag
    gmm<D, OPN, OPN>(T3, T4, cPPt, l, amask, true);       // + P^T M^dag PP^T
#pragma unroll
    for (int i = 0; i < D; ++i) {
        cplx v = T3[i + D * l];
        const double pp = P.PP[i + D * l];
        v.x += 2.0 * tau.x * pp; v.y += -2.0 * tau.y * pp;   // + 2 conj(tau) PP
        T3[i + D * l] = v;
    }
    gmm<D, OPN, OPC>(mK, T3, mU0, l, amask);              // K = R U0^dag
    double scale_out = 1.0;
    if (role > 0) {
        // F_d2err = 2 [ Re tr(PP ME P ME^dag) - (1+D) Re tr(PP E^dag E) + |tau_e|^2 ] / (D(D+1))
        gmm<D, OPC, OPN>(T4, mU, mU, l, amask);           // E^dag E
        const cplx tee = gtrace2<D, OPN, OPN>(cPP, T4, scratch, l, amask);
        Fval = 2.0 * (tr12.x - (1.0 + Dt) * tee.x + tau.x * tau.x + tau.y * tau.y) / DD1;
        // K' = K - (1+D) (PP + PP^T) E^dag
#pragma unroll
        for (int i = 0; i < D; ++i) T4[i + D * l] = cmk(P.PP[i + D * l] + P.PPt[i + D * l], 0.0);
        gmm<D, OPN, OPC>(mK, T4, mU, l, amask, true, -(1.0 + Dt));
        scale_out = 2.0;
    }
    // ---- x_add: target-derivative parts (src/FidelityCalculations.jl:35-40,72-74,102,105,109)
    for (int j = 0; j < P.a; ++j) {
        EvalCtx ec{nullptr, xadd, 0.0, P.table, P.N, 0};
        const double h = __dsub_rn(__dadd_rn(xadd[j], P.eps), xadd[j]);
        cplx* coef = mX;
        __syncwarp(amask);
        for (int t = l; t < P.ntt; t += D) {
            cplx bb, dl;
            term_coef(P.tterms[t], ec, RG_S_ADD, j, h, bb, dl);
            coef[t] = cscale(dl, P.inv_eps);
        }
        __syncwarp(amask);
        assemble_col<D>(P.tents, P.tcolptr, coef, mV, l);
        if (P.tstack) {                                       // (U0(x_add + eps e_j) - U0(x_add)) / eps  (:35-40)
#pragma unroll
            for (int i = 0; i < D; ++i)
                mV[i + D * l] = cscale(csub(P.tstack[(size_t)(1 + j) * D * D + i + D * l], P.tstack[i + D * l]), P.inv_eps);
        }
        gmm<D, OPC, OPN>(mW, mV, mU, l, amask);           // S1 = V^dag U
        const cplx t3 = gtrace2<D, OPN, OPN>(cPP, mW, scratch, l, amask);      // tr(PP V^dag U)
        gmm<D, OPN, OPN>(T4, mW, T2, l, amask);           // S1 P M^dag
        const cplx t1 = gtrace2<D, OPN, OPN>(cPP, T4, scratch, l, amask);
        gmm<D, OPC, OPN>(mW, mU, mV, l, amask);           // U^dag V
        gmm<D, OPN, OPN>(T4, cP, mW, l, amask);           // P U^dag V
        const cplx t2 = gtrace2<D, OPN, OPN>(T1, T4, scratch, l, amask);        // tr(PP M P U^dag V)
        const double val = scale_out * (t1.x + t2.x + 2.0 * (tau.x * t3.x + tau.y * t3.y)) / DD1;
        if (live && l == 0) addT_dst[j] = val;
    }
    __syncwarp(amask);
    return Fval;
}

// grid.y = role: 0 -> fidelity F and its co-state; 1+e -> sensitivity F_d2err[e] and its co-states
template <int D>
__global__ void __launch_bounds__(64)
k_scan(const DevProblem P, const double* __restrict__ X, int B, int nc,
       const cplx* __restrict__ Qb, const cplx* __restrict__ Wlb,
       cplx* __restrict__ Cb, cplx* __restrict__ Wb, cplx* __restrict__ Gb, cplx* __restrict__ G1b, cplx* __restrict__ H1b,
       double* __restrict__ Fout,      // [B]
       double* __restrict__ F2out,     // [B][e]
       double* __restrict__ addT,      // [B][1+e][a]  target-derivative parts of the x_add gradient
       int materialize,                // 1: co-state seeds are the identity (G = B_k, H' = dB_k/derr) and
       cplx* __restrict__ Uout,        //    U = C_N        [B][d*d]            are written out
       cplx* __restrict__ Eout)        //    E_e = W_N/eps  [B][e][d*d]
{
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < B;
    if (!live) item = B - 1;
    const int b = (int)item;
    const int role = blockIdx.y;
    const int es = role - 1;
    const int ne = P.e;

    extern __shared__ cplx smem[];
    cplx* base = smem + (size_t)(warp * G + g) * k2_group_stride(D);
    constexpr int DD = D * D;
    cplx* mX = base;            // load buffer / generic
    cplx* mW = base + DD;       // second load buffer
    cplx* mU = base + 2 * DD;   // U (role 0) or E (role e)
    cplx* mU0 = base + 3 * DD;
    cplx* mM = base + 4 * DD;
    cplx* T1 = base + 5 * DD;
    cplx* T2 = base + 6 * DD;
    cplx* T3 = base + 7 * DD;
    cplx* T4 = base + 8 * DD;
    cplx* mK = base + 9 * DD;
    cplx* cPP = base + 10 * DD;
    cplx* cPPt = base + 11 * DD;
    cplx* cP = base + 12 * DD;
    cplx* mV = base + 13 * DD;
    cplx* scratch = base + 14 * DD;   // D + 1

    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];

    // ---- forward over chunks
    cplx c[D], w[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { c[i] = cmk(i == l ? 1.0 : 0.0, 0.0); w[i] = cmk(0.0, 0.0); }
    // chunk aggregates stream through a depth-RG_SCAN_RING cp.async ring so DRAM latency overlaps the matvecs
    cplx* ringQ = scratch + D + 1;
    cplx* ringW = ringQ + RG_SCAN_RING * DD;
    auto issue = [&](int ch, int slot) {
        const cplx* q = Qb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) cp_async16(ringQ + slot * DD + l * D + i, q + i);
        if (role > 0) {
            const cplx* ww = Wlb + (((size_t)b * nc + ch) * ne + es) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) cp_async16(ringW + slot * DD + l * D + i, ww + i);
        }
    };
    for (int s2 = 0; s2 < RG_SCAN_RING - 1; ++s2) { if (s2 < nc) issue(s2, s2); cp_async_commit(); }
    for (int ch = 0; ch < nc; ++ch) {
        const int nx2 = ch + RG_SCAN_RING - 1;
        if (nx2 < nc) issue(nx2, nx2 % RG_SCAN_RING);
        cp_async_commit();
        cp_async_wait<RG_SCAN_RING - 1>();
        __syncwarp(amask);
        const cplx* cX = ringQ + (ch % RG_SCAN_RING) * DD;
        const cplx* cW = ringW + (ch % RG_SCAN_RING) * DD;
        cplx cn[D];
        matvec<D>(cX, c, cn);
        if (role > 0) {
            cplx wn[D];
            matvec<D>(cX, w, wn);         // Q w
            matvec_acc<D>(cW, c, wn);     // + Wl c_old
#pragma unroll
            for (int i = 0; i < D; ++i) w[i] = wn[i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = cn[i];
        if (live) {
            if (role == 0) {
                cplx* dst = Cb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = c[i];
            } else {
                cplx* dst = Wb + (((size_t)b * ne + es) * nc + ch) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = w[i];
            }
        }
        __syncwarp(amask);
    }
    cp_async_wait<0>();

    if (materialize) {
        if (live) {
            cplx* dst = (role == 0) ? Uout + (size_t)b * DD + l * D : Eout + ((size_t)b * ne + es) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) dst[i] = (role == 0) ? c[i] : cscale(w[i], P.inv_eps);
        }
        // backward with identity seeds
        cplx gr[D], hr[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { gr[j] = cmk(j == l ? 1.0 : 0.0, 0.0); hr[j] = cmk(0.0, 0.0); }
        for (int ch = nc - 1; ch >= 0; --ch) {
            if (live) {
                if (role == 0) {
                    cplx* dst = Gb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
                    for (int i = 0; i < D; ++i) dst[i] = gr[i];
                } else {
                    cplx* dst = G1b + (((size_t)b * ne + es) * nc + ch) * DD + l * D;
                    cplx* dsh = H1b + (((size_t)b * ne + es) * nc + ch) * DD + l * D;
#pragma unroll
                    for (int i = 0; i < D; ++i) { dst[i] = gr[i]; dsh[i] = hr[i]; }
                }
            }
            const cplx* q = Qb + ((size_t)b * nc + ch) * DD + l * D;
            __syncwarp(amask);
#pragma unroll
            for (int i = 0; i < D; ++i) mX[i + D * l] = q[i];
            if (role > 0) {
                const cplx* ww = Wlb + (((size_t)b * nc + ch) * ne + es) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) mW[i + D * l] = ww[i];
            }
            __syncwarp(amask);
            cplx gn[D];
            vecmat<D>(gr, mX, gn);
            if (role > 0) {
                cplx hn[D];
                vecmat<D>(hr, mX, hn);
                vecmat_acc<D>(gr, mW, hn);
#pragma unroll
                for (int i = 0; i < D; ++i) hr[i] = hn[i];
            }
#pragma unroll
            for (int i = 0; i < D; ++i) gr[i] = gn[i];
        }
        return;
    }

    // ---- fidelity algebra
#pragma unroll
    for (int i = 0; i < D; ++i) mU[i + D * l] = (role == 0) ? c[i] : cscale(w[i], P.inv_eps);   // E = W_N / eps
    {
        const double Fval = fid_algebra<D>(P, base, xadd, role, l, amask, addT + ((size_t)b * (1 + ne) + role) * P.a, live);
        if (live && l == 0) {
            if (role == 0) Fout[b] = Fval; else F2out[(size_t)b * ne + es] = Fval;
        }
    }
    __syncwarp(amask);

    // ---- backward over chunks: co-states at chunk ends
    cplx gr[D], hr[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { gr[j] = mK[l + D * j]; hr[j] = cmk(0.0, 0.0); }
    __syncwarp(amask);
    for (int s2 = 0; s2 < RG_SCAN_RING - 1; ++s2) { if (nc - 1 - s2 >= 0) issue(nc - 1 - s2, s2); cp_async_commit(); }
    for (int it = 0; it < nc; ++it) {
        const int ch = nc - 1 - it;
        if (live) {
            if (role == 0) {
                cplx* dst = Gb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = gr[i];
            } else {
                cplx* dst = G1b + (((size_t)b * ne + es) * nc + ch) * DD + l * D;
                cplx* dsh = H1b + (((size_t)b * ne + es) * nc + ch) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) { dst[i] = gr[i]; dsh[i] = hr[i]; }
            }
        }
        const int nx2 = it + RG_SCAN_RING - 1;
        if (nx2 < nc) issue(nc - 1 - nx2, nx2 % RG_SCAN_RING);
        cp_async_commit();
        cp_async_wait<RG_SCAN_RING - 1>();
        __syncwarp(amask);
        const cplx* cX = ringQ + (it % RG_SCAN_RING) * DD;
        const cplx* cW = ringW + (it % RG_SCAN_RING) * DD;
        cplx gn[D];
        vecmat<D>(gr, cX, gn);
        if (role > 0) {
            cplx hn[D];
            vecmat<D>(hr, cX, hn);        // h Q
            vecmat_acc<D>(gr, cW, hn);    // + g_old Wl
#pragma unroll
            for (int i = 0; i < D; ++i) hr[i] = hn[i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) gr[i] = gn[i];
        __syncwarp(amask);
    }
    cp_async_wait<0>();
}

// ======================================================================================
// K3: backward gradient sweep per (pulse, chunk)
// ======================================================================================

// sum of a real over the D lanes of a group (lanes base..base+D-1), result valid in lane l==0
template <int D>
__device__ __forceinline__ double group_sum0(double v, int lane, int l, unsigned amask) {
    double s = v;
#pragma unroll
    for (int i = 1; i < D; ++i) {
        const double o = __shfl_sync(amask, v, (lane - l + i) & 31);
        s += o;
    }
    return s;
}

// ERR=false (role 0): out0[b*nx + p*k + v] = scale0 * Re tr(G_k dU_k^v C_{k-1})          (F_dx or -F_dx)
// ERR=true  (role 1+e, e = blockIdx.y):
//            out1[(b*ne+e)*nx + ...] = (2/DD1) * Re{[g' dU w + h' dU c]/eps^2 + g' d2U c/eps2^2}
//            (w and h' are built from raw, un-normalised differences, hence 1/eps^2)
// additional-parameter variables go to addS[((b*(1+ne)+role)*a + j)*N + k] for a later sum over k.
template <int D, bool ERR, u64 CM, u64 CMS>
__global__ void __launch_bounds__(128)
k_grad(const DevProblem P, int B, int L, int nc, const cplx* __restrict__ ws,
       const cplx* __restrict__ Cb, const cplx* __restrict__ Wb, const cplx* __restrict__ Gb,
       const cplx* __restrict__ G1b, const cplx* __restrict__ H1b,
       double* __restrict__ out0, double scale0, double* __restrict__ out1, double* __restrict__ addS) {
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
    const int es = ERR ? (int)blockIdx.y : -1;
    const int role = es + 1;
    const int nv = P.nvar, ne = P.e;
    const int ninv = P.hermitian ? 0 : 1;                 // extra slot: U_k^{-1} for non-Hermitian H
    const int nload = (ERR ? (2 + 2 * nv) : (1 + nv)) + ninv;

    extern __shared__ cplx smem[];
    cplx* base = smem + (size_t)(warp * G + g) * k3_group_stride(D, nload);
    cplx* buf0 = base;
    cplx* buf1 = base + nload * DD;
    typedef Pat<D, CMS> PT;            // stored (compact) pattern; CM is the closure used by the products
    // elements outside the stored pattern are never loaded: set them once (0, or 1 on the diagonal of inert levels
    // in the propagator slots) -- they keep that value
    if (!PT::full) {
        for (int s = 0; s < 2 * nload; ++s) {
            const int sl = s % nload;
            const bool uslot = (sl == 0) || (ninv && sl == nload - 1);
#pragma unroll
            for (int i = 0; i < D; ++i) base[s * DD + l * D + i] = cmk((uslot && i == l && !PT::has(l, l)) ? 1.0 : 0.0, 0.0);
        }
        __syncwarp(amask);
    }

    // stored objects this role needs: slot 0 = U, 1..nv = dU^v, [nv+1 = D_e, nv+2.. = d2U^{v,e}]
    const cplx* wsb = ws + (size_t)b * P.N * (size_t)PT::nnz;
    const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
    // compact offset of the first stored element of column l, and which rows are stored
    int coff = 0; unsigned rows = 0;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (PT::has(i, j)) { if (j < l) ++coff; if (j == l) rows |= 1u << i; }
    auto issue = [&](int k, cplx* dstbuf) {
        const cplx* wsk = wsb + (size_t)k * PT::nnz;
        for (int s = 0; s < nload; ++s) {
            int obj = s;
            if (ninv && s == nload - 1) obj = P.nstore - 1;
            else if (s == nv + 1) obj = 1 + nv + es;
            else if (s > nv + 1) obj = 1 + nv + ne + es * nv + (s - nv - 2);
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
    cplx c[D], gr[D];
    cplx w[ERR ? D : 1], hr[ERR ? D : 1];
    {
        const cplx* src = Cb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = src[i];
        if (!ERR) {
            const cplx* sg = Gb + ((size_t)b * nc + ch) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) gr[i] = sg[i];
        } else {
            const size_t off = (((size_t)b * ne + es) * nc + ch) * DD + l * D;
#pragma unroll
            for (int i = 0; i < (ERR ? D : 1); ++i) { w[i] = Wb[off + i]; gr[i] = G1b[off + i]; hr[i] = H1b[off + i]; }
        }
    }
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    const double f1 = 2.0 / DD1 * P.inv_eps * P.inv_eps, f2 = 2.0 / DD1 * P.inv_eps2sq;

    // Uniform trip count (see k_steps): iterations with k >= k1 are ghosts that load step k1-1 and
    // commit nothing.
    issue(min(k0 + L - 1, k1 - 1), buf0);
    for (int kk = L - 1; kk >= 0; --kk) {
        const bool ghost = (k0 + kk >= k1);
        const int k = min(k0 + kk, k1 - 1);
        cplx* cur = ((L - 1 - kk) & 1) ? buf1 : buf0;
        cplx* nxt = ((L - 1 - kk) & 1) ? buf0 : buf1;
        if (kk > 0) { issue(min(k0 + kk - 1, k1 - 1), nxt); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp(amask);
        const cplx* mZ = cur;
        const cplx* mDe = cur + (nv + 1) * DD;
        cplx cp[D], gn[D];
        const cplx* mInv = cur + (nload - 1) * DD;
        if (P.hermitian) {
            rewind_advance<D, CM>(mZ, c, gr, cp, gn);       // c_{k-1} = U_k^dag c_k  and  g_{k-1} = g_k U_k in one pass
        } else {
            matvec<D, CM>(mInv, c, cp);                     // c_{k-1} = U_k^{-1} c_k
            vecmat<D, CM>(gr, mZ, gn);
        }
        cplx wp[ERR ? D : 1];
        if (ERR) {
            cplx t[D];
            matvec<D, CM>(mDe, cp, t);                      // D_k c_{k-1}
#pragma unroll
            for (int i = 0; i < D; ++i) t[i] = csub(w[ERR ? i : 0], t[i]);
            cplx t2[D];
            if (P.hermitian) matvec_adj<D, CM>(mZ, t, t2);  // w_{k-1} = U_k^{-1} (w_k - D_k c_{k-1})
            else matvec<D, CM>(mInv, t, t2);
#pragma unroll
            for (int i = 0; i < (ERR ? D : 1); ++i) wp[i] = t2[i];
        }
        for (int v = 0; v < nv; ++v) {
            const cplx* mDv = cur + (1 + v) * DD;
            double s;
            {
                cplx t[D];
                matvec<D, CM>(mDv, cp, t);                  // dU^v c_{k-1}
                if (!ERR) {
                    s = redot<D>(gr, t) * scale0;
                } else {
                    cplx hh[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) hh[i] = hr[ERR ? i : 0];
                    s = redot<D>(hh, t);
                }
            }
            if (ERR) {
                cplx ww[D], t[D];
#pragma unroll
                for (int i = 0; i < D; ++i) ww[i] = wp[ERR ? i : 0];
                matvec<D, CM>(mDv, ww, t);                  // dU^v w_{k-1}
                s += redot<D>(gr, t);
                const cplx* mD2 = cur + (nv + 2 + v) * DD;
                matvec<D, CM>(mD2, cp, t);                  // d2U^{v,e} c_{k-1}
                s = f1 * s + f2 * redot<D>(gr, t);
            }
            s = group_sum0<D>(s, lane, l, amask);
            if (live && !ghost && l == 0) {
                if (P.var_space[v] == RG_S_MAIN) {
                    const size_t idx = (size_t)P.p * k + P.var_index[v];
                    if (!ERR) out0[(size_t)b * P.nx + idx] = s;
                    else out1[((size_t)b * ne + es) * P.nx + idx] = s;
                } else {
                    addS[(((size_t)b * (1 + ne) + role) * P.a + P.var_index[v]) * P.N + k] = s;
                }
            }
        }
        if (!ghost) {
            // advance co-states: h' = h' U + g' D ; g' = g' U ; and commit the rewound forward state
            if (ERR) {
                cplx hh[D], hn[D];
#pragma unroll
                for (int i = 0; i < D; ++i) hh[i] = hr[ERR ? i : 0];
                vecmat<D, CM>(hh, mZ, hn);
                vecmat_acc<D, CM>(gr, mDe, hn);
#pragma unroll
                for (int i = 0; i < (ERR ? D : 1); ++i) { hr[i] = hn[i]; w[i] = wp[i]; }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) { gr[i] = gn[i]; c[i] = cp[i]; }
        }
        __syncwarp(amask);
    }
}


// ======================================================================================
// Materialising variant of the backward sweep: emits the matrices of calculate_unitary_and_derivatives
// (src/UnitaryCalculations.jl:114-152) instead of contracting them.  Co-states come from k_scan run with
// identity seeds, so gr = row l of B_k = U_N ... U_{k+1} and hr = row l of dB_k/derr (un-normalised).
//   role 0     : U_dx[:,:,v,k]      = B_k dU^v C_{k-1} / eps
//   role 1 + e : U_derr_dx[:,:,v,k,e] = [B_k dU^v W_{k-1} + V_k dU^v C_{k-1}] / eps^2 + B_k d2U^{v,e} C_{k-1} / eps2^2
// Additional-parameter variables go to addM[((role*a + j)*N + k)*d*d ...] for a later sum over k.
template <int D, u64 CM>
__global__ void __launch_bounds__(128)
k_materialize(const DevProblem P, int L, int nc, const cplx* __restrict__ ws,
              const cplx* __restrict__ Cb, const cplx* __restrict__ Wb, const cplx* __restrict__ Gb,
              const cplx* __restrict__ G1b, const cplx* __restrict__ H1b,
              cplx* __restrict__ U_dx, cplx* __restrict__ U_derr_dx, cplx* __restrict__ addM) {
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    constexpr int DD = D * D;
    typedef Pat<D, CM> PT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < nc;
    if (!live) item = nc - 1;
    const int ch = (int)item;
    const int role = blockIdx.y, es = role - 1;
    const int nv = P.nvar, ne = P.e;
    const int ninv = P.hermitian ? 0 : 1;
    const int nload = ((role == 0) ? (1 + nv) : (2 + 2 * nv)) + ninv;
    const int nload_max = ((ne > 0) ? (2 + 2 * nv) : (1 + nv)) + ninv;

    extern __shared__ cplx smem[];
    cplx* base = smem + (size_t)(warp * G + g) * kmat_group_stride(D, nload_max);
    cplx* mats = base;                         // nload matrices, dense d x d (zeros outside the pattern)
    cplx* mB = base + nload_max * DD;
    cplx* mV = mB + DD;
    for (int s2 = 0; s2 < nload_max; ++s2)
#pragma unroll
        for (int i = 0; i < D; ++i) mats[s2 * DD + l * D + i] = cmk(0.0, 0.0);
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    cplx c[D], w[D], gr[D], hr[D];
    {
        const cplx* src = Cb + (size_t)ch * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) { c[i] = src[i]; w[i] = cmk(0, 0); hr[i] = cmk(0, 0); }
        if (role == 0) {
            const cplx* sg = Gb + (size_t)ch * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) gr[i] = sg[i];
        } else {
            const size_t off = ((size_t)es * nc + ch) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) { w[i] = Wb[off + i]; gr[i] = G1b[off + i]; hr[i] = H1b[off + i]; }
        }
    }
    for (int kk = L - 1; kk >= 0; --kk) {
        const bool ghost = (k0 + kk >= k1);
        const int k = min(k0 + kk, k1 - 1);
        __syncwarp(amask);
        const cplx* wsk = ws + (size_t)k * PT::nnz;
        const size_t objS = (size_t)P.wsB * P.N * PT::nnz;
        for (int s2 = 0; s2 < nload; ++s2) {
            int obj = s2;
            if (ninv && s2 == nload - 1) obj = P.nstore - 1;
            else if (s2 == nv + 1) obj = 1 + nv + es;
            else if (s2 > nv + 1) obj = 1 + nv + ne + es * nv + (s2 - nv - 2);
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (pat_has(CM, D, i, l)) mats[s2 * DD + l * D + i] = wsk[(size_t)obj * objS + pat_idx(CM, D, i, l)];
        }
        // rows of B_k and V_k for the left multiplications
#pragma unroll
        for (int j = 0; j < D; ++j) { mB[l + D * j] = gr[j]; mV[l + D * j] = hr[j]; }
        __syncwarp(amask);
        const cplx* mZ = mats;
        const cplx* mDe = mats + (nv + 1) * DD;
        cplx cp[D], gn[D], wp[D];
        const cplx* mInv = mats + (nload - 1) * DD;
        if (P.hermitian) rewind_advance<D>(mZ, c, gr, cp, gn);
        else { matvec<D>(mInv, c, cp); vecmat<D>(gr, mZ, gn); }
        if (role > 0) {
            cplx t[D];
            matvec<D>(mDe, cp, t);
#pragma unroll
            for (int i = 0; i < D; ++i) t[i] = csub(w[i], t[i]);
            if (P.hermitian) matvec_adj<D>(mZ, t, wp);
            else matvec<D>(mInv, t, wp);
        }
        for (int v = 0; v < nv; ++v) {
            const cplx* mDv = mats + (1 + v) * DD;
            cplx t1[D], out[D];
            matvec<D>(mDv, cp, t1);
            if (role == 0) {
                matvec<D>(mB, t1, out);
#pragma unroll
                for (int i = 0; i < D; ++i) out[i] = cscale(out[i], P.inv_eps);
            } else {
                cplx t2[D], a1[D], a2[D];
                matvec<D>(mDv, wp, t2);
                matvec<D>(mB, t2, a1);
                matvec_acc<D>(mV, t1, a1);
                matvec<D>(mats + (nv + 2 + v) * DD, cp, t2);
                matvec<D>(mB, t2, a2);
#pragma unroll
                for (int i = 0; i < D; ++i) out[i] = cadd(cscale(a1[i], P.inv_eps * P.inv_eps), cscale(a2[i], P.inv_eps2sq));
            }
            if (live && !ghost) {
                cplx* dst;
                if (P.var_space[v] == RG_S_MAIN) {
                    const size_t m = (size_t)P.var_index[v] + (size_t)P.p * k;
                    dst = (role == 0) ? U_dx + m * DD : U_derr_dx + (m + (size_t)P.p * P.N * es) * DD;
                } else {
                    dst = addM + (((size_t)role * P.a + P.var_index[v]) * P.N + k) * DD;
                }
#pragma unroll
                for (int i = 0; i < D; ++i) dst[l * D + i] = out[i];
            }
        }
        if (!ghost) {
            if (role > 0) {
                cplx hn[D];
                vecmat<D>(hr, mZ, hn);
                vecmat_acc<D>(gr, mDe, hn);
#pragma unroll
                for (int i = 0; i < D; ++i) { hr[i] = hn[i]; w[i] = wp[i]; }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) { gr[i] = gn[i]; c[i] = cp[i]; }
        }
    }
}
// U_dx_add[:,:,j] / U_derr_dx_add[:,:,j,e] = sum over time of the per-step matrices (src/UnitaryCalculations.jl:119-121,140-151);
// exact zero for additional parameters the Hamiltonian does not depend on.
static __global__ void k_reduce_add(const DevProblem P, const cplx* __restrict__ addM, cplx* __restrict__ U_dx_add, cplx* __restrict__ U_derr_dx_add) {
    const int DD = P.d * P.d;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrole = 1 + P.e;
    if (idx >= nrole * P.a * DD) return;
    const int el = idx % DD, j = (idx / DD) % P.a, role = idx / (DD * P.a);
    cplx s = cmk(0, 0);
    if (P.add_var[j] >= 0)
        for (int k = 0; k < P.N; ++k) s = cadd(s, addM[(((size_t)role * P.a + j) * P.N + k) * DD + el]);
    if (role == 0) U_dx_add[(size_t)j * DD + el] = s;
    else U_derr_dx_add[((size_t)j + (size_t)P.a * (role - 1)) * DD + el] = s;
}

// ======================================================================================
// K4: final assembly
// ======================================================================================
// x_add gradient entries: target part (from K2) + sum over time of the H-dependence part (from K3).
// role 0 part is scaled by scale0 (already applied to addS role 0 entries in K3; applied here to addT).
static __global__ void k_add_params(const DevProblem P, int B, const double* __restrict__ addT, const double* __restrict__ addS,
                             double* __restrict__ out0, double scale0T, double* __restrict__ out1) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrole = 1 + P.e;
    if (idx >= B * nrole * P.a) return;
    const int j = idx % P.a, role = (idx / P.a) % nrole, b = idx / (P.a * nrole);
    double s = 0.0;
    if (P.add_var[j] >= 0) {
        const double* src = addS + (((size_t)b * nrole + role) * P.a + j) * P.N;
        for (int k = 0; k < P.N; ++k) s += src[k];
    }
    const double t = addT[((size_t)b * nrole + role) * P.a + j];
    const size_t o = (size_t)P.p * P.N + j;
    if (role == 0) out0[(size_t)b * P.nx + o] = s + scale0T * t;
    else out1[((size_t)b * P.e + (role - 1)) * P.nx + o] = s + t;
}

// cost = 1 - F + sum_e c_e F2_e^2 ; grad = -F_dx + 2 sum_e c_e F2_e F2dx_e   (src/FidelityCalculations.jl:178-184)
// grad holds F_dx on entry when ne > 0 (written by K3/K4 with scale +1).
static __global__ void k_cost_grad(int B, int nx, int ne, const double* __restrict__ F, const double* __restrict__ F2,
                            const double* __restrict__ F2dx, const double* __restrict__ coeff /* device, ne */,
                            double* __restrict__ cost, double* __restrict__ grad) {
    const size_t n = (size_t)B * nx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / nx);
        const size_t r = i - (size_t)b * nx;
        double gacc = -grad[i];
        for (int e = 0; e < ne; ++e)
            gacc += 2.0 * coeff[e] * F2[(size_t)b * ne + e] * F2dx[((size_t)b * ne + e) * nx + r];
        grad[i] = gacc;
        if (r == 0) {
            double c = 1.0 - F[b];
            for (int e = 0; e < ne; ++e) { const double f2 = F2[(size_t)b * ne + e]; c += coeff[e] * f2 * f2; }
            cost[b] = c;
        }
    }
}
static __global__ void k_cost_only(int B, const double* __restrict__ F, double* __restrict__ cost) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) cost[b] = 1.0 - F[b];
}
