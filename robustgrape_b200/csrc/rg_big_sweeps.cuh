// rg_big_sweeps.cuh -- time-parallel scan and gradient sweeps of the dense path (11 <= d <= 64), written as sequences of
// CTA-level DMMA products (rg_big.cuh).  Same algorithm as rg_smalld.cuh / rg_steps_t.cuh: chunk aggregates, per-pulse scan
// over chunks with the fidelity algebra in the middle, backward co-state sweeps (SURVEY section 8 rows a-3 ... a-8;
// src/UnitaryCalculations.jl:44-152 and src/FidelityCalculations.jl:47-118 in co-state form), with two changes that suit
// d x d products: the trace Re tr(G dU C) is taken as Re tr((C G) dU), one product per step instead of one per variable,
// and all states are planar matrices in global memory (both orientations) so every operand is a row-contiguous read.
#pragma once
#include "rg_big.cuh"

// chunk-boundary states: matrix m of pulse b at (role / error e, chunk c); 4 planes each
template <int DP> __device__ __forceinline__ BMat big_state(double* base, size_t idx) {
    return bmat_at(base + idx * 4 * DP * DP, (size_t)DP * DP, true);
}
struct BigBufs {
    double* ws;       // step workspace (k_big_steps)
    double* Qb;       // [B][nc]
    double* Wlb;      // [B][nc][ne]
    double* Cb;       // [B][nc]      C at the end of chunk c
    double* Wb;       // [B][ne][nc]
    double* Gb;       // [B][nc]      co-state at the end of chunk c
    double* G1b;      // [B][ne][nc]
    double* H1b;      // [B][ne][nc]
    double* scratch;  // per CTA: RG_BIG_SWEEP_MATS matrices of 4 planes
};
#define RG_BIG_SWEEP_MATS 16

// ---- chunk aggregates: Q <- U_k Q ; Wl_e <- U_k Wl_e + D_k^e Q_old ----------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT)
k_big_agg(const DevProblem P, const BigBufs bb, int B, int L, int nc) {
    constexpr int NT = BigGemm<DP>::NT;
    extern __shared__ double smd[];
    const int nv = P.nvar, ne = P.e, d = P.d;
    const size_t wstep = big_ws_step_doubles(DP, nv, ne);
    double* scr = bb.scratch + (size_t)blockIdx.x * RG_BIG_SWEEP_MATS * 4 * DP * DP;
    BigGemm<DP> G;
    for (long long item = blockIdx.x; item < (long long)B * nc; item += gridDim.x) {
        const int b = (int)(item / nc), ch = (int)(item % nc);
        const int k0 = ch * L, k1 = min(P.N, k0 + L);
        __syncthreads();
        // ping-pong: slot 0/1 = Q, 2+2e / 3+2e = Wl_e
        bmat_set_identity<DP>(big_state<DP>(scr, 0), d, NT);
        for (int e = 0; e < ne; ++e) bmat_zero<DP>(big_state<DP>(scr, 2 + 2 * e), NT);
        __syncthreads();
        int cur = 0;
        for (int k = k0; k < k1; ++k) {
            double* wsk = bb.ws + ((size_t)b * P.N + k) * wstep;
            const BMat U = big_ws_U<DP>(wsk);
            const BMat Qo = big_state<DP>(scr, cur), Qn = (k == k1 - 1) ? big_state<DP>(bb.Qb, (size_t)b * nc + ch) : big_state<DP>(scr, cur ^ 1);
            for (int e = 0; e < ne; ++e) {
                const BMat Wo = big_state<DP>(scr, 2 + 2 * e + cur);
                const BMat Wn = (k == k1 - 1) ? big_state<DP>(bb.Wlb, ((size_t)b * nc + ch) * ne + e) : big_state<DP>(scr, 2 + 2 * e + (cur ^ 1));
                const BMat De = big_ws_D<DP>(wsk, e);
                G.zero();
                G.mac(smd, U, BOP_N, &Wo, 1, BOP_N);
                G.mac(smd, De, BOP_N, &Qo, 1, BOP_N);
                G.store(Wn, 1.0, nullptr, 0.0, 0.0, d);
            }
            G.zero();
            G.mac(smd, U, BOP_N, &Qo, 1, BOP_N);
            G.store(Qn, 1.0, nullptr, 0.0, 0.0, d);
            __syncthreads();
            cur ^= 1;
        }
    }
}

// ---- forward prefix over chunks: Cb[c] = Q_c Cb[c-1]  (one CTA per pulse) ------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT)
k_big_prefix(const DevProblem P, const BigBufs bb, int B, int nc) {
    extern __shared__ double smd[];
    BigGemm<DP> G;
    const int b = blockIdx.x;
    if (b >= B) return;
    bmat_copy<DP>(big_state<DP>(bb.Cb, (size_t)b * nc), big_state<DP>(bb.Qb, (size_t)b * nc), BigGemm<DP>::NT);
    __syncthreads();
    for (int c = 1; c < nc; ++c) {
        const BMat Cp = big_state<DP>(bb.Cb, (size_t)b * nc + c - 1);
        G.zero();
        G.mac(smd, big_state<DP>(bb.Qb, (size_t)b * nc + c), BOP_N, &Cp, 1, BOP_N);
        G.store(big_state<DP>(bb.Cb, (size_t)b * nc + c), 1.0, nullptr, 0.0, 0.0, P.d);
        __syncthreads();
    }
}

// real d x d matrix (column-major, as DevProblem holds them) -> planar BMat in scratch, both orientations, zero imaginary part
template <int DP>
__device__ __forceinline__ void bmat_from_real(const BMat& Z, const double* __restrict__ M, int d, bool transpose, int nt) {
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        const int r = idx / DP, c = idx % DP;
        const double v = (r < d && c < d) ? (transpose ? M[c + d * r] : M[r + d * c]) : 0.0;
        Z.re[idx] = v; Z.im[idx] = 0.0;
        Z.reT[(size_t)c * DP + r] = v; Z.imT[(size_t)c * DP + r] = 0.0;
    }
}
// Z = sum_t coef_t Tgt_t (dense target term matrices, canonical planes), both orientations
template <int DP>
__device__ __forceinline__ void bmat_from_terms(const BMat& Z, const double* __restrict__ tm, const cplx* coef, int nterms, int nt) {
    const size_t plane = (size_t)DP * DP;
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        double vr = 0.0, vi = 0.0;
        for (int t = 0; t < nterms; ++t) {
            const cplx c = coef[t];
            const double mr = tm[(size_t)t * 2 * plane + idx], mi = tm[(size_t)t * 2 * plane + plane + idx];
            vr += c.x * mr - c.y * mi; vi += c.x * mi + c.y * mr;
        }
        const int r = idx / DP, c2 = idx % DP;
        Z.re[idx] = vr; Z.im[idx] = vi;
        Z.reT[(size_t)c2 * DP + r] = vr; Z.imT[(size_t)c2 * DP + r] = vi;
    }
}
template <int DP>
__device__ __forceinline__ cplx bmat_trace(const BMat& A, double* red, int nt) {
    cplx s = cmk(0.0, 0.0);
    for (int i = threadIdx.x; i < DP; i += nt) { s.x += A.re[(size_t)i * DP + i]; s.y += A.im[(size_t)i * DP + i]; }
    return cta_sum(s, red, nt);
}

// ---- per (pulse, role): [W prefix,] fidelity algebra, backward co-states at chunk ends -------------------------------------
// role 0: F, K, G chain.  role 1 + e: W chain, E = W_N / eps, F_d2err[e], K', (G', H') chains.
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT)
k_big_scan(const DevProblem P, const BigData Bd, const BigBufs bb, const double* __restrict__ X, int B, int nc,
           double* __restrict__ Fout, double* __restrict__ F2out, double* __restrict__ addT) {
    constexpr int NT = BigGemm<DP>::NT;
    extern __shared__ double smd[];
    double* red = smd + BigGemm<DP>::SMEM_DOUBLES;
    cplx* coef = reinterpret_cast<cplx*>(red + 64);
    const int b = blockIdx.x, role = blockIdx.y, es = role - 1;
    const int ne = P.e, d = P.d;
    if (b >= B) return;
    BigGemm<DP> G;
    double* scr = bb.scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * RG_BIG_SWEEP_MATS * 4 * DP * DP;
    const BMat mU0 = big_state<DP>(scr, 0), mM = big_state<DP>(scr, 1), T1 = big_state<DP>(scr, 2), T2 = big_state<DP>(scr, 3),
               T3 = big_state<DP>(scr, 4), T4 = big_state<DP>(scr, 5), mK = big_state<DP>(scr, 6), mE = big_state<DP>(scr, 7),
               mPP = big_state<DP>(scr, 8), mP = big_state<DP>(scr, 9), mPPt = big_state<DP>(scr, 10), mSum = big_state<DP>(scr, 11),
               mV = big_state<DP>(scr, 12), S1 = big_state<DP>(scr, 13), mK2 = big_state<DP>(scr, 14);
    const double* xp = X + (size_t)b * P.nx;
    double xadd[RG_MAX_ADD];
    for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
    // ---- forward W chain (role e): W_c = Q_c W_{c-1} + Wl_c C_{c-1}
    if (role > 0) {
        for (int c = 0; c < nc; ++c) {
            const BMat Wn = big_state<DP>(bb.Wb, ((size_t)b * ne + es) * nc + c);
            const BMat Wl = big_state<DP>(bb.Wlb, ((size_t)b * nc + c) * ne + es);
            if (c == 0) { bmat_copy<DP>(Wn, Wl, NT); __syncthreads(); continue; }
            const BMat Wp = big_state<DP>(bb.Wb, ((size_t)b * ne + es) * nc + c - 1), Cp = big_state<DP>(bb.Cb, (size_t)b * nc + c - 1);
            G.zero();
            G.mac(smd, big_state<DP>(bb.Qb, (size_t)b * nc + c), BOP_N, &Wp, 1, BOP_N);
            G.mac(smd, Wl, BOP_N, &Cp, 1, BOP_N);
            G.store(Wn, 1.0, nullptr, 0.0, 0.0, d);
            __syncthreads();
        }
    }
    // ---- fidelity algebra (src/FidelityCalculations.jl:47-114; same sequence as fid_algebra in rg_smalld.cuh)
    const BMat CN = big_state<DP>(bb.Cb, (size_t)b * nc + nc - 1);
    BMat mU = CN;
    if (role > 0) {
        const BMat WN = big_state<DP>(bb.Wb, ((size_t)b * ne + es) * nc + nc - 1);
        for (int idx = threadIdx.x; idx < 4 * DP * DP; idx += NT) mE.re[idx] = WN.re[idx] * P.inv_eps;      // four contiguous planes
        mU = mE;
    }
    bmat_from_real<DP>(mPP, P.PP, d, false, NT);
    bmat_from_real<DP>(mPPt, P.PP, d, true, NT);
    bmat_from_real<DP>(mP, P.Pm, d, false, NT);
    for (int idx = threadIdx.x; idx < DP * DP; idx += NT) {
        const int r = idx / DP, c = idx % DP;
        const double v = (r < d && c < d) ? P.PP[r + d * c] + P.PP[c + d * r] : 0.0;
        mSum.re[idx] = v; mSum.im[idx] = 0.0; mSum.reT[idx] = v; mSum.imT[idx] = 0.0;
    }
    {
        EvalCtx ec{nullptr, xadd, 0.0, P.table, P.N, 0};
        for (int t = threadIdx.x; t < P.ntt; t += NT) { cplx bs, dl; term_coef(P.tterms[t], ec, RG_S_NONE, 0, 0.0, bs, dl); coef[t] = bs; }
        __syncthreads();
        bmat_from_terms<DP>(mU0, Bd.tgtM, coef, P.ntt, NT);
    }
    __syncthreads();
    const double Dt = P.Dtr, DD1 = Dt * (Dt + 1.0);
    G.zero(); G.mac(smd, mU0, BOP_H, &mU, 1, BOP_N); G.store(mM, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();        // M = U0^dag U
    G.zero(); G.mac(smd, mPP, BOP_N, &mM, 1, BOP_N); G.store(T1, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();        // T1 = PP M
    const cplx tau = bmat_trace<DP>(T1, red, NT);
    G.zero(); G.mac(smd, mP, BOP_N, &mM, 1, BOP_H); G.store(T2, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();         // T2 = P M^dag
    const cplx tr12 = bmat_trace_prod<DP>(T1, T2, red, NT);
    double Fval = (tr12.x + tau.x * tau.x + tau.y * tau.y) / DD1;
    G.zero(); G.mac(smd, mP, BOP_H, &mM, 1, BOP_H); G.store(T4, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();         // T4 = P^T M^dag
    G.zero(); G.mac(smd, T2, BOP_N, &mPP, 1, BOP_N); G.mac(smd, T4, BOP_N, &mPPt, 1, BOP_N);
    G.store(T3, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();                                                        // P M^dag PP + P^T M^dag PP^T
    for (int idx = threadIdx.x; idx < DP * DP; idx += NT) {                                                         // + 2 conj(tau) PP
        const int r = idx / DP, c = idx % DP;
        const double pp = mPP.re[idx];
        T3.re[idx] += 2.0 * tau.x * pp; T3.im[idx] -= 2.0 * tau.y * pp;
        T3.reT[(size_t)c * DP + r] += 2.0 * tau.x * pp; T3.imT[(size_t)c * DP + r] -= 2.0 * tau.y * pp;
    }
    __syncthreads();
    G.zero(); G.mac(smd, T3, BOP_N, &mU0, 1, BOP_H); G.store(mK, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();        // K = R U0^dag
    BMat Kfin = mK;
    double scale_out = 1.0;
    if (role > 0) {
        G.zero(); G.mac(smd, mU, BOP_H, &mU, 1, BOP_N); G.store(T4, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();     // E^dag E
        const cplx tee = bmat_trace_prod<DP>(mPP, T4, red, NT);
        Fval = 2.0 * (tr12.x - (1.0 + Dt) * tee.x + tau.x * tau.x + tau.y * tau.y) / DD1;
        G.zero(); G.mac(smd, mSum, BOP_N, &mU, 1, BOP_H); G.store(mK2, -(1.0 + Dt), &mK, 1.0, 0.0, d); __syncthreads();   // K' = K - (1+D)(PP+PP^T) E^dag
        Kfin = mK2;
        scale_out = 2.0;
    }
    if (threadIdx.x == 0) { if (role == 0) Fout[b] = Fval; else F2out[(size_t)b * ne + es] = Fval; }
    for (int j = 0; j < P.a; ++j) {                 // target-derivative parts of the x_add gradient
        EvalCtx ec{nullptr, xadd, 0.0, P.table, P.N, 0};
        const double h = __dsub_rn(__dadd_rn(xadd[j], P.eps), xadd[j]);
        __syncthreads();
        for (int t = threadIdx.x; t < P.ntt; t += NT) { cplx bs, dl; term_coef(P.tterms[t], ec, RG_S_ADD, j, h, bs, dl); coef[t] = cscale(dl, P.inv_eps); }
        __syncthreads();
        bmat_from_terms<DP>(mV, Bd.tgtM, coef, P.ntt, NT);
        __syncthreads();
        G.zero(); G.mac(smd, mV, BOP_H, &mU, 1, BOP_N); G.store(S1, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();     // S1 = V^dag U
        const cplx t3 = bmat_trace_prod<DP>(mPP, S1, red, NT);
        G.zero(); G.mac(smd, S1, BOP_N, &T2, 1, BOP_N); G.store(T4, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();
        const cplx t1 = bmat_trace_prod<DP>(mPP, T4, red, NT);
        G.zero(); G.mac(smd, mU, BOP_H, &mV, 1, BOP_N); G.store(S1, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();     // U^dag V
        G.zero(); G.mac(smd, mP, BOP_N, &S1, 1, BOP_N); G.store(T4, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();     // P U^dag V
        const cplx t2 = bmat_trace_prod<DP>(T1, T4, red, NT);
        const double val = scale_out * (t1.x + t2.x + 2.0 * (tau.x * t3.x + tau.y * t3.y)) / DD1;
        if (threadIdx.x == 0) addT[((size_t)b * (1 + ne) + role) * P.a + j] = val;
    }
    __syncthreads();
    // ---- backward over chunks: co-states at chunk ends.  G_{c-1} = G_c Q_c ; H_{c-1} = H_c Q_c + G_c Wl_c
    double* Gdst = role == 0 ? bb.Gb : bb.G1b;
    const size_t gbase = role == 0 ? (size_t)b * nc : ((size_t)b * ne + es) * nc;
    bmat_copy<DP>(big_state<DP>(Gdst, gbase + nc - 1), Kfin, NT);
    if (role > 0) bmat_zero<DP>(big_state<DP>(bb.H1b, gbase + nc - 1), NT);
    __syncthreads();
    for (int c = nc - 1; c >= 1; --c) {
        const BMat Gc = big_state<DP>(Gdst, gbase + c), Qc = big_state<DP>(bb.Qb, (size_t)b * nc + c);
        if (role > 0) {
            const BMat Hc = big_state<DP>(bb.H1b, gbase + c), Wl = big_state<DP>(bb.Wlb, ((size_t)b * nc + c) * ne + es);
            G.zero(); G.mac(smd, Hc, BOP_N, &Qc, 1, BOP_N); G.mac(smd, Gc, BOP_N, &Wl, 1, BOP_N);
            G.store(big_state<DP>(bb.H1b, gbase + c - 1), 1.0, nullptr, 0.0, 0.0, d);
        }
        G.zero(); G.mac(smd, Gc, BOP_N, &Qc, 1, BOP_N);
        G.store(big_state<DP>(Gdst, gbase + c - 1), 1.0, nullptr, 0.0, 0.0, d);
        __syncthreads();
    }
}

// Re tr(L M) with M canonical only: sum_ij Re(L_ij M_ji)
template <int DP>
__device__ __forceinline__ double bmat_retrace_prod(const BMat& Lm, const BMat& M, double* red, int nt) {
    cplx s = cmk(0.0, 0.0);
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        const int r = idx / DP, c = idx % DP;
        s.x += Lm.re[idx] * M.re[(size_t)c * DP + r] - Lm.im[idx] * M.im[(size_t)c * DP + r];
    }
    return cta_sum(s, red, nt).x;
}

// ---- backward gradient sweep, fidelity role: out0[b*nx + p*k + v] = scale0 * Re tr((C_{k-1} G_k) dU_k^v) ----------------------
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT)
k_big_grad(const DevProblem P, const BigBufs bb, int B, int L, int nc, double* __restrict__ out0, double scale0, double* __restrict__ addS) {
    constexpr int NT = BigGemm<DP>::NT;
    extern __shared__ double smd[];
    double* red = smd + BigGemm<DP>::SMEM_DOUBLES;
    const int nv = P.nvar, ne = P.e, d = P.d;
    const size_t wstep = big_ws_step_doubles(DP, nv, ne);
    double* scr = bb.scratch + (size_t)blockIdx.x * RG_BIG_SWEEP_MATS * 4 * DP * DP;
    BigGemm<DP> G;
    for (long long item = blockIdx.x; item < (long long)B * nc; item += gridDim.x) {
        const int b = (int)(item / nc), ch = (int)(item % nc);
        const int k0 = ch * L, k1 = min(P.N, k0 + L);
        __syncthreads();
        bmat_copy<DP>(big_state<DP>(scr, 0), big_state<DP>(bb.Cb, (size_t)b * nc + ch), NT);
        bmat_copy<DP>(big_state<DP>(scr, 2), big_state<DP>(bb.Gb, (size_t)b * nc + ch), NT);
        __syncthreads();
        int cur = 0;
        const BMat Lam = big_state<DP>(scr, 4);
        for (int k = k1 - 1; k >= k0; --k) {
            double* wsk = bb.ws + ((size_t)b * P.N + k) * wstep;
            const BMat U = big_ws_U<DP>(wsk);
            const BMat C = big_state<DP>(scr, cur), Cp = big_state<DP>(scr, cur ^ 1), Gk = big_state<DP>(scr, 2 + cur), Gn = big_state<DP>(scr, 2 + (cur ^ 1));
            G.zero(); G.mac(smd, U, BOP_H, &C, 1, BOP_N); G.store(Cp, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();      // C_{k-1} = U^dag C_k
            G.zero(); G.mac(smd, Cp, BOP_N, &Gk, 1, BOP_N); G.store(Lam, 1.0, nullptr, 0.0, 0.0, d);                    // Lambda = C_{k-1} G_k
            G.zero(); G.mac(smd, Gk, BOP_N, &U, 1, BOP_N); G.store(Gn, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();      // G_{k-1} = G_k U_k
            for (int v = 0; v < nv; ++v) {
                const double s = bmat_retrace_prod<DP>(Lam, big_ws_dU<DP>(wsk, ne, v), red, NT) * scale0;
                if (threadIdx.x == 0) {
                    if (P.var_space[v] == RG_S_MAIN) out0[(size_t)b * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
                    else addS[(((size_t)b * (1 + ne)) * P.a + P.var_index[v]) * P.N + k] = s;
                }
            }
            __syncthreads();
            cur ^= 1;
        }
    }
}

// ---- backward sweep of the sensitivity gradient, error source e = blockIdx.y ---------------------------------------------------
//   out1 = (2/DD1) Re{ tr((C H' + W G') dU)/eps^2 + tr((C G') d2U)/eps2^2 }   (C, W at step k-1; G', H' at step k)
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT)
k_big_grad_err(const DevProblem P, const BigBufs bb, int B, int L, int nc, double* __restrict__ out1, double* __restrict__ addS) {
    constexpr int NT = BigGemm<DP>::NT;
    extern __shared__ double smd[];
    double* red = smd + BigGemm<DP>::SMEM_DOUBLES;
    const int nv = P.nvar, ne = P.e, d = P.d, es = blockIdx.y;
    const size_t wstep = big_ws_step_doubles(DP, nv, ne);
    double* scr = bb.scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * RG_BIG_SWEEP_MATS * 4 * DP * DP;
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    const double f1 = 2.0 / DD1 * P.inv_eps * P.inv_eps, f2 = 2.0 / DD1 * P.inv_eps2sq;
    BigGemm<DP> G;
    for (long long item = blockIdx.x; item < (long long)B * nc; item += gridDim.x) {
        const int b = (int)(item / nc), ch = (int)(item % nc);
        const int k0 = ch * L, k1 = min(P.N, k0 + L);
        const size_t sidx = ((size_t)b * ne + es) * nc + ch;
        __syncthreads();
        // scratch: 0/1 C, 2/3 W, 4/5 G', 6/7 H', 8 -T, 9 Lambda1, 10 Lambda2
        bmat_copy<DP>(big_state<DP>(scr, 0), big_state<DP>(bb.Cb, (size_t)b * nc + ch), NT);
        bmat_copy<DP>(big_state<DP>(scr, 2), big_state<DP>(bb.Wb, sidx), NT);
        bmat_copy<DP>(big_state<DP>(scr, 4), big_state<DP>(bb.G1b, sidx), NT);
        bmat_copy<DP>(big_state<DP>(scr, 6), big_state<DP>(bb.H1b, sidx), NT);
        __syncthreads();
        int cur = 0;
        const BMat Tn = big_state<DP>(scr, 8), L1 = big_state<DP>(scr, 9), L2 = big_state<DP>(scr, 10);
        for (int k = k1 - 1; k >= k0; --k) {
            double* wsk = bb.ws + ((size_t)b * P.N + k) * wstep;
            const BMat U = big_ws_U<DP>(wsk), De = big_ws_D<DP>(wsk, es);
            const BMat C = big_state<DP>(scr, cur), Cp = big_state<DP>(scr, cur ^ 1), W = big_state<DP>(scr, 2 + cur), Wp = big_state<DP>(scr, 2 + (cur ^ 1));
            const BMat Gk = big_state<DP>(scr, 4 + cur), Gn = big_state<DP>(scr, 4 + (cur ^ 1)), Hk = big_state<DP>(scr, 6 + cur), Hn = big_state<DP>(scr, 6 + (cur ^ 1));
            G.zero(); G.mac(smd, U, BOP_H, &C, 1, BOP_N); G.store(Cp, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();        // C_{k-1}
            G.zero(); G.mac(smd, De, BOP_N, &Cp, 1, BOP_N); G.store(Tn, -1.0, nullptr, 0.0, 0.0, d); __syncthreads();     // -D_k C_{k-1}
            { const BMat ys[2] = {W, Tn}; G.zero(); G.mac(smd, U, BOP_H, ys, 2, BOP_N); G.store(Wp, 1.0, nullptr, 0.0, 0.0, d); }   // W_{k-1}
            G.zero(); G.mac(smd, Cp, BOP_N, &Gk, 1, BOP_N); G.store(L2, 1.0, nullptr, 0.0, 0.0, d);                        // Lambda2 = C G'
            G.zero(); G.mac(smd, Hk, BOP_N, &U, 1, BOP_N); G.mac(smd, Gk, BOP_N, &De, 1, BOP_N); G.store(Hn, 1.0, nullptr, 0.0, 0.0, d);   // H' U + G' D
            G.zero(); G.mac(smd, Gk, BOP_N, &U, 1, BOP_N); G.store(Gn, 1.0, nullptr, 0.0, 0.0, d); __syncthreads();        // G' U
            G.zero(); G.mac(smd, Cp, BOP_N, &Hk, 1, BOP_N); G.mac(smd, Wp, BOP_N, &Gk, 1, BOP_N); G.store(L1, 1.0, nullptr, 0.0, 0.0, d);   // Lambda1
            __syncthreads();
            for (int v = 0; v < nv; ++v) {
                const double s1 = bmat_retrace_prod<DP>(L1, big_ws_dU<DP>(wsk, ne, v), red, NT);
                const double s2 = bmat_retrace_prod<DP>(L2, big_ws_d2U<DP>(wsk, nv, ne, v, es), red, NT);
                const double s = f1 * s1 + f2 * s2;
                if (threadIdx.x == 0) {
                    if (P.var_space[v] == RG_S_MAIN) out1[((size_t)b * ne + es) * P.nx + (size_t)P.p * k + P.var_index[v]] = s;
                    else addS[(((size_t)b * (1 + ne) + 1 + es) * P.a + P.var_index[v]) * P.N + k] = s;
                }
            }
            __syncthreads();
            cur ^= 1;
        }
    }
}
