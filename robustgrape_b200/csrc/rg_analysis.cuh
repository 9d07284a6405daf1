// rg_analysis.cuh -- kernels for the analysis entry points:
//   k_interaction_ops : calculate_interaction_error_operators (src/UnitaryCalculations.jl:180-204)
//   k_response        : calculate_fidelity_response / _fft      (src/FidelityCalculations.jl:246-280, 306-343)
//   k_expectation     : calculate_expectation_values            (src/FidelityCalculations.jl:368-390)
#pragma once
#include "rg_smalld.cuh"

// O_k[e] = C_{k-1}^{-1} (Herr_e(k, x_k, x_add, eps)/eps) C_{k-1},  C_k = U_k C_{k-1}  (uses the cumulative
// operator *before* step k, :194-200).  One group (D lanes of one warp) per pulse, sequential in time;
// C^{-1} = C^dagger (Hermitian problems).  O layout: (d, d, N, e) column-major.
template <int D>
__global__ void __launch_bounds__(32)
k_interaction_ops(const DevProblem P, const double* __restrict__ x, cplx* __restrict__ O, int* __restrict__ status) {
    constexpr unsigned amask = (D == 32) ? 0xffffffffu : ((1u << D) - 1u);
    constexpr int DD = D * D;
    extern __shared__ cplx smem[];
    const StagedDesc sd = stage_desc(P, reinterpret_cast<unsigned char*>(smem));
    const int l = threadIdx.x;
    if (l >= D) return;
    cplx* base = smem + staged_desc_bytes(P.nterms, P.nent, D) / sizeof(cplx);
    cplx* mA = base; cplx* mD = base + DD; cplx* mX = base + 2 * DD; cplx* mC = base + 3 * DD; cplx* mE = base + 4 * DD;
    cplx* mCi = base + 5 * DD;       // C^{-1} (non-Hermitian H), assembled from the rows each lane carries
    cplx* coef = base + 6 * DD;      // 2 * nterms
    const int nt = P.nterms;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = x[(size_t)P.p * P.N + j];
    cplx c[D], ci[D];                // column l of C, row l of C^{-1}
#pragma unroll
    for (int i = 0; i < D; ++i) { c[i] = cmk(i == l ? 1.0 : 0.0, 0.0); ci[i] = c[i]; }
    for (int t = l; t < nt; t += D) coef[nt + t] = cmk(0.0, 0.0);
    __syncwarp(amask);
    assemble_col<D>(sd.ents, sd.colptr, coef + nt, mD, l, true);     // zero perturbation: only U is needed
    for (int k = 0; k < P.N; ++k) {
        for (int i = 0; i < P.p; ++i) xk[i] = x[(size_t)k * P.p + i];
#pragma unroll
        for (int i = 0; i < D; ++i) { mC[i + D * l] = c[i]; mCi[l + D * i] = ci[i]; }
        __syncwarp(amask);
        for (int e = 0; e < P.e; ++e) {
            // Herr_e(eps) / eps : value coefficients without the -i dt factor
            EvalCtx ec{xk, xadd, P.eps, P.table, P.N, k};
            for (int t = l; t < nt; t += D) {
                cplx b = cmk(0, 0), dl;
                if (sd.terms[t].owner == e) term_coef(sd.terms[t], ec, RG_S_NONE, 0, 0.0, b, dl);
                coef[t] = cscale(b, P.inv_eps);
            }
            __syncwarp(amask);
            assemble_col<D>(sd.ents, sd.colptr, coef, mE, l);
            __syncwarp(amask);
            cplx t1[D], o[D];
            matvec<D>(mE, c, t1);            // column l of Oerr C
            if (P.hermitian) matvec_adj<D>(mC, t1, o);        // column l of C^dagger Oerr C
            else matvec<D>(mCi, t1, o);                       // column l of C^{-1} Oerr C
            cplx* dst = O + ((size_t)e * P.N + k) * DD + l * D;
#pragma unroll
            for (int i = 0; i < D; ++i) dst[i] = o[i];
            __syncwarp(amask);
        }
        fill_coefs<D>(P, sd.terms, coef, VK_BASE, 0, 0.0, RG_S_NONE, 0, 0.0, xk, xadd, k, l);
        __syncwarp(amask);
        assemble_col<D>(sd.ents, sd.colptr, coef, mA, l);
        double nrm = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { const cplx a = mA[i + D * l]; nrm += sqrt(a.x * a.x + a.y * a.y); }
        int m, sq;
        {
            unsigned long long nb = __double_as_longlong(nrm * 1.001);
            nb = __reduce_max_sync(amask, (unsigned)(nb >> 32));
            expm_plan(__longlong_as_double((long long)((nb + 1ull) << 32)), m, sq);
            if (m == 99) { if (l == 0) atomicOr(status, 1); m = 12; }
        }
        if (sq) {
            const double sc = scalbn(1.0, -sq);
#pragma unroll
            for (int i = 0; i < D; ++i) mA[i + D * l] = cscale(mA[i + D * l], sc);
        }
        __syncwarp(amask);
        cplx y[D], dl[D];
        horner_fo<D>(mA, mD, l, m, y, dl);
        for (int q2 = 0; q2 < sq; ++q2) {
            __syncwarp(amask);
#pragma unroll
            for (int i = 0; i < D; ++i) mX[i + D * l] = y[i];
            __syncwarp(amask);
            cplx yn[D];
            matvec<D>(mX, y, yn);
#pragma unroll
            for (int i = 0; i < D; ++i) y[i] = yn[i];
        }
        __syncwarp(amask);
#pragma unroll
        for (int i = 0; i < D; ++i) mX[i + D * l] = y[i];
        __syncwarp(amask);
        cplx cn[D];
        matvec<D>(mX, c, cn);
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = cn[i];
        __syncwarp(amask);
        if (!P.hermitian) {
            // C_k^{-1} = C_{k-1}^{-1} U_k^{-1},  U_k^{-1} = exp(-A) (scaled A is still in mA)
#pragma unroll
            for (int i = 0; i < D; ++i) { mX[i + D * l] = cmk(-mA[i + D * l].x, -mA[i + D * l].y); }
            __syncwarp(amask);
            cplx yi[D], dli[D];
            horner_fo<D>(mX, mD, l, m, yi, dli);
            for (int q2 = 0; q2 < sq; ++q2) {
                __syncwarp(amask);
#pragma unroll
                for (int i = 0; i < D; ++i) mE[i + D * l] = yi[i];
                __syncwarp(amask);
                cplx yn[D];
                matvec<D>(mE, yi, yn);
#pragma unroll
                for (int i = 0; i < D; ++i) yi[i] = yn[i];
            }
            __syncwarp(amask);
#pragma unroll
            for (int i = 0; i < D; ++i) mE[i + D * l] = yi[i];
            __syncwarp(amask);
            cplx cin[D];
            vecmat<D>(ci, mE, cin);
#pragma unroll
            for (int i = 0; i < D; ++i) ci[i] = cin[i];
            __syncwarp(amask);
        }
    }
}


// Time-parallel version for Hermitian problems: one group per chunk of L steps.  The forward state at the chunk start
// comes from the chunk scan (Cb holds C at chunk *ends*), the step propagators from the workspace (object 0, dense
// layout), and the error Hamiltonians are assembled on the fly:  O_k[e] = C_{k-1}^dagger (Herr_e(eps)/eps) C_{k-1}.
template <int D>
__global__ void __launch_bounds__(128)
k_interaction_ops_par(const DevProblem P, const double* __restrict__ x, int L, int nc, const cplx* __restrict__ ws,
                      const cplx* __restrict__ Cb, cplx* __restrict__ O) {
    constexpr int G = GroupInfo<D>::G;
    constexpr unsigned amask = GroupInfo<D>::amask;
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ cplx smem[];
    const StagedDesc sd = stage_desc(P, reinterpret_cast<unsigned char*>(smem));
    if (lane >= G * D) return;
    const int g = lane / D, l = lane - g * D;
    long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * G + g;
    const bool live = item < nc;
    if (!live) item = nc - 1;
    const int ch = (int)item;
    const int nt = P.nterms;
    cplx* base = smem + staged_desc_bytes(P.nterms, P.nent, D) / sizeof(cplx) + (size_t)(warp * G + g) * rg_odd(3 * DD + nt);
    cplx* mC = base; cplx* mE = base + DD; cplx* mU = base + 2 * DD; cplx* coef = base + 3 * DD;
    double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
    for (int j = 0; j < P.a; ++j) xadd[j] = x[(size_t)P.p * P.N + j];
    cplx c[D];
    if (ch == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = cmk(i == l ? 1.0 : 0.0, 0.0);
    } else {
        const cplx* src = Cb + (size_t)(ch - 1) * DD + l * D;
#pragma unroll
        for (int i = 0; i < D; ++i) c[i] = src[i];
    }
    const int k0 = ch * L, k1 = min(P.N, k0 + L);
    for (int kk = 0; kk < L; ++kk) {
        const bool ghost = (k0 + kk >= k1);
        const int k = min(k0 + kk, k1 - 1);
        for (int i = 0; i < P.p; ++i) xk[i] = x[(size_t)k * P.p + i];
        __syncwarp(amask);
#pragma unroll
        for (int i = 0; i < D; ++i) { mC[i + D * l] = c[i]; mU[i + D * l] = ws[(size_t)k * DD + l * D + i]; }
        __syncwarp(amask);
        for (int e = 0; e < P.e; ++e) {
            EvalCtx ec{xk, xadd, P.eps, P.table, P.N, k};
            for (int t = l; t < nt; t += D) {
                cplx b = cmk(0, 0), dl;
                if (sd.terms[t].owner == e) term_coef(sd.terms[t], ec, RG_S_NONE, 0, 0.0, b, dl);
                coef[t] = cscale(b, P.inv_eps);
            }
            __syncwarp(amask);
            assemble_col<D>(sd.ents, sd.colptr, coef, mE, l);
            __syncwarp(amask);
            cplx t1[D], o[D];
            matvec<D>(mE, c, t1);
            matvec_adj<D>(mC, t1, o);
            if (live && !ghost) {
                cplx* dst = O + ((size_t)e * P.N + k) * DD + l * D;
#pragma unroll
                for (int i = 0; i < D; ++i) dst[i] = o[i];
            }
            __syncwarp(amask);
        }
        if (!ghost) {
            cplx cn[D];
            matvec<D>(mU, c, cn);
#pragma unroll
            for (int i = 0; i < D; ++i) c[i] = cn[i];
        }
    }
}

// Response function for one (frequency, error source) per block:
//   S = sum_{j=0}^{N-1} e^{-i w dt j} O_{j+1} ,  T = sum_{k} e^{+i w dt (k - 1 + shift)} O_k   (k = 1..N)
//   R = dt^2 [ Re tr_mod(T S P)/D - Re tr_mod(T P S P)/(D(D+1)) - Re(tr_mod(T P) tr_mod(S P))/(D(D+1)) ]
// shift = 1 reproduces calculate_fidelity_response (whose outer weight runs from 1, :269-271);
// shift = 0 with w_n = 2 pi n / (M dt) reproduces calculate_fidelity_response_fft (fft / M*ifft, :328-339).
// grid_mode: 0 = explicit frequencies; M > 0 = uniform FFT grid of M points (phases reduced exactly mod M).
// Mapping: NS time slices x DD matrix elements per block, strided over the threads (d*d may exceed the block).  The phase of a
// slice advances by a complex multiplication with e^{-i w dt NS} and is re-anchored with sincos every 32 terms, so a thread
// spends 2 complex FMAs + 1 complex multiplication per term instead of a sincos; the trace epilogue is spread over d*d threads.
template <int D>
__global__ void __launch_bounds__(128)
k_response(const DevProblem P, const cplx* __restrict__ O, const double* __restrict__ freqs, int first, int count,
           int M, int shift, double* __restrict__ R) {
    constexpr int DD = D * D;
    constexpr int NS = 128 / DD > 0 ? 128 / DD : 1;     // time slices per block
    __shared__ cplx sS[NS][DD], sT[NS][DD];
    __shared__ cplx mS[DD], mT[DD], SP[DD], TP[DD];
    __shared__ double red[4][8];
    const int f = blockIdx.x, e = blockIdx.y;
    const int tid = threadIdx.x;
    const int n = first + f;
    const double w = M > 0 ? 0.0 : freqs[n];
    const double wdt = w * P.dt;
    auto phase = [&](int j, double& sn, double& cs) {     // e^{+i w dt j}
        if (M > 0) { const long long r = ((long long)j * n) % M; sincospi(2.0 * (double)r / (double)M, &sn, &cs); }
        else sincos(wdt * (double)j, &sn, &cs);
    };
    for (int idx = tid; idx < NS * DD; idx += blockDim.x) {
        const int el = idx % DD, sl = idx / DD;
        cplx s = cmk(0, 0), t = cmk(0, 0);
        const cplx* Oe = O + (size_t)e * P.N * DD;
        double sn, cs, sn2, cs2;
        phase(NS, sn2, cs2);                             // step of the slice: e^{+i w dt NS}
        int cnt = 0;
        cplx ph = cmk(1, 0);
        for (int j = sl; j < P.N; j += NS, ++cnt) {
            if ((cnt & 31) == 0) { phase(j, sn, cs); ph = cmk(cs, sn); }
            const cplx o = Oe[(size_t)j * DD + el];
            cfma(s, cconj(ph), o);           // e^{-i w dt j}
            cfma(t, ph, o);                  // e^{+i w dt j}
            ph = cmul(ph, cmk(cs2, sn2));
        }
        sS[sl][el] = s; sT[sl][el] = t;
    }
    __syncthreads();
    for (int el = tid; el < DD; el += blockDim.x) {
        cplx a = cmk(0, 0), b = cmk(0, 0);
        for (int q = 0; q < NS; ++q) { a = cadd(a, sS[q][el]); b = cadd(b, sT[q][el]); }
        if (shift) {
            double sn, cs;
            sincos(wdt, &sn, &cs);
            b = cmul(b, cmk(cs, sn));
        }
        mS[el] = a; mT[el] = b;
    }
    __syncthreads();
    // SP = S P, TP = T P with P = (P0 != 0) (tr_mod(A) = tr(P0 A); :47-51)
    for (int el = tid; el < DD; el += blockDim.x) {
        const int i = el % D, j = el / D;
        cplx a = cmk(0, 0), b = cmk(0, 0);
        for (int q = 0; q < D; ++q) { const double p = P.Pm[q + D * j]; a.x += mS[i + D * q].x * p; a.y += mS[i + D * q].y * p; b.x += mT[i + D * q].x * p; b.y += mT[i + D * q].y * p; }
        SP[el] = a; TP[el] = b;
    }
    __syncthreads();
    // t1 = tr(P0 T SP), t2 = tr(P0 TP SP), t3 = tr(P0 TP), t4 = tr(P0 SP): entry (a, b) of P0 multiplies X[b][a]
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int el = tid; el < DD; el += blockDim.x) {
        const int a = el % D, b = el / D;
        const double p0 = P.P0raw[a + D * b];
        if (p0 == 0.0) continue;
        cplx x1 = cmk(0, 0), x2 = cmk(0, 0);
        for (int q = 0; q < D; ++q) { cfma(x1, mT[b + D * q], SP[q + D * a]); cfma(x2, TP[b + D * q], SP[q + D * a]); }
        v[0] += p0 * x1.x; v[1] += p0 * x1.y; v[2] += p0 * x2.x; v[3] += p0 * x2.y;
        v[4] += p0 * TP[b + D * a].x; v[5] += p0 * TP[b + D * a].y; v[6] += p0 * SP[b + D * a].x; v[7] += p0 * SP[b + D * a].y;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    if ((tid & 31) == 0)
        for (int q = 0; q < 8; ++q) red[tid >> 5][q] = v[q];
    __syncthreads();
    if (tid == 0) {
        double t[8];
        for (int q = 0; q < 8; ++q) t[q] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
        const double Dt = P.Dtr, DD1 = Dt * (Dt + 1.0);
        const double r = t[0] / Dt - t[2] / DD1 - (t[4] * t[6] - t[5] * t[7]) / DD1;
        R[(size_t)e * count + f] = P.dt * P.dt * r;
    }
}

// s[k,e] = tr(P0 O_k^e); out[k,e] = Re(dt * sum_{j<=k} s[j,e] / D)   (cumsum is linear in the trace, :374,384-388).
// One block per error source: threads take the traces of contiguous runs of steps, then a block scan of the run totals.
template <int D>
__global__ void __launch_bounds__(256)
k_expectation(const DevProblem P, const cplx* __restrict__ O, double* __restrict__ out) {
    constexpr int DD = D * D;
    __shared__ double tot[256];
    const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int per = (P.N + nt - 1) / nt;
    const int k0 = min(P.N, tid * per), k1 = min(P.N, k0 + per);
    const double* P0 = P.P0raw;
    auto tr = [&](int k) {
        const cplx* o = O + ((size_t)e * P.N + k) * DD;
        double t = 0.0;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) t += P0[j + D * i] * o[i + D * j].x;       // Re tr(P0 O) = sum P0[j][i] Re O[i][j]
        return t;
    };
    double run = 0.0;
    for (int k = k0; k < k1; ++k) run += tr(k);
    tot[tid] = run;
    __syncthreads();
    if (tid == 0) { double acc = 0.0; for (int q = 0; q < nt; ++q) { const double v = tot[q]; tot[q] = acc; acc += v; } }   // exclusive prefix
    __syncthreads();
    double acc = tot[tid];
    for (int k = k0; k < k1; ++k) { acc += tr(k); out[(size_t)e * P.N + k] = P.dt * acc / P.Dtr; }
}
