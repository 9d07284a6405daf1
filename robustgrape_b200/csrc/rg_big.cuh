// rg_big.cuh -- dense path for 11 <= d <= 64 (BASELINE.json configs[4]: d = 64, N = 1e4, p = 8, e = 4).
//
// Everything is built on one primitive: a CTA-level complex FP64 matrix product on the FP64 tensor path
// (mma.sync.m8n8k4.f64, SASS DMMA; tcgen05 has no FP64 kind).  Matrices live in global memory (L2 / HBM) as *planar*
// DP x DP arrays (DP = d rounded up to a multiple of 16, zero padded): re[i*DP + j], im[i*DP + j], and -- for matrices that
// are used as operands again -- the transposed planes reT, imT, so that every operand a product needs ("left" as [i][k],
// "right" as [n][k]) is a plain row-contiguous read, conjugation being a sign flip at load time.  A product streams its
// operands through shared memory in K-chunks of 16 and accumulates the DP x DP complex result in registers; several
// products can accumulate into the same output (output stationary), and a right operand can be the sum of up to four
// matrices formed while loading -- exactly what the difference arithmetic below needs.
//
// k_big_steps: step propagators and all their exact finite differences.  exp(A) is a degree-12 Taylor polynomial of the
// scaled matrix B = A / 2^s in Paterson-Stockmeyer form (5 products) followed by s squarings.  Every matrix is carried as
// a *jet* of 1 + 2(nv + ne) + nv ne slots -- value, first differences at eps and at eps2 for every variable / error source,
// and the mixed second differences -- and every product is the exact jet product
//      Z0 = X0 Y0;   Zf = Xf (Y0 + Yf) + X0 Yf;   Zve = Xve (Y0 + Yv + Ye + Yve) + Xv (Ye + Yve) + Xe (Yv + Yve) + X0 Yve
// (no rounded exponentials are ever subtracted: src/UnitaryCalculations.jl:52,60,70,80-83 to ~1e-15, as in rg_smalld.cuh).
#pragma once
#include "rg_common.cuh"

#define RG_BIG_KC 16          // K-chunk staged in shared memory
#define RG_BIG_LDK 20         // padded row length of a staged chunk (doubles): fragment loads hit 16 distinct 8-byte banks

struct BMat { double* re; double* im; double* reT; double* imT; };     // planar DP x DP; reT/imT may be null
__host__ __device__ inline BMat bmat_at(double* base, size_t plane, bool both) {
    BMat m; m.re = base; m.im = base + plane; m.reT = both ? base + 2 * plane : nullptr; m.imT = both ? base + 3 * plane : nullptr;
    return m;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

enum { BOP_N = 0, BOP_H = 1 };       // operand as stored, or its conjugate transpose

template <int DP>
struct BigGemm {
    static constexpr int WM = DP / 16, WN = 2;              // warp tile: (8 WM) x 16
    static constexpr int NWARP = 2 * (DP / 16), NT = 32 * NWARP;
    static constexpr int SMEM_DOUBLES = 4 * DP * RG_BIG_LDK;     // left re/im + right re/im chunks
    double cr[WM][WN][2], ci[WM][WN][2];

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int a = 0; a < WM; ++a)
#pragma unroll
            for (int b = 0; b < WN; ++b) { cr[a][b][0] = cr[a][b][1] = ci[a][b][0] = ci[a][b][1] = 0.0; }
    }
    // acc += op(X) * op(sum_i Y_i).  Left needs rows [i][k]: N -> (re, im), H -> (reT, -imT).  Right needs rows [n][k]:
    // N -> (reT, imT), H -> (re, -im).  All threads of the CTA must call; sm holds SMEM_DOUBLES doubles.
    template <int NY>
    __device__ __forceinline__ void mac(double* sm, const BMat& X, int opX, const BMat* Y, int opY) {
        const double* xr = opX == BOP_N ? X.re : X.reT;
        const double* xi = opX == BOP_N ? X.im : X.imT;
        const double* yrp[NY]; const double* yip[NY];
#pragma unroll
        for (int q = 0; q < NY; ++q) { yrp[q] = opY == BOP_N ? Y[q].reT : Y[q].re; yip[q] = opY == BOP_N ? Y[q].imT : Y[q].im; }
        const double sx = opX == BOP_N ? 1.0 : -1.0, sy = opY == BOP_N ? 1.0 : -1.0;
        double* sLr = sm; double* sLi = sm + DP * RG_BIG_LDK; double* sRr = sm + 2 * DP * RG_BIG_LDK; double* sRi = sm + 3 * DP * RG_BIG_LDK;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int row0 = (warp / (DP / 16)) * (DP / 2), col0 = (warp % (DP / 16)) * 16;
        constexpr int ITER = DP * (RG_BIG_KC / 2) / NT;          // = 2: staging items per thread and chunk
        static_assert(ITER * NT == DP * (RG_BIG_KC / 2), "staging must divide evenly over the CTA");
        for (int k0 = 0; k0 < DP; k0 += RG_BIG_KC) {
            // all global loads of the chunk are issued before the first use (independent 16-byte loads in flight)
            double2 a[ITER], b[ITER], u[ITER][NY], v[ITER][NY];
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int idx = threadIdx.x + it * NT;
                const int r = idx / (RG_BIG_KC / 2), c = (idx % (RG_BIG_KC / 2)) * 2;
                const size_t go = (size_t)r * DP + k0 + c;
                a[it] = *reinterpret_cast<const double2*>(xr + go);
                b[it] = *reinterpret_cast<const double2*>(xi + go);
#pragma unroll
                for (int q = 0; q < NY; ++q) {
                    u[it][q] = *reinterpret_cast<const double2*>(yrp[q] + go);
                    v[it][q] = *reinterpret_cast<const double2*>(yip[q] + go);
                }
            }
            __syncthreads();                       // previous chunk consumed
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int idx = threadIdx.x + it * NT;
                const int r = idx / (RG_BIG_KC / 2), c = (idx % (RG_BIG_KC / 2)) * 2;
                b[it].x *= sx; b[it].y *= sx;
                *reinterpret_cast<double2*>(sLr + r * RG_BIG_LDK + c) = a[it];
                *reinterpret_cast<double2*>(sLi + r * RG_BIG_LDK + c) = b[it];
                double2 yr = u[it][0], yi = v[it][0];
#pragma unroll
                for (int q = 1; q < NY; ++q) { yr.x += u[it][q].x; yr.y += u[it][q].y; yi.x += v[it][q].x; yi.y += v[it][q].y; }
                yi.x *= sy; yi.y *= sy;
                *reinterpret_cast<double2*>(sRr + r * RG_BIG_LDK + c) = yr;
                *reinterpret_cast<double2*>(sRi + r * RG_BIG_LDK + c) = yi;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < RG_BIG_KC; kk += 4) {
                double ar[WM], ai[WM], nai[WM], br[WN], bi[WN];
#pragma unroll
                for (int a2 = 0; a2 < WM; ++a2) {
                    ar[a2] = sLr[(row0 + a2 * 8 + g) * RG_BIG_LDK + kk + t];
                    ai[a2] = sLi[(row0 + a2 * 8 + g) * RG_BIG_LDK + kk + t];
                    nai[a2] = -ai[a2];
                }
#pragma unroll
                for (int b2 = 0; b2 < WN; ++b2) {
                    br[b2] = sRr[(col0 + b2 * 8 + g) * RG_BIG_LDK + kk + t];
                    bi[b2] = sRi[(col0 + b2 * 8 + g) * RG_BIG_LDK + kk + t];
                }
#pragma unroll
                for (int a2 = 0; a2 < WM; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < WN; ++b2) {
                        dmma884(cr[a2][b2][0], cr[a2][b2][1], ar[a2], br[b2]);
                        dmma884(cr[a2][b2][0], cr[a2][b2][1], nai[a2], bi[b2]);
                        dmma884(ci[a2][b2][0], ci[a2][b2][1], ar[a2], bi[b2]);
                        dmma884(ci[a2][b2][0], ci[a2][b2][1], ai[a2], br[b2]);
                    }
            }
        }
    }
    // run-time source count (sweeps): dispatch to the unrolled versions
    __device__ __forceinline__ void mac(double* sm, const BMat& X, int opX, const BMat* Y, int nY, int opY) {
        if (nY == 1) mac<1>(sm, X, opX, Y, opY);
        else if (nY == 2) mac<2>(sm, X, opX, Y, opY);
        else mac<4>(sm, X, opX, Y, opY);
    }
    // Z = alpha * acc + sum_q beta[q] * Add[q] (canonical planes, q < nadd <= 3) [+ gamma * I on the first d diagonal entries];
    // writes the transposed planes when present.
    __device__ __forceinline__ void store(const BMat& Z, double alpha, const BMat* Add, const double* beta, int nadd, double gamma, int d) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int row0 = (warp / (DP / 16)) * (DP / 2), col0 = (warp % (DP / 16)) * 16;
#pragma unroll
        for (int a = 0; a < WM; ++a)
#pragma unroll
            for (int b = 0; b < WN; ++b) {
                const int r = row0 + a * 8 + g, c = col0 + b * 8 + 2 * t;
                double2 vr = make_double2(alpha * cr[a][b][0], alpha * cr[a][b][1]);
                double2 vi = make_double2(alpha * ci[a][b][0], alpha * ci[a][b][1]);
                for (int q = 0; q < nadd; ++q) {
                    const double2 pr = *reinterpret_cast<const double2*>(Add[q].re + (size_t)r * DP + c);
                    const double2 pi = *reinterpret_cast<const double2*>(Add[q].im + (size_t)r * DP + c);
                    vr.x = fma(beta[q], pr.x, vr.x); vr.y = fma(beta[q], pr.y, vr.y); vi.x = fma(beta[q], pi.x, vi.x); vi.y = fma(beta[q], pi.y, vi.y);
                }
                if (gamma != 0.0 && r < d) { if (c == r) vr.x += gamma; if (c + 1 == r) vr.y += gamma; }
                *reinterpret_cast<double2*>(Z.re + (size_t)r * DP + c) = vr;
                *reinterpret_cast<double2*>(Z.im + (size_t)r * DP + c) = vi;
                if (Z.reT) {
                    Z.reT[(size_t)c * DP + r] = vr.x; Z.reT[(size_t)(c + 1) * DP + r] = vr.y;
                    Z.imT[(size_t)c * DP + r] = vi.x; Z.imT[(size_t)(c + 1) * DP + r] = vi.y;
                }
            }
    }
    __device__ __forceinline__ void store(const BMat& Z, double alpha, const BMat* Add, double beta, double gamma, int d) {
        store(Z, alpha, Add, &beta, Add ? 1 : 0, gamma, d);
    }
};

// elementwise helpers over planar matrices (all threads of the CTA)
template <int DP>
__device__ __forceinline__ void bmat_set_identity(const BMat& Z, int d, int nt) {
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        const int r = idx / DP, c = idx % DP;
        const double v = (r == c && r < d) ? 1.0 : 0.0;
        Z.re[idx] = v; Z.im[idx] = 0.0;
        if (Z.reT) { Z.reT[idx] = v; Z.imT[idx] = 0.0; }
    }
}
template <int DP>
__device__ __forceinline__ void bmat_zero(const BMat& Z, int nt) {
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        Z.re[idx] = 0.0; Z.im[idx] = 0.0;
        if (Z.reT) { Z.reT[idx] = 0.0; Z.imT[idx] = 0.0; }
    }
}
template <int DP>
__device__ __forceinline__ void bmat_copy(const BMat& Z, const BMat& A, int nt) {
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        Z.re[idx] = A.re[idx]; Z.im[idx] = A.im[idx];
        if (Z.reT) { Z.reT[idx] = A.reT[idx]; Z.imT[idx] = A.imT[idx]; }
    }
}
// sum over the CTA of a complex value; result valid in every thread.  red: >= 2 * 32 doubles of shared memory.
__device__ __forceinline__ cplx cta_sum(cplx v, double* red, int nt) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, off); v.y += __shfl_xor_sync(0xffffffffu, v.y, off); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[2 * (threadIdx.x >> 5)] = v.x; red[2 * (threadIdx.x >> 5) + 1] = v.y; }
    __syncthreads();
    cplx s = cmk(0.0, 0.0);
    for (int w = 0; w < nt / 32; ++w) { s.x += red[2 * w]; s.y += red[2 * w + 1]; }
    return s;
}
// tr(A B) = sum_ij A_ij B_ji  (uses B's canonical planes at the transposed index)
template <int DP>
__device__ __forceinline__ cplx bmat_trace_prod(const BMat& A, const BMat& B, double* red, int nt) {
    cplx s = cmk(0.0, 0.0);
    for (int idx = threadIdx.x; idx < DP * DP; idx += nt) {
        const int r = idx / DP, c = idx % DP;
        const double ar = A.re[idx], ai = A.im[idx], br = B.re[(size_t)c * DP + r], bi = B.im[(size_t)c * DP + r];
        s.x += ar * br - ai * bi; s.y += ar * bi + ai * br;
    }
    return cta_sum(s, red, nt);
}

// ---- problem data of the dense path ------------------------------------------------------------------------------
struct BigData {
    int DP, nslots;                 // slots of a jet: 1 + 2 nf + nv ne, nf = nv + ne
    const double* termM;            // [nterms][2 planes][DP*DP] dense term matrices (canonical planes only)
    const double* tgtM;             // [ntt][2][DP*DP] dense target term matrices
    const double* projPP;           // [3][DP*DP]: PP, P (0/1), PP + PP^T as real planes
    double* scratch;                // per-CTA scratch
    size_t scratch_per_cta;         // doubles
};

// slot numbering of a jet
//   0                         value
//   1 + f                     first difference at eps   (f < nv: variable f; f >= nv: error source f - nv at err = eps)
//   1 + nf + f                first difference at eps2
//   1 + 2 nf + e * nv + v     mixed second difference (variable v, error source e) at (eps2, eps2)
__host__ __device__ inline int big_nslots(int nv, int ne) { return 1 + 2 * (nv + ne) + nv * ne; }

// A jet in scratch: slot s at base + s * 4 * DP*DP (planes re, im, reT, imT)
template <int DP>
__device__ __forceinline__ BMat jet_slot(double* base, int s) { return bmat_at(base + (size_t)s * 4 * DP * DP, (size_t)DP * DP, true); }

// Where the final stage of the propagator jet goes: the per-step workspace of the sweeps.
//   U (both forms) | D^e, e < ne (both forms) | dU^v, v < nv (canonical) | d2U^{v,e} (canonical, index e * nv + v)
__host__ __device__ inline size_t big_ws_step_doubles(int DP, int nv, int ne) { return (size_t)DP * DP * (4 + 4 * ne + 2 * nv + 2 * nv * ne); }
template <int DP>
__device__ __forceinline__ BMat big_ws_U(double* ws) { return bmat_at(ws, (size_t)DP * DP, true); }
template <int DP>
__device__ __forceinline__ BMat big_ws_D(double* ws, int e) { return bmat_at(ws + (size_t)DP * DP * (4 + 4 * e), (size_t)DP * DP, true); }
template <int DP>
__device__ __forceinline__ BMat big_ws_dU(double* ws, int ne, int v) { return bmat_at(ws + (size_t)DP * DP * (4 + 4 * ne + 2 * v), (size_t)DP * DP, false); }
template <int DP>
__device__ __forceinline__ BMat big_ws_d2U(double* ws, int nv, int ne, int v, int e) {
    return bmat_at(ws + (size_t)DP * DP * (4 + 4 * ne + 2 * nv + 2 * (e * nv + v)), (size_t)DP * DP, false);
}

// Z = X (x) Y [+ a1 A1 + a2 A2 + a3 A3 + gamma I], the exact jet product with a fused linear-combination epilogue (slot-wise;
// the identity only in the value slot).  `final_ws` != nullptr: last stage -- only the slots the sweeps need are formed and they
// go to the step workspace instead of scratch.  xz / yz: the mixed slots of X / Y are identically zero (B itself, when no error
// term depends on a control), so the products with them are skipped.
template <int DP>
__device__ __noinline__ void jet_product(double* sm, double* Z, double* X, double* Y, double* A1, double* A2, double* A3, double a1,
                                         double a2, double a3, double gamma, int nv, int ne, int d, double* final_ws, bool xz, bool yz) {
    BigGemm<DP> G;
    const int nf = nv + ne, nslots = big_nslots(nv, ne);
    for (int s = 0; s < nslots; ++s) {
        BMat dst;
        if (final_ws) {
            if (s == 0) dst = big_ws_U<DP>(final_ws);
            else if (s <= nf) dst = (s - 1 < nv) ? big_ws_dU<DP>(final_ws, ne, s - 1) : big_ws_D<DP>(final_ws, s - 1 - nv);
            else if (s <= 2 * nf) continue;                              // first differences at eps2: intermediates only
            else { const int q = s - 1 - 2 * nf; dst = big_ws_d2U<DP>(final_ws, nv, ne, q % nv, q / nv); }
        } else {
            dst = jet_slot<DP>(Z, s);
        }
        const BMat X0 = jet_slot<DP>(X, 0), Y0 = jet_slot<DP>(Y, 0);
        G.zero();
        if (s == 0) {
            G.template mac<1>(sm, X0, BOP_N, &Y0, BOP_N);
        } else if (s <= 2 * nf) {
            const BMat Xf = jet_slot<DP>(X, s), Yf = jet_slot<DP>(Y, s);
            const BMat ys[2] = {Y0, Yf};
            G.template mac<2>(sm, Xf, BOP_N, ys, BOP_N);
            G.template mac<1>(sm, X0, BOP_N, &Yf, BOP_N);
        } else {
            const int q = s - 1 - 2 * nf, v = q % nv, e = q / nv;
            const int sa = 1 + nf + v, sb = 1 + nf + nv + e;
            const BMat Xa = jet_slot<DP>(X, sa), Xb = jet_slot<DP>(X, sb), Xab = jet_slot<DP>(X, s);
            const BMat Ya = jet_slot<DP>(Y, sa), Yb = jet_slot<DP>(Y, sb), Yab = jet_slot<DP>(Y, s);
            if (!xz) { const BMat y4[4] = {Y0, Ya, Yb, Yab}; G.template mac<4>(sm, Xab, BOP_N, y4, BOP_N); }
            if (yz) {
                G.template mac<1>(sm, Xa, BOP_N, &Yb, BOP_N);
                G.template mac<1>(sm, Xb, BOP_N, &Ya, BOP_N);
            } else {
                const BMat y2a[2] = {Yb, Yab};
                G.template mac<2>(sm, Xa, BOP_N, y2a, BOP_N);
                const BMat y2b[2] = {Ya, Yab};
                G.template mac<2>(sm, Xb, BOP_N, y2b, BOP_N);
                G.template mac<1>(sm, X0, BOP_N, &Yab, BOP_N);
            }
        }
        BMat adds[3]; double bc[3]; int na = 0;
        if (A1) { adds[na] = jet_slot<DP>(A1, s); bc[na++] = a1; }
        if (A2) { adds[na] = jet_slot<DP>(A2, s); bc[na++] = a2; }
        if (A3) { adds[na] = jet_slot<DP>(A3, s); bc[na++] = a3; }
        G.store(dst, 1.0, adds, bc, na, s == 0 ? gamma : 0.0, d);
    }
    __syncthreads();
}

// Transposed planes of Z = c0 I + c1 J1 + c2 J2 + c3 J3 + c4 J4, slot-wise (identity only in the value slot).  Z is only ever
// used as a right operand (which reads reT / imT), so its canonical planes are not formed.
template <int DP>
__device__ __noinline__ void jet_lincomb_T(double* Z, double c0, double c1, double* J1, double c2, double* J2, double c3, double* J3,
                                           double c4, double* J4, int nslots, int d, int nt) {
    const size_t plane = (size_t)DP * DP, half = 2 * plane;
    for (size_t it = threadIdx.x; it < half * nslots; it += nt) {
        const size_t s = it / half, e2 = it % half;
        const size_t idx = s * 4 * plane + 2 * plane + e2;            // planes 2, 3 of slot s
        double v = c1 * J1[idx] + c2 * J2[idx] + c3 * J3[idx] + c4 * J4[idx];
        if (s == 0 && e2 < plane) {
            const int r = (int)(e2 / DP), c = (int)(e2 % DP);
            if (r == c && r < d) v += c0;
        }
        Z[idx] = v;
    }
    __syncthreads();
}

// ---- k_big_steps: one persistent CTA per (pulse, step) task ---------------------------------------------------------------
#ifndef RG_BIG16_CTAS
#define RG_BIG16_CTAS 8       // d <= 16: 64-thread CTAs; measured on B200 (d16 workload): uncapped (164 registers, 6 CTAs/SM) 30.2 ms,
#endif                        // 8 CTAs/SM (128 registers, no spills) 28.6 ms, 9-10 CTAs/SM 35.4 ms, 12 CTAs/SM (80 registers, spills) 37.8 ms
#ifndef RG_BIG32_CTAS
#define RG_BIG32_CTAS 4       // d <= 32: 128-thread CTAs; measured (d32 workload): uncapped (190 registers, 2 CTAs/SM) 200 ms, 3 CTAs/SM 164 ms,
#endif                        // 4 CTAs/SM (128 registers, no spills) 161 ms = 0.57 of DMMA peak
#ifndef RG_BIG48_CTAS
#define RG_BIG48_CTAS 2
#endif
template <int DP>
__global__ void __launch_bounds__(BigGemm<DP>::NT, DP <= 16 ? RG_BIG16_CTAS : (DP <= 32 ? RG_BIG32_CTAS : (DP <= 48 ? RG_BIG48_CTAS : 2)))
k_big_steps(const DevProblem P, const BigData Bd, const double* __restrict__ X, int B, double* __restrict__ ws, int* __restrict__ status) {
    constexpr int NT = BigGemm<DP>::NT;
    extern __shared__ double smd[];
    double* sm = smd;                                               // gemm chunks
    double* red = smd + BigGemm<DP>::SMEM_DOUBLES;                  // 64 doubles
    cplx* coef = reinterpret_cast<cplx*>(red + 64);                 // [nslots][nterms]
    __shared__ int s_sq;
    const int nv = P.nvar, ne = P.e, nf = nv + ne, nslots = Bd.nslots, nt = P.nterms, d = P.d;
    const size_t plane = (size_t)DP * DP, jet = (size_t)nslots * 4 * plane;
    double* scr = Bd.scratch + (size_t)blockIdx.x * Bd.scratch_per_cta;
    double* JB = scr; double* J2 = scr + jet; double* J3 = scr + 2 * jet; double* J4 = scr + 3 * jet; double* JR = scr + 4 * jet;
    double* JT = scr + 5 * jet;
    const long long tasks = (long long)B * P.N;
    for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const int b = (int)(task / P.N), k = (int)(task % P.N);
        const double* xp = X + (size_t)b * P.nx;
        double xadd[RG_MAX_ADD], xk[RG_MAX_MAIN];
        for (int j = 0; j < P.a; ++j) xadd[j] = xp[(size_t)P.p * P.N + j];
        for (int i = 0; i < P.p; ++i) xk[i] = xp[(size_t)k * P.p + i];
        __syncthreads();
        // ---- coefficients of every (slot, term), times -i dt (src/UnitaryCalculations.jl:45-97 perturbation pattern)
        for (int it = threadIdx.x; it < nslots * nt; it += NT) {
            const int s = it / nt, t = it % nt;
            const DevTerm& tm = P.terms[t];
            const bool isH0 = tm.owner == RG_OWNER_H0;
            int kind = 0, v = -1, e = -1;            // kind 0 value, 1 var, 2 err, 3 mixed
            double ee = P.eps;
            if (s == 0) kind = 0;
            else if (s <= 2 * nf) { const int f = (s - 1) % nf; ee = (s <= nf) ? P.eps : P.eps2; if (f < nv) { kind = 1; v = f; } else { kind = 2; e = f - nv; } }
            else { const int q = s - 1 - 2 * nf; kind = 3; v = q % nv; e = q / nv; ee = P.eps2; }
            cplx out = cmk(0.0, 0.0);
            const bool use = (kind == 0 || kind == 1) ? isH0 : (tm.owner == e);
            if (use) {
                int sp_ = RG_S_NONE, ix = 0; double h = 0.0;
                if (kind == 1 || kind == 3) {
                    sp_ = P.var_space[v]; ix = P.var_index[v];
                    const double val = (sp_ == RG_S_MAIN) ? xk[ix] : xadd[ix];
                    h = __dsub_rn(__dadd_rn(val, ee), val);
                }
                EvalCtx ec{xk, xadd, (kind >= 2) ? ee : 0.0, P.table, P.N, k};
                cplx base, del;
                term_coef(tm, ec, sp_, ix, h, base, del);
                const cplx c = (kind == 0 || kind == 2) ? base : del;
                out = cmk(c.y * P.dt, -c.x * P.dt);
            }
            coef[it] = out;
        }
        __syncthreads();
        // ---- A = sum_t coef[0][t] M_t (unscaled) into the value slot of JB, its 1-norm, number of squarings
        {
            BMat A0 = jet_slot<DP>(JB, 0);
            double rowmax = 0.0;
            for (int idx = threadIdx.x; idx < DP * DP; idx += NT) {
                double vr = 0.0, vi = 0.0;
                for (int t = 0; t < nt; ++t) {
                    const cplx c = coef[t];
                    if (c.x == 0.0 && c.y == 0.0) continue;
                    const double mr = Bd.termM[(size_t)t * 2 * plane + idx], mi = Bd.termM[(size_t)t * 2 * plane + plane + idx];
                    vr += c.x * mr - c.y * mi; vi += c.x * mi + c.y * mr;
                }
                A0.re[idx] = vr; A0.im[idx] = vi;
            }
            __syncthreads();
            for (int r = threadIdx.x; r < DP; r += NT) {            // A is skew-Hermitian: max row sum = 1-norm
                double sum = 0.0;
                for (int c = 0; c < DP; ++c) sum += sqrt(A0.re[(size_t)r * DP + c] * A0.re[(size_t)r * DP + c] + A0.im[(size_t)r * DP + c] * A0.im[(size_t)r * DP + c]);
                rowmax = fmax(rowmax, sum);
            }
            cplx mx = cmk(rowmax, 0.0);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) mx.x = fmax(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, off));
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx.x;
            __syncthreads();
            if (threadIdx.x == 0) {
                double nrm = 0.0;
                for (int w = 0; w < NT / 32; ++w) nrm = fmax(nrm, red[w]);
                nrm = nrm * 1.001 + 2.0 * P.eps2 * P.dt;
                int sq = 0;
                if (nrm > 0.31) sq = (int)ceil(log2(nrm / 0.31));
                if (sq > RG_MAX_SQUARINGS) { sq = RG_MAX_SQUARINGS; atomicOr(status, 1); }
                s_sq = sq;
            }
            __syncthreads();
        }
        const int sq = s_sq;
        const double sc = scalbn(1.0, -sq);
        // ---- jets of B = A / 2^s: every slot is a linear combination of the dense term matrices
        for (size_t it = threadIdx.x; it < (size_t)nslots * plane; it += NT) {
            const int s = (int)(it / plane); const size_t idx = it % plane;
            double vr = 0.0, vi = 0.0;
            for (int t = 0; t < nt; ++t) {
                const cplx c = coef[s * nt + t];
                if (c.x == 0.0 && c.y == 0.0) continue;
                const double mr = Bd.termM[(size_t)t * 2 * plane + idx], mi = Bd.termM[(size_t)t * 2 * plane + plane + idx];
                vr += c.x * mr - c.y * mi; vi += c.x * mi + c.y * mr;
            }
            vr *= sc; vi *= sc;
            const int r = (int)(idx / DP), c = (int)(idx % DP);
            double* base = JB + (size_t)s * 4 * plane;
            base[idx] = vr; base[plane + idx] = vi;
            base[2 * plane + (size_t)c * DP + r] = vr; base[3 * plane + (size_t)c * DP + r] = vi;
        }
        __syncthreads();
        // ---- degree-12 Taylor polynomial, Paterson-Stockmeyer with B^2, B^3, B^4 (5 jet products), then s squarings
        double* wsk = ws + (size_t)task * big_ws_step_doubles(DP, nv, ne);
        const bool bz = P.mixed_zero != 0;
        jet_product<DP>(sm, J2, JB, JB, nullptr, nullptr, nullptr, 0, 0, 0, 0.0, nv, ne, d, nullptr, bz, bz);
        jet_product<DP>(sm, J3, J2, JB, nullptr, nullptr, nullptr, 0, 0, 0, 0.0, nv, ne, d, nullptr, false, bz);
        jet_product<DP>(sm, J4, J2, J2, nullptr, nullptr, nullptr, 0, 0, 0, 0.0, nv, ne, d, nullptr, false, false);
        // R2 = c8 I + c9 B + c10 B2 + c11 B3 + c12 B4 (right operand only)
        jet_lincomb_T<DP>(JR, 1.0 / 40320, 1.0 / 362880, JB, 1.0 / 3628800, J2, 1.0 / 39916800, J3, 1.0 / 479001600, J4, nslots, d, NT);
        // R1 = P1 + B4 R2,  P1 = c4 I + c5 B + c6 B2 + c7 B3 fused into the epilogue
        jet_product<DP>(sm, JT, J4, JR, JB, J2, J3, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 24, nv, ne, d, nullptr, false, false);
        // T = P0 + B4 R1,  P0 = I + B + B2/2 + B3/6
        jet_product<DP>(sm, JR, J4, JT, JB, J2, J3, 1.0, 0.5, 1.0 / 6, 1.0, nv, ne, d, sq == 0 ? wsk : nullptr, false, false);
        double* cur = JR; double* nxt = JT;
        for (int q2 = 0; q2 < sq; ++q2) {
            jet_product<DP>(sm, nxt, cur, cur, nullptr, nullptr, nullptr, 0, 0, 0, 0.0, nv, ne, d, q2 == sq - 1 ? wsk : nullptr, false, false);
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
}
