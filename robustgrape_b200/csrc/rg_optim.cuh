// rg_optim.cuh -- the optimiser side of the hot path, kept on the device:
//   k_regularize : the enumerated regularisation terms of calculate_common! (src/FidelityCalculations.jl:186-195) added to
//                  [cost | grad] on the device: regularization_cost (src/Regularization.jl:26-47), regularization_cost_phase
//                  (:111-115, i.e. :78-83 applied to cos and sin) and the sin^2-of-differences form of test/runtests.jl:9-45;
//   k_lbfgs_*    : batched L-BFGS (one independent optimiser per pulse): two-loop recursion, backtracking line search with the
//                  Armijo condition, curvature-safeguarded history update.  Every pulse runs its own state machine: one round =
//                  [new direction for the pulses whose last trial was accepted | next trial point for everyone | one batched
//                  evaluation | Armijo test], so a pulse never waits for another pulse's line search (in lockstep, with 8192 pulses
//                  some pulse always needed all 8 halvings: 7.7 evaluations per iteration instead of 1.3).  The iterate X,
//                  the gradient and the (s, y) history never leave HBM; the host only reads a 12-byte progress record per
//                  line-search round.  Role of Optim.optimize(...; method = LBFGS()) in src/FidelityCalculations.jl:199-217.
#pragma once
#include <cuda_runtime.h>
#include "rg_common.cuh"

// RG_REG_NONE / PLAIN / PHASE / SIN2 are declared in include/robustgrape_b200.h

// block-wide sum (blockDim.x <= 1024, multiple of 32); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

// first / second difference penalties of a sequence y_0..y_{n-1} and their gradient w.r.t. y_i:
//   reg1 = sum (y_{i+1} - y_i)^2 ,  reg2 = sum dd_i^2 , dd_i = y_{i+2} - 2 y_{i+1} + y_i
//   d reg1 / d y_i = 2 (y_i - y_{i-1}) [i > 0] - 2 (y_{i+1} - y_i) [i < n-1]     (src/Regularization.jl:34-36)
//   d reg2 / d y_i = 2 (dd_{i-2} - 2 dd_{i-1} + dd_i)  with dd_j = 0 outside 0..n-3   (:37-43)
// One tile of the sequence staged in shared memory (x and, for the phase form, cos x and sin x, each evaluated once per element:
// the first version called cos/sin from the stencils, ~30 trigonometric evaluations per element, and at 0.20 ms per call on C4 cost
// more than the fused evaluation kernel itself; staged it is one sincos per element).
#define RG_REG_TILE 1024
struct RegSeq {
    const double* y0; int t0, n;               // y0[j] = y(t0 - 2 + j)
    __device__ __forceinline__ double y(int i) const { return y0[i - t0 + 2]; }
    __device__ __forceinline__ double dd(int j) const { return (j < 0 || j > n - 3) ? 0.0 : y(j + 2) - 2.0 * y(j + 1) + y(j); }
};

// grid (B, p), block 128: cost[b] += c1 reg1 + c2 reg2 ; grad[b*nx + p*i + row] += c1 jac1_i + c2 jac2_i
static __global__ void __launch_bounds__(128)
k_regularize(int B, int nx, int p, int N, const int* __restrict__ kinds, const double* __restrict__ c1s, const double* __restrict__ c2s,
             const double* __restrict__ X, double* __restrict__ cost, double* __restrict__ grad) {
    __shared__ double red[4];
    __shared__ double sx[RG_REG_TILE + 4], sc[RG_REG_TILE + 4], ss[RG_REG_TILE + 4];
    const int b = blockIdx.x, row = blockIdx.y;
    const int kind = kinds[row];
    const double c1 = c1s[row], c2 = c2s[row];
    if (kind == RG_REG_NONE || (c1 == 0.0 && c2 == 0.0)) return;
    const double* x = X + (size_t)b * nx + row;
    double* g = grad + (size_t)b * nx + row;
    double r1 = 0.0, r2 = 0.0;
    const int n = N;
    for (int t0 = 0; t0 < n; t0 += RG_REG_TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < RG_REG_TILE + 4; j += blockDim.x) {
            const int i = t0 - 2 + j;
            if (i >= 0 && i < n) {
                const double v = x[(size_t)i * p];
                sx[j] = v;
                if (kind == RG_REG_PHASE) rg_sincos(v, ss[j], sc[j]);
            }
        }
        __syncthreads();
        const int t1 = min(n, t0 + RG_REG_TILE);
        if (kind == RG_REG_PLAIN || kind == RG_REG_PHASE) {
            const int nseq = kind == RG_REG_PLAIN ? 1 : 2;
            for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
                double gi = 0.0;
                for (int q = 0; q < nseq; ++q) {
                    const int sk = kind == RG_REG_PLAIN ? 0 : 1 + q;          // 0 y = x, 1 y = cos x, 2 y = sin x
                    RegSeq s{sk == 0 ? sx : (sk == 1 ? sc : ss), t0, n};
                    const double yi = s.y(i);
                    double j1 = 0.0;
                    if (i > 0) j1 += 2.0 * (yi - s.y(i - 1));
                    if (i < n - 1) { const double d = s.y(i + 1) - yi; j1 -= 2.0 * d; r1 += d * d; }
                    const double ddi = s.dd(i);
                    if (i <= n - 3) r2 += ddi * ddi;
                    const double j2 = 2.0 * (s.dd(i - 2) - 2.0 * s.dd(i - 1) + ddi);
                    // chain rule of regularization_cost(x, f, df) (:78-83): df = 1 | -sin x | cos x
                    const double df = sk == 0 ? 1.0 : (sk == 1 ? -ss[i - t0 + 2] : sc[i - t0 + 2]);
                    gi += df * (c1 * j1 + c2 * j2);
                }
                g[(size_t)i * p] += gi;
            }
        } else {
            // test/runtests.jl:9-45 (shadows the exported function in the reference's tests), restated index for index (1-based i):
            //   reg1 = sum sin^2(dx_i / 2), reg2 = sum sin^2(ddx_i / 2)
            //   jac1[i] = -0.5 sin(dx_i) [i < n-1] + 0.5 sin(dx_{i-1}) [i > 1], for i = 1..n-1 only (jac1[n] stays 0)
            //   jac2[i] = -0.5 sin(ddx_i) [i < n-2] + sin(ddx_{i-1}) [1 < i < n-1] - 0.5 sin(ddx_{i-2}) [i > 2]
            // sin(dx_i), sin(ddx_i) for i in [t0 - 1, t0 + RG_REG_TILE] staged once each (sc, ss are free in this form): one sincos of
            // the half angle gives sin^2(d/2) for the cost and sin d = 2 sin(d/2) cos(d/2) for the Jacobian (7 sin calls per element before)
            auto xv = [&](int i0) { return sx[i0 - t0 + 2]; };                                              // x at 0-based index
            for (int j = threadIdx.x; j < RG_REG_TILE + 2; j += blockDim.x) {
                const int i = t0 - 1 + j;                                                                   // 1-based difference index
                const bool own = (i >= t0 + 1 && i <= t1);                                                  // element i0 = i - 1 is in this tile
                if (i >= 1 && i <= n - 1) {
                    double sh, ch; rg_sincos(0.5 * (xv(i) - xv(i - 1)), sh, ch);
                    sc[j] = 2.0 * sh * ch;
                    if (own) r1 += sh * sh;
                }
                if (i >= 1 && i <= n - 2) {
                    double sh, ch; rg_sincos(0.5 * (xv(i + 1) - 2.0 * xv(i) + xv(i - 1)), sh, ch);
                    ss[j] = 2.0 * sh * ch;
                    if (own) r2 += sh * sh;
                }
            }
            __syncthreads();
            auto sdx = [&](int i1) { return sc[i1 - t0 + 1]; };                                            // sin(diff_x[i1])
            auto sddx = [&](int i1) { return ss[i1 - t0 + 1]; };                                           // sin(diff_diff_x[i1])
            for (int i0 = t0 + threadIdx.x; i0 < t1; i0 += blockDim.x) {
                const int i = i0 + 1;
                double j1 = 0.0, j2 = 0.0;
                if (i <= n - 1) {
                    if (i < n - 1) j1 -= 0.5 * sdx(i);
                    if (i > 1) j1 += 0.5 * sdx(i - 1);
                }
                if (i < n - 2) j2 -= 0.5 * sddx(i);
                if (i > 1 && i < n - 1) j2 += sddx(i - 1);
                if (i > 2) j2 -= 0.5 * sddx(i - 2);
                g[(size_t)i0 * p] += c1 * j1 + c2 * j2;
            }
        }
    }
    const double t1s = block_sum(r1, red), t2s = block_sum(r2, red);
    if (threadIdx.x == 0) atomicAdd(cost + b, c1 * t1s + c2 * t2s);
}

// ---- batched L-BFGS ------------------------------------------------------------------------------------------------
struct LbfgsState {
    int B, nx, m;
    double* X;        // [B][nx] current iterate
    double* G;        // [B][nx] gradient at X
    double* F;        // [B]     cost at X
    double* Xt;       // [B][nx] trial point
    double* Gt;       // [B][nx] gradient at the trial point
    double* Ft;       // [B]     cost at the trial point
    double* Dir;      // [B][nx] search direction
    double* S;        // [B][m][nx]
    double* Y;        // [B][m][nx]
    double* rho;      // [B][m]
    double* alpha;    // [B]     current step length
    double* gd;       // [B]     g . d at X
    int* hist;        // [B]     number of stored pairs (<= m)
    int* head;        // [B]     ring position of the next pair
    int* active;      // [B]     1 while the pulse still searches along Dir in this iteration
    int* done;        // [B]     1 once converged (gradient norm below g_tol)
    int* iters;       // [B]     iterations taken
    int* lsr;         // [B]     rejected trial steps of the current line search
    int max_iters;    //         iterations allowed per pulse
    int* counters;    // [0] pulses still active in the line search, [1] pulses not finished, [2] line-search failures, [3] max iterations
};

// one CTA per pulse: d = -H g by the two-loop recursion; gd = g . d; initial step of the iteration
static __global__ void __launch_bounds__(256)
k_lbfgs_direction(LbfgsState st, double g_tol) {
    __shared__ double red[8];
    __shared__ double al[32];
    const int b = blockIdx.x;
    const int nx = st.nx, m = st.m;
    const double* g = st.G + (size_t)b * nx;
    double* d = st.Dir + (size_t)b * nx;
    if (st.done[b]) { if (threadIdx.x == 0) st.active[b] = 0; return; }
    if (st.active[b]) return;                              // still inside a line search: keeps its direction
    // convergence test on the infinity norm of the gradient (Optim's g_tol)
    double gmax = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) gmax = fmax(gmax, fabs(g[i]));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gmax;
    __syncthreads();
    gmax = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) gmax = fmax(gmax, red[w]);
    if (gmax <= g_tol) { if (threadIdx.x == 0) { st.done[b] = 1; st.active[b] = 0; } return; }
    for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] = g[i];
    __syncthreads();
    const int h = st.hist[b], head = st.head[b];
    for (int j = 0; j < h; ++j) {                          // newest to oldest
        const int slot = (head - 1 - j + 2 * m) % m;
        const double* s = st.S + ((size_t)b * m + slot) * nx;
        const double* y = st.Y + ((size_t)b * m + slot) * nx;
        double dot = 0.0;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) dot += s[i] * d[i];
        const double a = st.rho[(size_t)b * m + slot] * block_sum(dot, red);
        if (threadIdx.x == 0) al[j] = a;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] -= a * y[i];
        __syncthreads();
    }
    if (h > 0) {                                           // H0 = (s.y / y.y) I with the newest pair
        const int slot = (head - 1 + m) % m;
        const double* y = st.Y + ((size_t)b * m + slot) * nx;
        double yy = 0.0;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) yy += y[i] * y[i];
        yy = block_sum(yy, red);
        const double gamma = 1.0 / (st.rho[(size_t)b * m + slot] * yy);
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] *= gamma;
        __syncthreads();
    }
    for (int j = h - 1; j >= 0; --j) {                     // oldest to newest
        const int slot = (head - 1 - j + 2 * m) % m;
        const double* s = st.S + ((size_t)b * m + slot) * nx;
        const double* y = st.Y + ((size_t)b * m + slot) * nx;
        double dot = 0.0;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) dot += y[i] * d[i];
        const double beta = st.rho[(size_t)b * m + slot] * block_sum(dot, red);
        const double a = al[j];
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] += (a - beta) * s[i];
        __syncthreads();
    }
    double gd = 0.0, gg = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { d[i] = -d[i]; gd += g[i] * d[i]; gg += g[i] * g[i]; }
    gd = block_sum(gd, red); gg = block_sum(gg, red);
    if (!(gd < 0.0)) {                                     // not a descent direction (stale curvature): restart with steepest descent
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] = -g[i];
        gd = -gg;
        if (threadIdx.x == 0) { st.hist[b] = 0; st.head[b] = 0; }
    }
    if (threadIdx.x == 0) {
        st.gd[b] = gd;
        st.alpha[b] = (h == 0 || !(gd < 0.0)) ? fmin(1.0, 1.0 / sqrt(gg)) : 1.0;
        st.active[b] = 1;
        st.lsr[b] = 0;
    }
}

// trial point of every active pulse; inactive pulses keep Xt = X (their trial evaluation is ignored)
static __global__ void k_lbfgs_trial(LbfgsState st) {
    const size_t n = (size_t)st.B * st.nx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / st.nx);
        st.Xt[i] = st.active[b] ? st.X[i] + st.alpha[b] * st.Dir[i] : st.X[i];
    }
}

// one CTA per pulse: Armijo test f(x + a d) <= f(x) + c1 a g.d.  Accept: store the (s, y) pair, move X, G, F.  Reject: shrink a
// (quadratic interpolation clamped to [0.1 a, 0.5 a]); after max_ls rejections the pulse takes no step in this iteration.
static __global__ void __launch_bounds__(256)
k_lbfgs_check(LbfgsState st, int max_ls) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    if (!st.active[b]) return;
    const int nx = st.nx, m = st.m;
    const double f0 = st.F[b], ft = st.Ft[b], a = st.alpha[b], gd = st.gd[b];
    const bool ok = isfinite(ft) && ft <= f0 + 1e-4 * a * gd + 1e-15 * fabs(f0);
    if (ok) {
        const int slot = st.head[b];
        double* s = st.S + ((size_t)b * m + slot) * nx;
        double* y = st.Y + ((size_t)b * m + slot) * nx;
        double* x = st.X + (size_t)b * nx; double* g = st.G + (size_t)b * nx;
        const double* xt = st.Xt + (size_t)b * nx; const double* gt = st.Gt + (size_t)b * nx;
        double sy = 0.0;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) {
            const double si = xt[i] - x[i], yi = gt[i] - g[i];
            s[i] = si; y[i] = yi; sy += si * yi;
            x[i] = xt[i]; g[i] = gt[i];
        }
        sy = block_sum(sy, red);
        if (threadIdx.x == 0) {
            st.F[b] = ft;
            if (sy > 1e-300) {                                 // curvature condition: keep the pair
                st.rho[(size_t)b * m + slot] = 1.0 / sy;
                st.head[b] = (slot + 1) % m;
                st.hist[b] = min(st.hist[b] + 1, m);
            }
            st.active[b] = 0;
            const int it = st.iters[b] + 1;
            st.iters[b] = it;
            if (it >= st.max_iters) st.done[b] = 1;            // its budget is spent: no further direction
        }
    } else if (threadIdx.x == 0) {
        // no acceptable step along a descent direction within max_ls halvings: the pulse sits at the resolution of the cost
        // (rounding) -- it is finished
        const int r = st.lsr[b] + 1;
        st.lsr[b] = r;
        if (r >= max_ls) { st.active[b] = 0; st.done[b] = 1; atomicAdd(st.counters + 2, 1); }
        else {
            double an = 0.5 * a;
            if (isfinite(ft)) {
                const double q = -gd * a * a / (2.0 * (ft - f0 - gd * a));      // minimiser of the interpolating parabola
                if (q > 0.1 * a && q < 0.5 * a) an = q;
            }
            st.alpha[b] = an;
        }
    }
}

static __global__ void k_lbfgs_count(LbfgsState st) {
    int act = 0, open = 0, mit = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < st.B; b += gridDim.x * blockDim.x) {
        act += st.active[b]; open += 1 - st.done[b]; mit = max(mit, st.iters[b]);
    }
    if (act) atomicAdd(st.counters + 0, act);
    if (open) atomicAdd(st.counters + 1, open);
    atomicMax(st.counters + 3, mit);
}
