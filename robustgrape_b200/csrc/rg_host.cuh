// rg_host.cuh -- host-side structs and the per-dimension launch templates.  Every rg_dims_*.cu translation
// unit instantiates these for a few Hilbert-space dimensions (compiled in parallel); rg_api.cu holds the C ABI.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

enum { RG_K_STEPS = 0, RG_K_STEPS_SO = 1, RG_K_SCAN = 2, RG_K_GRAD = 3, RG_K_GRAD_ERR = 4, RG_K_EPILOGUE = 5, RG_K_ANALYSIS = 6, RG_K_AGG = 7, RG_NKERNELS = 8 };
#include "rg_smalld.cuh"
#include "rg_steps_t.cuh"
#include "rg_analysis.cuh"

struct rg_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int* d_status = nullptr;
    int* h_status = nullptr;    // pinned
    int sm_count = 148;
    size_t ws_limit = (size_t)64 << 30;
    // optional per-kernel timing (CUDA events on the launch stream), see rg_ctx_set_timing
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the pipelined host entry points
    int host_slabs = 8, host_slabs_forced = 0;        // upper bound on the number of slabs; RG_HOST_SLABS forces an exact count
    int host_slab_min = 256;                          // RG_HOST_SLAB_MIN (smallest slab, pulses)
    std::vector<cudaEvent_t> slab_events;             // reused by the pipelined host entry points
#define RG_PEER_STREAMS 8
    cudaStream_t s_peer[RG_PEER_STREAMS] = {};        // side streams of rg_gather_to_peers (created on first use)
    cudaEvent_t ev_src = nullptr, ev_peer_join[RG_PEER_STREAMS] = {}, ev_gather[2] = {nullptr, nullptr};
    int gather_split = 0;                             // RG_GATHER_SPLIT: pieces per peer copy (0 = automatic)
    bool gather_pending[2] = {false, false};
    bool timing = false;
    struct Span { int kernel; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    double kern_ms[RG_NKERNELS] = {0};
    int64_t kern_n[RG_NKERNELS] = {0};
};

struct KTimer {
    rg_ctx* c; int k; cudaEvent_t e0 = nullptr, e1 = nullptr;
    KTimer(rg_ctx* c_, int k_) : c(c_), k(k_) {
        if (c->timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    }
    ~KTimer() {
        c->launches++;
        if (c->timing) { cudaEventRecord(e1, c->stream); c->spans.push_back({k, e0, e1}); }
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct rg_problem {
    rg_ctx* ctx = nullptr;
    DevProblem dp{};
    std::vector<void*> owned;          // device allocations of the descriptor
    int any_add_dep = 0;
    int chunk_override = 0;
    int force_dense = 0;      // RG_DENSE=1: treat H as dense (no structural-zero skipping)
    int force_group = 0;      // RG_GROUP=1: force the group-per-chunk k_steps kernel
    int fused_agg = 0;            // RG_FUSED_AGG=1: form the e = 0 chunk aggregates inside k_steps_t (measured slower: the tree
                                  // product issues 5 x 64 DFMA warp-instructions with most lanes idle)
    int force_group_sweeps = 0;   // RG_GROUP_SWEEPS=1: group (shared-memory) versions of k_chunk_agg / k_grad
    int force_sequential_analysis = 0;   // RG_SEQ_ANALYSIS=1: time-sequential interaction-operator kernel
    int force_ws = 0;             // RG_WS=1: step-matrix workspace path even where the workspace-free block-2 path applies
    int b2_agg_ctas = 0, b2_grad_ctas = 0;   // resident CTAs/SM of the block-2 sweeps (occupancy query, cached)
    int force_b2 = 0;             // RG_B2=1: three-kernel block-2 path even where the one-launch fused quaternion path applies
    // dense path (rg_big.cuh, ndim > 10): dense planar term matrices, padded dimension, occupancy of k_big_steps
    DevBuf big_termM, big_tgtM;
    int big_dp = 0, big_ctas = 0;
    int is_hstack = 0;            // closure problem: Hamiltonians arrive as host-evaluated stacks (rg_*_from_hstack)
    DevBuf dHs, dTs;
    DevBuf regbuf, lbfgs, dXopt;   // regularisation table, L-BFGS state, staged iterate (rg_api_optim.inl)
    std::vector<unsigned char> reg_host;   // host copy of the table in regbuf
    int diag_alg = 0;             // projector and target are diagonal: elementwise fidelity algebra in the fused kernel
    int force_dense_alg = 0;      // RG_DENSE_ALG=1: dense fidelity algebra even then (A/B and tests)
    int wpp_override = 0;         // RG_WPP=1|2|4: warps per pulse of the fused quaternion kernel
    // fused quaternion kernel, per warps-per-pulse choice (1, 2, 4): resident CTAs/SM by role, shared memory, controls staged or not
    int fq_ready = 0, fq_occ[2][2][3] = {}, fq_xs[2][3] = {};      // [variant: default | gradient staged out][role][wpp choice]
    size_t fq_smem[2][3] = {};
    PeerOut peer_out{};           // set by rg_cost_and_grad_batch_dev_scatter around one evaluation; peer_out_done: the kernel took it
    int peer_out_done = 0;
    int no_accum = 0;             // RG_NO_ACCUM=1: separate dF2/dx buffer and k_cost_grad epilogue instead of in-kernel cost/gradient assembly
    int stage_xs = 0;             // RG_XS=1: stage the controls (and the gradient) through shared-memory rows.  Measured on B200 and
                                  // rejected as the default: C4 0.240 ms staged vs 0.183 ms direct (the lane-strided global loads hit L1
                                  // three times out of four, and staging adds two CTA-wide barriers and a serial load/store phase)
    int pc_ready = 0;             // phase-only class: constants evaluated (k_fqc_consts); dp.pc is cleared if they are out of range
    TriPlanDev tri{};         // upper-triangle assembly plan (Hermitian fast path)
    int tri_ok = 0;
    double tri_density = 1.0;
    unsigned tri_union = 0;   // union of all structural masks
    int costate_in_pattern = 0;   // projector and target lie inside the closure pattern
    // workspaces
    DevBuf ws, Qb, Wlb, Cb, Wb, Gb, G1b, H1b, F, F2, addT, addS, F2dx, Fdx, coeff, dX, dOut, dOut2, dO, dFreq, dM;
    int has_target = 0;
};

#define RG_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        char _b[512];                                             \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                    \
        (ctx)->err = _b;                                          \
        return (code);                                            \
    } while (0)

#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            RG_FAIL(ctx, RG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                      \
    } while (0)


// ---- workspace-free block-2 path (rg_block2.cu): closed-form 2 x 2 block propagators recomputed inside the sweeps
int rg_b2_pattern(const rg_problem* pr);          // 0 = not eligible, else the instantiated pattern
int rg_b2_launch_agg(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX);
int rg_b2_launch_grad(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out0, double scale0);
int rg_b2_launch_grad_err(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out1);
void rg_b2_occupancy(const rg_problem* pr, int* agg_ctas, int* grad_ctas);
// one-launch fused quaternion path (rg_fusedq.cuh): block-2 patterns without diagonal terms
int rg_fq_pattern(const rg_problem* pr);
int rg_fq_launch(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
                 double scale0, double scale0T, int do_grad, const PeerOut* po = nullptr, const FQAccum* ac = nullptr);
int rg_fq_prepare(rg_problem* pr);
static inline bool rg_use_b2(const rg_problem* pr) {
    return !pr->force_group && !pr->force_dense && !pr->force_group_sweeps && !pr->fused_agg && rg_b2_pattern(pr) != 0;
}

static inline bool rg_use_fq(const rg_problem* pr) { return rg_use_b2(pr) && !pr->force_b2 && rg_fq_pattern(pr) != 0; }

// ---- per-dimension operations table (defined by RG_DEFINE_DIM in rg_dims_*.cu) --------------------------
struct Plan;
struct DimOps {
    int (*run_slab)(rg_problem*, int, const Plan&, const double*, int, const double*, double*, double*, double*, double*, bool);
    int (*materialize)(rg_problem*, const double*, cplx*, cplx*, cplx*, cplx*, cplx*, cplx*);
    int (*interaction)(rg_problem*, const double*, cplx*);
    int (*response)(rg_problem*, const double*, int, int, int, int, double*);
    int (*expectation)(rg_problem*, double*);
};
const DimOps* rg_dim_ops(int d);

// ------------------------------------------------------------------------------------------
struct Plan { int L, nc, slab; };

static inline Plan make_plan(rg_problem* pr, int B) {
    const DevProblem& P = pr->dp;
    Plan pl;
    if (rg_use_b2(pr)) {
        // Workspace-free path: nothing of size N is stored, so the whole batch is one slab.  The chunk length trades the
        // tail of the last wave of the two thread-per-(pulse, chunk) sweeps against the sequential chunk scan:
        //   cost(L) = L * (ceil(items / cap_grad) + 0.5 * ceil(items / cap_agg)) + 2 * nc * ceil(B / scan_cap)
        // in units of one sweep step; capacities come from occupancy queries of the kernels that will run, the aggregate
        // sweep does about half the work of the gradient sweep per step, and one scan step costs about two sweep steps.
        if (!pr->b2_grad_ctas) rg_b2_occupancy(pr, &pr->b2_agg_ctas, &pr->b2_grad_ctas);
        const double cap_g = (double)pr->ctx->sm_count * pr->b2_grad_ctas * 128, cap_a = (double)pr->ctx->sm_count * pr->b2_agg_ctas * 128;
        const double scan_cap = (double)pr->ctx->sm_count * 5 * (32 / P.d);
        pl.slab = B;
        int L = std::min(P.N, 32);
        double best = 1e300;
        for (int cand = 4; cand <= 128 && cand <= P.N; ++cand) {
            const int ncc = (P.N + cand - 1) / cand;
            if (cand > 4 && (P.N + cand - 2) / (cand - 1) == ncc) continue;          // same chunk count as a shorter chunk
            const double items = (double)B * ncc;
            const double cost = cand * (std::ceil(items / cap_g) + 0.5 * std::ceil(items / cap_a)) + 2.0 * ncc * std::ceil(B / scan_cap);
            if (cost < best) { best = cost; L = cand; }
        }
        if (pr->chunk_override > 0) L = pr->chunk_override;
        L = std::max(1, std::min(L, P.N));
        pl.L = L;
        pl.nc = (P.N + L - 1) / L;
        return pl;
    }
    if (pr->big_dp) {
        // dense path: per-step workspace of planar matrices; chunks sized so that (pulse, chunk) items fill the GPU a few times
        const size_t per_pulse = (size_t)P.N * (size_t)pr->big_dp * pr->big_dp * (4 + 4 * P.e + 2 * P.nvar + 2 * P.nvar * P.e) * 8;
        pl.slab = (int)std::min<size_t>(std::max<size_t>(1, pr->ctx->ws_limit / std::max<size_t>(per_pulse, 1)), (size_t)B);
        const long long want = 2LL * pr->ctx->sm_count;
        int L = (int)std::max<long long>(1, std::min<long long>(32, (long long)P.N * pl.slab / want));
        if (pr->chunk_override > 0) L = pr->chunk_override;
        L = std::max(1, std::min(L, P.N));
        pl.L = L; pl.nc = (P.N + L - 1) / L;
        return pl;
    }
    const size_t per_pulse = (size_t)P.N * P.nstore * P.d * P.d * sizeof(cplx);
    size_t slab = std::max<size_t>(1, pr->ctx->ws_limit / std::max<size_t>(per_pulse, 1));
    pl.slab = (int)std::min<size_t>(slab, (size_t)B);
    // enough (pulse, chunk) work items for ~8 waves of resident groups
    const long long target_items = (long long)pr->ctx->sm_count * 72 * 8;
    long long want_nc = (target_items + pl.slab - 1) / pl.slab;
    // chunk length: long enough that the per-pulse sequential chunk scan (k_scan, latency bound) stays short,
    // short enough to expose (pulse, chunk) parallelism for small batches
    const long long lmin = (pl.slab >= 64) ? 16 : 4;
    int L = (int)std::max<long long>(lmin, std::min<long long>(32, P.N / std::max<long long>(1, want_nc)));
    if (pr->tri_ok && P.d <= 5 && !pr->force_group && !pr->force_group_sweeps) {
        // thread-per-chunk sweeps over the step-matrix workspace (structured d <= 5 problems outside the block-2 path):
        // longer chunks shorten the sequential k_scan; 2 (gradient) and 3 (aggregate) resident CTAs/SM of 128 threads
        if (pl.slab >= 2048) {
            const double per_wave_g = 2.0 * pr->ctx->sm_count * 128, per_wave_a = 3.0 * pr->ctx->sm_count * 128;
            double best = 1e300;
            for (int cand = 32; cand <= 96 && cand <= P.N; ++cand) {
                const int ncc = (P.N + cand - 1) / cand;
                if (cand > 32 && (P.N + cand - 2) / (cand - 1) == ncc) continue;      // same chunk count as a shorter chunk
                const double wg = (double)pl.slab * ncc / per_wave_g, wa = (double)pl.slab * ncc / per_wave_a;
                const double cost = 2.0 * std::ceil(wg) / wg + std::ceil(wa) / wa + 0.02 * ncc;
                if (cost < best) { best = cost; L = cand; }
            }
        } else if (pl.slab >= 512) L = 32;
    }
    if (pr->chunk_override > 0) L = pr->chunk_override;
    L = std::min(L, P.N);
    pl.L = L;
    pl.nc = (P.N + L - 1) / L;
    return pl;
}

// Opt in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute is a host-side call of a few microseconds; the
// largest size requested so far is cached per kernel so that steady-state launches skip it.
template <class K>
static int set_smem(rg_ctx* ctx, K kern, size_t bytes) {
    if (bytes > 227 * 1024) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "kernel needs %zu bytes of shared memory", bytes);
    // keyed by (kernel address, device): instantiations with the same signature share the type K
    struct Granted { const void* fn; int dev; size_t bytes; };
    static std::vector<Granted> granted;
    static std::mutex granted_mutex;                 // contexts on different host threads share this cache
    std::lock_guard<std::mutex> lock(granted_mutex);
    const void* fn = reinterpret_cast<const void*>(kern);
    for (auto& g : granted)
        if (g.fn == fn && g.dev == ctx->device) {
            if (bytes <= g.bytes) return RG_OK;
            CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            g.bytes = bytes;
            return RG_OK;
        }
    CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    granted.push_back({fn, ctx->device, bytes});
    return RG_OK;
}

// Structural patterns instantiated ahead of time (upper-triangle bit = k(k+1)/2 + i, i <= k).  A problem
// whose union mask fits one of them gets kernels with the zero positions removed at compile time; anything
// else runs the full (dense) instantiation.  (Arbitrary patterns: NVRTC specialisation, see DESIGN.md.)
//   PAT_M5_DRIVE: 5-level symmetric-blockaded Rydberg model (src/RydbergTools.jl:31-39), drive only: (1,3),(2,4)
//   PAT_M5_FULL : same plus the Rydberg detuning diagonal (3,3),(4,4) (frequency error / delta != 0)
enum { PAT_FULL = 0, PAT_M5_DRIVE = 1, PAT_M5_FULL = 2 };
template <int D, int PID> constexpr unsigned tri_mask_of() {
    return (D == 5 && PID == PAT_M5_DRIVE) ? ((1u << 7) | (1u << 12))
         : (D == 5 && PID == PAT_M5_FULL) ? ((1u << 7) | (1u << 12) | (1u << 9) | (1u << 14))
         : ((D * (D + 1) / 2 >= 32) ? ~0u : ((1u << (D * (D + 1) / 2)) - 1u));
}
template <int D, int PID> constexpr u64 cmask_of() {
    return (PID == PAT_FULL) ? full_cmask<D>() : closure_from_tri(D, tri_mask_of<D, PID>());
}
template <int D, int PID> constexpr u64 smask_of() {      // stored pattern: closure minus inert diagonals
    return (PID == PAT_FULL) ? full_cmask<D>() : stored_from_tri(D, tri_mask_of<D, PID>());
}

// Run the fused path for one slab of pulses already resident on the device.
//   mode 0: fidelity + derivatives  -> dF, dFdx (+1 scale), dF2, dF2dx
//   mode 1: cost + grad             -> dcost (in dF slot), dgrad (in dFdx slot)
template <int D, int PID>
static int run_slab(rg_problem* pr, int B, const Plan& pl, const double* dX, int mode, const double* d_coeff,
                    double* dF, double* dFdx, double* dF2, double* dF2dx, bool want_grad) {
    rg_ctx* ctx = pr->ctx;
    constexpr u64 CM = cmask_of<D, PID>();
    constexpr u64 CMS = smask_of<D, PID>();
    constexpr int WSM = Pat<D, CMS>::nnz;
    DevProblem P = pr->dp;
    P.wsm = WSM; P.cmask = CMS; P.wsB = B;
    constexpr int G = GroupInfo<D>::G;
    const int DD = D * D, ne = P.e, nc = pl.nc, L = pl.L;
    cudaStream_t st = ctx->stream;
    if (!pr->has_target) RG_FAIL(ctx, RG_ERR_INVALID, "problem has no target/projector: fidelity entry points unavailable");

    const size_t cb = sizeof(cplx);
    const bool b2 = rg_use_b2(pr);
    if (pr->ws.ensure(b2 ? 16 : (size_t)B * P.N * P.nstore * WSM * cb) || pr->Qb.ensure((size_t)B * nc * DD * cb) ||
        pr->Wlb.ensure(std::max<size_t>(16, (size_t)B * nc * ne * DD * cb)) || pr->Cb.ensure((size_t)B * nc * DD * cb) ||
        pr->Wb.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) || pr->Gb.ensure((size_t)B * nc * DD * cb) ||
        pr->G1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) ||
        pr->H1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) ||
        pr->F.ensure((size_t)B * 8) || pr->F2.ensure(std::max<size_t>(16, (size_t)B * ne * 8)) ||
        pr->addT.ensure(std::max<size_t>(16, (size_t)B * (1 + ne) * P.a * 8)) ||
        pr->addS.ensure(std::max<size_t>(16, pr->any_add_dep ? (size_t)B * (1 + ne) * P.a * P.N * 8 : 16)))
        RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed (B=%d)", B);
    double* iF = dF ? dF : pr->F.as<double>();
    double* iF2 = dF2 ? dF2 : pr->F2.as<double>();
    if (mode == 1) { iF = pr->F.as<double>(); iF2 = pr->F2.as<double>(); }
    double* iF2dx = dF2dx;
    if (want_grad && ne > 0 && (mode == 1 || !dF2dx)) {
        if (pr->F2dx.ensure((size_t)B * ne * P.nx * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed");
        iF2dx = pr->F2dx.as<double>();
    }
    double* iFdx = dFdx;
    if (want_grad && !iFdx) {
        if (pr->Fdx.ensure((size_t)B * P.nx * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed");
        iFdx = pr->Fdx.as<double>();
    }

    if (rg_use_fq(pr)) {
        // one launch per role: forward sweep, scan, fidelity algebra, backward sweep (rg_fusedq.cuh)
        const double DD1q = P.Dtr * (P.Dtr + 1.0);
        const double sgn = (mode == 1 && ne == 0) ? -1.0 : 1.0;
        const bool cost_fused = (mode == 1 && ne == 0);
        // fused evaluation + gather (rg_cost_and_grad_batch_dev_scatter): only where this launch writes the final cost and gradient
        const PeerOut* po = (cost_fused && pr->peer_out.n > 0) ? &pr->peer_out : nullptr;      // the launcher sets peer_out_done if it takes it
        if (mode == 1 && ne > 0 && want_grad && !pr->stage_xs && !pr->no_accum && rg_fq_pattern(pr) == 1) {
            // cost and gradient assembled by the launches themselves: the fidelity role leaves 1 - F and -dF/dx in the outputs, every
            // error source's launch (one after the other: they update the same entries) adds c_e F2^2 and 2 c_e F2 dF2/dx.  No dF2/dx
            // buffer, no epilogue kernel.  Measured on B200 (C4', 8192 x 1000, e = 1): 0.449 ms vs 0.487 ms with k_cost_grad; for the
            // 7-level model the read-modify-write in the three-block error sweep costs more than the epilogue saves (0.727 vs 0.700 ms),
            // so only the two-block pattern takes this path.
            int rc = rg_fq_launch(pr, P, B, dX, 0, dF, 1, iFdx, -P.inv_eps / DD1q, -1.0, 1);
            for (int e = 0; e < ne && !rc; ++e) {
                const FQAccum ac{1, e, d_coeff, dF};
                rc = rg_fq_launch(pr, P, B, dX, 1, iF2, 0, iFdx, 0.0, 1.0, 1, nullptr, &ac);
            }
            if (rc) return rc;
            CU(ctx, cudaGetLastError());
            return RG_OK;
        }
        int rc = rg_fq_launch(pr, P, B, dX, 0, cost_fused ? dF : iF, cost_fused ? 1 : 0, iFdx, sgn * P.inv_eps / DD1q, sgn, want_grad ? 1 : 0, po);
        if (rc) return rc;
        if (ne > 0) { rc = rg_fq_launch(pr, P, B, dX, 1, iF2, 0, iF2dx, 0.0, 1.0, want_grad ? 1 : 0); if (rc) return rc; }
        if (mode == 1 && ne > 0) {
            if (!want_grad) RG_FAIL(ctx, RG_ERR_INVALID, "cost without gradient is not exposed");
            KTimer kt(ctx, RG_K_EPILOGUE);
            const size_t n = (size_t)B * P.nx;
            const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
            k_cost_grad<<<grid, 256, 0, st>>>(B, P.nx, ne, iF, iF2, iF2dx, d_coeff, dF, iFdx);
        }
        CU(ctx, cudaGetLastError());
        return RG_OK;
    }
    // ---- K1: step propagators + first-order differences (+ chunk aggregates)
    constexpr bool kThreadOK = (D <= 5);
    constexpr bool kSparseThread = (PID != PAT_FULL) && (Pat<D, CMS>::nnz <= 12);   // state fits one thread's registers
    const bool fast = kThreadOK && pr->tri_ok && !pr->force_group;
    if (!fast && PID != PAT_FULL) RG_FAIL(ctx, RG_ERR_INVALID, "internal: structural pattern without the fast path");
    if (b2) {
        // workspace-free: chunk aggregates from recomputed closed-form block propagators (rg_block2.cuh)
        int rc = rg_b2_launch_agg(pr, P, B, L, nc, dX);
        if (rc) return rc;
    } else if (fast) {
        // Hermitian fast path: one thread per time step, triangles in registers; aggregates in a second kernel.
        constexpr int DT = kThreadOK ? D : 2;
        constexpr unsigned UM = tri_mask_of<DT, (kThreadOK ? PID : PAT_FULL)>();
        size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
        // opt-in (RG_FUSED_AGG=1): e = 0, sparse pattern, power-of-two chunk length -> chunk aggregates formed inside k_steps_t
        const bool fuse_agg = kSparseThread && ne == 0 && !pr->force_group_sweeps && pr->fused_agg && L <= 32 && (L & (L - 1)) == 0;
        const int agg_off = (int)smem;
        if (fuse_agg) smem += (size_t)128 * WSM * cb;
        const long long items = (long long)B * (fuse_agg ? (long long)nc * L : (long long)P.N);
        const int grid = (int)((items + 127) / 128);
        {
            KTimer kt(ctx, RG_K_STEPS);
            if (fuse_agg)
                k_steps_t<DT, UM><<<grid, 128, smem, st>>>(P, pr->tri, dX, B, pr->ws.as<cplx>(), ctx->d_status, L, nc,
                                                           pr->Qb.as<cplx>(), agg_off);
            else
                k_steps_t<DT, UM><<<grid, 128, smem, st>>>(P, pr->tri, dX, B, pr->ws.as<cplx>(), ctx->d_status);
        }
        const long long citems = (long long)B * nc;
        bool agg_done = fuse_agg;
        if constexpr (kSparseThread) {
            if (!agg_done && !pr->force_group_sweeps) {
                // sparse pattern: whole matrices in one thread's registers, no shared memory
                KTimer kt(ctx, RG_K_AGG);
                k_chunk_agg_t<D, CMS><<<(int)((citems + 127) / 128), 128, 0, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(), pr->Wlb.as<cplx>());
                agg_done = true;
            }
        }
        if (!agg_done) {
            const int gs = kagg_group_stride(D, ne);
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
            const size_t smem2 = (size_t)wpc * G * gs * cb;
            int rc = set_smem(ctx, k_chunk_agg<D, CM, CMS>, smem2);
            if (rc) return rc;
            const int grid2 = (int)((citems + (long long)wpc * G - 1) / ((long long)wpc * G));
            KTimer kt(ctx, RG_K_AGG);
            k_chunk_agg<D, CM, CMS><<<grid2, wpc * 32, smem2, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(), pr->Wlb.as<cplx>());
        }
    } else {
        const int gs = k1_group_stride(D, P.nterms, ne);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps<D>, smem);
        if (rc) return rc;
        const long long items = (long long)B * nc;
        const int grid = (int)((items + (long long)wpc * G - 1) / ((long long)wpc * G));
        KTimer kt(ctx, RG_K_STEPS);
        k_steps<D><<<grid, wpc * 32, smem, st>>>(P, dX, B, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(),
                                                       pr->Wlb.as<cplx>(), ctx->d_status);
    }
    // ---- K1b: mixed second differences (only needed for the sensitivity gradient)
    if (b2) {
        // mixed second differences are recomputed inside the sensitivity-gradient sweep
    } else if (ne > 0 && want_grad && P.nvar > 0 && fast && PID != PAT_FULL) {
        // structured fast path: thread per step, four triangles in registers
        constexpr int DT = kThreadOK ? D : 2;
        constexpr unsigned UM = tri_mask_of<DT, (kThreadOK ? PID : PAT_FULL)>();
        const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
        const long long items = (long long)B * P.N;
        KTimer kt(ctx, RG_K_STEPS_SO);
        if constexpr (PID != PAT_FULL)
            k_steps_so_t<DT, UM><<<(int)((items + 127) / 128), 128, smem, st>>>(P, pr->tri, dX, B, pr->ws.as<cplx>(), ctx->d_status);
    } else if (ne > 0 && want_grad && P.nvar > 0) {
        const int gs = k1b_group_stride(D, P.nterms);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps_so<D>, smem);
        if (rc) return rc;
        const long long items = (long long)B * P.N;
        const int grid = (int)((items + (long long)wpc * G - 1) / ((long long)wpc * G));
        KTimer kt(ctx, RG_K_STEPS_SO);
        k_steps_so<D><<<grid, wpc * 32, smem, st>>>(P, dX, B, pr->ws.as<cplx>(), ctx->d_status);
    }
    // ---- K2
    {
        const int gs = k2_group_stride(D);
        int wpc = RG_SCAN_WPC;
        while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb;
        int rc = set_smem(ctx, k_scan<D>, smem);
        if (rc) return rc;
        dim3 grid((B + wpc * G - 1) / (wpc * G), 1 + ne);
        KTimer kt(ctx, RG_K_SCAN);
        k_scan<D><<<grid, wpc * 32, smem, st>>>(P, dX, B, nc, pr->Qb.as<cplx>(), pr->Wlb.as<cplx>(), pr->Cb.as<cplx>(),
                                               pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(),
                                               iF, iF2, pr->addT.as<double>(), 0, nullptr, nullptr);
    }
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    // mode 1 with no error sources writes -F_dx straight into grad
    const double sign0 = (mode == 1 && ne == 0) ? -1.0 : 1.0;
    if (want_grad) {
        // ---- K3: backward gradient sweeps (fidelity role, then one role per error source)
        const long long items = (long long)B * nc;
        bool grad_done = false, graderr_done = false;
        if (b2) {
            int rc = rg_b2_launch_grad(pr, P, B, L, nc, dX, iFdx, sign0 * P.inv_eps / DD1);
            if (rc) return rc;
            if (ne > 0) { rc = rg_b2_launch_grad_err(pr, P, B, L, nc, dX, iF2dx); if (rc) return rc; }
            grad_done = graderr_done = true;
        }
        if constexpr (kSparseThread) {
            if (!grad_done && fast && !pr->force_group_sweeps && pr->costate_in_pattern) {
                {
                    KTimer kt(ctx, RG_K_GRAD);
                    k_grad_t<D, CMS><<<(int)((items + 127) / 128), 128, 0, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(),
                        pr->Gb.as<cplx>(), iFdx, sign0 * P.inv_eps / DD1, pr->addS.as<double>());
                }
                if (ne > 0) {
                    KTimer kt(ctx, RG_K_GRAD_ERR);
                    dim3 grid((unsigned)((items + 127) / 128), ne);
                    k_grad_err_t<D, CMS><<<grid, 128, 0, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(), pr->Wb.as<cplx>(),
                                                            pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), iF2dx, pr->addS.as<double>());
                }
                grad_done = graderr_done = true;
            }
        }
        if (!grad_done) {
            const int gs = k3_group_stride(D, 1 + P.nvar + (P.hermitian ? 0 : 1));
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
            const size_t smem = (size_t)wpc * G * gs * cb;
            int rc = set_smem(ctx, k_grad<D, false, CM, CMS>, smem);
            if (rc) return rc;
            dim3 grid((unsigned)((items + (long long)wpc * G - 1) / ((long long)wpc * G)), 1);
            KTimer kt(ctx, RG_K_GRAD);
            k_grad<D, false, CM, CMS><<<grid, wpc * 32, smem, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(),
                pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), iFdx,
                sign0 * P.inv_eps / DD1, iF2dx, pr->addS.as<double>());
        }
        if (ne > 0 && !graderr_done) {
            const int gs = k3_group_stride(D, 2 + 2 * P.nvar + (P.hermitian ? 0 : 1));
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
            const size_t smem = (size_t)wpc * G * gs * cb;
            int rc = set_smem(ctx, k_grad<D, true, CM, CMS>, smem);
            if (rc) return rc;
            dim3 grid((unsigned)((items + (long long)wpc * G - 1) / ((long long)wpc * G)), ne);
            KTimer kt(ctx, RG_K_GRAD_ERR);
            k_grad<D, true, CM, CMS><<<grid, wpc * 32, smem, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(),
                pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), iFdx,
                0.0, iF2dx, pr->addS.as<double>());
        }
        // ---- K4: additional parameters
        if (P.a > 0) {
            const int n = B * (1 + ne) * P.a;
            KTimer kt(ctx, RG_K_EPILOGUE);
            k_add_params<<<(n + 127) / 128, 128, 0, st>>>(P, B, pr->addT.as<double>(), pr->addS.as<double>(), iFdx, sign0, iF2dx);
        }
    }
    if (mode == 1) {
        KTimer kt(ctx, RG_K_EPILOGUE);
        if (ne > 0 && want_grad) {
            const size_t n = (size_t)B * P.nx;
            const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
            k_cost_grad<<<grid, 256, 0, st>>>(B, P.nx, ne, iF, iF2, iF2dx, d_coeff, dF, iFdx);
        } else if (ne > 0) {
            RG_FAIL(ctx, RG_ERR_INVALID, "cost without gradient is not exposed");
        } else {
            k_cost_only<<<(B + 255) / 256, 256, 0, st>>>(B, iF, dF);
        }
    }
    CU(ctx, cudaGetLastError());
    return RG_OK;
}

// Per-dimension entry: picks the structural pattern instantiation for this problem.
template <int D>
static int dim_run_slab(rg_problem* pr, int B, const Plan& pl, const double* dX, int mode, const double* d_coeff,
                        double* dF, double* dFdx, double* dF2, double* dF2dx, bool want_grad) {
    const bool fast = pr->tri_ok && !pr->force_group && !pr->force_dense;
    if (rg_use_b2(pr)) return run_slab<D, PAT_FULL>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
    if constexpr (D == 5) {
        if (fast) {
            if ((pr->tri_union & ~tri_mask_of<5, PAT_M5_DRIVE>()) == 0)
                return run_slab<5, PAT_M5_DRIVE>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
            if ((pr->tri_union & ~tri_mask_of<5, PAT_M5_FULL>()) == 0)
                return run_slab<5, PAT_M5_FULL>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
        }
    }
    return run_slab<D, PAT_FULL>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
}

// ---- calculate_unitary_and_derivatives, materialised (one pulse) ---------------------------------------
template <int D>
static int materialize_impl(rg_problem* pr, const double* dx, cplx* dU, cplx* dU_dx, cplx* dU_dx_add, cplx* dU_derr,
                            cplx* dU_derr_dx, cplx* dU_derr_dx_add) {
    rg_ctx* ctx = pr->ctx;
    constexpr u64 CM = full_cmask<D>();
    DevProblem P = pr->dp;
    P.wsm = D * D; P.cmask = CM; P.wsB = 1;
    constexpr int G = GroupInfo<D>::G;
    const int DD = D * D, ne = P.e;
    const size_t cb = sizeof(cplx);
    cudaStream_t st = ctx->stream;
    int L = std::max(1, std::min(16, P.N / 64));
    if (pr->chunk_override > 0) L = std::min(pr->chunk_override, P.N);
    const int nc = (P.N + L - 1) / L;
    if (pr->ws.ensure((size_t)P.N * P.nstore * DD * cb) || pr->Qb.ensure((size_t)nc * DD * cb) ||
        pr->Wlb.ensure(std::max<size_t>(16, (size_t)nc * ne * DD * cb)) || pr->Cb.ensure((size_t)nc * DD * cb) ||
        pr->Wb.ensure(std::max<size_t>(16, (size_t)ne * nc * DD * cb)) || pr->Gb.ensure((size_t)nc * DD * cb) ||
        pr->G1b.ensure(std::max<size_t>(16, (size_t)ne * nc * DD * cb)) || pr->H1b.ensure(std::max<size_t>(16, (size_t)ne * nc * DD * cb)) ||
        pr->dM.ensure(std::max<size_t>(16, (size_t)(1 + ne) * P.a * P.N * DD * cb)))
        RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed");
    {   // general group kernel: propagators, first-order differences, chunk aggregates (with scaling-and-squaring)
        const int gs = k1_group_stride(D, P.nterms, ne);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_STEPS);
        k_steps<D><<<(nc + wpc * G - 1) / (wpc * G), wpc * 32, smem, st>>>(P, dx, 1, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(),
                                                                                pr->Wlb.as<cplx>(), ctx->d_status);
    }
    if (ne > 0 && P.nvar > 0) {
        const int gs = k1b_group_stride(D, P.nterms);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps_so<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_STEPS_SO);
        k_steps_so<D><<<(P.N + wpc * G - 1) / (wpc * G), wpc * 32, smem, st>>>(P, dx, 1, pr->ws.as<cplx>(), ctx->d_status);
    }
    {
        const int gs = k2_group_stride(D);
        const size_t smem = (size_t)G * gs * cb;
        int rc = set_smem(ctx, k_scan<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_SCAN);
        k_scan<D><<<dim3(1, 1 + ne), 32, smem, st>>>(P, dx, 1, nc, pr->Qb.as<cplx>(), pr->Wlb.as<cplx>(), pr->Cb.as<cplx>(),
                                                   pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(),
                                                   nullptr, nullptr, nullptr, 1, dU, dU_derr);
    }
    {
        const int nload_max = ((ne > 0) ? (2 + 2 * P.nvar) : (1 + P.nvar)) + (P.hermitian ? 0 : 1);
        const int gs = kmat_group_stride(D, nload_max);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb;
        int rc = set_smem(ctx, k_materialize<D, CM>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_ANALYSIS);
        k_materialize<D, CM><<<dim3((nc + wpc * G - 1) / (wpc * G), 1 + ne), wpc * 32, smem, st>>>(
            P, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(), pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(),
            pr->H1b.as<cplx>(), dU_dx, dU_derr_dx, pr->dM.as<cplx>());
    }
    if (P.a > 0) {
        const int n = (1 + ne) * P.a * DD;
        KTimer kt(ctx, RG_K_ANALYSIS);
        k_reduce_add<<<(n + 127) / 128, 128, 0, st>>>(P, pr->dM.as<cplx>(), dU_dx_add, dU_derr_dx_add);
    }
    CU(ctx, cudaGetLastError());
    return RG_OK;
}

// ---- interaction-picture error operators on the device (shared by the three analysis entry points)
template <int D>
static int launch_interaction(rg_problem* pr, const double* dx, cplx* dO) {
    rg_ctx* ctx = pr->ctx;
    DevProblem P = pr->dp;
    constexpr int G = GroupInfo<D>::G;
    const int DD = D * D;
    const size_t cb = sizeof(cplx);
    cudaStream_t st = ctx->stream;
    if (!P.hermitian || P.N < 64 || pr->force_sequential_analysis) {
        // time-sequential kernel: general (carries C^{-1} for non-Hermitian H), one warp
        const size_t smem = staged_desc_bytes(P.nterms, P.nent, D) + (size_t)(6 * D * D + 2 * P.nterms) * sizeof(cplx);
        int rc = set_smem(ctx, k_interaction_ops<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_ANALYSIS);
        k_interaction_ops<D><<<1, 32, smem, st>>>(P, dx, dO, ctx->d_status);
        return RG_OK;
    }
    // time-parallel path: step propagators -> chunk products -> prefix at chunk ends -> per-chunk forward sweep.
    // Only U_k is needed: run the general step kernel on a copy of the problem without error sources / variables.
    DevProblem Pu = P;
    Pu.e = 0; Pu.nvar = 0; Pu.nstore = 1; Pu.wsm = DD; Pu.cmask = full_cmask<D>(); Pu.wsB = 1;
    const int L = 16;
    const int nc = (P.N + L - 1) / L;
    if (pr->ws.ensure((size_t)P.N * DD * cb) || pr->Qb.ensure((size_t)nc * DD * cb) || pr->Wlb.ensure(16) ||
        pr->Cb.ensure((size_t)nc * DD * cb) || pr->Wb.ensure(16) || pr->Gb.ensure((size_t)nc * DD * cb) ||
        pr->G1b.ensure(16) || pr->H1b.ensure(16) || pr->dM.ensure((size_t)DD * cb))
        RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed");
    {
        const int gs = k1_group_stride(D, P.nterms, 0);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_STEPS);
        k_steps<D><<<(nc + wpc * G - 1) / (wpc * G), wpc * 32, smem, st>>>(Pu, dx, 1, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(),
                                                                        pr->Wlb.as<cplx>(), ctx->d_status);
    }
    {
        const int gs = k2_group_stride(D);
        const size_t smem = (size_t)G * gs * cb;
        int rc = set_smem(ctx, k_scan<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_SCAN);
        k_scan<D><<<dim3(1, 1), 32, smem, st>>>(Pu, dx, 1, nc, pr->Qb.as<cplx>(), pr->Wlb.as<cplx>(), pr->Cb.as<cplx>(),
                                              pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(),
                                              nullptr, nullptr, nullptr, 1, pr->dM.as<cplx>(), nullptr);
    }
    {
        const int gs = rg_odd(3 * DD + P.nterms);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_interaction_ops_par<D>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_ANALYSIS);
        k_interaction_ops_par<D><<<(nc + wpc * G - 1) / (wpc * G), wpc * 32, smem, st>>>(P, dx, L, nc, pr->ws.as<cplx>(),
                                                                                      pr->Cb.as<cplx>(), dO);
    }
    CU(ctx, cudaGetLastError());
    return RG_OK;
}
template <int D>
static int launch_response(rg_problem* pr, const double* dfreqs, int first, int count, int M, int shift, double* dR) {
    rg_ctx* ctx = pr->ctx;
    KTimer kt(ctx, RG_K_ANALYSIS);
    dim3 grid(count, pr->dp.e);
    k_response<D><<<grid, 128, 0, ctx->stream>>>(pr->dp, pr->dO.as<cplx>(), dfreqs, first, count, M, shift, dR);
    return RG_OK;
}
template <int D>
static int launch_expectation(rg_problem* pr, double* dOut) {
    KTimer kt(pr->ctx, RG_K_ANALYSIS);
    k_expectation<D><<<pr->dp.e, 256, 0, pr->ctx->stream>>>(pr->dp, pr->dO.as<cplx>(), dOut);
    return RG_OK;
}

#define RG_DEFINE_DIM(D)                                                                                         \
    extern const DimOps rg_ops_d##D = {dim_run_slab<D>, materialize_impl<D>, launch_interaction<D>, launch_response<D>, \
                                       launch_expectation<D>};
