// rg_block2.cu -- launchers of the workspace-free block-2 path (rg_block2.cuh).  A separate translation unit: the jets
// unroll into long straight-line kernels, and nothing here depends on the per-dimension group kernels.
#include "rg_host.cuh"
#include "rg_block2.cuh"
#include "rg_fusedq.cuh"

// Patterns instantiated ahead of time (upper-triangle bit = k(k+1)/2 + i):
//   d = 5 symmetric-blockaded model (src/RydbergTools.jl:31-39): drive (1,3),(2,4) [+ Rydberg diagonal (3,3),(4,4)]
//   d = 7 full-blockaded model (src/RydbergTools.jl:71-81): drive (1,4),(2,5),(3,6) [+ diagonal (4,4),(5,5),(6,6)]
constexpr unsigned B2_M5_DRIVE = (1u << 7) | (1u << 12);
constexpr unsigned B2_M5_FULL = B2_M5_DRIVE | (1u << 9) | (1u << 14);
constexpr unsigned B2_M7_DRIVE = (1u << 11) | (1u << 17) | (1u << 24);
constexpr unsigned B2_M7_FULL = B2_M7_DRIVE | (1u << 14) | (1u << 20) | (1u << 27);
static_assert(b2_eligible(5, B2_M5_FULL) && b2_eligible(7, B2_M7_FULL), "patterns must decompose into blocks of <= 2 levels");

int rg_b2_pattern(const rg_problem* pr) {
    const DevProblem& P = pr->dp;
    if (!pr->tri_ok || !P.hermitian || !pr->costate_in_pattern || pr->force_ws) return 0;
    const unsigned u = pr->tri_union;
    if (P.d == 5) {
        if ((u & ~B2_M5_DRIVE) == 0) return 1;
        if ((u & ~B2_M5_FULL) == 0) return 2;
    } else if (P.d == 7) {
        if ((u & ~B2_M7_DRIVE) == 0) return 3;
        if ((u & ~B2_M7_FULL) == 0) return 4;
    }
    return 0;
}

template <int D, unsigned UM>
static int launch_agg(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_AGG);
    k_agg_b2<D, UM><<<(int)((items + 127) / 128), 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Qb.as<cplx>(),
                                                                            pr->Wlb.as<cplx>(), ctx->d_status);
    return RG_OK;
}
template <int D, unsigned UM>
static int launch_grad(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out0, double scale0) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_GRAD);
    k_grad_b2<D, UM><<<(int)((items + 127) / 128), 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Cb.as<cplx>(),
                                                                             pr->Gb.as<cplx>(), out0, scale0, pr->addS.as<double>());
    return RG_OK;
}
template <int D, unsigned UM>
static int launch_grad_err(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out1) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_GRAD_ERR);
    dim3 grid((unsigned)((items + 127) / 128), P.e);
    k_grad_err_b2<D, UM><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Cb.as<cplx>(), pr->Wb.as<cplx>(),
                                                           pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), out1, pr->addS.as<double>());
    return RG_OK;
}

#define B2_DISPATCH(fn, ...)                                              \
    switch (rg_b2_pattern(pr)) {                                          \
    case 1: return fn<5, B2_M5_DRIVE>(__VA_ARGS__);                       \
    case 2: return fn<5, B2_M5_FULL>(__VA_ARGS__);                        \
    case 3: return fn<7, B2_M7_DRIVE>(__VA_ARGS__);                       \
    case 4: return fn<7, B2_M7_FULL>(__VA_ARGS__);                        \
    default: pr->ctx->err = "internal: block-2 path without an eligible pattern"; return RG_ERR_INVALID; \
    }

int rg_b2_launch_agg(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX) {
    B2_DISPATCH(launch_agg, pr, P, B, L, nc, dX)
}
int rg_b2_launch_grad(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out0, double scale0) {
    B2_DISPATCH(launch_grad, pr, P, B, L, nc, dX, out0, scale0)
}
int rg_b2_launch_grad_err(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out1) {
    B2_DISPATCH(launch_grad_err, pr, P, B, L, nc, dX, out1)
}
// resident CTAs per SM of the sweeps (occupancy query; feeds the chunk planner)
void rg_b2_occupancy(const rg_problem* pr, int* agg_ctas, int* grad_ctas) {
    const size_t smem = staged_plan_bytes(pr->dp.nterms, pr->tri.nent, pr->dp.d);
    int a = 1, g = 1;
    switch (rg_b2_pattern(pr)) {
    case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_agg_b2<5, B2_M5_DRIVE>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_grad_b2<5, B2_M5_DRIVE>, 128, smem); break;
    case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_agg_b2<5, B2_M5_FULL>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_grad_b2<5, B2_M5_FULL>, 128, smem); break;
    case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_agg_b2<7, B2_M7_DRIVE>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_grad_b2<7, B2_M7_DRIVE>, 128, smem); break;
    case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_agg_b2<7, B2_M7_FULL>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_grad_b2<7, B2_M7_FULL>, 128, smem); break;
    default: break;
    }
    *agg_ctas = std::max(1, a); *grad_ctas = std::max(1, g);
}

// ---- one-launch fused quaternion path (rg_fusedq.cuh): patterns without diagonal terms
static_assert(b2_quat(5, B2_M5_DRIVE) && b2_quat(7, B2_M7_DRIVE) && !b2_quat(5, B2_M5_FULL), "quaternion eligibility");
int rg_fq_pattern(const rg_problem* pr) {
    const int p = rg_b2_pattern(pr);
    return (p == 1 || p == 3) ? p : 0;
}
template <int D, unsigned UM>
static int launch_fq(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
                     double scale0, double scale0T, int do_grad) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = fq_smem_bytes(D, b2_nblocks(D, UM), P.nterms, pr->tri.nent);
    if (!pr->fq_ctas[0]) {
        int rc = set_smem(ctx, k_fused_q<D, UM, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true>, smem); if (rc) return rc;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[0], k_fused_q<D, UM, false>, 128, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[1], k_fused_q<D, UM, true>, 128, smem);
        pr->fq_ctas[0] = std::max(1, pr->fq_ctas[0]); pr->fq_ctas[1] = std::max(1, pr->fq_ctas[1]);
    }
    // warps per pulse: cost = waves * (sweep steps per lane + fixed scan/algebra overhead of ~24 sweep steps)
    const double cap = (double)ctx->sm_count * pr->fq_ctas[err_role ? 1 : 0];
    int wpp = 1; double best = 1e300;
    for (int w = 1; w <= 4; w <<= 1) {
        const int Lw = (P.N + 32 * w - 1) / (32 * w);
        const double ctas = std::ceil((double)B * w / 4.0) * (err_role ? P.e : 1);
        const double cost = std::ceil(ctas / cap) * (Lw + 24.0);
        if (cost < best) { best = cost; wpp = w; }
    }
    if (pr->chunk_override > 0) wpp = std::max(1, std::min(4, pr->chunk_override >= 4 ? 4 : pr->chunk_override));
    if (wpp == 3) wpp = 2;
    const int L = (P.N + 32 * wpp - 1) / (32 * wpp);
    const int ppc = 4 / wpp;
    dim3 grid((unsigned)((B + ppc - 1) / ppc), err_role ? P.e : 1);
    KTimer kt(ctx, err_role ? RG_K_GRAD_ERR : RG_K_GRAD);
    if (err_role)
        k_fused_q<D, UM, true><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
    else
        k_fused_q<D, UM, false><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
    return RG_OK;
}
int rg_fq_launch(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
                 double scale0, double scale0T, int do_grad) {
    switch (rg_fq_pattern(pr)) {
    case 1: return launch_fq<5, B2_M5_DRIVE>(pr, P, B, dX, err_role, Fout, fmode, out, scale0, scale0T, do_grad);
    case 3: return launch_fq<7, B2_M7_DRIVE>(pr, P, B, dX, err_role, Fout, fmode, out, scale0, scale0T, do_grad);
    default: pr->ctx->err = "internal: fused quaternion path without an eligible pattern"; return RG_ERR_INVALID;
    }
}
