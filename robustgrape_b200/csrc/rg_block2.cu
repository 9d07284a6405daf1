// rg_block2.cu -- launchers of the workspace-free block-2 path (rg_block2.cuh).  A separate translation unit: the jets
// unroll into long straight-line kernels, and nothing here depends on the per-dimension group kernels.
#include "rg_host.cuh"
#include "rg_block2.cuh"

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
