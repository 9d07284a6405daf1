// rg_b2_impl.inl -- launchers of the three-kernel block-2 path for one pattern (RG_B2_D, RG_B2_MASK, RG_B2_ID); included by
// the rg_b2_p*.cu translation units so the long unrolled kernels compile in parallel.
#include "rg_host.cuh"
#include "rg_block2.cuh"
#include "rg_block2_patterns.cuh"
static_assert(b2_eligible(RG_B2_D, RG_B2_MASK), "pattern must decompose into blocks of <= 2 levels");
namespace {
constexpr int D = RG_B2_D;
constexpr unsigned UM = RG_B2_MASK;
int launch_agg(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_AGG);
    k_agg_b2<D, UM><<<(int)((items + 127) / 128), 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Qb.as<cplx>(),
                                                                            pr->Wlb.as<cplx>(), ctx->d_status);
    return RG_OK;
}
int launch_grad(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out0, double scale0) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_GRAD);
    k_grad_b2<D, UM><<<(int)((items + 127) / 128), 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Cb.as<cplx>(),
                                                                             pr->Gb.as<cplx>(), out0, scale0, pr->addS.as<double>());
    return RG_OK;
}
int launch_grad_err(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out1) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    const long long items = (long long)B * nc;
    KTimer kt(ctx, RG_K_GRAD_ERR);
    dim3 grid((unsigned)((items + 127) / 128), P.e);
    k_grad_err_b2<D, UM><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, L, nc, pr->Cb.as<cplx>(), pr->Wb.as<cplx>(),
                                                           pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), out1, pr->addS.as<double>());
    return RG_OK;
}
// resident CTAs per SM of the sweeps (occupancy query; feeds the chunk planner)
void occupancy(const rg_problem* pr, int* a, int* g) {
    const size_t smem = staged_plan_bytes(pr->dp.nterms, pr->tri.nent, D);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(a, k_agg_b2<D, UM>, 128, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(g, k_grad_b2<D, UM>, 128, smem);
}
}  // namespace
#define RG_B2_CAT2(a, b) a##b
#define RG_B2_CAT(a, b) RG_B2_CAT2(a, b)
extern const B2Ops RG_B2_CAT(rg_b2_ops_p, RG_B2_ID) = {launch_agg, launch_grad, launch_grad_err, occupancy};
