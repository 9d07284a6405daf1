// rg_fq_impl.inl -- launcher of the one-launch fused quaternion path for one pattern (RG_B2_D, RG_B2_MASK, RG_B2_ID).
#include "rg_host.cuh"
#include "rg_fusedq.cuh"
#include "rg_block2_patterns.cuh"
static_assert(b2_quat(RG_B2_D, RG_B2_MASK), "pattern must be quaternion-eligible (blocks of <= 2 levels, no diagonal terms)");
namespace {
constexpr int D = RG_B2_D;
constexpr unsigned UM = RG_B2_MASK;
int launch_fq(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
              double scale0, double scale0T, int do_grad) {
    rg_ctx* ctx = pr->ctx;
    const size_t smem = fq_smem_bytes(D, b2_nblocks(D, UM), P.nterms, pr->tri.nent);
    const bool da = pr->diag_alg && !pr->force_dense_alg;
    if (!pr->fq_ctas[0]) {
        int rc = set_smem(ctx, k_fused_q<D, UM, false, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, false, true>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, true>, smem); if (rc) return rc;
        if (da) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[0], k_fused_q<D, UM, false, true>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[1], k_fused_q<D, UM, true, true>, 128, smem);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[0], k_fused_q<D, UM, false, false>, 128, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[1], k_fused_q<D, UM, true, false>, 128, smem);
        }
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
    if (pr->wpp_override > 0) wpp = pr->wpp_override >= 4 ? 4 : (pr->wpp_override >= 2 ? 2 : 1);
    const int L = (P.N + 32 * wpp - 1) / (32 * wpp);
    const int ppc = 4 / wpp;
    dim3 grid((unsigned)((B + ppc - 1) / ppc), err_role ? P.e : 1);
    KTimer kt(ctx, err_role ? RG_K_GRAD_ERR : RG_K_GRAD);
    if (err_role) {
        if (da) k_fused_q<D, UM, true, true><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
        else k_fused_q<D, UM, true, false><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
    } else {
        if (da) k_fused_q<D, UM, false, true><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
        else k_fused_q<D, UM, false, false><<<grid, 128, smem, ctx->stream>>>(P, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status);
    }
    return RG_OK;
}
}  // namespace
#define RG_B2_CAT2(a, b) a##b
#define RG_B2_CAT(a, b) RG_B2_CAT2(a, b)
extern const FQOps RG_B2_CAT(rg_fq_ops_p, RG_B2_ID) = {launch_fq};
