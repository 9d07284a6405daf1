// rg_fq_impl.inl -- launcher of the one-launch fused quaternion path for one pattern (RG_B2_D, RG_B2_MASK, RG_B2_ID).
#include "rg_host.cuh"
#include "rg_fusedq.cuh"
#include "rg_block2_patterns.cuh"
static_assert(b2_quat(RG_B2_D, RG_B2_MASK), "pattern must be quaternion-eligible (blocks of <= 2 levels, no diagonal terms)");
namespace {
constexpr int D = RG_B2_D;
constexpr unsigned UM = RG_B2_MASK;
// phase-only class: evaluate the step-independent constants once (generic closed-form path at the reference point)
int prepare_fq(rg_problem* pr) {
    rg_ctx* ctx = pr->ctx;
    constexpr int NB = b2_nblocks(D, UM);
    if (!pr->dp.pc || pr->pc_ready) return RG_OK;
    const DevProblem& P = pr->dp;
    const size_t n = (size_t)(1 + 2 * P.e) * 2 * NB;
    void* buf = nullptr;
    CU(ctx, cudaMalloc(&buf, n * sizeof(cplx) + 16));
    pr->owned.push_back(buf);
    int* flag = reinterpret_cast<int*>(static_cast<cplx*>(buf) + n);
    const size_t sm = staged_plan_bytes(P.nterms, pr->tri.nent, D);
    int rc = set_smem(ctx, k_fqc_consts<D, UM>, sm); if (rc) return rc;
    k_fqc_consts<D, UM><<<1, 32, sm, ctx->stream>>>(pr->dp, pr->tri, static_cast<cplx*>(buf), flag);
    int hflag = 1;
    CU(ctx, cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (hflag) pr->dp.pc = 0;          // out of the closed-form range: the generic kernel runs and reports it
    else pr->dp.pc_consts = static_cast<const cplx*>(buf);
    pr->pc_ready = 1;
    pr->fq_ctas[0] = 0;
    return RG_OK;
}
int launch_fq(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
              double scale0, double scale0T, int do_grad) {
    rg_ctx* ctx = pr->ctx;
    constexpr int NB = b2_nblocks(D, UM);
    { int rc = prepare_fq(pr); if (rc) return rc; }
    const bool pc = pr->dp.pc != 0;
    DevProblem Pl = P;
    Pl.pc = pr->dp.pc; Pl.pc_consts = pr->dp.pc_consts;
    const size_t smem = fq_smem_bytes(D, NB, P.nterms, pr->tri.nent, pc);
    const bool da = pr->diag_alg && !pr->force_dense_alg;
    if (!pr->fq_ctas[0]) {
        int rc = set_smem(ctx, k_fused_q<D, UM, false, false, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, false, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, false, true, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, true, false>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, false, false, true>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, false, true>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, false, true, true>, smem); if (rc) return rc;
        rc = set_smem(ctx, k_fused_q<D, UM, true, true, true>, smem); if (rc) return rc;
#define RG_FQ_OCC(i, ERRR, DAA, PCC) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->fq_ctas[i], k_fused_q<D, UM, ERRR, DAA, PCC>, 128, smem)
        if (pc) { if (da) { RG_FQ_OCC(0, false, true, true); RG_FQ_OCC(1, true, true, true); } else { RG_FQ_OCC(0, false, false, true); RG_FQ_OCC(1, true, false, true); } }
        else { if (da) { RG_FQ_OCC(0, false, true, false); RG_FQ_OCC(1, true, true, false); } else { RG_FQ_OCC(0, false, false, false); RG_FQ_OCC(1, true, false, false); } }
#undef RG_FQ_OCC
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
#define RG_FQ_GO(ERRR, DAA, PCC) k_fused_q<D, UM, ERRR, DAA, PCC><<<grid, 128, smem, ctx->stream>>>(Pl, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, ctx->d_status)
    if (pc) {
        if (err_role) { if (da) RG_FQ_GO(true, true, true); else RG_FQ_GO(true, false, true); }
        else { if (da) RG_FQ_GO(false, true, true); else RG_FQ_GO(false, false, true); }
    } else {
        if (err_role) { if (da) RG_FQ_GO(true, true, false); else RG_FQ_GO(true, false, false); }
        else { if (da) RG_FQ_GO(false, true, false); else RG_FQ_GO(false, false, false); }
    }
#undef RG_FQ_GO
    return RG_OK;
}
}  // namespace
#define RG_B2_CAT2(a, b) a##b
#define RG_B2_CAT(a, b) RG_B2_CAT2(a, b)
extern const FQOps RG_B2_CAT(rg_fq_ops_p, RG_B2_ID) = {launch_fq, prepare_fq};
