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
    pr->fq_ready = 0;
    return RG_OK;
}
int launch_fq(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
              double scale0, double scale0T, int do_grad, const PeerOut* po_in, const FQAccum* ac_in) {
    rg_ctx* ctx = pr->ctx;
    constexpr int NB = b2_nblocks(D, UM);
    { int rc = prepare_fq(pr); if (rc) return rc; }
    const bool pc = pr->dp.pc != 0;
    DevProblem Pl = P;
    Pl.pc = pr->dp.pc; Pl.pc_consts = pr->dp.pc_consts;
    const bool da = pr->diag_alg && !pr->force_dense_alg;
    const int role = err_role ? 1 : 0;
    // Per warps-per-pulse choice w: chunk length, shared memory (the staged controls need (N p + 32 w) doubles per pulse; a pulse
    // that does not fit is read from global memory instead) and resident CTAs/SM from the occupancy query -- all fixed per problem.
    PeerOut po = po_in ? *po_in : PeerOut{};
    int vsel = (po.n > 0 && po.grads) ? 1 : 0;                // variant 1: gradient staged through shared-memory rows (peer gradients)
    if (!pr->fq_ready) {
        for (int v = 0; v < 2; ++v)
            for (int wi = 0; wi < 3; ++wi) {
                const int w = 1 << wi, Lw = (P.N + 32 * w - 1) / (32 * w);
                int flags = (pr->stage_xs == 1 ? 3 : (pr->stage_xs & 3)) | (v ? 2 : 0);      // RG_XS: 1 = stage in and out, 2 = out only, 3 = both
                size_t sm = fq_smem_bytes(D, NB, P.nterms, pr->tri.nent, pc, da, P.a, P.p, w, Lw, flags != 0);
                if (sm > 200 * 1024) { flags = v ? -1 : 0; sm = fq_smem_bytes(D, NB, P.nterms, pr->tri.nent, pc, da, P.a, P.p, w, Lw, false); }
                pr->fq_xs[v][wi] = flags; pr->fq_smem[v][wi] = sm;
                for (int r = 0; r < 2; ++r) {
                    int occ = 1, rc = RG_OK;
#define RG_FQ_SET(ERRR, DAA, PCC) { rc = set_smem(ctx, k_fused_q<D, UM, ERRR, DAA, PCC>, sm); if (!rc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_q<D, UM, ERRR, DAA, PCC>, 128, sm); }
                    if (pc) { if (r) { if (da) RG_FQ_SET(true, true, true) else RG_FQ_SET(true, false, true) } else { if (da) RG_FQ_SET(false, true, true) else RG_FQ_SET(false, false, true) } }
                    else { if (r) { if (da) RG_FQ_SET(true, true, false) else RG_FQ_SET(true, false, false) } else { if (da) RG_FQ_SET(false, true, false) else RG_FQ_SET(false, false, false) } }
#undef RG_FQ_SET
                    if (rc) return rc;
                    pr->fq_occ[v][r][wi] = std::max(1, occ);
                }
            }
        pr->fq_ready = 1;
    }
    if (vsel == 1 && pr->fq_xs[1][0] < 0 && pr->fq_xs[1][1] < 0 && pr->fq_xs[1][2] < 0) {
        // a pulse this long does not fit the staged rows in any configuration: evaluate without the fused gather (peer_out_done
        // stays 0, so rg_cost_and_grad_batch_dev_scatter pushes the block with the copy engines afterwards)
        po = PeerOut{}; vsel = 0;
    }
    if (po.n > 0) pr->peer_out_done = 1;
    // warps per pulse: cost = waves * (sweep steps per lane + fixed scan/algebra overhead of ~24 sweep steps)
    int wpp = 1, wsel = 0; double best = 1e300;
    for (int wi = 0; wi < 3; ++wi) {
        const int w = 1 << wi;
        const int Lw = (P.N + 32 * w - 1) / (32 * w);
        const double cap = (double)ctx->sm_count * pr->fq_occ[vsel][role][wi];
        if (pr->fq_xs[vsel][wi] < 0) continue;                 // the staged rows of this choice do not fit
        const double ctas = std::ceil((double)B * w / 4.0) * (err_role ? ((ac_in && ac_in->on) ? 1 : P.e) : 1);
        const double cost = std::ceil(ctas / cap) * (Lw + 24.0);
        if (cost < best) { best = cost; wpp = w; wsel = wi; }
    }
    if (pr->wpp_override > 0) { wpp = pr->wpp_override >= 4 ? 4 : (pr->wpp_override >= 2 ? 2 : 1); wsel = wpp == 4 ? 2 : (wpp == 2 ? 1 : 0); }
    const int L = (P.N + 32 * wpp - 1) / (32 * wpp);
    const int ppc = 4 / wpp;
    if (pr->fq_xs[vsel][wsel] < 0) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "fused gather of gradients: a pulse of %d steps does not fit the staged rows", P.N);
    const size_t smem = pr->fq_smem[vsel][wsel];
    const int use_xs = pr->fq_xs[vsel][wsel];
    FQAccum ac = ac_in ? *ac_in : FQAccum{0, 0, nullptr, nullptr};
    if (use_xs != 0) ac.on = 0;                                   // (the host only asks for it on the unstaged path)
    if (ac_in && ac_in->on && !ac.on) RG_FAIL(ctx, RG_ERR_INVALID, "internal: in-kernel cost assembly on the staged path");
    dim3 grid((unsigned)((B + ppc - 1) / ppc), err_role ? (ac.on ? 1 : P.e) : 1);
    KTimer kt(ctx, err_role ? RG_K_GRAD_ERR : RG_K_GRAD);
#define RG_FQ_GO(ERRR, DAA, PCC) k_fused_q<D, UM, ERRR, DAA, PCC><<<grid, 128, smem, ctx->stream>>>(Pl, pr->tri, dX, B, wpp, L, Fout, fmode, out, scale0, scale0T, do_grad, use_xs, po, ac, ctx->d_status)
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
