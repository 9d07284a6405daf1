// rg_big_impl.inl -- host side of the dense path for one padded dimension RG_BIG_DP (included by rg_big_dp*.cu).
#include "rg_host.cuh"
#include "rg_big_sweeps.cuh"

namespace {
constexpr int DP = RG_BIG_DP;
typedef BigGemm<DP> GM;

int big_run_slab(rg_problem* pr, int B, const Plan& pl, const double* dX, int mode, const double* d_coeff, double* dF, double* dFdx,
                 double* dF2, double* dF2dx, bool want_grad) {
    rg_ctx* ctx = pr->ctx;
    DevProblem P = pr->dp;
    cudaStream_t st = ctx->stream;
    const int nv = P.nvar, ne = P.e, nslots = big_nslots(nv, ne);
    if (!pr->has_target) RG_FAIL(ctx, RG_ERR_INVALID, "problem has no target/projector: fidelity entry points unavailable");
    if (!P.hermitian) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "the dense path (ndim > 10) needs a Hermitian Hamiltonian");
    const size_t plane = (size_t)DP * DP, mat = 4 * plane, cb = sizeof(double);
    const int L = pl.L, nc = pl.nc;
    // occupancy-sized persistent grids
    const size_t smem_steps = (size_t)(GM::SMEM_DOUBLES + 64) * 8 + (size_t)nslots * P.nterms * sizeof(cplx);
    const size_t smem_sweep = (size_t)(GM::SMEM_DOUBLES + 64) * 8 + (size_t)std::max(P.ntt, 1) * sizeof(cplx);
    int rc = set_smem(ctx, k_big_steps<DP>, smem_steps); if (rc) return rc;
    rc = set_smem(ctx, k_big_scan<DP>, smem_sweep); if (rc) return rc;
    if (!pr->big_ctas) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr->big_ctas, k_big_steps<DP>, GM::NT, smem_steps);
        pr->big_ctas = std::max(1, std::min(pr->big_ctas, DP <= 16 ? 16 : (DP <= 32 ? 8 : 4)));
    }
    const long long tasks = (long long)B * P.N;
    const int grid_steps = (int)std::min<long long>(tasks, (long long)ctx->sm_count * pr->big_ctas);
    const int grid_sweep = (int)std::min<long long>((long long)B * nc, (long long)ctx->sm_count * 4);
    const size_t jet = (size_t)nslots * mat;
    const size_t scr_steps = (size_t)grid_steps * 6 * jet;
    const size_t scr_sweep = (size_t)std::max<long long>((long long)grid_sweep * std::max(ne, 1), (long long)B * (1 + ne)) * RG_BIG_SWEEP_MATS * mat;
    if (pr->ws.ensure((size_t)tasks * big_ws_step_doubles(DP, nv, ne) * cb) || pr->Qb.ensure((size_t)B * nc * mat * cb) ||
        pr->Wlb.ensure(std::max<size_t>(16, (size_t)B * nc * ne * mat * cb)) || pr->Cb.ensure((size_t)B * nc * mat * cb) ||
        pr->Wb.ensure(std::max<size_t>(16, (size_t)B * ne * nc * mat * cb)) || pr->Gb.ensure((size_t)B * nc * mat * cb) ||
        pr->G1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * mat * cb)) || pr->H1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * mat * cb)) ||
        pr->dM.ensure(std::max(scr_steps, scr_sweep) * cb) ||
        pr->F.ensure((size_t)B * 8) || pr->F2.ensure(std::max<size_t>(16, (size_t)B * ne * 8)) ||
        pr->addT.ensure(std::max<size_t>(16, (size_t)B * (1 + ne) * P.a * 8)) ||
        pr->addS.ensure(std::max<size_t>(16, pr->any_add_dep ? (size_t)B * (1 + ne) * P.a * P.N * 8 : 16)))
        RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed (dense path, B=%d)", B);
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
    BigData Bd{DP, nslots, pr->big_termM.as<double>(), pr->big_tgtM.as<double>(), nullptr, pr->dM.as<double>(), 6 * jet};
    BigBufs bb{pr->ws.as<double>(), pr->Qb.as<double>(), pr->Wlb.as<double>(), pr->Cb.as<double>(), pr->Wb.as<double>(),
               pr->Gb.as<double>(), pr->G1b.as<double>(), pr->H1b.as<double>(), pr->dM.as<double>()};
    { KTimer kt(ctx, RG_K_STEPS); k_big_steps<DP><<<grid_steps, GM::NT, smem_steps, st>>>(P, Bd, dX, B, bb.ws, ctx->d_status); }
    { KTimer kt(ctx, RG_K_AGG); k_big_agg<DP><<<grid_sweep, GM::NT, smem_sweep, st>>>(P, bb, B, L, nc); }
    { KTimer kt(ctx, RG_K_SCAN); k_big_prefix<DP><<<B, GM::NT, smem_sweep, st>>>(P, bb, B, nc); }
    { KTimer kt(ctx, RG_K_SCAN); k_big_scan<DP><<<dim3(B, 1 + ne), GM::NT, smem_sweep, st>>>(P, Bd, bb, dX, B, nc, iF, iF2, pr->addT.as<double>()); }
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    const double sign0 = (mode == 1 && ne == 0) ? -1.0 : 1.0;
    if (want_grad) {
        { KTimer kt(ctx, RG_K_GRAD); k_big_grad<DP><<<grid_sweep, GM::NT, smem_sweep, st>>>(P, bb, B, L, nc, iFdx, sign0 * P.inv_eps / DD1, pr->addS.as<double>()); }
        if (ne > 0) { KTimer kt(ctx, RG_K_GRAD_ERR); k_big_grad_err<DP><<<dim3(grid_sweep, ne), GM::NT, smem_sweep, st>>>(P, bb, B, L, nc, iF2dx, pr->addS.as<double>()); }
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
int big_unsupported_mat(rg_problem* pr, const double*, cplx*, cplx*, cplx*, cplx*, cplx*, cplx*) {
    RG_FAIL(pr->ctx, RG_ERR_UNSUPPORTED, "calculate_unitary_and_derivatives (materialised) is implemented for ndim <= 10 only");
}
int big_unsupported_int(rg_problem* pr, const double*, cplx*) {
    RG_FAIL(pr->ctx, RG_ERR_UNSUPPORTED, "the analysis entry points are implemented for ndim <= 10 only");
}
int big_unsupported_resp(rg_problem* pr, const double*, int, int, int, int, double*) {
    RG_FAIL(pr->ctx, RG_ERR_UNSUPPORTED, "the analysis entry points are implemented for ndim <= 10 only");
}
int big_unsupported_exp(rg_problem* pr, double*) {
    RG_FAIL(pr->ctx, RG_ERR_UNSUPPORTED, "the analysis entry points are implemented for ndim <= 10 only");
}
}  // namespace
#define RG_BIG_CAT2(a, b) a##b
#define RG_BIG_CAT(a, b) RG_BIG_CAT2(a, b)
extern const DimOps RG_BIG_CAT(rg_ops_big, RG_BIG_DP) = {big_run_slab, big_unsupported_mat, big_unsupported_int, big_unsupported_resp, big_unsupported_exp};
