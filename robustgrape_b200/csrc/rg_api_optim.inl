// rg_api_optim.inl -- host side of the device-resident optimiser loop (included by rg_api.cu): regularisation epilogue and
// batched L-BFGS (rg_optim.cuh).
#include "rg_optim.cuh"

static int upload_reg(rg_problem* pr, const int32_t* kind, const double* c1, const double* c2, bool* any) {
    rg_ctx* ctx = pr->ctx;
    const int p = pr->dp.p;
    *any = false;
    if (!kind || p == 0) return RG_OK;
    for (int i = 0; i < p; ++i) {
        if (kind[i] < RG_REG_NONE || kind[i] > RG_REG_SIN2) RG_FAIL(ctx, RG_ERR_INVALID, "unknown regularisation kind %d", kind[i]);
        if (kind[i] != RG_REG_NONE && (kind[i] == RG_REG_SIN2 ? pr->dp.N < 3 : pr->dp.N < 4)) RG_FAIL(ctx, RG_ERR_INVALID, "regularisation needs ntimes >= 4");
        if (kind[i] != RG_REG_NONE) *any = true;
    }
    if (!*any) return RG_OK;
    if (!c1 || !c2) RG_FAIL(ctx, RG_ERR_INVALID, "regularisation coefficients missing");
    if (pr->regbuf.ensure((size_t)p * (sizeof(int) + 16))) RG_FAIL(ctx, RG_ERR_NOMEM, "alloc");
    std::vector<unsigned char> h((size_t)p * (sizeof(int) + 16));
    double* hc1 = reinterpret_cast<double*>(h.data()); double* hc2 = hc1 + p; int* hk = reinterpret_cast<int*>(hc2 + p);
    for (int i = 0; i < p; ++i) { hc1[i] = c1[i]; hc2[i] = c2[i]; hk[i] = kind[i]; }
    if (h == pr->reg_host) return RG_OK;                   // same table as the last call (every evaluation of an optimiser run): no copy, no sync
    CU(ctx, cudaMemcpyAsync(pr->regbuf.p, h.data(), h.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));           // h goes out of scope
    pr->reg_host = h;
    return RG_OK;
}
static int launch_reg(rg_problem* pr, int B, const double* dX, double* dcost, double* dgrad) {
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    const double* c1 = pr->regbuf.as<double>(); const double* c2 = c1 + P.p; const int* k = reinterpret_cast<const int*>(c2 + P.p);
    KTimer kt(ctx, RG_K_EPILOGUE);
    k_regularize<<<dim3(B, P.p), 128, 0, ctx->stream>>>(B, P.nx, P.p, P.N, k, c1, c2, dX, dcost, dgrad);
    CU(ctx, cudaGetLastError());
    return RG_OK;
}

extern "C" int rg_cost_and_grad_batch_reg_dev(rg_problem* pr, int32_t B, const double* dX, const double* err_coeff, const int32_t* reg_kind,
                                              const double* reg_c1, const double* reg_c2, double* dcost, double* dgrad) {
    if (!pr) return RG_ERR_INVALID;
    if (B < 0 || (B > 0 && (!dX || !dcost || !dgrad))) RG_FAIL(pr->ctx, RG_ERR_INVALID, "bad batch arguments");
    bool any = false;
    int rc = upload_reg(pr, reg_kind, reg_c1, reg_c2, &any);
    if (rc) return rc;
    rc = run_dev(pr, B, dX, 1, err_coeff, dcost, dgrad, nullptr, nullptr);
    if (rc || !any || B == 0) return rc;
    return launch_reg(pr, B, dX, dcost, dgrad);
}

extern "C" int rg_cost_and_grad_batch_reg(rg_problem* pr, int32_t B, const double* X, const double* err_coeff, const int32_t* reg_kind,
                                          const double* reg_c1, const double* reg_c2, double* cost, double* grad) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (B < 0 || (B > 0 && (!X || !cost || !grad))) RG_FAIL(ctx, RG_ERR_INVALID, "bad batch arguments");
    if (B == 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t nx = pr->dp.nx;
    if (pr->dX.ensure(B * nx * 8) || pr->dOut.ensure((B + B * nx) * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device staging allocation failed");
    double* o = pr->dOut.as<double>();
    CU(ctx, cudaMemcpyAsync(pr->dX.p, X, B * nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rg_cost_and_grad_batch_reg_dev(pr, B, pr->dX.as<double>(), err_coeff, reg_kind, reg_c1, reg_c2, o, o + B);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(cost, o, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(grad, o + B, (size_t)B * nx * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}

// Batched L-BFGS with the iterate resident on the device.  dX (nx, B) is updated in place; dcost (B) receives the final cost;
// iters_out (host, B, may be NULL) the iterations each pulse took, info_out (host, 3 ints, may be NULL): evaluations, iterations of
// the longest-running pulse, line-search failures.
extern "C" int rg_lbfgs_batch_dev(rg_problem* pr, int32_t B, double* dX, const double* err_coeff, const int32_t* reg_kind, const double* reg_c1,
                                  const double* reg_c2, int32_t history, int32_t iterations, double g_tol, double* dcost, int32_t* iters_out,
                                  int32_t* info_out) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (B <= 0 || !dX || !dcost || history < 1 || history > 32 || iterations < 0) RG_FAIL(ctx, RG_ERR_INVALID, "bad L-BFGS arguments (1 <= history <= 32)");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t nx = pr->dp.nx, m = history;
    const size_t nd = (size_t)B * nx * 5 + 2 * (size_t)B * m * nx + (size_t)B * m + (size_t)B * 4;       // doubles
    const size_t ni = (size_t)B * 6 + 4;
    if (pr->lbfgs.ensure(nd * 8 + ni * 4)) RG_FAIL(ctx, RG_ERR_NOMEM, "L-BFGS state allocation failed (%zu MB)", (nd * 8) >> 20);
    double* q = pr->lbfgs.as<double>();
    LbfgsState st;
    st.B = B; st.nx = (int)nx; st.m = (int)m;
    st.X = dX; st.F = dcost;
    st.G = q; q += (size_t)B * nx; st.Xt = q; q += (size_t)B * nx; st.Gt = q; q += (size_t)B * nx; st.Dir = q; q += (size_t)B * nx;
    q += (size_t)B * nx;                                      // spare
    st.S = q; q += (size_t)B * m * nx; st.Y = q; q += (size_t)B * m * nx; st.rho = q; q += (size_t)B * m;
    st.Ft = q; q += B; st.alpha = q; q += B; st.gd = q; q += 2 * (size_t)B;
    int* qi = reinterpret_cast<int*>(q);
    st.hist = qi; st.head = qi + B; st.active = qi + 2 * B; st.done = qi + 3 * B; st.iters = qi + 4 * B; st.lsr = qi + 5 * B;
    st.counters = qi + 6 * B;
    st.max_iters = iterations;
    cudaStream_t s = ctx->stream;
    CU(ctx, cudaMemsetAsync(qi, 0, ni * 4, s));
    int nev = 0, rc;
    const int max_ls = 8;
    rc = rg_cost_and_grad_batch_reg_dev(pr, B, st.X, err_coeff, reg_kind, reg_c1, reg_c2, st.F, st.G); ++nev;
    if (rc) return rc;
    // One round = one batched evaluation.  Pulses whose last trial was accepted get a new direction, the others the next (shorter) step
    // along theirs; nobody waits.  The host reads 16 bytes per round to learn whether any pulse is still open.
    int h_cnt[4] = {0, 0, 0, 0};
    const long long max_rounds = (long long)iterations * (max_ls + 1) + 1;
    for (long long round = 0; iterations > 0 && round < max_rounds; ++round) {
        k_lbfgs_direction<<<B, 256, 0, s>>>(st, g_tol);
        k_lbfgs_trial<<<std::min(4 * ctx->sm_count, (int)(((size_t)B * nx + 255) / 256)), 256, 0, s>>>(st);
        ctx->launches += 2;
        rc = rg_cost_and_grad_batch_reg_dev(pr, B, st.Xt, err_coeff, reg_kind, reg_c1, reg_c2, st.Ft, st.Gt); ++nev;
        if (rc) return rc;
        k_lbfgs_check<<<B, 256, 0, s>>>(st, max_ls);
        CU(ctx, cudaMemsetAsync(st.counters, 0, 2 * sizeof(int), s));
        k_lbfgs_count<<<std::min(64, (B + 255) / 256), 256, 0, s>>>(st);
        ctx->launches += 2;
        CU(ctx, cudaMemcpyAsync(h_cnt, st.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        if (h_cnt[1] == 0) break;                          // every pulse converged, failed its line search or spent its iterations
    }
    const int its = h_cnt[3], ls_fail = h_cnt[2];
    CU(ctx, cudaGetLastError());
    if (iters_out) CU(ctx, cudaMemcpyAsync(iters_out, st.iters, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (info_out) { info_out[0] = nev; info_out[1] = its; info_out[2] = ls_fail; }
    return rg_ctx_synchronize(ctx);
}

extern "C" int rg_lbfgs_batch(rg_problem* pr, int32_t B, double* X, const double* err_coeff, const int32_t* reg_kind, const double* reg_c1,
                              const double* reg_c2, int32_t history, int32_t iterations, double g_tol, double* cost, int32_t* iters_out,
                              int32_t* info_out) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (B <= 0 || !X || !cost) RG_FAIL(ctx, RG_ERR_INVALID, "bad L-BFGS arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t nx = pr->dp.nx;
    if (pr->dXopt.ensure((size_t)B * (nx + 1) * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device staging allocation failed");
    double* dX = pr->dXopt.as<double>(); double* dc = dX + (size_t)B * nx;
    CU(ctx, cudaMemcpyAsync(dX, X, (size_t)B * nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rg_lbfgs_batch_dev(pr, B, dX, err_coeff, reg_kind, reg_c1, reg_c2, history, iterations, g_tol, dc, iters_out, info_out);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(X, dX, (size_t)B * nx * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(cost, dc, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}
