// rg_api_analysis.inl -- host side of the analysis entry points (included by rg_api.cu).

extern "C" int rg_problem_path(rg_problem* pr, char* buf, int32_t len) {
    if (!pr || !buf || len <= 0) return RG_ERR_INVALID;
    cudaSetDevice(pr->ctx->device);
    const char* s = "group";
    if (pr->is_hstack) s = "hstack";
    else if (pr->big_dp) s = "dense";
    else if (rg_use_fq(pr)) {
        const int rc = rg_fq_prepare(pr);
        if (rc) return rc;
        s = pr->dp.pc ? "fused_q_pc" : "fused_q";
    } else if (rg_use_b2(pr)) s = "block2";
    else if (pr->tri_ok && pr->dp.d <= 5 && !pr->force_group) s = "steps_t";
    snprintf(buf, (size_t)len, "%s", s);
    return RG_OK;
}

extern "C" int rg_measure_fp64_peak(rg_ctx* ctx, double seconds, double* dfma_tflops, double* dmma_tflops) {
    if (!ctx) return RG_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;
    const int blocks = ctx->sm_count * 8, threads = 256;
    CU(ctx, cudaMalloc(&d, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaStream_t st = ctx->stream;
    for (int which = 0; which < 2; ++which) {
        int iters = 2000;
        double best = 0.0, elapsed = 0.0;
        for (int rep = 0; rep < 200 && elapsed < seconds; ++rep) {
            cudaEventRecord(e0, st);
            if (which == 0) k_peak_dfma<<<blocks, threads, 0, st>>>(d, iters, 1.0000001, 1e-9);
            else k_peak_dmma<<<blocks, threads, 0, st>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1, st);
            CU(ctx, cudaEventSynchronize(e1));
            ctx->launches++;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops = (which == 0) ? (double)blocks * threads * iters * 16 * 2
                                              : (double)blocks * (threads / 32) * iters * 8 * 512.0;
            const double tf = flops / (ms * 1e-3) / 1e12;
            if (rep > 0) { best = std::max(best, tf); elapsed += ms * 1e-3; }
            if (ms < 20.0f) iters *= 2;
        }
        if (which == 0 && dfma_tflops) *dfma_tflops = best;
        if (which == 1 && dmma_tflops) *dmma_tflops = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return RG_OK;
}

extern "C" int rg_unitary_and_derivatives(rg_problem* pr, const double* x, double* U, double* U_dx, double* U_dx_add,
                                          double* U_derr, double* U_derr_dx, double* U_derr_dx_add) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (!x) RG_FAIL(ctx, RG_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    const DevProblem& P = pr->dp;
    const size_t DD = (size_t)P.d * P.d, cb = sizeof(cplx);
    const size_t nU = DD, nUdx = DD * P.p * P.N, nUa = DD * P.a, nUe = DD * P.e, nUex = DD * P.p * P.N * P.e, nUea = DD * P.a * P.e;
    const size_t oU = 0, oUdx = oU + nU, oUa = oUdx + nUdx, oUe = oUa + nUa, oUex = oUe + nUe, oUea = oUex + nUex, tot = oUea + nUea;
    if (pr->dX.ensure((size_t)P.nx * 8) || pr->dOut2.ensure(tot * cb)) RG_FAIL(ctx, RG_ERR_NOMEM, "device allocation failed");
    cplx* o = pr->dOut2.as<cplx>();
    CU(ctx, cudaMemsetAsync(o, 0, tot * cb, ctx->stream));
    CU(ctx, cudaMemcpyAsync(pr->dX.p, x, (size_t)P.nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    const DimOps* ops = rg_dim_ops(P.d);
    if (!ops) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
    int rc = ops->materialize(pr, pr->dX.as<double>(), o + oU, o + oUdx, o + oUa, o + oUe, o + oUex, o + oUea);
    if (rc) return rc;
    struct { double* h; size_t off, n; } outs[] = {{U, oU, nU}, {U_dx, oUdx, nUdx}, {U_dx_add, oUa, nUa}, {U_derr, oUe, nUe},
                                                  {U_derr_dx, oUex, nUex}, {U_derr_dx_add, oUea, nUea}};
    for (auto& t : outs)
        if (t.h && t.n) CU(ctx, cudaMemcpyAsync(t.h, o + t.off, t.n * cb, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}
static int interaction_on_device(rg_problem* pr, const double* x) {
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    if (P.e == 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    if (pr->dX.ensure((size_t)P.nx * 8) || pr->dO.ensure((size_t)P.d * P.d * P.N * P.e * sizeof(cplx))) RG_FAIL(ctx, RG_ERR_NOMEM, "device allocation failed");
    CU(ctx, cudaMemcpyAsync(pr->dX.p, x, (size_t)P.nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    const DimOps* ops = rg_dim_ops(P.d);
    if (!ops) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
    return ops->interaction(pr, pr->dX.as<double>(), pr->dO.as<cplx>());
}

extern "C" int rg_interaction_error_operators(rg_problem* pr, const double* x, double* O) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (!x || (!O && pr->dp.e > 0)) RG_FAIL(ctx, RG_ERR_INVALID, "null argument");
    int rc = interaction_on_device(pr, x);
    if (rc) return rc;
    const DevProblem& P = pr->dp;
    if (P.e > 0) CU(ctx, cudaMemcpyAsync(O, pr->dO.p, (size_t)P.d * P.d * P.N * P.e * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}

static int response_common(rg_problem* pr, const double* x, const double* freqs, int nfreq, int first, int count, int M,
                           int shift, double* R) {
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    if (!pr->has_target) RG_FAIL(ctx, RG_ERR_INVALID, "problem has no projector");
    if (count <= 0 || P.e == 0) return RG_OK;
    int rc = interaction_on_device(pr, x);
    if (rc) return rc;
    if (pr->dOut.ensure((size_t)count * P.e * 8) || pr->dFreq.ensure(std::max<size_t>(16, (size_t)nfreq * 8))) RG_FAIL(ctx, RG_ERR_NOMEM, "device allocation failed");
    if (freqs) CU(ctx, cudaMemcpyAsync(pr->dFreq.p, freqs, (size_t)nfreq * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        const DimOps* ops = rg_dim_ops(P.d);
        if (!ops) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
        rc = ops->response(pr, pr->dFreq.as<double>(), first, count, M, shift, pr->dOut.as<double>());
    }
    if (rc) return rc;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(R, pr->dOut.p, (size_t)count * P.e * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}

extern "C" int rg_fidelity_response(rg_problem* pr, const double* x, const double* freqs, int32_t nfreq, int32_t first,
                                    int32_t count, double* R) {
    if (!pr) return RG_ERR_INVALID;
    if (!x || !freqs || !R || first < 0 || count < 0 || first + count > nfreq) RG_FAIL(pr->ctx, RG_ERR_INVALID, "bad frequency range");
    return response_common(pr, x, freqs, nfreq, first, count, 0, 1, R);
}

extern "C" int rg_fidelity_response_fft(rg_problem* pr, const double* x, int32_t oversampling, double* R, double* freqs_out) {
    if (!pr) return RG_ERR_INVALID;
    if (!x || !R || oversampling < 1) RG_FAIL(pr->ctx, RG_ERR_INVALID, "bad arguments");
    const DevProblem& P = pr->dp;
    const int M = P.N * oversampling;
    if (freqs_out) for (int n = 0; n < M; ++n) freqs_out[n] = (2.0 * M_PI / (M * P.dt)) * n;      // :341
    return response_common(pr, x, nullptr, 0, 0, M, M, 0, R);
}

extern "C" int rg_expectation_values(rg_problem* pr, const double* x, double* out) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    if (!x || (!out && P.e > 0)) RG_FAIL(ctx, RG_ERR_INVALID, "null argument");
    if (!pr->has_target) RG_FAIL(ctx, RG_ERR_INVALID, "problem has no projector");
    if (P.e == 0) return RG_OK;
    int rc = interaction_on_device(pr, x);
    if (rc) return rc;
    if (pr->dOut.ensure((size_t)P.N * P.e * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device allocation failed");
    {
        const DimOps* ops = rg_dim_ops(P.d);
        if (!ops) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
        rc = ops->expectation(pr, pr->dOut.as<double>());
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, pr->dOut.p, (size_t)P.N * P.e * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}

