// rg_api_analysis.inl -- host side of the analysis entry points (included by rg_api.cu).

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

// ---- calculate_unitary_and_derivatives, materialised (one pulse) ---------------------------------------
template <int D>
static int materialize_impl(rg_problem* pr, const double* dx, cplx* dU, cplx* dU_dx, cplx* dU_dx_add, cplx* dU_derr,
                            cplx* dU_derr_dx, cplx* dU_derr_dx_add) {
    rg_ctx* ctx = pr->ctx;
    constexpr u64 CM = full_cmask<D>();
    DevProblem P = pr->dp;
    P.wsm = D * D; P.cmask = CM;
    constexpr int G = GroupInfo<D>::G;
    const int DD = D * D, ne = P.e;
    const size_t cb = sizeof(cplx);
    cudaStream_t st = ctx->stream;
    if (!P.hermitian) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "non-Hermitian Hamiltonians are not supported yet");
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
        int rc = set_smem(ctx, k_steps<D, false>, smem);
        if (rc) return rc;
        KTimer kt(ctx, RG_K_STEPS);
        k_steps<D, false><<<(nc + wpc * G - 1) / (wpc * G), wpc * 32, smem, st>>>(P, dx, 1, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(),
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
        const int nload_max = (ne > 0) ? (2 + 2 * P.nvar) : (1 + P.nvar);
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
    int rc = RG_ERR_UNSUPPORTED;
    switch (P.d) {
#define RG_CASE(D) case D: rc = materialize_impl<D>(pr, pr->dX.as<double>(), o + oU, o + oUdx, o + oUa, o + oUe, o + oUex, o + oUea); break;
        RG_CASE(2) RG_CASE(3) RG_CASE(4) RG_CASE(5) RG_CASE(6) RG_CASE(7) RG_CASE(8) RG_CASE(9)
#undef RG_CASE
    default: RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
    }
    if (rc) return rc;
    struct { double* h; size_t off, n; } outs[] = {{U, oU, nU}, {U_dx, oUdx, nUdx}, {U_dx_add, oUa, nUa}, {U_derr, oUe, nUe},
                                                  {U_derr_dx, oUex, nUex}, {U_derr_dx_add, oUea, nUea}};
    for (auto& t : outs)
        if (t.h && t.n) CU(ctx, cudaMemcpyAsync(t.h, o + t.off, t.n * cb, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}
// ---- interaction-picture error operators on the device (shared by the three analysis entry points)
template <int D>
static int launch_interaction(rg_problem* pr, const double* dx, cplx* dO) {
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    const size_t smem = staged_desc_bytes(P.nterms, P.nent, D) + (size_t)(5 * D * D + 2 * P.nterms) * sizeof(cplx);
    int rc = set_smem(ctx, k_interaction_ops<D>, smem);
    if (rc) return rc;
    KTimer kt(ctx, RG_K_ANALYSIS);
    k_interaction_ops<D><<<1, 32, smem, ctx->stream>>>(P, dx, dO, ctx->d_status);
    return RG_OK;
}
static int interaction_on_device(rg_problem* pr, const double* x) {
    rg_ctx* ctx = pr->ctx;
    const DevProblem& P = pr->dp;
    if (!P.hermitian) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "non-Hermitian Hamiltonians are not supported yet");
    if (P.e == 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    if (pr->dX.ensure((size_t)P.nx * 8) || pr->dO.ensure((size_t)P.d * P.d * P.N * P.e * sizeof(cplx))) RG_FAIL(ctx, RG_ERR_NOMEM, "device allocation failed");
    CU(ctx, cudaMemcpyAsync(pr->dX.p, x, (size_t)P.nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    switch (P.d) {
#define RG_CASE(D) case D: return launch_interaction<D>(pr, pr->dX.as<double>(), pr->dO.as<cplx>());
        RG_CASE(2) RG_CASE(3) RG_CASE(4) RG_CASE(5) RG_CASE(6) RG_CASE(7) RG_CASE(8) RG_CASE(9)
#undef RG_CASE
    default: break;
    }
    RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
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

template <int D>
static int launch_response(rg_problem* pr, const double* dfreqs, int first, int count, int M, int shift, double* dR) {
    rg_ctx* ctx = pr->ctx;
    KTimer kt(ctx, RG_K_ANALYSIS);
    dim3 grid(count, pr->dp.e);
    k_response<D><<<grid, 128, 0, ctx->stream>>>(pr->dp, pr->dO.as<cplx>(), dfreqs, first, count, M, shift, dR);
    return RG_OK;
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
    switch (P.d) {
#define RG_CASE(D) case D: rc = launch_response<D>(pr, pr->dFreq.as<double>(), first, count, M, shift, pr->dOut.as<double>()); break;
        RG_CASE(2) RG_CASE(3) RG_CASE(4) RG_CASE(5) RG_CASE(6) RG_CASE(7) RG_CASE(8) RG_CASE(9)
#undef RG_CASE
    default: RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
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

template <int D>
static int launch_expectation(rg_problem* pr, double* dOut) {
    KTimer kt(pr->ctx, RG_K_ANALYSIS);
    k_expectation<D><<<pr->dp.e, 32, 0, pr->ctx->stream>>>(pr->dp, pr->dO.as<cplx>(), dOut);
    return RG_OK;
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
    switch (P.d) {
#define RG_CASE(D) case D: rc = launch_expectation<D>(pr, pr->dOut.as<double>()); break;
        RG_CASE(2) RG_CASE(3) RG_CASE(4) RG_CASE(5) RG_CASE(6) RG_CASE(7) RG_CASE(8) RG_CASE(9)
#undef RG_CASE
    default: RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "unsupported ndim");
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, pr->dOut.p, (size_t)P.N * P.e * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return rg_ctx_synchronize(ctx);
}

