// rg_api.cu -- C ABI of librobustgrape_b200.so (see include/robustgrape_b200.h).
// Host logic only: descriptor flattening, workspace management, kernel sequencing.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

enum { RG_K_STEPS = 0, RG_K_STEPS_SO = 1, RG_K_SCAN = 2, RG_K_GRAD = 3, RG_K_GRAD_ERR = 4, RG_K_EPILOGUE = 5, RG_K_ANALYSIS = 6, RG_K_AGG = 7, RG_NKERNELS = 8 };
#include "rg_smalld.cuh"
#include "rg_steps_t.cuh"
#include "rg_analysis.cuh"
#include "rg_peak.cuh"

struct rg_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int* d_status = nullptr;
    int* h_status = nullptr;    // pinned
    int sm_count = 148;
    size_t ws_limit = (size_t)64 << 30;
    // optional per-kernel timing (CUDA events on the launch stream), see rg_ctx_set_timing
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the pipelined host entry points
    int host_slabs = 4;                               // RG_HOST_SLABS (upper bound; slabs hold >= 2048 pulses)
    bool timing = false;
    struct Span { int kernel; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    double kern_ms[RG_NKERNELS] = {0};
    int64_t kern_n[RG_NKERNELS] = {0};
};

struct KTimer {
    rg_ctx* c; int k; cudaEvent_t e0 = nullptr, e1 = nullptr;
    KTimer(rg_ctx* c_, int k_) : c(c_), k(k_) {
        if (c->timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    }
    ~KTimer() {
        c->launches++;
        if (c->timing) { cudaEventRecord(e1, c->stream); c->spans.push_back({k, e0, e1}); }
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct rg_problem {
    rg_ctx* ctx = nullptr;
    DevProblem dp{};
    std::vector<void*> owned;          // device allocations of the descriptor
    int any_add_dep = 0;
    int chunk_override = 0;
    int force_dense = 0;      // RG_DENSE=1: treat H as dense (no structural-zero skipping)
    int force_group = 0;      // RG_GROUP=1: force the group-per-chunk k_steps kernel
    TriPlanDev tri{};         // upper-triangle assembly plan (Hermitian fast path)
    int tri_ok = 0;
    double tri_density = 1.0;
    unsigned tri_union = 0;   // union of all structural masks
    // workspaces
    DevBuf ws, Qb, Wlb, Cb, Wb, Gb, G1b, H1b, F, F2, addT, addS, F2dx, Fdx, coeff, dX, dOut, dOut2, dO, dFreq, dM;
    int has_target = 0;
};

#define RG_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        char _b[512];                                             \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                    \
        (ctx)->err = _b;                                          \
        return (code);                                            \
    } while (0)

#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            RG_FAIL(ctx, RG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                      \
    } while (0)

static std::string g_global_err;

extern "C" int rg_ctx_create(rg_ctx** out, int device) {
    if (!out) return RG_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_global_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        cudaGetLastError();
        return RG_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_global_err = "device index out of range"; return RG_ERR_INVALID; }
    rg_ctx* c = new rg_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; g_global_err = "cudaSetDevice failed"; return RG_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c; g_global_err = "cudaStreamCreate failed"; return RG_ERR_CUDA;
    }
    c->stream = c->own_stream;
    cudaMalloc(&c->d_status, sizeof(int));
    cudaMemset(c->d_status, 0, sizeof(int));
    cudaMallocHost(&c->h_status, sizeof(int));
    *c->h_status = 0;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    double inv[32];
    inv[0] = 0.0;
    for (int j = 1; j < 32; ++j) inv[j] = 1.0 / j;
    cudaMemcpyToSymbol(c_inv_j, inv, sizeof(inv));
    if (const char* s = getenv("RG_HOST_SLABS")) c->host_slabs = std::max(1, atoi(s));
    cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    if (const char* s = getenv("RG_WS_LIMIT_GB")) c->ws_limit = (size_t)atof(s) * ((size_t)1 << 30);
    if (cudaGetLastError() != cudaSuccess) { g_global_err = "context initialisation failed"; delete c; return RG_ERR_CUDA; }
    *out = c;
    return RG_OK;
}

extern "C" void rg_ctx_destroy(rg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->d_status) cudaFree(c->d_status);
    if (c->h_status) cudaFreeHost(c->h_status);
    delete c;
}

extern "C" const char* rg_last_error(const rg_ctx* c) { return c ? c->err.c_str() : g_global_err.c_str(); }

extern "C" int rg_ctx_set_stream(rg_ctx* c, void* s) {
    if (!c) return RG_ERR_INVALID;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return RG_OK;
}
extern "C" int64_t rg_ctx_launch_count(const rg_ctx* c) { return c ? c->launches : 0; }

extern "C" int rg_ctx_synchronize(rg_ctx* c) {
    if (!c) return RG_ERR_INVALID;
    CU(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (*c->h_status != 0) {
        const int fl = *c->h_status;
        *c->h_status = 0;
        cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream);
        if (fl & 1) RG_FAIL(c, RG_ERR_NORM, "||dt*H||_1 needs more than %d squarings; reduce dt", RG_MAX_SQUARINGS);
        RG_FAIL(c, RG_ERR_NORM, "||dt*H||_1 > 1.1 needs scaling-and-squaring, which the thread-per-step fast path does not do: "
                                "host-buffer entry points fall back automatically; with device buffers set RG_GROUP=1");
    }
    return RG_OK;
}

// Per-kernel timing: when enabled every kernel launch is bracketed by CUDA events on the launch stream.
extern "C" int rg_ctx_set_timing(rg_ctx* c, int enable) {
    if (!c) return RG_ERR_INVALID;
    c->timing = enable != 0;
    return RG_OK;
}
// Collect finished spans (synchronises the stream); returns accumulated milliseconds and launch count
// for kernel class `kernel` (RG_K_*), and clears them when `reset` is set.
extern "C" int rg_ctx_get_timing(rg_ctx* c, int kernel, int reset, double* ms, int64_t* count) {
    if (!c || kernel < 0 || kernel >= RG_NKERNELS) return RG_ERR_INVALID;
    CU(c, cudaStreamSynchronize(c->stream));
    for (auto& s : c->spans) {
        float t = 0;
        cudaEventElapsedTime(&t, s.e0, s.e1);
        c->kern_ms[s.kernel] += t; c->kern_n[s.kernel]++;
        cudaEventDestroy(s.e0); cudaEventDestroy(s.e1);
    }
    c->spans.clear();
    if (ms) *ms = c->kern_ms[kernel];
    if (count) *count = c->kern_n[kernel];
    if (reset) { c->kern_ms[kernel] = 0; c->kern_n[kernel] = 0; }
    return RG_OK;
}

extern "C" int rg_host_alloc(void** p, uint64_t bytes) {
    if (!p) return RG_ERR_INVALID;
    return cudaMallocHost(p, bytes) == cudaSuccess ? RG_OK : RG_ERR_NOMEM;
}
extern "C" void rg_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------
template <class T>
static T* upload(rg_problem* pr, const std::vector<T>& v) {
    T* d = nullptr;
    const size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    if (!v.empty()) cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    pr->owned.push_back(d);
    return d;
}

static bool supported_dim(int d) { return d >= 2 && d <= 9; }

static int flatten_terms(const rg_term* terms, int n, bool target, int d, std::vector<DevTerm>& dt,
                         std::vector<DevEntry>& ents, std::vector<int>& colptr, std::string& why) {
    std::vector<DevEntry> all;
    for (int t = 0; t < n; ++t) {
        const rg_term& s = terms[t];
        if (s.nfactors < 0 || s.nfactors > RG_MAX_FACTORS) { why = "term with too many factors"; return -1; }
        if (target != (s.owner == RG_OWNER_TARGET)) { why = "term owner does not match its list"; return -1; }
        DevTerm o{};
        o.owner = s.owner; o.nf = s.nfactors; o.cr = s.coef_re; o.ci = s.coef_im;
        for (int f = 0; f < s.nfactors; ++f) {
            const rg_factor& ff = s.factors[f];
            if (ff.kind < RG_F_VAR || ff.kind > RG_F_TABLE) { why = "unknown factor kind"; return -1; }
            o.f[f] = DevFactor{ff.kind, ff.space, ff.index, 0, ff.scale, ff.offset};
        }
        for (int z = 0; z < s.nnz; ++z) {
            if (s.rows[z] < 0 || s.rows[z] >= d || s.cols[z] < 0 || s.cols[z] >= d) { why = "matrix entry out of range"; return -1; }
            all.push_back(DevEntry{s.rows[z], s.cols[z], t, 0, s.vals[2 * z], s.vals[2 * z + 1]});
        }
        dt.push_back(o);
    }
    std::stable_sort(all.begin(), all.end(), [](const DevEntry& a, const DevEntry& b) { return a.col < b.col; });
    colptr.assign(d + 1, 0);
    for (auto& e : all) colptr[e.col + 1]++;
    for (int c = 0; c < d; ++c) colptr[c + 1] += colptr[c];
    ents = all;
    return 0;
}

extern "C" int rg_problem_create(rg_ctx* ctx, const rg_problem_desc* desc, rg_problem** out) {
    if (!ctx) return RG_ERR_INVALID;
    if (!desc || !out) RG_FAIL(ctx, RG_ERR_INVALID, "null argument");
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    const int d = desc->ndim;
    if (!supported_dim(d)) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "ndim=%d not supported by the small-d path (2..9)", d);
    if (desc->ntimes < 1) RG_FAIL(ctx, RG_ERR_INVALID, "ntimes must be >= 1");
    if (desc->nparam < 0 || desc->nparam > RG_MAX_MAIN) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "nparam=%d (max %d)", desc->nparam, RG_MAX_MAIN);
    if (desc->nb_additional_param < 0 || desc->nb_additional_param > RG_MAX_ADD)
        RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "nb_additional_param=%d (max %d)", desc->nb_additional_param, RG_MAX_ADD);
    if (desc->nerr < 0 || desc->nerr > RG_MAX_ERR) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "nerr=%d (max %d)", desc->nerr, RG_MAX_ERR);
    if (!(desc->eps > 0) || !(desc->eps2 > 0)) RG_FAIL(ctx, RG_ERR_INVALID, "eps and eps2 must be positive");

    rg_problem* pr = new rg_problem();
    pr->ctx = ctx;
    DevProblem& P = pr->dp;
    P.d = d; P.N = desc->ntimes; P.p = desc->nparam; P.a = desc->nb_additional_param; P.e = desc->nerr;
    P.t0 = desc->t0; P.dt = desc->t0 / desc->ntimes;            // src/UnitaryCalculations.jl:30
    P.eps = desc->eps; P.eps2 = desc->eps2;
    P.inv_eps = 1.0 / desc->eps;                                 // (1/eps) as the reference forms it (:52)
    P.inv_eps2sq = 1.0 / (desc->eps2 * desc->eps2);              // (1/eps2^2) (:80)
    P.nx = P.p * P.N + P.a;
    P.hermitian = desc->hermitian;

    std::string why;
    std::vector<DevTerm> ht, tt;
    std::vector<DevEntry> he, te;
    std::vector<int> hc, tc;
    auto fail = [&](int code, const std::string& m) { ctx->err = m; rg_problem_destroy(pr); return code; };
    if (flatten_terms(desc->terms, desc->nterms, false, d, ht, he, hc, why)) return fail(RG_ERR_INVALID, why);
    if (flatten_terms(desc->target_terms, desc->ntarget_terms, true, d, tt, te, tc, why)) return fail(RG_ERR_INVALID, why);
    if ((int)tt.size() > d * d) return fail(RG_ERR_UNSUPPORTED, "too many target terms");
    for (auto& t : ht)
        if (t.owner < RG_OWNER_H0 || t.owner >= P.e) return fail(RG_ERR_INVALID, "term owner out of range");
    // range checks on variable references
    auto check = [&](const std::vector<DevTerm>& v, bool target) -> bool {
        for (auto& t : v)
            for (int f = 0; f < t.nf; ++f) {
                const DevFactor& ff = t.f[f];
                if (ff.kind <= RG_F_EXPI) {
                    if (ff.space == RG_S_MAIN) { if (target || ff.index < 0 || ff.index >= P.p) return false; }
                    else if (ff.space == RG_S_ADD) { if (ff.index < 0 || ff.index >= P.a) return false; }
                    else return false;
                }
                if (ff.kind == RG_F_TABLE && (ff.index < 0 || ff.index >= desc->ntable_cols || !desc->table)) return false;
            }
        return true;
    };
    if (!check(ht, false) || !check(tt, true)) return fail(RG_ERR_INVALID, "factor references a variable out of range");

    // perturbation variables
    P.nvar = 0;
    for (int i = 0; i < P.p; ++i) { P.var_space[P.nvar] = RG_S_MAIN; P.var_index[P.nvar] = i; P.nvar++; }
    for (int j = 0; j < RG_MAX_ADD; ++j) P.add_var[j] = -1;
    for (int j = 0; j < P.a; ++j) {
        bool dep = false;
        for (auto& t : ht)
            for (int f = 0; f < t.nf; ++f)
                if (t.f[f].kind <= RG_F_EXPI && t.f[f].space == RG_S_ADD && t.f[f].index == j) dep = true;
        if (dep) {
            if (P.nvar >= RG_MAX_VARS) return fail(RG_ERR_UNSUPPORTED, "too many perturbation variables");
            P.add_var[j] = P.nvar;
            P.var_space[P.nvar] = RG_S_ADD; P.var_index[P.nvar] = j; P.nvar++;
            pr->any_add_dep = 1;
        }
    }
    P.nstore = 1 + P.nvar + P.e + P.nvar * P.e;

    P.nterms = (int)ht.size(); P.terms = upload(pr, ht);
    P.nent = (int)he.size(); P.ents = upload(pr, he); P.colptr = upload(pr, hc);
    P.ntt = (int)tt.size(); P.tterms = upload(pr, tt);
    P.ntent = (int)te.size(); P.tents = upload(pr, te); P.tcolptr = upload(pr, tc);
    pr->has_target = (desc->ntarget_terms > 0 && desc->projector != nullptr);

    std::vector<double> P0(d * d, 0.0), Pm(d * d, 0.0), PP(d * d, 0.0), PPt(d * d, 0.0);
    if (desc->projector) {
        for (int i = 0; i < d * d; ++i) { P0[i] = desc->projector[i]; Pm[i] = (P0[i] != 0.0) ? 1.0 : 0.0; }   // src/FidelityCalculations.jl:47-50
    } else {
        for (int i = 0; i < d; ++i) { P0[i + d * i] = 1.0; Pm[i + d * i] = 1.0; }
    }
    double tr = 0.0;
    for (int i = 0; i < d; ++i) tr += P0[i + d * i];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += P0[i + d * k] * Pm[k + d * j];
            PP[i + d * j] = s; PPt[j + d * i] = s;
        }
    P.Dtr = tr;
    P.PP = upload(pr, PP); P.PPt = upload(pr, PPt); P.Pm = upload(pr, Pm); P.P0raw = upload(pr, P0);
    P.ntab = desc->ntable_cols;
    std::vector<double> tab;
    if (desc->table && desc->ntable_cols > 0) tab.assign(desc->table, desc->table + (size_t)desc->ntable_cols * P.N);
    P.table = upload(pr, tab);
    if (const char* s = getenv("RG_CHUNK")) pr->chunk_override = atoi(s);
    if (const char* s = getenv("RG_DENSE")) pr->force_dense = atoi(s);
    if (const char* s = getenv("RG_GROUP")) pr->force_group = atoi(s);
    // ---- upper-triangle plan for the thread-per-step kernel (Hermitian, d <= 5, few terms)
    if (P.hermitian && d <= 5 && P.nterms <= RG_T_MAX_TERMS) {
        const int npos = d * (d + 1) / 2;
        std::vector<std::vector<std::pair<int, std::pair<double, double>>>> lists(npos);
        std::vector<double> colw((size_t)std::max(1, P.nterms) * d, 0.0);
        std::vector<int> used(std::max(1, P.nterms), 0);
        TriPlanDev& tp = pr->tri;
        tp.maskA = 0;
        for (int v = 0; v < RG_MAX_VARS; ++v) tp.maskVar[v] = 0;
        for (int e = 0; e < RG_MAX_ERR; ++e) tp.maskErr[e] = 0;
        for (auto& en : he) {
            colw[(size_t)en.term * d + en.col] += std::sqrt(en.vr * en.vr + en.vi * en.vi);
            if (en.row > en.col) continue;
            const int pos = en.col * (en.col + 1) / 2 + en.row;
            lists[pos].push_back({en.term, {en.vr, en.vi}});
            used[en.term] = 1;
            const DevTerm& t = ht[en.term];
            if (t.owner == RG_OWNER_H0) {
                tp.maskA |= 1u << pos;
                for (int v = 0; v < P.nvar; ++v)
                    for (int f = 0; f < t.nf; ++f)
                        if (t.f[f].kind <= RG_F_EXPI && t.f[f].space == P.var_space[v] && t.f[f].index == P.var_index[v]) tp.maskVar[v] |= 1u << pos;
            } else {
                tp.maskErr[t.owner] |= 1u << pos;
            }
        }
        std::vector<int> ptr(npos + 1, 0), term;
        std::vector<double> val;
        for (int pos = 0; pos < npos; ++pos) {
            for (auto& x : lists[pos]) { term.push_back(x.first); val.push_back(x.second.first); val.push_back(x.second.second); }
            ptr[pos + 1] = (int)term.size();
        }
        tp.nent = (int)term.size();
        tp.ptr = upload(pr, ptr); tp.term = upload(pr, term); tp.val = upload(pr, val);
        tp.colw = upload(pr, colw); tp.used = upload(pr, used);
        unsigned all = tp.maskA;
        for (int v = 0; v < P.nvar; ++v) all |= tp.maskVar[v];
        for (int e = 0; e < P.e; ++e) all |= tp.maskErr[e];
        pr->tri_density = (double)__builtin_popcount(all) / npos;
        pr->tri_union = all;
        pr->tri_ok = 1;
    }
    if (cudaGetLastError() != cudaSuccess || !P.terms || !P.table) return fail(RG_ERR_CUDA, "descriptor upload failed");
    *out = pr;
    return RG_OK;
}

extern "C" void rg_problem_destroy(rg_problem* pr) {
    if (!pr) return;
    cudaSetDevice(pr->ctx->device);
    for (void* p : pr->owned) cudaFree(p);
    DevBuf* bufs[] = {&pr->ws, &pr->Qb, &pr->Wlb, &pr->Cb, &pr->Wb, &pr->Gb, &pr->G1b, &pr->H1b, &pr->F, &pr->F2,
                      &pr->addT, &pr->addS, &pr->F2dx, &pr->Fdx, &pr->coeff, &pr->dX, &pr->dOut, &pr->dOut2, &pr->dO, &pr->dFreq, &pr->dM};
    for (DevBuf* b : bufs) b->release();
    delete pr;
}

// ------------------------------------------------------------------------------------------
struct Plan { int L, nc, slab; };

static Plan make_plan(const rg_problem* pr, int B) {
    const DevProblem& P = pr->dp;
    Plan pl;
    const size_t per_pulse = (size_t)P.N * P.nstore * P.d * P.d * sizeof(cplx);
    size_t slab = std::max<size_t>(1, pr->ctx->ws_limit / std::max<size_t>(per_pulse, 1));
    pl.slab = (int)std::min<size_t>(slab, (size_t)B);
    // enough (pulse, chunk) work items for ~8 waves of resident groups
    const long long target_items = (long long)pr->ctx->sm_count * 72 * 8;
    long long want_nc = (target_items + pl.slab - 1) / pl.slab;
    int L = (int)std::max<long long>(4, std::min<long long>(32, P.N / std::max<long long>(1, want_nc)));
    if (pr->chunk_override > 0) L = pr->chunk_override;
    L = std::min(L, P.N);
    pl.L = L;
    pl.nc = (P.N + L - 1) / L;
    return pl;
}

template <class K>
static int set_smem(rg_ctx* ctx, K kern, size_t bytes) {
    if (bytes > 227 * 1024) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "kernel needs %zu bytes of shared memory", bytes);
    CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RG_OK;
}

// Structural patterns instantiated ahead of time (upper-triangle bit = k(k+1)/2 + i, i <= k).  A problem
// whose union mask fits one of them gets kernels with the zero positions removed at compile time; anything
// else runs the full (dense) instantiation.  (Arbitrary patterns: NVRTC specialisation, see DESIGN.md.)
//   PAT_M5_DRIVE: 5-level symmetric-blockaded Rydberg model (src/RydbergTools.jl:31-39), drive only: (1,3),(2,4)
//   PAT_M5_FULL : same plus the Rydberg detuning diagonal (3,3),(4,4) (frequency error / delta != 0)
enum { PAT_FULL = 0, PAT_M5_DRIVE = 1, PAT_M5_FULL = 2 };
template <int D, int PID> constexpr unsigned tri_mask_of() {
    return (D == 5 && PID == PAT_M5_DRIVE) ? ((1u << 7) | (1u << 12))
         : (D == 5 && PID == PAT_M5_FULL) ? ((1u << 7) | (1u << 12) | (1u << 9) | (1u << 14))
         : ((D * (D + 1) / 2 >= 32) ? ~0u : ((1u << (D * (D + 1) / 2)) - 1u));
}
template <int D, int PID> constexpr u64 cmask_of() {
    return (PID == PAT_FULL) ? full_cmask<D>() : closure_from_tri(D, tri_mask_of<D, PID>());
}

// Run the fused path for one slab of pulses already resident on the device.
//   mode 0: fidelity + derivatives  -> dF, dFdx (+1 scale), dF2, dF2dx
//   mode 1: cost + grad             -> dcost (in dF slot), dgrad (in dFdx slot)
template <int D, int PID>
static int run_slab(rg_problem* pr, int B, const Plan& pl, const double* dX, int mode, const double* d_coeff,
                    double* dF, double* dFdx, double* dF2, double* dF2dx, bool want_grad) {
    rg_ctx* ctx = pr->ctx;
    constexpr u64 CM = cmask_of<D, PID>();
    constexpr int WSM = Pat<D, CM>::nnz;
    DevProblem P = pr->dp;
    P.wsm = WSM; P.cmask = CM;
    constexpr int G = GroupInfo<D>::G;
    const int DD = D * D, ne = P.e, nc = pl.nc, L = pl.L;
    cudaStream_t st = ctx->stream;
    if (!pr->has_target) RG_FAIL(ctx, RG_ERR_INVALID, "problem has no target/projector: fidelity entry points unavailable");
    if (!P.hermitian) RG_FAIL(ctx, RG_ERR_UNSUPPORTED, "non-Hermitian Hamiltonians are not supported by the fused path yet");

    const size_t cb = sizeof(cplx);
    if (pr->ws.ensure((size_t)B * P.N * P.nstore * WSM * cb) || pr->Qb.ensure((size_t)B * nc * DD * cb) ||
        pr->Wlb.ensure(std::max<size_t>(16, (size_t)B * nc * ne * DD * cb)) || pr->Cb.ensure((size_t)B * nc * DD * cb) ||
        pr->Wb.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) || pr->Gb.ensure((size_t)B * nc * DD * cb) ||
        pr->G1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) ||
        pr->H1b.ensure(std::max<size_t>(16, (size_t)B * ne * nc * DD * cb)) ||
        pr->F.ensure((size_t)B * 8) || pr->F2.ensure(std::max<size_t>(16, (size_t)B * ne * 8)) ||
        pr->addT.ensure(std::max<size_t>(16, (size_t)B * (1 + ne) * P.a * 8)) ||
        pr->addS.ensure(std::max<size_t>(16, pr->any_add_dep ? (size_t)B * (1 + ne) * P.a * P.N * 8 : 16)))
        RG_FAIL(ctx, RG_ERR_NOMEM, "device workspace allocation failed (B=%d)", B);
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

    // ---- K1: step propagators + first-order differences (+ chunk aggregates)
    constexpr bool kThreadOK = (D <= 5);
    const bool fast = kThreadOK && pr->tri_ok && !pr->force_group;
    if (!fast && PID != PAT_FULL) RG_FAIL(ctx, RG_ERR_INVALID, "internal: structural pattern without the fast path");
    if (fast) {
        // Hermitian fast path: one thread per time step, triangles in registers; aggregates in a second kernel.
        constexpr int DT = kThreadOK ? D : 2;
        constexpr unsigned UM = tri_mask_of<DT, (kThreadOK ? PID : PAT_FULL)>();
        const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
        const long long items = (long long)B * P.N;
        const int grid = (int)((items + 127) / 128);
        {
            KTimer kt(ctx, RG_K_STEPS);
            k_steps_t<DT, UM><<<grid, 128, smem, st>>>(P, pr->tri, dX, B, pr->ws.as<cplx>(), ctx->d_status);
        }
        const int gs = kagg_group_stride(D, ne);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
        const size_t smem2 = (size_t)wpc * G * gs * cb;
        int rc = set_smem(ctx, k_chunk_agg<D, CM>, smem2);
        if (rc) return rc;
        const long long citems = (long long)B * nc;
        const int grid2 = (int)((citems + (long long)wpc * G - 1) / ((long long)wpc * G));
        KTimer kt(ctx, RG_K_AGG);
        k_chunk_agg<D, CM><<<grid2, wpc * 32, smem2, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(), pr->Wlb.as<cplx>());
    } else {
        const int gs = k1_group_stride(D, P.nterms, ne);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps<D, false>, smem);
        if (rc) return rc;
        const long long items = (long long)B * nc;
        const int grid = (int)((items + (long long)wpc * G - 1) / ((long long)wpc * G));
        KTimer kt(ctx, RG_K_STEPS);
        k_steps<D, false><<<grid, wpc * 32, smem, st>>>(P, dX, B, L, nc, pr->ws.as<cplx>(), pr->Qb.as<cplx>(),
                                                       pr->Wlb.as<cplx>(), ctx->d_status);
    }
    // ---- K1b: mixed second differences (only needed for the sensitivity gradient)
    if (ne > 0 && want_grad && P.nvar > 0 && fast && PID != PAT_FULL) {
        // structured fast path: thread per step, four triangles in registers
        constexpr int DT = kThreadOK ? D : 2;
        constexpr unsigned UM = tri_mask_of<DT, (kThreadOK ? PID : PAT_FULL)>();
        const size_t smem = staged_plan_bytes(P.nterms, pr->tri.nent, D);
        const long long items = (long long)B * P.N;
        KTimer kt(ctx, RG_K_STEPS_SO);
        if constexpr (PID != PAT_FULL)
            k_steps_so_t<DT, UM><<<(int)((items + 127) / 128), 128, smem, st>>>(P, pr->tri, dX, B, pr->ws.as<cplx>(), ctx->d_status);
    } else if (ne > 0 && want_grad && P.nvar > 0) {
        const int gs = k1b_group_stride(D, P.nterms);
        const size_t dbytes = staged_desc_bytes(P.nterms, P.nent, D);
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * G * gs * cb + dbytes > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb + dbytes;
        int rc = set_smem(ctx, k_steps_so<D>, smem);
        if (rc) return rc;
        const long long items = (long long)B * P.N;
        const int grid = (int)((items + (long long)wpc * G - 1) / ((long long)wpc * G));
        KTimer kt(ctx, RG_K_STEPS_SO);
        k_steps_so<D><<<grid, wpc * 32, smem, st>>>(P, dX, B, pr->ws.as<cplx>(), ctx->d_status);
    }
    // ---- K2
    {
        const int gs = k2_group_stride(D);
        int wpc = 2;
        while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * G * gs * cb;
        int rc = set_smem(ctx, k_scan<D>, smem);
        if (rc) return rc;
        dim3 grid((B + wpc * G - 1) / (wpc * G), 1 + ne);
        KTimer kt(ctx, RG_K_SCAN);
        k_scan<D><<<grid, wpc * 32, smem, st>>>(P, dX, B, nc, pr->Qb.as<cplx>(), pr->Wlb.as<cplx>(), pr->Cb.as<cplx>(),
                                               pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(),
                                               iF, iF2, pr->addT.as<double>(), 0, nullptr, nullptr);
    }
    const double DD1 = P.Dtr * (P.Dtr + 1.0);
    // mode 1 with no error sources writes -F_dx straight into grad
    const double sign0 = (mode == 1 && ne == 0) ? -1.0 : 1.0;
    if (want_grad) {
        // ---- K3: backward gradient sweeps (fidelity role, then one role per error source)
        const long long items = (long long)B * nc;
        {
            const int gs = k3_group_stride(D, 1 + P.nvar);
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
            const size_t smem = (size_t)wpc * G * gs * cb;
            int rc = set_smem(ctx, k_grad<D, false, CM>, smem);
            if (rc) return rc;
            dim3 grid((unsigned)((items + (long long)wpc * G - 1) / ((long long)wpc * G)), 1);
            KTimer kt(ctx, RG_K_GRAD);
            k_grad<D, false, CM><<<grid, wpc * 32, smem, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(),
                pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), iFdx,
                sign0 * P.inv_eps / DD1, iF2dx, pr->addS.as<double>());
        }
        if (ne > 0) {
            const int gs = k3_group_stride(D, 2 + 2 * P.nvar);
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * G * gs * cb > 200 * 1024) wpc >>= 1;
            const size_t smem = (size_t)wpc * G * gs * cb;
            int rc = set_smem(ctx, k_grad<D, true, CM>, smem);
            if (rc) return rc;
            dim3 grid((unsigned)((items + (long long)wpc * G - 1) / ((long long)wpc * G)), ne);
            KTimer kt(ctx, RG_K_GRAD_ERR);
            k_grad<D, true, CM><<<grid, wpc * 32, smem, st>>>(P, B, L, nc, pr->ws.as<cplx>(), pr->Cb.as<cplx>(),
                pr->Wb.as<cplx>(), pr->Gb.as<cplx>(), pr->G1b.as<cplx>(), pr->H1b.as<cplx>(), iFdx,
                0.0, iF2dx, pr->addS.as<double>());
        }
        // ---- K4: additional parameters
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

static int dispatch_slab(rg_problem* pr, int B, const Plan& pl, const double* dX, int mode, const double* d_coeff,
                         double* dF, double* dFdx, double* dF2, double* dF2dx, bool want_grad) {
    const bool fast = pr->tri_ok && !pr->force_group && !pr->force_dense;
    if (pr->dp.d == 5 && fast) {
        if ((pr->tri_union & ~tri_mask_of<5, PAT_M5_DRIVE>()) == 0)
            return run_slab<5, PAT_M5_DRIVE>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
        if ((pr->tri_union & ~tri_mask_of<5, PAT_M5_FULL>()) == 0)
            return run_slab<5, PAT_M5_FULL>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
    }
    switch (pr->dp.d) {
#define RG_CASE(D) case D: return run_slab<D, PAT_FULL>(pr, B, pl, dX, mode, d_coeff, dF, dFdx, dF2, dF2dx, want_grad);
        RG_CASE(2) RG_CASE(3) RG_CASE(4) RG_CASE(5) RG_CASE(6) RG_CASE(7) RG_CASE(8) RG_CASE(9)
#undef RG_CASE
    default: break;
    }
    pr->ctx->err = "unsupported ndim";
    return RG_ERR_UNSUPPORTED;
}

// Device-buffer drivers: loop over slabs of pulses so the step workspace stays below ws_limit.
static int run_dev(rg_problem* pr, int B, const double* dX, int mode, const double* h_coeff, double* dF, double* dFdx,
                   double* dF2, double* dF2dx) {
    rg_ctx* ctx = pr->ctx;
    if (B <= 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const DevProblem& P = pr->dp;
    const Plan pl = make_plan(pr, B);
    const double* d_coeff = nullptr;
    if (mode == 1 && P.e > 0) {
        if (!h_coeff) RG_FAIL(ctx, RG_ERR_INVALID, "err_coeff is required when nerr > 0");
        if (pr->coeff.ensure(P.e * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "alloc");
        CU(ctx, cudaMemcpyAsync(pr->coeff.p, h_coeff, P.e * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_coeff = pr->coeff.as<double>();
    }
    const bool want_grad = (mode == 1) || dFdx || dF2dx;
    for (int b0 = 0; b0 < B; b0 += pl.slab) {
        const int bs = std::min(pl.slab, B - b0);
        int rc = dispatch_slab(pr, bs, pl, dX + (size_t)b0 * P.nx, mode, d_coeff, dF ? dF + b0 : nullptr,
                               dFdx ? dFdx + (size_t)b0 * P.nx : nullptr, dF2 ? dF2 + (size_t)b0 * P.e : nullptr,
                               dF2dx ? dF2dx + (size_t)b0 * P.e * P.nx : nullptr, want_grad);
        if (rc) return rc;
    }
    return RG_OK;
}

extern "C" int rg_fidelity_and_derivatives_batch_dev(rg_problem* pr, int32_t B, const double* dX, double* dF,
                                                     double* dF_dx, double* dF_d2err, double* dF_d2err_dx) {
    if (!pr) return RG_ERR_INVALID;
    if (B < 0 || (B > 0 && !dX)) RG_FAIL(pr->ctx, RG_ERR_INVALID, "bad batch arguments");
    return run_dev(pr, B, dX, 0, nullptr, dF, dF_dx, dF_d2err, dF_d2err_dx);
}

extern "C" int rg_cost_and_grad_batch_dev(rg_problem* pr, int32_t B, const double* dX, const double* err_coeff,
                                          double* dcost, double* dgrad) {
    if (!pr) return RG_ERR_INVALID;
    if (B < 0 || (B > 0 && (!dX || !dcost || !dgrad))) RG_FAIL(pr->ctx, RG_ERR_INVALID, "bad batch arguments");
    return run_dev(pr, B, dX, 1, err_coeff, dcost, dgrad, nullptr, nullptr);
}

// Host-buffer entry points.  The batch is cut into slabs and pipelined over three streams:
// H2D of slab i+1 and D2H of slab i-1 overlap the kernels of slab i (full-duplex PCIe), so the
// end-to-end time approaches max(copy-in, compute, copy-out) instead of their sum.
struct SlabPipe {
    rg_ctx* ctx; int nslab, per; std::vector<cudaEvent_t> ev;
    SlabPipe(rg_ctx* c, int B) : ctx(c) {
        nslab = std::max(1, std::min(c->host_slabs, B / 2048));
        per = (B + nslab - 1) / nslab;
        nslab = (B + per - 1) / per;
        ev.resize(2 * nslab);
        for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    ~SlabPipe() { for (auto& e : ev) cudaEventDestroy(e); }
};

extern "C" int rg_fidelity_and_derivatives_batch(rg_problem* pr, int32_t B, const double* X, double* F, double* F_dx,
                                                 double* F_d2err, double* F_d2err_dx) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (B < 0 || (B > 0 && !X)) RG_FAIL(ctx, RG_ERR_INVALID, "bad batch arguments");
    if (B == 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const DevProblem& P = pr->dp;
    const size_t nx = P.nx, ne = P.e;
    const size_t oF = 0, oFdx = oF + B, oF2 = oFdx + (F_dx ? B * nx : 0), oF2dx = oF2 + B * ne,
                 tot = oF2dx + (F_d2err_dx ? B * ne * nx : 0);
    if (pr->dX.ensure(B * nx * 8) || pr->dOut.ensure(std::max<size_t>(16, tot * 8))) RG_FAIL(ctx, RG_ERR_NOMEM, "device staging allocation failed");
    double* o = pr->dOut.as<double>();
    double* dX = pr->dX.as<double>();
    SlabPipe sp(ctx, B);
    // order the copy streams after whatever is already queued on the compute stream
    CU(ctx, cudaEventRecord(sp.ev[0], ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->s_in, sp.ev[0], 0));
    for (int s = 0; s < sp.nslab; ++s) {
        const size_t b0 = (size_t)s * sp.per, bs = std::min<size_t>(sp.per, B - b0);
        CU(ctx, cudaMemcpyAsync(dX + b0 * nx, X + b0 * nx, bs * nx * 8, cudaMemcpyHostToDevice, ctx->s_in));
        CU(ctx, cudaEventRecord(sp.ev[2 * s], ctx->s_in));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, sp.ev[2 * s], 0));
        int rc = run_dev(pr, (int)bs, dX + b0 * nx, 0, nullptr, o + oF + b0, F_dx ? o + oFdx + b0 * nx : nullptr,
                         o + oF2 + b0 * ne, F_d2err_dx ? o + oF2dx + b0 * ne * nx : nullptr);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CU(ctx, cudaEventRecord(sp.ev[2 * s + 1], ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->s_out, sp.ev[2 * s + 1], 0));
        if (F) CU(ctx, cudaMemcpyAsync(F + b0, o + oF + b0, bs * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        if (F_dx) CU(ctx, cudaMemcpyAsync(F_dx + b0 * nx, o + oFdx + b0 * nx, bs * nx * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        if (F_d2err && ne) CU(ctx, cudaMemcpyAsync(F_d2err + b0 * ne, o + oF2 + b0 * ne, bs * ne * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        if (F_d2err_dx && ne) CU(ctx, cudaMemcpyAsync(F_d2err_dx + b0 * ne * nx, o + oF2dx + b0 * ne * nx, bs * ne * nx * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CU(ctx, cudaStreamSynchronize(ctx->s_out));
    int rcs = rg_ctx_synchronize(ctx);
    if (rcs == RG_ERR_NORM && pr->tri_ok && !pr->force_group) {      // fast path out of range: general kernels square
        pr->force_group = 1;
        return rg_fidelity_and_derivatives_batch(pr, B, X, F, F_dx, F_d2err, F_d2err_dx);
    }
    return rcs;
}

extern "C" int rg_cost_and_grad_batch(rg_problem* pr, int32_t B, const double* X, const double* err_coeff,
                                      double* cost, double* grad) {
    if (!pr) return RG_ERR_INVALID;
    rg_ctx* ctx = pr->ctx;
    if (B < 0 || (B > 0 && (!X || !cost || !grad))) RG_FAIL(ctx, RG_ERR_INVALID, "bad batch arguments");
    if (B == 0) return RG_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const DevProblem& P = pr->dp;
    const size_t nx = P.nx;
    if (pr->dX.ensure(B * nx * 8) || pr->dOut.ensure((B + B * nx) * 8)) RG_FAIL(ctx, RG_ERR_NOMEM, "device staging allocation failed");
    double* o = pr->dOut.as<double>();
    double* dX = pr->dX.as<double>();
    SlabPipe sp(ctx, B);
    CU(ctx, cudaEventRecord(sp.ev[0], ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->s_in, sp.ev[0], 0));
    for (int s = 0; s < sp.nslab; ++s) {
        const size_t b0 = (size_t)s * sp.per, bs = std::min<size_t>(sp.per, B - b0);
        CU(ctx, cudaMemcpyAsync(dX + b0 * nx, X + b0 * nx, bs * nx * 8, cudaMemcpyHostToDevice, ctx->s_in));
        CU(ctx, cudaEventRecord(sp.ev[2 * s], ctx->s_in));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, sp.ev[2 * s], 0));
        int rc = run_dev(pr, (int)bs, dX + b0 * nx, 1, err_coeff, o + b0, o + B + b0 * nx, nullptr, nullptr);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CU(ctx, cudaEventRecord(sp.ev[2 * s + 1], ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->s_out, sp.ev[2 * s + 1], 0));
        CU(ctx, cudaMemcpyAsync(cost + b0, o + b0, bs * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpyAsync(grad + b0 * nx, o + B + b0 * nx, bs * nx * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CU(ctx, cudaStreamSynchronize(ctx->s_out));
    int rcs = rg_ctx_synchronize(ctx);
    if (rcs == RG_ERR_NORM && pr->tri_ok && !pr->force_group) {
        pr->force_group = 1;
        return rg_cost_and_grad_batch(pr, B, X, err_coeff, cost, grad);
    }
    return rcs;
}

#include "rg_api_analysis.inl"
