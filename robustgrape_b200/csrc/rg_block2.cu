// rg_block2.cu -- pattern selection and dispatch of the workspace-free paths (rg_block2.cuh, rg_fusedq.cuh).  The kernels
// themselves are instantiated one pattern per translation unit (rg_b2_p*.cu, rg_fq_p*.cu): the jets unroll into long
// straight-line code, and separate units compile in parallel.
#include "rg_host.cuh"
#include "rg_block2_patterns.cuh"

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
static const B2Ops* b2_ops(const rg_problem* pr) {
    switch (rg_b2_pattern(pr)) {
    case 1: return &rg_b2_ops_p1; case 2: return &rg_b2_ops_p2; case 3: return &rg_b2_ops_p3; case 4: return &rg_b2_ops_p4;
    default: return nullptr;
    }
}
#define B2_OPS_OR_FAIL                                                                                           \
    const B2Ops* ops = b2_ops(pr);                                                                               \
    if (!ops) { pr->ctx->err = "internal: block-2 path without an eligible pattern"; return RG_ERR_INVALID; }
int rg_b2_launch_agg(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX) {
    B2_OPS_OR_FAIL
    return ops->agg(pr, P, B, L, nc, dX);
}
int rg_b2_launch_grad(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out0, double scale0) {
    B2_OPS_OR_FAIL
    return ops->grad(pr, P, B, L, nc, dX, out0, scale0);
}
int rg_b2_launch_grad_err(rg_problem* pr, const DevProblem& P, int B, int L, int nc, const double* dX, double* out1) {
    B2_OPS_OR_FAIL
    return ops->grad_err(pr, P, B, L, nc, dX, out1);
}
void rg_b2_occupancy(const rg_problem* pr, int* agg_ctas, int* grad_ctas) {
    int a = 1, g = 1;
    if (const B2Ops* ops = b2_ops(pr)) ops->occupancy(pr, &a, &g);
    *agg_ctas = std::max(1, a); *grad_ctas = std::max(1, g);
}
// one-launch fused quaternion path: patterns without diagonal terms
int rg_fq_pattern(const rg_problem* pr) {
    const int p = rg_b2_pattern(pr);
    return (p == 1 || p == 3) ? p : 0;
}
int rg_fq_launch(rg_problem* pr, const DevProblem& P, int B, const double* dX, int err_role, double* Fout, int fmode, double* out,
                 double scale0, double scale0T, int do_grad, const PeerOut* po, const FQAccum* ac) {
    switch (rg_fq_pattern(pr)) {
    case 1: return rg_fq_ops_p1.launch(pr, P, B, dX, err_role, Fout, fmode, out, scale0, scale0T, do_grad, po, ac);
    case 3: return rg_fq_ops_p3.launch(pr, P, B, dX, err_role, Fout, fmode, out, scale0, scale0T, do_grad, po, ac);
    default: pr->ctx->err = "internal: fused quaternion path without an eligible pattern"; return RG_ERR_INVALID;
    }
}
// evaluates the phase-only constants if the problem is in that class (clears dp.pc when they are out of range)
int rg_fq_prepare(rg_problem* pr) {
    switch (rg_fq_pattern(pr)) {
    case 1: return rg_fq_ops_p1.prepare(pr);
    case 3: return rg_fq_ops_p3.prepare(pr);
    default: return RG_OK;
    }
}
