/* robustgrape_b200.h -- C ABI of librobustgrape_b200.so
 *
 * Drop-in boundary for the GRAPE propagator hot path of RobustGRAPE.jl.  The reference
 * has no FFI (it is pure Julia); these entry points are what a `ccall` shim would bind
 * in place of the Julia functions cited on each declaration (paths relative to the
 * reference repository).  julia/RobustGRAPEB200.jl holds that shim; INTEGRATION.md
 * shows the binding.
 *
 * Conventions
 *  - Column-major arrays everywhere; complex numbers are interleaved (re, im) doubles,
 *    i.e. Julia `ComplexF64` / numpy complex128 memory layout.  Tensor index order is
 *    exactly what the reference functions return.
 *  - Every function returns 0 on success or a negative rg_status; nothing throws,
 *    aborts or prints.  rg_last_error() gives a message for the last failure.
 *  - The caller owns all host buffers.  The library owns contexts, device memory and
 *    streams.  A context is bound to one CUDA device and may be used by one host thread
 *    at a time.  Calls with host buffers are blocking.
 *  - There is no CPU fallback: without a CUDA device rg_ctx_create fails.
 */
#ifndef ROBUSTGRAPE_B200_H
#define ROBUSTGRAPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rg_status {
    RG_OK = 0,
    RG_ERR_INVALID = -1,      /* bad argument / shape */
    RG_ERR_CUDA = -2,         /* CUDA runtime error (message in rg_last_error) */
    RG_ERR_UNSUPPORTED = -3,  /* descriptor outside what the kernels implement */
    RG_ERR_NORM = -4,         /* ||dt*H||_1 outside the supported range */
    RG_ERR_NOMEM = -5
} rg_status;

/* ---- declarative operators -------------------------------------------------------
 * The reference's H0 / Herror / target_unitary are opaque closures (src/Types.jl:13,35,55).
 * Across the ABI they are term lists:  op = sum_t coef_t * prod_f factor_f * M_t.        */
enum { RG_F_VAR = 0, RG_F_COS = 1, RG_F_SIN = 2, RG_F_EXPI = 3, RG_F_ERR = 4,
       RG_F_ERR1P_M1 = 5, RG_F_TABLE = 6 };
enum { RG_S_MAIN = 0, RG_S_ADD = 1, RG_S_NONE = 2 };
enum { RG_OWNER_H0 = -1, RG_OWNER_TARGET = -2 };   /* owner >= 0: error source index */
#define RG_MAX_FACTORS 4

typedef struct rg_factor {
    int32_t kind;     /* RG_F_*: VAR u; COS cos u; SIN sin u; EXPI exp(i u); ERR err;
                         ERR1P_M1 fl(1+err)-1; TABLE table[time_step][index]              */
    int32_t space;    /* RG_S_MAIN: x[index] at the current time step; RG_S_ADD: x_add[index] */
    int32_t index;
    int32_t reserved;
    double scale;     /* u = scale * v + offset */
    double offset;
} rg_factor;

typedef struct rg_term {
    int32_t owner;            /* RG_OWNER_H0, RG_OWNER_TARGET or error-source index */
    int32_t nfactors;
    double coef_re, coef_im;
    rg_factor factors[RG_MAX_FACTORS];
    int32_t nnz;
    int32_t reserved;
    const int32_t* rows;      /* 0-based */
    const int32_t* cols;
    const double* vals;       /* nnz interleaved complex */
} rg_term;

/* Fields mirror UnitaryRobustGRAPEProblem / FidelityRobustGRAPEProblem (src/Types.jl:31-40,52-56). */
typedef struct rg_problem_desc {
    int32_t ndim, ntimes, nparam, nb_additional_param, nerr;
    double t0, eps, eps2;
    int32_t nterms;           /* H0 terms and error-source terms */
    const rg_term* terms;
    int32_t ntarget_terms;    /* target_unitary(x_add) terms; 0 for unitary-only problems */
    const rg_term* target_terms;
    const double* projector;  /* ndim x ndim real, column-major; may be NULL if ntarget_terms==0 */
    int32_t ntable_cols;
    const double* table;      /* ntimes x ntable_cols real, column-major; may be NULL */
    int32_t hermitian;        /* 1: H0 + error terms Hermitian for real variables (unitary propagators) */
    int32_t hstack;           /* 1: H-stack problem -- H0 / Herror / target_unitary are opaque closures (src/Types.jl:13,35,55)
                                 evaluated by the caller; nterms = ntarget_terms = 0; use the *_from_hstack entry points    */
} rg_problem_desc;

typedef struct rg_ctx rg_ctx;
typedef struct rg_problem rg_problem;

int rg_ctx_create(rg_ctx** out, int device);
void rg_ctx_destroy(rg_ctx* ctx);
const char* rg_last_error(const rg_ctx* ctx);
/* Use an external CUDA stream (cudaStream_t as void*) for the *_dev entry points; NULL = the context's own. */
int rg_ctx_set_stream(rg_ctx* ctx, void* cuda_stream);
/* Wait for everything enqueued on the context stream; reports deferred kernel-side errors (RG_ERR_NORM). */
int rg_ctx_synchronize(rg_ctx* ctx);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t rg_ctx_launch_count(const rg_ctx* ctx);

/* Per-kernel timing with CUDA events on the launch stream (bench.py's live roofline).
 * Kernel classes: 0 step propagators (k_steps), 1 mixed second differences (k_steps_so), 2 chunk scan,
 * 3 fidelity-gradient sweep, 4 sensitivity-gradient sweep, 5 epilogues, 6 analysis kernels.          */
int rg_ctx_set_timing(rg_ctx* ctx, int enable);
int rg_ctx_get_timing(rg_ctx* ctx, int kernel, int reset, double* ms, int64_t* count);

int rg_problem_create(rg_ctx* ctx, const rg_problem_desc* desc, rg_problem** out);
void rg_problem_destroy(rg_problem* prob);

/* nx = nparam*ntimes + nb_additional_param.  X is (nx, B): one pulse per column, each ordered
 * [x_1(t_1)..x_p(t_1), x_1(t_2).. ; x_add] (src/UnitaryCalculations.jl:21-26).                   */

/* Batched calculate_fidelity_and_derivatives (src/FidelityCalculations.jl:19-119), host buffers.
 * F (B), F_dx (nx,B), F_d2err (nerr,B), F_d2err_dx (nx,nerr,B).  Output pointers may be NULL.   */
int rg_fidelity_and_derivatives_batch(rg_problem* prob, int32_t B, const double* X,
                                      double* F, double* F_dx, double* F_d2err, double* F_d2err_dx);

/* Batched calculate_common! without regularisation (src/FidelityCalculations.jl:174-184):
 * cost = 1-F + sum_e c_e F_d2err[e]^2 ; grad = -F_dx + 2 sum_e c_e F_d2err[e] F_d2err_dx[:,e].
 * err_coeff (nerr, host) may be NULL when nerr==0.  cost (B), grad (nx,B).                       */
int rg_cost_and_grad_batch(rg_problem* prob, int32_t B, const double* X, const double* err_coeff,
                           double* cost, double* grad);

/* Same two calls with DEVICE buffers, enqueued on the context stream without synchronising
 * (err_coeff stays a host pointer).  For pipelines that keep pulses resident in HBM.             */
int rg_fidelity_and_derivatives_batch_dev(rg_problem* prob, int32_t B, const double* dX,
                                          double* dF, double* dF_dx, double* dF_d2err, double* dF_d2err_dx);
int rg_cost_and_grad_batch_dev(rg_problem* prob, int32_t B, const double* dX, const double* err_coeff,
                               double* dcost, double* dgrad);

/* --- the optimiser loop on the device --------------------------------------------------------------
 * Regularisation terms of calculate_common! (src/FidelityCalculations.jl:186-195) as an enumerated device epilogue, one kind per
 * control row:  RG_REG_PLAIN = regularization_cost (src/Regularization.jl:26-47), RG_REG_PHASE = regularization_cost_phase
 * (:111-115), RG_REG_SIN2 = the sin^2-of-differences form the reference's tests use (test/runtests.jl:9-45).
 *   cost += sum_np c1[np] reg1 + c2[np] reg2 ;  grad[main parameters] += c1[np] jac1 + c2[np] jac2.
 * reg_kind / reg_c1 / reg_c2 have nparam entries (host); reg_kind may be NULL (no regularisation).                            */
enum { RG_REG_NONE = 0, RG_REG_PLAIN = 1, RG_REG_PHASE = 2, RG_REG_SIN2 = 3 };
int rg_cost_and_grad_batch_reg(rg_problem* prob, int32_t B, const double* X, const double* err_coeff, const int32_t* reg_kind,
                               const double* reg_c1, const double* reg_c2, double* cost, double* grad);
int rg_cost_and_grad_batch_reg_dev(rg_problem* prob, int32_t B, const double* dX, const double* err_coeff, const int32_t* reg_kind,
                                   const double* reg_c1, const double* reg_c2, double* dcost, double* dgrad);
/* Batched L-BFGS, one independent optimiser per pulse, in place of Optim.optimize(f, g!, x0; method = LBFGS(), iterations, g_tol)
 * (src/FidelityCalculations.jl:199-217) for multi-start runs: the iterates, gradients and curvature history stay in HBM and only a
 * 12-byte progress record crosses PCIe per line-search round.  X (nx, B) is updated in place (host or device buffer); cost (B);
 * iters_out (B, host, may be NULL); info_out (3 ints, host, may be NULL): evaluations, iterations, line-search failures.
 * history <= 32 curvature pairs; stops when every pulse has ||grad||_inf <= g_tol or after `iterations`.                    */
int rg_lbfgs_batch(rg_problem* prob, int32_t B, double* X, const double* err_coeff, const int32_t* reg_kind, const double* reg_c1,
                   const double* reg_c2, int32_t history, int32_t iterations, double g_tol, double* cost, int32_t* iters_out,
                   int32_t* info_out);
int rg_lbfgs_batch_dev(rg_problem* prob, int32_t B, double* dX, const double* err_coeff, const int32_t* reg_kind, const double* reg_c1,
                       const double* reg_c2, int32_t history, int32_t iterations, double g_tol, double* dcost, int32_t* iters_out,
                       int32_t* info_out);

/* calculate_unitary_and_derivatives (src/UnitaryCalculations.jl:20-155), one pulse, materialised:
 * U (d,d), U_dx (d,d,p,N), U_dx_add (d,d,a), U_derr (d,d,e), U_derr_dx (d,d,p,N,e),
 * U_derr_dx_add (d,d,a,e); complex interleaved.  Output pointers may be NULL.                    */
int rg_unitary_and_derivatives(rg_problem* prob, const double* x, double* U, double* U_dx,
                               double* U_dx_add, double* U_derr, double* U_derr_dx, double* U_derr_dx_add);

/* Closure problems: the caller evaluates its closures into the Hamiltonians whose exponentials the reference
 * forms (src/UnitaryCalculations.jl:45-97) and the library does everything after that (exponentials, scans, contractions,
 * reductions).  nvar = nparam + nb_additional_param, n_exp = 1 + 2 nvar + nerr (2 + nvar).
 *   Hstack (d, d, n_exp, N) complex: per step, in this order,
 *       H0(x) | H0(x + eps e_v), v < nvar | H0(x + eps2 e_v) | H0(x) + Herr_e(x, eps), e < nerr | H0(x) + Herr_e(x, eps2) |
 *       H0(x + eps2 e_v) + Herr_e(x + eps2 e_v, eps2)   (index e * nvar + v)
 *     where e_v perturbs main parameter v of this step for v < nparam and additional parameter v - nparam otherwise;
 *   Tstack (d, d, 1 + nb_additional_param) complex: target_unitary(x_add) | target_unitary(x_add + eps e_j).
 * Outputs as rg_fidelity_and_derivatives_batch with B = 1 / rg_unitary_and_derivatives.  ndim <= 10.                          */
int rg_fidelity_and_derivatives_from_hstack(rg_problem* prob, const double* Hstack, const double* Tstack,
                                            double* F, double* F_dx, double* F_d2err, double* F_d2err_dx);
int rg_unitary_and_derivatives_from_hstack(rg_problem* prob, const double* Hstack, double* U, double* U_dx, double* U_dx_add,
                                           double* U_derr, double* U_derr_dx, double* U_derr_dx_add);

/* calculate_interaction_error_operators (src/UnitaryCalculations.jl:180-204): O (d,d,N,e) complex. */
int rg_interaction_error_operators(rg_problem* prob, const double* x, double* O);

/* calculate_fidelity_response (src/FidelityCalculations.jl:246-280): R (nfreq, nerr).
 * Frequencies [first, first+count) of `freqs` are evaluated (shard over ranks); R receives
 * `count` rows per error source laid out (count, nerr).                                          */
int rg_fidelity_response(rg_problem* prob, const double* x, const double* freqs, int32_t nfreq,
                         int32_t first, int32_t count, double* R);

/* calculate_fidelity_response_fft (src/FidelityCalculations.jl:306-343): R (N*os, nerr), freqs_out (N*os). */
int rg_fidelity_response_fft(rg_problem* prob, const double* x, int32_t oversampling,
                             double* R, double* freqs_out);

/* calculate_expectation_values (src/FidelityCalculations.jl:368-390): out (N, nerr). */
int rg_expectation_values(rg_problem* prob, const double* x, double* out);

/* Pinned host allocation helpers so callers can stage X / grad without pageable-copy overhead. */
int rg_host_alloc(void** ptr, uint64_t bytes);
void rg_host_free(void* ptr);

/* --- Gather of per-shard results through peer memory (one process per GPU, one NVLink/NVSwitch node) --------------
 * The reference is a single CPU process and has no counterpart; this is north_star's "allgather of per-shard costs and
 * gradients" for the multi-start batch, written so that it never competes with the evaluation kernels for SMs.
 *   rg_peer_buffer_create : cudaMalloc a buffer on the context device and export its 64-byte CUDA IPC handle.
 *   rg_peer_buffer_open   : map another process's buffer (handle obtained out of band) into this process.
 *   rg_gather_to_peers    : copy src[0, bytes) to peer_base[p] + dst_offset for p < npeers (a base may be the caller's
 *                           own buffer). Ordered after everything queued on the context stream so far; runs on the
 *                           library's side streams, i.e. asynchronously to later work on the context stream.
 *                           mode 0: copy engines (one cudaMemcpyAsync per peer); mode 1: one kernel storing to all peers.
 *                           slot (0/1) names the completion event, for double-buffered sources.
 *   rg_gather_wait        : the context stream waits (on the device) for the gather last issued on `slot`.
 * rg_ctx_synchronize() also waits for outstanding gathers. Cross-process completion (all peers have written into my
 * buffer) is the caller's barrier.                                                                                  */
int rg_peer_buffer_create(rg_ctx* ctx, uint64_t bytes, void** dptr, unsigned char handle[64]);
int rg_peer_buffer_open(rg_ctx* ctx, const unsigned char handle[64], void** dptr);
int rg_peer_buffer_close(rg_ctx* ctx, void* dptr);
int rg_peer_buffer_destroy(rg_ctx* ctx, void* dptr);
int rg_gather_to_peers(rg_ctx* ctx, const void* src, uint64_t bytes, int32_t npeers, void* const* peer_base,
                       uint64_t dst_offset, int32_t slot, int32_t mode);
int rg_gather_wait(rg_ctx* ctx, int32_t slot);
/* Same wait, on a caller-chosen stream (cudaStream_t as void*): used to order a cross-rank barrier after this rank's pushes
 * without stalling the evaluation stream. */
int rg_gather_wait_on(rg_ctx* ctx, int32_t slot, void* cuda_stream);

/* Evaluation and gather in one call: rg_cost_and_grad_batch_dev (calculate_common!, src/FidelityCalculations.jl:174-184, for a batch
 * of device-resident pulses) whose results also land in the gathered buffers of `npeers` other ranks.  peer_base[q] is peer q's
 * gathered buffer (rg_peer_buffer_open); this rank's block [cost (B) | grad (B, nx)] starts `dst_offset` bytes into it; what = 0
 * gathers the costs only (8 B per pulse -- all a sharded multi-start optimisation exchanges), what = 1 costs and gradients
 * (north_star's all-gather).  For block-2 problems without error sources the evaluation kernel itself stores to the peers over
 * NVLink (rg_fusedq.cuh: full-line stores from staged rows, overlapped with the other CTAs' arithmetic); other problems are
 * evaluated and then pushed by the copy engines.  The transfer is ordered on the context stream; completion on the *receiving*
 * ranks is the caller's cross-rank barrier, as for rg_gather_to_peers.  No counterpart in the reference (single CPU process). */
int rg_cost_and_grad_batch_dev_scatter(rg_problem* prob, int32_t B, const double* dX, const double* err_coeff, double* dcost, double* dgrad,
                                       int32_t npeers, void* const* peer_base, uint64_t dst_offset, int32_t what);

/* Which kernel family a cost/gradient evaluation of this problem takes, as a short string (diagnostics, benchmarks and tests; the
 * reference has no counterpart -- there is one generic path, src/UnitaryCalculations.jl:20-155):
 *   "fused_q_pc"  one launch per role, phase-only drive class (step constants evaluated once, one sincos per step and sweep)
 *   "fused_q"     one launch per role, closed-form 2 x 2 blocks recomputed per step
 *   "block2"      workspace-free three-kernel path (blocks with diagonal terms)
 *   "steps_t"     thread-per-step propagators with a step-matrix workspace (d <= 5)
 *   "group"       general shared-memory kernels (scaling and squaring, non-Hermitian, d <= 10)
 *   "dense"       planar DMMA path (d > 10)        "hstack"  host-evaluated Hamiltonian stacks
 * Returns RG_OK; the string is truncated to len - 1 characters. */
int rg_problem_path(rg_problem* prob, char* buf, int32_t len);

/* FP64 peak microbenchmarks (DFMA loop, DMMA m8n8k4 loop) on the context device: TFLOP/s. */
int rg_measure_fp64_peak(rg_ctx* ctx, double seconds, double* dfma_tflops, double* dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* ROBUSTGRAPE_B200_H */
