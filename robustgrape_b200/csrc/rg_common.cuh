// rg_common.cuh -- device-side problem description, complex helpers and the
// finite-difference coefficient algebra shared by all kernels.
//
// Semantics follow the reference's finite-difference formulas
// (src/UnitaryCalculations.jl:45-97): the perturbed input is fl(x + eps) and the quotient
// divides by the nominal eps.  Differences of coefficients are formed analytically
// (e^{i(u+h)} - e^{iu} = e^{iu}(e^{ih} - 1), ...) so that no cancellation is amplified by 1/eps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/robustgrape_b200.h"

#define RG_MAX_VARS 8     // perturbation variables handled per problem (main + dependent additional)
#define RG_MAX_ERR 4
#define RG_MAX_ADD 8
#define RG_MAX_MAIN 8

typedef double2 cplx;

#define RG_MAX_TRIG 2     // distinct trigonometric arguments per step shared between factor evaluations (see TrigSlots)
struct DevFactor { int kind, space, index, slot; double scale, offset; };   // slot: trig slot of this factor's argument, or -1
struct DevTerm { int owner, nf; double cr, ci; DevFactor f[RG_MAX_FACTORS]; };
struct DevEntry { int row, col, term, pad; double vr, vi; };

struct DevProblem {
    int d, N, p, a, e;
    double t0, dt, eps, eps2, inv_eps, inv_eps2sq;
    int nterms; const DevTerm* terms;
    int nent; const DevEntry* ents; const int* colptr;          // H0 + error entries, sorted by column
    int ntt; const DevTerm* tterms; int ntent; const DevEntry* tents; const int* tcolptr;   // target
    const double* PP;    // P0 * P      (d x d, column-major, real)
    const double* PPt;   // (P0 * P)^T
    const double* Pm;    // P = (P0 != 0)
    const double* P0raw; // the projector itself
    double Dtr;          // tr(P0)
    int nvar;            // perturbation variables: main 0..p-1, then additional params H depends on
    int var_space[RG_MAX_VARS], var_index[RG_MAX_VARS];
    int add_var[RG_MAX_ADD];   // index into vars for each additional parameter, or -1
    int ntab; const double* table;
    int hermitian;
    int ntrig;           // trig slots: distinct arguments u = scale * v + offset of the COS/SIN/EXPI factors of H0 and error terms
    int trig_space[RG_MAX_TRIG], trig_index[RG_MAX_TRIG];
    double trig_scale[RG_MAX_TRIG], trig_offset[RG_MAX_TRIG];
    // H-stack problems (closures evaluated on the host, src/Types.jl:13,35,55): per step the n_exp Hamiltonians whose exponentials
    // the reference forms (src/UnitaryCalculations.jl:45-97), column-major d x d complex, order: H0(x) | H0(x + eps e_v) |
    // H0(x + eps2 e_v) | H0 + Herr_e(eps) | H0 + Herr_e(eps2) | H0(x + eps2 e_v) + Herr_e(x + eps2 e_v, eps2) [e major]; and the
    // target stack U0(x_add) | U0(x_add + eps e_j).  Null for descriptor problems.
    const double2* hstack; const double2* tstack; int nexp;
    int mixed_zero;      // 1: no error term depends on a perturbation variable (the mixed slots of dt*H itself vanish)
    int nx;              // p*N + a
    int nstore;          // matrices stored per time step: 1 + nvar + e + nvar*e
    int wsm;             // complex elements stored per step matrix (d*d, or the closure pattern's nnz)
    int wsB;             // pulses in the current launch: the workspace is object-major, ws[obj][b][k][wsm]
    unsigned long long cmask;   // closure pattern of the stored matrices: bit (i + d*j)
    // Phase-only drive class (rg_fusedq.cuh, PC kernels): every upper-triangle entry of H0 and of every error Hamiltonian is
    // (constant | error amplitude) x E(x), with E the same product of EXPI factors everywhere.  Then |w| of every 2 x 2 block is
    // the same at every time step and so are cos/sinc and all their finite differences: they are evaluated once per problem
    // (k_fqc_consts) and a step costs one sincos.
    int pc;                     // 1: class detected on the host
    int pc_nf;                  // EXPI factors of E
    DevFactor pc_f[RG_MAX_FACTORS];
    double pc_vscale[RG_MAX_VARS];   // d(arg of E)/d(var v): scale of the factor that depends on v, 0 if none
    const double2* pc_consts;   // [(1 + 2 e)][2 NB]: reference propagator (a_n, b_n conj(E_ref)); per error source its difference at
                                // eps and at eps2
};

// Destinations of a fused evaluation + gather (rg_cost_and_grad_batch_dev_scatter): besides its own output the kernel stores the
// cost of every pulse -- and, if `grads`, the gradient -- into the gathered buffers of up to RG_MAX_PEER_OUT peers (CUDA IPC
// mappings: plain stores that travel over NVLink).  cost[q] / grad[q] point at this rank's slot in peer q's buffer.
#define RG_MAX_PEER_OUT 15
struct PeerOut { int n, grads; double* cost[RG_MAX_PEER_OUT]; double* grad[RG_MAX_PEER_OUT]; };

// Cost/gradient assembly inside the error-role launch of the fused kernel (calculate_common!, src/FidelityCalculations.jl:178-184):
// with `on`, the launch of error source es0 adds coeff[es0] F2^2 to cost[b] and 2 coeff[es0] F2 dF2/dx to the gradient the
// fidelity-role launch left in `out` (= -dF/dx), instead of writing dF2/dx for a separate epilogue kernel to combine.
struct FQAccum { int on, es0; const double* coeff; double* cost; };

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ cplx cmk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmk(a.x, -a.y); }
// acc += a*b  (4 DFMA)
// 256-bit global accesses (sm_100: LDG.256 / STG.256): two complex doubles per instruction; p must be 32-byte aligned.
// Halves the L1 tag look-ups of the thread-per-chunk sweeps, whose lanes each stream their own 128-byte lines.
__device__ __forceinline__ void ld256(const cplx* __restrict__ p, cplx& a, cplx& b) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
// Software prefetch of one 128-byte line (a compact step matrix) for the thread-per-chunk sweeps: costs no registers,
// unlike loading the next step's operands early. RG_PREFETCH_LEVEL: 0 off, 1 into L1, 2 into L2.
#ifndef RG_PREFETCH_LEVEL
#define RG_PREFETCH_LEVEL 1
#endif
__device__ __forceinline__ void prefetch_line(const void* p) {
#if RG_PREFETCH_LEVEL == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif RG_PREFETCH_LEVEL == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void st256(cplx* __restrict__ p, const cplx a, const cplx b) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cfma_conj(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

// acc -= conj(a) * b   (the lower-triangle partner of a skew-Hermitian element)
__device__ __forceinline__ void cfma_nconj(cplx& acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// sin and cos of one argument for the per-step phase: Cody-Waite reduction by pi/2 in three fused steps, the two minimax
// polynomials of fdlibm's k_sin / k_cos on [-pi/4, pi/4], quadrant fix-up with selects.  Measured against long-double
// sinl/cosl on 2e7 random arguments per range (|x| < 7, 100, 1e5): max error 1.58 ulp.  The library sincos() costs ~250 SASS
// instructions per call here (coefficient tables through LDC, constants through UMOV pairs under the register cap), and after
// the closed-form step constants the two calls per step were half of the kernel; this one is ~35.  |x| > 1e5 takes sincos().
__constant__ double c_sc[16] = {
    0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,                      // 2/pi, pi/2 split in three
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,                                                       // S6 .. S1
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02};                                                      // C6 .. C1
static __device__ __noinline__ double2 rg_sincos_slow(double x) { double2 r; sincos(x, &r.y, &r.x); return r; }
// (coefficients come from the constant bank as DFMA operands: as literals each one costs a UMOV pair per use)
__device__ __forceinline__ void rg_sincos_core(double x, double& s, double& c) {      // branch-free; valid for |x| <= 1e5
    const double MAGIC = 6755399441055744.0;                     // 1.5 * 2^52: the fma rounds x * 2/pi to the nearest integer
    const double t = fma(x, c_sc[0], MAGIC);
    const int q = __double2loint(t);
    const double j = t - MAGIC;
    double r = fma(-j, c_sc[1], x);
    r = fma(-j, c_sc[2], r);
    r = fma(-j, c_sc[3], r);
    const double z = r * r;
    double ps = fma(z, c_sc[4], c_sc[5]);
    ps = fma(z, ps, c_sc[6]); ps = fma(z, ps, c_sc[7]); ps = fma(z, ps, c_sc[8]); ps = fma(z, ps, c_sc[9]);
    const double sr = fma(r * z, ps, r);
    double pc = fma(z, c_sc[10], c_sc[11]);
    pc = fma(z, pc, c_sc[12]); pc = fma(z, pc, c_sc[13]); pc = fma(z, pc, c_sc[14]); pc = fma(z, pc, c_sc[15]);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    // sign flips on the high words: s negative in quadrants 2, 3; c negative in quadrants 1, 2
    s = __hiloint2double(__double2hiint(a) ^ ((q & 2) << 30), __double2loint(a));
    c = __hiloint2double(__double2hiint(b) ^ (((q + 1) & 2) << 30), __double2loint(b));
}
__device__ __forceinline__ void rg_sincos(double x, double& s, double& c) {
    if (fabs(x) > 1.0e5) { const double2 r = rg_sincos_slow(x); s = r.y; c = r.x; return; }
    rg_sincos_core(x, s, c);
}
// ---------------------------------------------------------------------------------------
// Factor value and accurate finite difference.
//   value  f(v)           (err factors use `errv`)
//   delta  f(v + h) - f(v) when the factor depends on the perturbed variable (pspace, pindex)
// h is the step actually taken in the variable, fl(v + eps) - v.
struct EvalCtx {
    const double* xk;      // main parameters at this step (p)
    const double* xadd;    // additional parameters (a)
    double errv;           // error amplitude seen by ERR factors
    const double* table; int N; int k;   // per-step table, 0-based step
    int ntrig;             // > 0: sin/cos of the trig slots were evaluated once for this step (TrigSlots)
    double ts0, tc0, ts1, tc1;
};
// sin/cos of every distinct trigonometric argument of the step, evaluated once and shared by all coefficient evaluations
// of that step (the e^{-i phi} of the drive appears in H0 and in every amplitude-type error term, and each of the
// value / variable / error / mixed evaluations of a step would otherwise call sincos again).
struct TrigSlots { int n; double s0, c0, s1, c1; };
__device__ __forceinline__ TrigSlots trig_eval(const DevProblem& P, const double* xk, const double* xadd);

__device__ inline void factor_eval(const DevFactor& f, const EvalCtx& c, int pspace, int pindex, double h,
                                   cplx& val, cplx& del) {
    del = cmk(0.0, 0.0);
    switch (f.kind) {
    case RG_F_ERR: val = cmk(c.errv, 0.0); return;
    case RG_F_ERR1P_M1: val = cmk((1.0 + c.errv) - 1.0, 0.0); return;
    case RG_F_TABLE: val = cmk(c.table[(size_t)f.index * c.N + c.k], 0.0); return;
    default: break;
    }
    const double v = (f.space == RG_S_MAIN) ? c.xk[f.index] : c.xadd[f.index];
    const bool plain = (f.scale == 1.0 && f.offset == 0.0);
    const double u = plain ? v : fma(f.scale, v, f.offset);
    const bool dep = (f.space == pspace && f.index == pindex);
    const double hu = dep ? f.scale * h : 0.0;
    if (f.kind == RG_F_VAR) { val = cmk(u, 0.0); del = cmk(hu, 0.0); return; }
    double s, co;
    if (c.ntrig > 0 && f.slot >= 0) { s = f.slot == 0 ? c.ts0 : c.ts1; co = f.slot == 0 ? c.tc0 : c.tc1; }
    else rg_sincos(u, s, co);
    // e^{ihu} - 1 = -2 sin^2(hu/2) + i sin(hu)
    double er = 0.0, ei = 0.0;
    if (dep) {
        if (fabs(hu) < 1e-3) {
            // series: truncation < hu^6/5040 relative, i.e. < 2e-22 -- below double rounding
            const double h2 = hu * hu;
            ei = hu * (1.0 - h2 * (1.0 / 6.0) * (1.0 - h2 * (1.0 / 20.0)));
            er = -0.5 * h2 * (1.0 - h2 * (1.0 / 12.0) * (1.0 - h2 * (1.0 / 30.0)));
        } else {
            const double s2 = sin(0.5 * hu);
            ei = sin(hu); er = -2.0 * s2 * s2;
        }
    }
    if (f.kind == RG_F_EXPI) {
        val = cmk(co, s);
        del = cmk(co * er - s * ei, co * ei + s * er);
    } else if (f.kind == RG_F_COS) {
        val = cmk(co, 0.0);
        del = cmk(co * er - s * ei, 0.0);          // Re(e^{iu}(e^{ih}-1))
    } else {                                        // RG_F_SIN
        val = cmk(s, 0.0);
        del = cmk(co * ei + s * er, 0.0);          // Im(e^{iu}(e^{ih}-1))
    }
}

__device__ __forceinline__ TrigSlots trig_eval(const DevProblem& P, const double* xk, const double* xadd) {
#ifndef RG_TRIG_SLOTS
#define RG_TRIG_SLOTS 0      // measured on B200 (C4): sharing sincos through trig slots is slower (0.68 vs 0.57 ms; longer live ranges), so off
#endif
    TrigSlots t{RG_TRIG_SLOTS ? P.ntrig : 0, 0.0, 1.0, 0.0, 1.0};
    if (!RG_TRIG_SLOTS) return t;
    if (P.ntrig > 0) {
        const double v = (P.trig_space[0] == RG_S_MAIN) ? xk[P.trig_index[0]] : xadd[P.trig_index[0]];
        const bool plain = (P.trig_scale[0] == 1.0 && P.trig_offset[0] == 0.0);
        sincos(plain ? v : fma(P.trig_scale[0], v, P.trig_offset[0]), &t.s0, &t.c0);
    }
    if (P.ntrig > 1) {
        const double v = (P.trig_space[1] == RG_S_MAIN) ? xk[P.trig_index[1]] : xadd[P.trig_index[1]];
        const bool plain = (P.trig_scale[1] == 1.0 && P.trig_offset[1] == 0.0);
        sincos(plain ? v : fma(P.trig_scale[1], v, P.trig_offset[1]), &t.s1, &t.c1);
    }
    return t;
}

// Term coefficient and its finite difference in variable (pspace,pindex) with step h:
//   base = coef * prod f ;  delta = coef * (prod f(v+h) - prod f(v))   by telescoping.
__device__ inline void term_coef(const DevTerm& t, const EvalCtx& c, int pspace, int pindex, double h,
                                 cplx& base, cplx& delta) {
    cplx P = cmk(t.cr, t.ci), Dl = cmk(0.0, 0.0);
    for (int i = 0; i < t.nf; ++i) {
        cplx v, dv;
        factor_eval(t.f[i], c, pspace, pindex, h, v, dv);
        // Dl' = Dl * (v + dv) + P * dv ;  P' = P * v
        Dl = cadd(cmul(Dl, cadd(v, dv)), cmul(P, dv));
        P = cmul(P, v);
    }
    base = P; delta = Dl;
}

// column l of -i dt (Hs[obj] - Hs[ref]) (ref < 0: no subtraction) from the H-stack of step k of pulse b
template <int D>
__device__ __forceinline__ void hstack_col(const DevProblem& P, int b, int k, int obj, int ref, cplx* M, int l, bool zero = false) {
    const cplx* base = P.hstack + (((size_t)b * P.N + k) * P.nexp) * (D * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        cplx h = zero ? cmk(0.0, 0.0) : base[(size_t)obj * D * D + i + D * l];
        if (ref >= 0 && !zero) h = csub(h, base[(size_t)ref * D * D + i + D * l]);
        M[i + D * l] = cmk(h.y * P.dt, -h.x * P.dt);
    }
}

// Taylor degree for ||A||_1 <= theta with remainder below 2^-53 (expm section).
__device__ __forceinline__ int taylor_degree(double nrm) {
    // theta_m = ((m+1)! * 2^-53)^(1/(m+1)) with a 10 % margin
    if (nrm <= 1.5e-3) return 4;
    if (nrm <= 6.0e-3) return 5;
    if (nrm <= 1.6e-2) return 6;
    if (nrm <= 3.4e-2) return 7;
    if (nrm <= 6.5e-2) return 8;
    if (nrm <= 0.105) return 9;
    if (nrm <= 0.16) return 10;
    if (nrm <= 0.225) return 11;
    if (nrm <= 0.31) return 12;
    if (nrm <= 0.52) return 14;
    if (nrm <= 0.78) return 16;
    if (nrm <= 1.10) return 18;
    return 99;   // needs scaling and squaring
}
// Scaling-and-squaring plan: exp(A) = (T_m(A / 2^s))^(2^s).  Above the Taylor range the matrix is scaled into
// [0.155, 0.31] (degree 12) and squared s times; differences are squared with
//   d(X^2) = dX (X + dX) + X dX     (exact: (X+dX)^2 - X^2)
// so they stay free of cancellation.  s > RG_MAX_SQUARINGS is reported as RG_ERR_NORM.
#define RG_MAX_SQUARINGS 24
__device__ __forceinline__ void expm_plan(double nrm, int& m, int& s) {
    m = taylor_degree(nrm); s = 0;
    if (m == 99) {
        s = (int)ceil(log2(nrm / 0.31));
        if (s < 1) s = 1;
        m = 12;
        if (s > RG_MAX_SQUARINGS) { s = RG_MAX_SQUARINGS; m = 99; }
    }
}

// ---------------------------------------------------------------------------------------
// Structural pattern of the step matrices.  If H only couples certain pairs of levels, U = exp(-i dt H)
// and all its differences are confined to the reflexive-transitive closure of that coupling graph
// (block-diagonal for the Rydberg models).  CM has bit (i + D*j) set for every element that can be
// non-zero; only those are stored in the HBM workspace, loaded and multiplied.
typedef unsigned long long u64;
template <int D> __host__ __device__ constexpr u64 full_cmask() { return ~0ull; }   // all-ones = dense pattern, any d
__host__ __device__ constexpr int cx_popc(u64 v) { int c = 0; while (v) { v &= v - 1; ++c; } return c; }
// closure of an upper-triangle mask (bit k(k+1)/2 + i, i <= k) as a full d x d pattern
__host__ __device__ constexpr u64 closure_from_tri(int d, unsigned tri) {
    bool r[8][8] = {};
    for (int i = 0; i < d; ++i) r[i][i] = true;
    for (int k = 0; k < d; ++k)
        for (int i = 0; i < k; ++i)
            if ((tri >> (k * (k + 1) / 2 + i)) & 1u) { r[i][k] = true; r[k][i] = true; }
    for (int m = 0; m < d; ++m)
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                if (r[i][m] && r[m][j]) r[i][j] = true;
    u64 out = 0;
    for (int j = 0; j < d; ++j)
        for (int i = 0; i < d; ++i)
            if (r[i][j]) out |= 1ull << (i + d * j);
    return out;
}
// Stored pattern: the closure minus the diagonal of *inert* levels (levels no term touches, not even on the
// diagonal): there U(l,l) = 1 and every difference is 0 exactly, so nothing is computed or stored for them.
// In a stored pattern, level l is inert  <=>  the diagonal bit (l,l) is absent.
__host__ __device__ constexpr u64 stored_from_tri(int d, unsigned tri) {
    u64 cm = closure_from_tri(d, tri);
    for (int l = 0; l < d; ++l) {
        bool touched = false;
        for (int k = 0; k < d; ++k)
            for (int i = 0; i <= k; ++i)
                if (((tri >> (k * (k + 1) / 2 + i)) & 1u) && (i == l || k == l)) touched = true;
        if (!touched) cm &= ~(1ull << (l + d * l));
    }
    return cm;
}
template <int D, u64 CM> struct Pat {
    static constexpr bool full = (CM == full_cmask<D>());
    static constexpr int nnz = full ? D * D : cx_popc(CM);
    // patterns are only representable for d*d <= 64; larger d always use the full (dense) pattern
    __host__ __device__ static constexpr bool has(int i, int j) { return full || ((i + D * j) < 64 && ((CM >> ((i + D * j) & 63)) & 1ull)); }
    __host__ __device__ static constexpr int idx(int i, int j) { return full ? (i + D * j) : cx_popc(CM & ((1ull << ((i + D * j) & 63)) - 1ull)); }
};
// run-time versions (warp-uniform mask from DevProblem)
__device__ __forceinline__ bool pat_has(u64 cm, int d, int i, int j) { return cm == ~0ull || ((cm >> ((i + d * j) & 63)) & 1ull); }
__device__ __forceinline__ int pat_idx(u64 cm, int d, int i, int j) { return cm == ~0ull ? (i + d * j) : __popcll(cm & ((1ull << ((i + d * j) & 63)) - 1ull)); }
