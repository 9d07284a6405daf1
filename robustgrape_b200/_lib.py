"""ctypes binding of librobustgrape_b200.so (include/robustgrape_b200.h).

The product path fails loudly when the CUDA library is missing or no GPU is present:
there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import descriptors as D

# RG_LIB_PATH: load a differently-built copy of the library (kernel A/B experiments); default = the in-tree build
LIB_PATH = Path(os.environ.get("RG_LIB_PATH") or Path(__file__).resolve().parent / "lib" / "librobustgrape_b200.so")

RG_OK, RG_ERR_INVALID, RG_ERR_CUDA, RG_ERR_UNSUPPORTED, RG_ERR_NORM, RG_ERR_NOMEM = 0, -1, -2, -3, -4, -5
MAX_FACTORS = 4


class RGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librobustgrape_b200 error {code}: {msg}")
        self.code = code


class rg_factor(C.Structure):
    _fields_ = [("kind", C.c_int32), ("space", C.c_int32), ("index", C.c_int32), ("reserved", C.c_int32),
                ("scale", C.c_double), ("offset", C.c_double)]


class rg_term(C.Structure):
    _fields_ = [("owner", C.c_int32), ("nfactors", C.c_int32), ("coef_re", C.c_double), ("coef_im", C.c_double),
                ("factors", rg_factor * MAX_FACTORS), ("nnz", C.c_int32), ("reserved", C.c_int32),
                ("rows", C.POINTER(C.c_int32)), ("cols", C.POINTER(C.c_int32)), ("vals", C.POINTER(C.c_double))]


class rg_problem_desc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("ntimes", C.c_int32), ("nparam", C.c_int32),
                ("nb_additional_param", C.c_int32), ("nerr", C.c_int32),
                ("t0", C.c_double), ("eps", C.c_double), ("eps2", C.c_double),
                ("nterms", C.c_int32), ("terms", C.POINTER(rg_term)),
                ("ntarget_terms", C.c_int32), ("target_terms", C.POINTER(rg_term)),
                ("projector", C.POINTER(C.c_double)),
                ("ntable_cols", C.c_int32), ("table", C.POINTER(C.c_double)),
                ("hermitian", C.c_int32), ("hstack", C.c_int32)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# every symbol declared in include/robustgrape_b200.h, with its signature
SIGNATURES = {
    "rg_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "rg_ctx_destroy": (None, [_vp]),
    "rg_last_error": (C.c_char_p, [_vp]),
    "rg_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "rg_ctx_synchronize": (C.c_int, [_vp]),
    "rg_ctx_launch_count": (C.c_int64, [_vp]),
    "rg_ctx_set_timing": (C.c_int, [_vp, C.c_int]),
    "rg_ctx_get_timing": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.POINTER(C.c_int64)]),
    "rg_problem_create": (C.c_int, [_vp, C.POINTER(rg_problem_desc), C.POINTER(_vp)]),
    "rg_problem_destroy": (None, [_vp]),
    "rg_fidelity_and_derivatives_batch": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "rg_cost_and_grad_batch": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "rg_fidelity_and_derivatives_batch_dev": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "rg_cost_and_grad_batch_dev": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "rg_cost_and_grad_batch_reg": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_cost_and_grad_batch_reg_dev": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_lbfgs_batch": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_double, _vp, _vp, _vp]),
    "rg_lbfgs_batch_dev": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_double, _vp, _vp, _vp]),
    "rg_unitary_and_derivatives": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_fidelity_and_derivatives_from_hstack": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_unitary_and_derivatives_from_hstack": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_interaction_error_operators": (C.c_int, [_vp, _vp, _vp]),
    "rg_fidelity_response": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "rg_fidelity_response_fft": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp]),
    "rg_expectation_values": (C.c_int, [_vp, _vp, _vp]),
    "rg_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "rg_host_free": (None, [_vp]),
    "rg_peer_buffer_create": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp), C.c_char_p]),
    "rg_peer_buffer_open": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "rg_peer_buffer_close": (C.c_int, [_vp, _vp]),
    "rg_peer_buffer_destroy": (C.c_int, [_vp, _vp]),
    "rg_gather_to_peers": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int32, C.POINTER(_vp), C.c_uint64, C.c_int32, C.c_int32]),
    "rg_gather_wait": (C.c_int, [_vp, C.c_int32]),
    "rg_gather_wait_on": (C.c_int, [_vp, C.c_int32, _vp]),
    "rg_measure_fp64_peak": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "rg_problem_path": (C.c_int, [_vp, C.c_char_p, C.c_int32]),
    "rg_cost_and_grad_batch_dev_scatter": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_int32, C.POINTER(C.c_void_p), C.c_uint64, C.c_int32]),
}

_lib = None


def load_library():
    """dlopen the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RGError(RG_ERR_CUDA, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Context:
    """One CUDA device (rg_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.rg_ctx_create(C.byref(h), int(device))
        if rc != 0:
            raise RGError(rc, (self.lib.rg_last_error(None) or b"").decode())
        self.handle = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise RGError(rc, (self.lib.rg_last_error(self.handle) or b"").decode())

    def set_stream(self, cuda_stream):
        self.check(self.lib.rg_ctx_set_stream(self.handle, _vp(cuda_stream)))

    def synchronize(self):
        self.check(self.lib.rg_ctx_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.rg_ctx_launch_count(self.handle))

    KERNELS = ("k_steps", "k_steps_so", "k_scan", "k_grad", "k_grad_err", "epilogue", "analysis", "k_chunk_agg")

    def set_timing(self, enable):
        self.check(self.lib.rg_ctx_set_timing(self.handle, 1 if enable else 0))

    def get_timing(self, reset=True):
        """{kernel class: (total ms, launches)} measured with CUDA events on the launch stream."""
        out = {}
        for i, name in enumerate(self.KERNELS):
            ms, n = C.c_double(), C.c_int64()
            self.check(self.lib.rg_ctx_get_timing(self.handle, i, 1 if reset else 0, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def measure_fp64_peak(self, seconds=0.5):
        a, b = C.c_double(), C.c_double()
        self.check(self.lib.rg_measure_fp64_peak(self.handle, float(seconds), C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- peer-memory gather (include/robustgrape_b200.h: rg_peer_buffer_*, rg_gather_to_peers)
    def peer_buffer_create(self, nbytes):
        """cudaMalloc `nbytes` on this device; returns (device pointer, 64-byte IPC handle)."""
        p = _vp()
        h = C.create_string_buffer(64)
        self.check(self.lib.rg_peer_buffer_create(self.handle, int(nbytes), C.byref(p), h))
        return p.value, h.raw

    def peer_buffer_open(self, handle):
        p = _vp()
        self.check(self.lib.rg_peer_buffer_open(self.handle, bytes(handle), C.byref(p)))
        return p.value

    def peer_buffer_close(self, ptr):
        self.check(self.lib.rg_peer_buffer_close(self.handle, _vp(ptr)))

    def peer_buffer_destroy(self, ptr):
        self.check(self.lib.rg_peer_buffer_destroy(self.handle, _vp(ptr)))

    def gather_to_peers(self, src_ptr, nbytes, peer_ptrs, dst_offset, slot=0, mode=0):
        arr = (_vp * len(peer_ptrs))(*[_vp(p) for p in peer_ptrs])
        self.check(self.lib.rg_gather_to_peers(self.handle, _vp(src_ptr), int(nbytes), len(peer_ptrs), arr, int(dst_offset),
                                               int(slot), int(mode)))

    def gather_wait(self, slot=0):
        self.check(self.lib.rg_gather_wait(self.handle, int(slot)))

    def gather_wait_on(self, slot, cuda_stream):
        self.check(self.lib.rg_gather_wait_on(self.handle, int(slot), _vp(cuda_stream)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# --------------------------------------------------------------------------------------------
def _terms_to_c(terms, keep):
    arr = (rg_term * max(1, len(terms)))()
    for i, t in enumerate(terms):
        ct = arr[i]
        ct.owner = int(t.owner)
        ct.nfactors = len(t.factors)
        ct.coef_re, ct.coef_im = float(np.real(t.coef)), float(np.imag(t.coef))
        for j, f in enumerate(t.factors):
            cf = ct.factors[j]
            cf.kind, cf.space, cf.index = int(f.kind), int(f.space), int(f.index)
            cf.scale, cf.offset = float(f.scale), float(f.offset)
        rows = np.array([e[0] for e in t.entries], dtype=np.int32)
        cols = np.array([e[1] for e in t.entries], dtype=np.int32)
        vals = np.array([[e[2].real, e[2].imag] for e in t.entries], dtype=np.float64).reshape(-1)
        keep += [rows, cols, vals]
        ct.nnz = len(t.entries)
        ct.rows = rows.ctypes.data_as(C.POINTER(C.c_int32))
        ct.cols = cols.ctypes.data_as(C.POINTER(C.c_int32))
        ct.vals = vals.ctypes.data_as(_dp)
    return arr


class DescriptorError(TypeError):
    pass


def _require_terms(obj, cls, what):
    if not isinstance(obj, cls):
        raise DescriptorError(
            f"{what} must be a robustgrape_b200.descriptors.{cls.__name__} (a declarative term list): the CUDA path "
            f"cannot call an opaque closure (got {type(obj).__name__})")
    return obj


class Problem:
    """Device-resident problem (rg_problem) built from the reference-style problem structs."""

    def __init__(self, problem, ctx=None):
        from .types import FidelityRobustGRAPEProblem, UnitaryRobustGRAPEProblem
        self.ctx = ctx or default_context()
        if isinstance(problem, FidelityRobustGRAPEProblem):
            up, fp = problem.unitary_problem, problem
        elif isinstance(problem, UnitaryRobustGRAPEProblem):
            up, fp = problem, None
        else:
            raise TypeError("expected a UnitaryRobustGRAPEProblem or FidelityRobustGRAPEProblem")
        self.up, self.fp = up, fp
        H0 = _require_terms(up.H0, D.TermHamiltonian, "H0")
        if H0.ndim != up.ndim:
            raise ValueError("H0.ndim does not match problem.ndim")
        terms = list(H0.terms)
        table = H0.table
        for e, src in enumerate(up.error_sources):
            He = _require_terms(src.Herror, D.TermErrorHamiltonian, f"error_sources[{e}].Herror")
            for t in He.terms:
                terms.append(D.Term(t.coef, t.factors, t.entries, e))
            if He.table is not None:
                if table is not None and He.table is not table:
                    raise ValueError("H0 and error sources must share one per-step table")
                table = He.table
        herm = H0.is_hermitian() and all(s.Herror.is_hermitian() for s in up.error_sources)
        self._keep = []
        desc = rg_problem_desc()
        desc.ndim, desc.ntimes = int(up.ndim), int(up.ntimes)
        desc.nb_additional_param, desc.nerr = int(up.nb_additional_param), len(up.error_sources)
        nmain = 1 + max([f.index for t in terms for f in t.factors if f.space == D.S_MAIN and f.kind <= D.F_EXPI] + [-1])
        self.nparam_min = nmain
        desc.nparam = -1     # filled in per call (nparam is implied by len(x) in the reference)
        desc.t0, desc.eps, desc.eps2 = float(up.t0), float(up.eps), float(up.eps2)
        self._terms = _terms_to_c(terms, self._keep)
        desc.nterms, desc.terms = len(terms), self._terms
        if fp is not None:
            tgt = _require_terms(fp.target_unitary, D.TermTarget, "target_unitary")
            tterms = [D.Term(t.coef, t.factors, t.entries, D.OWNER_TARGET) for t in tgt.terms]
            self._tterms = _terms_to_c(tterms, self._keep)
            desc.ntarget_terms, desc.target_terms = len(tterms), self._tterms
            proj = np.asfortranarray(np.asarray(fp.projector, dtype=np.float64))
            if proj.shape != (up.ndim, up.ndim):
                raise ValueError("projector must be ndim x ndim")
            self._keep.append(proj)
            desc.projector = proj.ctypes.data_as(_dp)
        if table is not None:
            tab = np.asfortranarray(np.asarray(table, dtype=np.float64))
            if tab.shape[0] != up.ntimes:
                raise ValueError("table must have ntimes rows")
            self._keep.append(tab)
            desc.ntable_cols, desc.table = tab.shape[1], tab.ctypes.data_as(_dp)
        desc.hermitian = 1 if herm else 0
        self._desc = desc
        self._handles = {}     # nparam -> rg_problem*
        self.ndim, self.ntimes = int(up.ndim), int(up.ntimes)
        self.na, self.nerr = int(up.nb_additional_param), len(up.error_sources)

    # nparam is not a field of the reference structs: it is derived from len(x) at call time
    # (src/UnitaryCalculations.jl:21-25), so device problems are cached per nparam.
    def handle_for(self, nx):
        nmain = nx - self.na
        if nmain < 0 or nmain % self.ntimes != 0:
            raise AssertionError("Control parameter size must be a multiple of time steps")
        p = nmain // self.ntimes
        if p not in self._handles:
            if p < self.nparam_min:
                raise ValueError(f"H0 references x[{self.nparam_min - 1}] but only {p} control parameters per step were given")
            self._desc.nparam = p
            h = _vp()
            self.ctx.check(self.ctx.lib.rg_problem_create(self.ctx.handle, C.byref(self._desc), C.byref(h)))
            self._handles[p] = h
        return self._handles[p], p

    def path(self, nx):
        """Kernel family an evaluation with len(x) == nx takes (include/robustgrape_b200.h: rg_problem_path)."""
        h, _ = self.handle_for(nx)
        buf = C.create_string_buffer(32)
        self.ctx.check(self.ctx.lib.rg_problem_path(h, buf, 32))
        return buf.value.decode()

    def close(self):
        for h in self._handles.values():
            self.ctx.lib.rg_problem_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass

    # ---- batched hot path -------------------------------------------------------------------
    def _batch_in(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        X = np.asfortranarray(X)
        return X, X.shape[0], X.shape[1]

    def fidelity_and_derivatives_batch(self, X, want_grad=True):
        """X: (nx, B).  Returns F (B), F_dx (nx,B), F_d2err (nerr,B), F_d2err_dx (nx,nerr,B)."""
        X, nx, B = self._batch_in(X)
        h, p = self.handle_for(nx)
        F = np.zeros(B)
        F2 = np.zeros((self.nerr, B), order="F")
        Fdx = np.zeros((nx, B), order="F") if want_grad else None
        F2dx = np.zeros((nx, self.nerr, B), order="F") if want_grad else None
        self.ctx.check(self.ctx.lib.rg_fidelity_and_derivatives_batch(h, B, _ptr(X), _ptr(F), _ptr(Fdx), _ptr(F2), _ptr(F2dx)))
        return F, Fdx, F2, F2dx

    def cost_and_grad_batch(self, X, error_source_coeff=()):
        X, nx, B = self._batch_in(X)
        h, p = self.handle_for(nx)
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        if len(coeff) != self.nerr:
            raise AssertionError("error_source_coeff must have one entry per error source")
        cost = np.zeros(B)
        grad = np.zeros((nx, B), order="F")
        self.ctx.check(self.ctx.lib.rg_cost_and_grad_batch(h, B, _ptr(X), _ptr(coeff) if self.nerr else None, _ptr(cost), _ptr(grad)))
        return cost, grad

    def _reg_arrays(self, p, reg):
        """reg: None or a list of (kind, c1, c2) per control row -> ctypes-ready arrays (kept alive on self)."""
        if not reg:
            return None, None, None
        if len(reg) != p:
            raise AssertionError("one regularisation entry per control parameter")
        k = np.array([r[0] for r in reg], dtype=np.int32)
        c1 = np.array([r[1] for r in reg], dtype=np.float64)
        c2 = np.array([r[2] for r in reg], dtype=np.float64)
        self._keep_reg = (k, c1, c2)
        return k, c1, c2

    def cost_and_grad_batch_reg(self, X, error_source_coeff=(), reg=None):
        """calculate_common! including the enumerated regularisation terms, all on the device."""
        X, nx, B = self._batch_in(X)
        h, p = self.handle_for(nx)
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        k, c1, c2 = self._reg_arrays(p, reg)
        cost = np.zeros(B)
        grad = np.zeros((nx, B), order="F")
        self.ctx.check(self.ctx.lib.rg_cost_and_grad_batch_reg(h, B, _ptr(X), _ptr(coeff) if self.nerr else None, _ptr(k), _ptr(c1), _ptr(c2),
                                                             _ptr(cost), _ptr(grad)))
        return cost, grad

    def lbfgs_batch(self, X0, error_source_coeff=(), reg=None, history=10, iterations=100, g_tol=1e-8):
        """Batched device L-BFGS from the columns of X0 (nx, B).  Returns (X, cost, iterations per pulse, info dict)."""
        X, nx, B = self._batch_in(X0)
        X = X.copy(order="F")
        h, p = self.handle_for(nx)
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        if len(coeff) != self.nerr:
            raise AssertionError("error_source_coeff must have one entry per error source")
        k, c1, c2 = self._reg_arrays(p, reg)
        cost = np.zeros(B)
        iters = np.zeros(B, dtype=np.int32)
        info = np.zeros(3, dtype=np.int32)
        self.ctx.check(self.ctx.lib.rg_lbfgs_batch(h, B, _ptr(X), _ptr(coeff) if self.nerr else None, _ptr(k), _ptr(c1), _ptr(c2),
                                                 int(history), int(iterations), float(g_tol), _ptr(cost), _ptr(iters), _ptr(info)))
        return X, cost, iters, {"evaluations": int(info[0]), "iterations": int(info[1]), "line_search_failures": int(info[2])}

    def lbfgs_batch_dev(self, B, nx, dX_ptr, error_source_coeff, dcost_ptr, reg=None, history=10, iterations=100, g_tol=1e-8):
        """Device-pointer variant of lbfgs_batch: X (nx, B) is optimised in place in HBM.  Returns (iters, info)."""
        h, p = self.handle_for(nx)
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        k, c1, c2 = self._reg_arrays(p, reg)
        iters = np.zeros(B, dtype=np.int32)
        info = np.zeros(3, dtype=np.int32)
        self.ctx.check(self.ctx.lib.rg_lbfgs_batch_dev(h, B, _vp(dX_ptr), _ptr(coeff) if self.nerr else None, _ptr(k), _ptr(c1), _ptr(c2),
                                                     int(history), int(iterations), float(g_tol), _vp(dcost_ptr), _ptr(iters), _ptr(info)))
        return iters, {"evaluations": int(info[0]), "iterations": int(info[1]), "line_search_failures": int(info[2])}

    def cost_and_grad_batch_dev(self, B, nx, dX_ptr, error_source_coeff, dcost_ptr, dgrad_ptr):
        """Device-pointer variant (integers from e.g. torch.Tensor.data_ptr()); asynchronous."""
        h, p = self.handle_for(nx)
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        self._keep_coeff = coeff
        self.ctx.check(self.ctx.lib.rg_cost_and_grad_batch_dev(h, B, _vp(dX_ptr), _ptr(coeff) if self.nerr else None,
                                                             _vp(dcost_ptr), _vp(dgrad_ptr)))


def _scatter_dev(self, B, nx, dX_ptr, error_source_coeff, dcost_ptr, dgrad_ptr, peer_ptrs, dst_offset_bytes, what):
    """rg_cost_and_grad_batch_dev_scatter: evaluation whose [cost | grad] block also lands in the peers' gathered buffers
    (peer_ptrs: device pointers of the other ranks' buffers; what = 0 costs only, 1 costs and gradients)."""
    h, p = self.handle_for(nx)
    cf = np.ascontiguousarray(error_source_coeff, dtype=np.float64)
    self._keep_coeff = cf
    arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(q)) for q in peer_ptrs])
    self.ctx.check(self.ctx.lib.rg_cost_and_grad_batch_dev_scatter(h, int(B), _vp(dX_ptr), _ptr(cf) if self.nerr else None, _vp(dcost_ptr),
                                                                  _vp(dgrad_ptr), len(peer_ptrs), arr, int(dst_offset_bytes), int(what)))


Problem.cost_and_grad_batch_dev_scatter = _scatter_dev


class HStackProblem:
    """Closure problem (reference src/Types.jl:13,35,55: H0 / Herror / target_unitary are arbitrary callables).  The closures are
    evaluated here, on the host, into the stack of Hamiltonians whose exponentials the reference forms
    (src/UnitaryCalculations.jl:45-97); everything after that -- exponentials, scans, contractions, fidelity reductions --
    runs in the CUDA library (rg_*_from_hstack).  Finite differences of host-evaluated closures carry the reference's own
    FP64 noise floor (the perturbed and unperturbed Hamiltonians are subtracted after rounding), unlike descriptor problems."""

    def __init__(self, problem, ctx=None):
        from .types import FidelityRobustGRAPEProblem, UnitaryRobustGRAPEProblem
        self.ctx = ctx or default_context()
        if isinstance(problem, FidelityRobustGRAPEProblem):
            self.up, self.fp = problem.unitary_problem, problem
        elif isinstance(problem, UnitaryRobustGRAPEProblem):
            self.up, self.fp = problem, None
        else:
            raise TypeError("expected a UnitaryRobustGRAPEProblem or FidelityRobustGRAPEProblem")
        up = self.up
        if up.ndim > 10:
            raise RGError(RG_ERR_UNSUPPORTED, "closure (H-stack) problems are implemented for ndim <= 10")
        self.ndim, self.ntimes = int(up.ndim), int(up.ntimes)
        self.na, self.nerr = int(up.nb_additional_param), len(up.error_sources)
        self._handles = {}
        self._keep = []

    def _handle(self, p, hermitian):
        key = (p, hermitian)
        if key not in self._handles:
            up = self.up
            desc = rg_problem_desc()
            desc.ndim, desc.ntimes, desc.nparam = self.ndim, self.ntimes, p
            desc.nb_additional_param, desc.nerr = self.na, self.nerr
            desc.t0, desc.eps, desc.eps2 = float(up.t0), float(up.eps), float(up.eps2)
            desc.nterms, desc.ntarget_terms = 0, 0
            if self.fp is not None:
                proj = np.asfortranarray(np.asarray(self.fp.projector, dtype=np.float64))
                self._keep.append(proj)
                desc.projector = proj.ctypes.data_as(_dp)
            desc.hermitian, desc.hstack = (1 if hermitian else 0), 1
            h = _vp()
            self.ctx.check(self.ctx.lib.rg_problem_create(self.ctx.handle, C.byref(desc), C.byref(h)))
            self._handles[key] = h
        return self._handles[key]

    def stacks(self, x):
        """(Hstack (d,d,n_exp,N), Tstack (d,d,1+a) or None, p) for pulse x, with the reference's perturbation arithmetic
        (x + eps, evaluate, reset: src/UnitaryCalculations.jl:50-55,58-63,76-84,88-96; target: src/FidelityCalculations.jl:32-40)."""
        up = self.up
        x = np.asarray(x, dtype=np.float64)
        N, d, a, ne = self.ntimes, self.ndim, self.na, self.nerr
        nmain = len(x) - a
        if nmain < 0 or nmain % N != 0:
            raise AssertionError("Control parameter size must be a multiple of time steps")
        p = nmain // N
        nvar = p + a
        nexp = 1 + 2 * nvar + ne * (2 + nvar)
        xm = x[:nmain].reshape((p, N), order="F")
        xa = x[nmain:].copy()
        eps, eps2 = float(up.eps), float(up.eps2)
        Hs = np.zeros((d, d, nexp, N), dtype=np.complex128, order="F")
        H0 = up.H0

        def pert(k, v, h):
            xk, xad = xm[:, k].copy(), xa.copy()
            if v < p:
                xk[v] += h
            else:
                xad[v - p] += h
            return xk, xad

        for k in range(N):
            xk = xm[:, k].copy()
            base = np.asarray(H0(k + 1, xk, xa), dtype=np.complex128)
            Hs[:, :, 0, k] = base
            for v in range(nvar):
                Hs[:, :, 1 + v, k] = H0(k + 1, *pert(k, v, eps))
                Hs[:, :, 1 + nvar + v, k] = H0(k + 1, *pert(k, v, eps2))
            for e, src in enumerate(up.error_sources):
                Hs[:, :, 1 + 2 * nvar + e, k] = base + np.asarray(src.Herror(k + 1, xk, xa, eps))
                Hs[:, :, 1 + 2 * nvar + ne + e, k] = base + np.asarray(src.Herror(k + 1, xk, xa, eps2))
                for v in range(nvar):
                    xk2, xa2 = pert(k, v, eps2)
                    Hs[:, :, 1 + 2 * nvar + 2 * ne + e * nvar + v, k] = np.asarray(H0(k + 1, xk2, xa2)) + np.asarray(src.Herror(k + 1, xk2, xa2, eps2))
        Ts = None
        if self.fp is not None:
            Ts = np.zeros((d, d, 1 + a), dtype=np.complex128, order="F")
            Ts[:, :, 0] = self.fp.target_unitary(xa)
            for j in range(a):
                xa2 = xa.copy()
                xa2[j] += eps
                Ts[:, :, 1 + j] = self.fp.target_unitary(xa2)
        return Hs, Ts, p

    @staticmethod
    def _hermitian(Hs):
        return bool(np.abs(Hs - np.conj(np.swapaxes(Hs, 0, 1))).max() <= 1e-13 * max(1.0, np.abs(Hs).max()))

    def fidelity_and_derivatives(self, x):
        Hs, Ts, p = self.stacks(x)
        h = self._handle(p, self._hermitian(Hs))
        nx = len(x)
        F = np.zeros(1)
        Fdx = np.zeros((nx, 1), order="F")
        F2 = np.zeros((self.nerr, 1), order="F")
        F2dx = np.zeros((nx, self.nerr, 1), order="F")
        self.ctx.check(self.ctx.lib.rg_fidelity_and_derivatives_from_hstack(h, _ptr(Hs), _ptr(Ts), _ptr(F), _ptr(Fdx), _ptr(F2), _ptr(F2dx)))
        return float(F[0]), Fdx[:, 0].copy(), F2[:, 0].copy(), F2dx[:, :, 0].copy()

    def unitary_and_derivatives(self, x):
        Hs, _, p = self.stacks(x)
        h = self._handle(p, self._hermitian(Hs))
        d, N, a, e = self.ndim, self.ntimes, self.na, self.nerr
        z = lambda *s: np.zeros(s, dtype=np.complex128, order="F")
        U, U_dx, U_dx_add, U_derr = z(d, d), z(d, d, p, N), z(d, d, a), z(d, d, e)
        U_derr_dx, U_derr_dx_add = z(d, d, p, N, e), z(d, d, a, e)
        self.ctx.check(self.ctx.lib.rg_unitary_and_derivatives_from_hstack(h, _ptr(Hs), _ptr(U), _ptr(U_dx), _ptr(U_dx_add), _ptr(U_derr),
                                                                         _ptr(U_derr_dx), _ptr(U_derr_dx_add)))
        return U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add

    def close(self):
        for h in self._handles.values():
            self.ctx.lib.rg_problem_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass


def is_descriptor_problem(problem):
    """True when every callable of the problem is a declarative term list (device-evaluable)."""
    from .types import FidelityRobustGRAPEProblem
    up = problem.unitary_problem if isinstance(problem, FidelityRobustGRAPEProblem) else problem
    ok = isinstance(up.H0, D.TermHamiltonian) and all(isinstance(s.Herror, D.TermErrorHamiltonian) for s in up.error_sources)
    if isinstance(problem, FidelityRobustGRAPEProblem):
        ok = ok and isinstance(problem.target_unitary, D.TermTarget)
    return ok
