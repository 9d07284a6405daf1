# RobustGRAPEB200.jl -- ccall shim that puts librobustgrape_b200.so behind the RobustGRAPE.jl API.
#
# STATUS: written against include/robustgrape_b200.h, NOT EXECUTED -- Julia is not installed in the build
# image (no network), so this file has never been loaded.  The Python mirror (robustgrape_b200/*.py) binds the
# identical C ABI through ctypes and is what the test-suite exercises.  See INTEGRATION.md.
#
# All seven reference functions of the hot path are bound:
#   calculate_unitary_and_derivatives, calculate_interaction_error_operators            (src/UnitaryCalculations.jl:20,180)
#   calculate_fidelity_and_derivatives, calculate_fidelity_response, calculate_fidelity_response_fft,
#   calculate_expectation_values, optimize_fidelity_and_error_sources                   (src/FidelityCalculations.jl:19,161,246,306,368)
# Problems whose H0 / Herror / target_unitary are descriptors (below) are evaluated entirely on the GPU; problems with
# ordinary Julia closures are evaluated on the host into an H-stack and go through rg_*_from_hstack (any closure works).
#
# Usage (drop-in for the hot path):
#     using RobustGRAPE, RobustGRAPEB200
#     H0   = RobustGRAPEB200.rydberg_h0(:symmetric_blockaded)           # instead of a closure
#     Herr = RobustGRAPEB200.rydberg_amplitude_error(:symmetric_blockaded)
#     cz   = RobustGRAPEB200.cz_target(:symmetric_blockaded)
#     prob = FidelityRobustGRAPEProblem(UnitaryRobustGRAPEProblem(t0=7.613, ntimes=500, ndim=5, H0=H0,
#                nb_additional_param=1, error_sources=[ErrorSource(Herr)]), collect(Diagonal([1,2,1,0,0])), cz)
#     F, F_dx, F_d2err, F_d2err_dx = RobustGRAPEB200.calculate_fidelity_and_derivatives(prob, x)
#
# The descriptor structs below are `<: Function`, so they are valid values for the reference's
# `H0::Function`, `Herror::Function`, `target_unitary::Function` fields (src/Types.jl:13,35,55) and evaluate
# on the CPU to the same matrices as src/RydbergTools.jl, which keeps every reference function working on them.
module RobustGRAPEB200

using LinearAlgebra
using RobustGRAPE

const LIB = get(ENV, "ROBUSTGRAPE_B200_LIB", "librobustgrape_b200.so")

# ---- mirror of include/robustgrape_b200.h ------------------------------------------------------------
const RG_F_VAR, RG_F_COS, RG_F_SIN, RG_F_EXPI, RG_F_ERR, RG_F_ERR1P_M1, RG_F_TABLE = Int32.(0:6)
const RG_S_MAIN, RG_S_ADD, RG_S_NONE = Int32.(0:2)
const RG_OWNER_H0, RG_OWNER_TARGET = Int32(-1), Int32(-2)

struct CFactor            # rg_factor (32 bytes)
    kind::Int32; space::Int32; index::Int32; reserved::Int32
    scale::Float64; offset::Float64
end
struct CTerm              # rg_term
    owner::Int32; nfactors::Int32
    coef_re::Float64; coef_im::Float64
    factors::NTuple{4,CFactor}
    nnz::Int32; reserved::Int32
    rows::Ptr{Int32}; cols::Ptr{Int32}; vals::Ptr{Float64}
end
struct CProblemDesc       # rg_problem_desc
    ndim::Int32; ntimes::Int32; nparam::Int32; nb_additional_param::Int32; nerr::Int32
    t0::Float64; eps::Float64; eps2::Float64
    nterms::Int32; terms::Ptr{CTerm}
    ntarget_terms::Int32; target_terms::Ptr{CTerm}
    projector::Ptr{Float64}
    ntable_cols::Int32; table::Ptr{Float64}
    hermitian::Int32
    hstack::Int32
end

# ---- declarative operators ---------------------------------------------------------------------------
struct Factor
    kind::Int32; space::Int32; index::Int      # index is 1-based here, 0-based across the ABI
    scale::Float64; offset::Float64
end
expi(space, index; scale=1.0, offset=0.0) = Factor(RG_F_EXPI, space, index, scale, offset)
errfactor() = Factor(RG_F_ERR, RG_S_NONE, 1, 1.0, 0.0)
err1p_m1() = Factor(RG_F_ERR1P_M1, RG_S_NONE, 1, 1.0, 0.0)

struct Term
    coef::ComplexF64
    factors::Vector{Factor}
    entries::Vector{Tuple{Int,Int,ComplexF64}}   # (row, col, value), 1-based
end

function factor_value(f::Factor, x, x_add, err)
    f.kind == RG_F_ERR && return complex(float(err))
    f.kind == RG_F_ERR1P_M1 && return complex((1.0 + err) - 1.0)
    v = f.space == RG_S_MAIN ? x[f.index] : x_add[f.index]
    u = (f.scale == 1.0 && f.offset == 0.0) ? v : f.scale * v + f.offset
    f.kind == RG_F_VAR && return complex(u)
    f.kind == RG_F_COS && return complex(cos(u))
    f.kind == RG_F_SIN && return complex(sin(u))
    return cis(u)
end

function dense(ndim, terms::Vector{Term}, x, x_add, err)
    M = zeros(ComplexF64, ndim, ndim)
    for t in terms
        c = t.coef
        for f in t.factors
            c *= factor_value(f, x, x_add, err)
        end
        for (r, cidx, v) in t.entries
            M[r, cidx] += c * v
        end
    end
    return M
end

"Valid for `UnitaryRobustGRAPEProblem.H0`: callable as H0(time_step, x, x_add)."
struct TermHamiltonian <: Function
    ndim::Int; terms::Vector{Term}
end
(h::TermHamiltonian)(time_step, x, x_add) = dense(h.ndim, h.terms, x, x_add, 0.0)

"Valid for `ErrorSource.Herror`: callable as Herror(time_step, x, x_add, err)."
struct TermErrorHamiltonian <: Function
    ndim::Int; terms::Vector{Term}
end
(h::TermErrorHamiltonian)(time_step, x, x_add, err) = dense(h.ndim, h.terms, x, x_add, err)

"Valid for `FidelityRobustGRAPEProblem.target_unitary`: callable as U0(x_add)."
struct TermTarget <: Function
    ndim::Int; terms::Vector{Term}
end
(t::TermTarget)(x_add) = dense(t.ndim, t.terms, Float64[], x_add, 0.0)

# ---- RydbergTools as descriptors (src/RydbergTools.jl:31-39, 71-81, 160-162, 197-203) -------------------
function _model(model::Symbol)
    model == :symmetric_blockaded && return 5, [(2, 4, 0.5 + 0im), (3, 5, 1 / sqrt(2) + 0im)], [4, 5]
    model == :full_blockaded && return 7, [(2, 5, 0.5 + 0im), (3, 6, 0.5 + 0im), (4, 7, 1 / sqrt(2) + 0im)], [5, 6, 7]
    error("unknown model $model")
end
function _drive(up, extra)
    dn = [(c, r, v) for (r, c, v) in up]
    [Term(1.0, vcat([expi(RG_S_MAIN, 1; scale=-1.0)], extra), up), Term(1.0, vcat([expi(RG_S_MAIN, 1; scale=1.0)], extra), dn)]
end
function rydberg_h0(model::Symbol=:symmetric_blockaded; eps=0.0, delta=0.0)
    ndim, up, ryd = _model(model)
    terms = _drive([(r, c, v * (1 + eps)) for (r, c, v) in up], Factor[])
    delta != 0 && push!(terms, Term(delta, Factor[], [(r, r, 1.0 + 0im) for r in ryd]))
    TermHamiltonian(ndim, terms)
end
rydberg_amplitude_error(model::Symbol=:symmetric_blockaded) =
    (m = _model(model); TermErrorHamiltonian(m[1], _drive(m[2], [err1p_m1()])))
rydberg_frequency_error(model::Symbol=:symmetric_blockaded) =
    (m = _model(model); TermErrorHamiltonian(m[1], [Term(1.0, [errfactor()], [(r, r, 1.0 + 0im) for r in m[3]])]))
function cz_target(model::Symbol=:symmetric_blockaded)
    e1 = expi(RG_S_ADD, 1); e2 = expi(RG_S_ADD, 1; scale=2.0, offset=Float64(pi))
    if model == :symmetric_blockaded
        return TermTarget(5, [Term(1.0, Factor[], [(1, 1, 1.0 + 0im)]), Term(1.0, [e1], [(2, 2, 1.0 + 0im)]), Term(1.0, [e2], [(3, 3, 1.0 + 0im)])])
    end
    TermTarget(7, [Term(1.0, Factor[], [(1, 1, 1.0 + 0im)]), Term(1.0, [e1], [(2, 2, 1.0 + 0im), (3, 3, 1.0 + 0im)]), Term(1.0, [e2], [(4, 4, 1.0 + 0im)])])
end

# ---- library handles -------------------------------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rg_ctx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
        rc == 0 || error("rg_ctx_create failed ($rc): ", unsafe_string(ccall((:rg_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        c = new(h[])
        finalizer(x -> ccall((:rg_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), c)
        return c
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context(0)); DEFAULT_CTX[])
check(ctx::Context, rc) = rc == 0 || error("librobustgrape_b200 error $rc: ", unsafe_string(ccall((:rg_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))

mutable struct DeviceProblem
    ctx::Context
    handle::Ptr{Cvoid}
    nerr::Int; ntimes::Int; na::Int
end

function _cterms(terms::Vector{Term}, owner::Int32, keep::Vector{Any})
    out = CTerm[]
    for t in terms
        rows = Int32[e[1] - 1 for e in t.entries]; cols = Int32[e[2] - 1 for e in t.entries]
        vals = reinterpret(Float64, ComplexF64[e[3] for e in t.entries]) |> collect
        push!(keep, rows, cols, vals)
        fs = [CFactor(f.kind, f.space, Int32(f.index - 1), 0, f.scale, f.offset) for f in t.factors]
        while length(fs) < 4
            push!(fs, CFactor(0, 0, 0, 0, 0.0, 0.0))
        end
        push!(out, CTerm(owner, length(t.factors), real(t.coef), imag(t.coef), Tuple(fs), length(rows), 0,
                         pointer(rows), pointer(cols), pointer(vals)))
    end
    return out
end

_unitary(p::UnitaryRobustGRAPEProblem) = p
_unitary(p::FidelityRobustGRAPEProblem) = p.unitary_problem
is_descriptor_problem(p::UnitaryRobustGRAPEProblem) =
    p.H0 isa TermHamiltonian && all(s.Herror isa TermErrorHamiltonian for s in p.error_sources)
is_descriptor_problem(p::FidelityRobustGRAPEProblem) = is_descriptor_problem(p.unitary_problem) && p.target_unitary isa TermTarget

"Hermitian for real variables?  Checked numerically at a few random points (non-Hermitian H0, e.g. -i gamma/2 decay, is legal:
the reference uses inv, src/UnitaryCalculations.jl:47)."
function _is_hermitian(up::UnitaryRobustGRAPEProblem, nparam::Int)
    for _ in 1:3
        x = randn(max(nparam, 1)); xa = randn(max(up.nb_additional_param, 1))
        M = up.H0(1, x, xa)
        maximum(abs.(M - M')) <= 1e-13 * max(1.0, maximum(abs.(M))) || return false
        for s in up.error_sources
            E = s.Herror(1, x, xa, randn())
            maximum(abs.(E - E')) <= 1e-13 * max(1.0, maximum(abs.(E))) || return false
        end
    end
    return true
end

"Build the device-resident twin of a problem.  Descriptor problems carry their term lists; closure problems are created as
H-stack problems (`hstack = 1`, no terms) and are fed host-evaluated Hamiltonians per call."
function DeviceProblem(prob, nparam::Int; ctx::Context=default_context(), hermitian::Union{Nothing,Bool}=nothing)
    up = _unitary(prob)
    keep = Any[]
    terms = CTerm[]; tterms = CTerm[]
    descr = is_descriptor_problem(prob)
    if descr
        terms = _cterms(up.H0.terms, RG_OWNER_H0, keep)
        for (e, src) in enumerate(up.error_sources)
            append!(terms, _cterms(src.Herror.terms, Int32(e - 1), keep))
        end
        prob isa FidelityRobustGRAPEProblem && (tterms = _cterms(prob.target_unitary.terms, RG_OWNER_TARGET, keep))
    end
    proj = prob isa FidelityRobustGRAPEProblem ? Matrix{Float64}(prob.projector) : Float64[]
    herm = hermitian === nothing ? _is_hermitian(up, nparam) : hermitian
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep terms tterms proj begin
        desc = Ref(CProblemDesc(up.ndim, up.ntimes, nparam, up.nb_additional_param, length(up.error_sources),
                                up.t0, up.ϵ, up.ϵ2, length(terms), isempty(terms) ? C_NULL : pointer(terms),
                                length(tterms), isempty(tterms) ? C_NULL : pointer(tterms),
                                isempty(proj) ? C_NULL : pointer(proj), 0, C_NULL, herm ? 1 : 0, descr ? 0 : 1))
        check(ctx, ccall((:rg_problem_create, LIB), Cint, (Ptr{Cvoid}, Ref{CProblemDesc}, Ref{Ptr{Cvoid}}), ctx.handle, desc, h))
    end
    p = DeviceProblem(ctx, h[], length(up.error_sources), up.ntimes, up.nb_additional_param)
    finalizer(x -> ccall((:rg_problem_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), p)
    return p
end

"Host evaluation of closures into the stacks rg_*_from_hstack take (include/robustgrape_b200.h), with the reference's
perturbation arithmetic (src/UnitaryCalculations.jl:50-55,58-63,76-84,88-96; src/FidelityCalculations.jl:32-40)."
function hstacks(prob, x::Vector{Float64})
    up = _unitary(prob)
    N, d, a, ne = up.ntimes, up.ndim, up.nb_additional_param, length(up.error_sources)
    nmain = length(x) - a
    @assert mod(nmain, N) == 0 "Control parameter size must be a multiple of time steps"
    p = nmain ÷ N; nvar = p + a
    nexp = 1 + 2nvar + ne * (2 + nvar)
    xm = reshape(x[1:nmain], p, N); xa = x[nmain+1:end]
    Hs = zeros(ComplexF64, d, d, nexp, N)
    function pert(k, v, h)
        xk = xm[:, k]; xad = copy(xa)
        v <= p ? (xk[v] += h) : (xad[v-p] += h)
        return xk, xad
    end
    for k in 1:N
        xk = xm[:, k]
        base = ComplexF64.(up.H0(k, xk, xa))
        Hs[:, :, 1, k] = base
        for v in 1:nvar
            Hs[:, :, 1+v, k] = up.H0(k, pert(k, v, up.ϵ)...)
            Hs[:, :, 1+nvar+v, k] = up.H0(k, pert(k, v, up.ϵ2)...)
        end
        for (e, src) in enumerate(up.error_sources)
            Hs[:, :, 1+2nvar+e, k] = base + src.Herror(k, xk, xa, up.ϵ)
            Hs[:, :, 1+2nvar+ne+e, k] = base + src.Herror(k, xk, xa, up.ϵ2)
            for v in 1:nvar
                xk2, xa2 = pert(k, v, up.ϵ2)
                Hs[:, :, 1+2nvar+2ne+(e-1)*nvar+v, k] = up.H0(k, xk2, xa2) + src.Herror(k, xk2, xa2, up.ϵ2)
            end
        end
    end
    Ts = zeros(ComplexF64, d, d, 1 + a)
    if prob isa FidelityRobustGRAPEProblem
        Ts[:, :, 1] = prob.target_unitary(xa)
        for j in 1:a
            xa2 = copy(xa); xa2[j] += up.ϵ
            Ts[:, :, 1+j] = prob.target_unitary(xa2)
        end
    end
    herm = maximum(abs.(Hs .- conj.(permutedims(Hs, (2, 1, 3, 4))))) <= 1e-13 * max(1.0, maximum(abs.(Hs)))
    return Hs, Ts, p, herm
end

const _CACHE = IdDict{Any,DeviceProblem}()
function device_problem(prob, x)
    get!(_CACHE, prob) do
        up = _unitary(prob)
        DeviceProblem(prob, (length(x) - up.nb_additional_param) ÷ up.ntimes)
    end
end

# ---- the hot path ----------------------------------------------------------------------------------
"Replaces RobustGRAPE.calculate_fidelity_and_derivatives (src/FidelityCalculations.jl:19-119)."
function calculate_fidelity_and_derivatives(fp::FidelityRobustGRAPEProblem, x::Vector{Float64})
    if !is_descriptor_problem(fp)                       # ordinary closures: host-evaluated H-stack
        Hs, Ts, p, herm = hstacks(fp, x)
        dp = DeviceProblem(fp, p; hermitian=herm)
        nx = length(x); ne = dp.nerr
        F = zeros(1); Fdx = zeros(nx, 1); F2 = zeros(ne, 1); F2dx = zeros(nx, ne, 1)
        GC.@preserve Hs Ts F Fdx F2 F2dx begin
            check(dp.ctx, ccall((:rg_fidelity_and_derivatives_from_hstack, LIB), Cint,
                                (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                dp.handle, Hs, Ts, F, Fdx, F2, F2dx))
        end
        return F[1], Fdx[:, 1], F2[:, 1], F2dx[:, :, 1]
    end
    F, Fdx, F2, F2dx = calculate_fidelity_and_derivatives_batch(fp, reshape(x, :, 1))
    return F[1], Fdx[:, 1], F2[:, 1], F2dx[:, :, 1]
end

"Replaces RobustGRAPE.calculate_unitary_and_derivatives (src/UnitaryCalculations.jl:20-155); returns concrete ComplexF64 arrays
of the reference's shapes (the reference's containers have abstract eltype `Complex`, src/UnitaryCalculations.jl:106-110)."
function calculate_unitary_and_derivatives(up::UnitaryRobustGRAPEProblem, x::Vector{Float64})
    N, d, a, e = up.ntimes, up.ndim, up.nb_additional_param, length(up.error_sources)
    @assert mod(length(x) - a, N) == 0 "Control parameter size must be a multiple of time steps"
    p = (length(x) - a) ÷ N
    U = zeros(ComplexF64, d, d); U_dx = zeros(ComplexF64, d, d, p, N); U_dx_add = zeros(ComplexF64, d, d, a)
    U_derr = zeros(ComplexF64, d, d, e); U_derr_dx = zeros(ComplexF64, d, d, p, N, e); U_derr_dx_add = zeros(ComplexF64, d, d, a, e)
    T = Ptr{ComplexF64}
    if is_descriptor_problem(up)
        dp = device_problem(up, x)
        GC.@preserve x U U_dx U_dx_add U_derr U_derr_dx U_derr_dx_add begin
            check(dp.ctx, ccall((:rg_unitary_and_derivatives, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, T, T, T, T, T, T),
                                dp.handle, x, U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add))
        end
    else
        Hs, _, _, herm = hstacks(up, x)
        dp = DeviceProblem(up, p; hermitian=herm)
        GC.@preserve Hs U U_dx U_dx_add U_derr U_derr_dx U_derr_dx_add begin
            check(dp.ctx, ccall((:rg_unitary_and_derivatives_from_hstack, LIB), Cint, (Ptr{Cvoid}, T, T, T, T, T, T, T),
                                dp.handle, Hs, U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add))
        end
    end
    return U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add
end

"Replaces RobustGRAPE.calculate_interaction_error_operators (src/UnitaryCalculations.jl:180-204): (ndim, ndim, ntimes, nerr)."
function calculate_interaction_error_operators(up::UnitaryRobustGRAPEProblem, x::Vector{Float64})
    dp = device_problem(up, x)
    O = zeros(ComplexF64, up.ndim, up.ndim, up.ntimes, length(up.error_sources))
    GC.@preserve x O begin
        check(dp.ctx, ccall((:rg_interaction_error_operators, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{ComplexF64}), dp.handle, x, O))
    end
    return O
end

"Replaces RobustGRAPE.calculate_fidelity_response (src/FidelityCalculations.jl:246-280): (nfreq, nerr).  `first`/`count`
(0-based first row, number of rows) select a shard of the frequency grid for multi-GPU use."
function calculate_fidelity_response(fp::FidelityRobustGRAPEProblem, x::Vector{Float64}, normalized_frequencies::Vector{Float64};
                                     first::Int=0, count::Int=length(normalized_frequencies) - first)
    dp = device_problem(fp, x)
    R = zeros(count, dp.nerr)
    GC.@preserve x normalized_frequencies R begin
        check(dp.ctx, ccall((:rg_fidelity_response, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}),
                            dp.handle, x, normalized_frequencies, length(normalized_frequencies), first, count, R))
    end
    return R
end

"Replaces RobustGRAPE.calculate_fidelity_response_fft (src/FidelityCalculations.jl:306-343): (response (N*os, nerr), frequencies)."
function calculate_fidelity_response_fft(fp::FidelityRobustGRAPEProblem, x::Vector{Float64}; oversampling::Int=1)
    @assert oversampling >= 1
    dp = device_problem(fp, x)
    n = dp.ntimes * oversampling
    R = zeros(n, dp.nerr); fr = zeros(n)
    GC.@preserve x R fr begin
        check(dp.ctx, ccall((:rg_fidelity_response_fft, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
                            dp.handle, x, oversampling, R, fr))
    end
    return R, fr
end

"Replaces RobustGRAPE.calculate_expectation_values (src/FidelityCalculations.jl:368-390): (ntimes, nerr)."
function calculate_expectation_values(fp::FidelityRobustGRAPEProblem, x::Vector{Float64})
    dp = device_problem(fp, x)
    out = zeros(dp.ntimes, dp.nerr)
    GC.@preserve x out begin
        check(dp.ctx, ccall((:rg_expectation_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), dp.handle, x, out))
    end
    return out
end

"Batched variant: X is (nx, B), one pulse per column."
function calculate_fidelity_and_derivatives_batch(fp::FidelityRobustGRAPEProblem, X::Matrix{Float64})
    @assert mod(size(X, 1) - fp.unitary_problem.nb_additional_param, fp.unitary_problem.ntimes) == 0 "Control parameter size must be a multiple of time steps"
    dp = device_problem(fp, view(X, :, 1))
    nx, B = size(X)
    F = zeros(B); Fdx = zeros(nx, B); F2 = zeros(dp.nerr, B); F2dx = zeros(nx, dp.nerr, B)
    GC.@preserve X F Fdx F2 F2dx begin
        check(dp.ctx, ccall((:rg_fidelity_and_derivatives_batch, LIB), Cint,
                            (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                            dp.handle, B, X, F, Fdx, F2, F2dx))
    end
    return F, Fdx, F2, F2dx
end

"`calculate_common!` of src/FidelityCalculations.jl:174-184 for every column of X (no regularisation)."
function cost_and_gradient_batch(fp::FidelityRobustGRAPEProblem, X::Matrix{Float64}, error_source_coeff::Vector{Float64})
    @assert length(error_source_coeff) == length(fp.unitary_problem.error_sources)
    dp = device_problem(fp, view(X, :, 1))
    nx, B = size(X)
    cost = zeros(B); grad = zeros(nx, B)
    GC.@preserve X cost grad error_source_coeff begin
        check(dp.ctx, ccall((:rg_cost_and_grad_batch, LIB), Cint,
                            (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                            dp.handle, B, X, isempty(error_source_coeff) ? C_NULL : pointer(error_source_coeff), cost, grad))
    end
    return cost, grad
end

"Kernel family an evaluation of this problem takes (`rg_problem_path`): \"fused_q_pc\", \"fused_q\", \"block2\", \"steps_t\", \"group\", \"dense\", \"hstack\"."
function kernel_path(fp::FidelityRobustGRAPEProblem, x::AbstractVector{Float64})
    dp = device_problem(fp, x)
    buf = zeros(UInt8, 32)
    GC.@preserve buf check(dp.ctx, ccall((:rg_problem_path, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32), dp.handle, buf, 32))
    return unsafe_string(pointer(buf))
end

"Multi-start optimisation with the iterates resident on the device (`rg_lbfgs_batch`): one independent L-BFGS per column of X, the
role of `Optim.optimize(...; method = LBFGS())` in src/FidelityCalculations.jl:199-217 for a whole batch.  `reg_kind[i]` selects the
enumerated regularisation of control i (0 none, 1 regularization_cost, 2 regularization_cost_phase, 3 the tests' sin^2 form).
Returns (X_final, cost, iterations per pulse)."
function lbfgs_batch(fp::FidelityRobustGRAPEProblem, X0::Matrix{Float64}, error_source_coeff::Vector{Float64};
                     reg_kind::Vector{Int32} = Int32[], reg_c1::Vector{Float64} = Float64[], reg_c2::Vector{Float64} = Float64[],
                     history::Integer = 10, iterations::Integer = 100, g_tol::Float64 = 1e-8)
    dp = device_problem(fp, view(X0, :, 1))
    nx, B = size(X0)
    X = copy(X0); cost = zeros(B); iters = zeros(Int32, B); info = zeros(Int32, 3)
    GC.@preserve X cost iters info error_source_coeff reg_kind reg_c1 reg_c2 begin
        check(dp.ctx, ccall((:rg_lbfgs_batch, LIB), Cint,
                            (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64,
                             Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                            dp.handle, B, X, isempty(error_source_coeff) ? C_NULL : pointer(error_source_coeff),
                            isempty(reg_kind) ? C_NULL : pointer(reg_kind), isempty(reg_c1) ? C_NULL : pointer(reg_c1),
                            isempty(reg_c2) ? C_NULL : pointer(reg_c2), history, iterations, g_tol, cost, iters, info))
    end
    return X, cost, iters
end

"Drop-in for RobustGRAPE.optimize_fidelity_and_error_sources (src/FidelityCalculations.jl:161-218): same closure
structure, with the cost/gradient evaluation (:177-184) served by the GPU."
function optimize_fidelity_and_error_sources(fp::FidelityRobustGRAPEProblem, prm::FidelityRobustGRAPEParameters)
    up = fp.unitary_problem
    nerr = length(up.error_sources); ntimes = up.ntimes; na = up.nb_additional_param
    nparam = (length(prm.x_initial) - na) ÷ ntimes
    coeff = Float64.(prm.error_source_coeff)
    function common!(x, last_x, buffer)
        if x != last_x
            copy!(last_x, x)
            cost, grad = cost_and_gradient_batch(fp, reshape(Vector{Float64}(x), :, 1), coeff)
            buffer[1] = cost[1]; buffer[2:end] = grad[:, 1]
            x_main = reshape(x[1:end-na], nparam, ntimes)
            for np = 1:nparam                                        # :189-195 (host-side regularisation)
                r1, j1, r2, j2 = prm.regularization_functions[np](x_main[np, :])
                buffer[1] += prm.regularization_coeff1[np] * r1 + prm.regularization_coeff2[np] * r2
                buffer[1 .+ (np:nparam:nparam*ntimes)] .+= prm.regularization_coeff1[np] .* j1 .+ prm.regularization_coeff2[np] .* j2
            end
        end
    end
    buffer = zeros(length(prm.x_initial) + 1); last_x = similar(prm.x_initial); fill!(last_x, NaN)
    f(x) = (common!(x, last_x, buffer); buffer[1])
    g!(stor, x) = (common!(x, last_x, buffer); stor .= buffer[2:end])
    return RobustGRAPE.Optim.optimize(f, g!, prm.x_initial; method=prm.solver_algorithm, time_limit=prm.time_limit,
                                      iterations=prm.iterations, prm.additional_parameters...)
end

end # module
