# make_golden.jl -- pins the oracle against the reference itself.  NOT EXECUTED in the build image (no Julia).
#
#     julia --project=/path/to/RobustGRAPE tests/golden/make_golden.jl
#
# For every problem of tests/cases.py:golden_cases() that can be written with the reference's RydbergTools closures it
# evaluates, with the REFERENCE's code, calculate_fidelity_and_derivatives (src/FidelityCalculations.jl:19),
# calculate_unitary_and_derivatives (src/UnitaryCalculations.jl:20) and calculate_fidelity_response
# (src/FidelityCalculations.jl:246) on the seeded input stored in tests/golden/<case>.npz (exported to <case>_x.csv by
# `python tests/golden/make_golden.py --export-inputs`), and writes tests/golden/julia_<case>.csv:
#     line 1: F ; line 2: F_dx ; line 3: F_d2err ; following lines: F_d2err_dx columns ; then U re/im ; then response rows.
# tests/test_oracle_structure.py::test_julia_fixtures (skipped while no julia_*.csv exists) compares the oracle with these
# files: F at 1e-12, the finite-difference outputs at the measured FP64 noise floor.  Prints Threads.nthreads() and VERSION.
using RobustGRAPE, RobustGRAPE.RydbergTools, LinearAlgebra, Printf, DelimitedFiles

here = @__DIR__
println("Julia ", VERSION, ", threads = ", Threads.nthreads())

function problem(model::Symbol, ntimes, t0, errors; eps0=0.0, delta0=0.0)
    Hm = model == :symmetric_blockaded ? rydberg_hamiltonian_symmetric_blockaded : rydberg_hamiltonian_full_blockaded
    d = model == :symmetric_blockaded ? 5 : 7
    H0(k, ϕ, xa) = Hm(ϕ[1], eps0, delta0)
    srcs = ErrorSource[]
    for e in errors
        if e == :amp
            push!(srcs, ErrorSource((k, ϕ, xa, ϵ) -> Hm(ϕ[1], ϵ, 0) - Hm(ϕ[1], 0, 0)))
        else
            push!(srcs, ErrorSource((k, ϕ, xa, δ) -> Hm(ϕ[1], 0, δ) - Hm(ϕ[1], 0, 0)))
        end
    end
    up = UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=d, H0=H0, nb_additional_param=1, error_sources=srcs)
    proj = model == :symmetric_blockaded ? collect(Diagonal([1, 2, 1, 0, 0])) : collect(Diagonal([1, 1, 1, 1, 0, 0, 0]))
    tgt = model == :symmetric_blockaded ? (xa -> cz_with_1q_phase_symmetric(xa[1])) : (xa -> cz_with_1q_phase_full(xa[1]; rydberg_dimension=3))
    FidelityRobustGRAPEProblem(unitary_problem=up, projector=proj, target_unitary=tgt)
end

cases = Dict(
    "cz5_e0_N24" => problem(:symmetric_blockaded, 24, 7.613 * 24 / 200, Symbol[]),
    "cz5_e1_N17" => problem(:symmetric_blockaded, 17, 14.32 * 17 / 200, [:amp]),
    "cz5_e2_N12" => problem(:symmetric_blockaded, 12, 7.613 * 12 / 100, [:amp, :freq]),
    "cz7_e2_N10" => problem(:full_blockaded, 10, 7.613 * 10 / 100, [:amp, :freq]),
    "cz5_C2_e1_N200" => problem(:symmetric_blockaded, 200, 14.32, [:amp]),
    "cz5_detuned_e2_N31" => problem(:symmetric_blockaded, 31, 7.613 * 31 / 100, [:freq, :amp]; eps0=0.02, delta0=0.37),
    "cz7_detuned_e2_N19" => problem(:full_blockaded, 19, 7.613 * 19 / 60, [:amp, :freq]; delta0=-0.21),
)
for (name, fp) in cases
    xf = joinpath(here, name * "_x.csv")
    isfile(xf) || (println("skip ", name, ": run `python tests/golden/make_golden.py --export-inputs` first"); continue)
    x = vec(readdlm(xf, ','))
    F, F_dx, F_d2err, F_d2err_dx = calculate_fidelity_and_derivatives(fp, x)
    U = calculate_unitary_and_derivatives(fp.unitary_problem, x)[1]
    open(joinpath(here, "julia_" * name * ".csv"), "w") do io
        @printf(io, "%.17g\n", F)
        println(io, join([@sprintf("%.17g", v) for v in F_dx], ","))
        println(io, join([@sprintf("%.17g", v) for v in F_d2err], ","))
        for e in 1:size(F_d2err_dx, 2)
            println(io, join([@sprintf("%.17g", v) for v in F_d2err_dx[:, e]], ","))
        end
        println(io, join([@sprintf("%.17g", real(v)) for v in vec(U)], ","))
        println(io, join([@sprintf("%.17g", imag(v)) for v in vec(U)], ","))
        if length(fp.unitary_problem.error_sources) > 0
            R = calculate_fidelity_response(fp, x, collect(range(0, 3, length=9)))
            for r in 1:size(R, 1)
                println(io, join([@sprintf("%.17g", v) for v in R[r, :]], ","))
            end
        end
    end
    println(name, ": F = ", F)
end
