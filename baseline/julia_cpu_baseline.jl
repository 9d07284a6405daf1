# julia_cpu_baseline.jl -- the reference's own CPU path, timed (BASELINE.md section 3-4, SURVEY 8d).
#
# NOT EXECUTED in the build image (no Julia, no network).  For anyone with Julia >= 1.9 and the reference checked out:
#     julia --project=/path/to/RobustGRAPE -t auto baseline/julia_cpu_baseline.jl [npulses] [ntimes] [nerr]
# It evaluates `calculate_fidelity_and_derivatives` (src/FidelityCalculations.jl:19) -- the body of `calculate_common!`
# (src/FidelityCalculations.jl:174-184) -- on the same synthetic pulses as bench.py's C4 workload (phi_k ~ 2 pi U(0,1),
# theta ~ 2 pi U(0,1), 5-level symmetric-blockaded CZ, t0 = 7.613, projector diag(1,2,1,0,0)) and prints evals/s, the thread
# count and one JSON line in bench.py's `--impl reference` format with kind = "reference".
# The pulses are read from tests/golden/c4_pulses.csv when present (written by tests/golden/make_golden.jl / bench.py
# --dump-pulses) so that the GPU and the Julia run see bit-identical inputs; otherwise Julia's own RNG is used.
using RobustGRAPE, RobustGRAPE.RydbergTools, LinearAlgebra, Random, Printf

npulses = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 64
ntimes = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1000
nerr = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 0
t0 = 7.613

H0(time_step, ϕ, x_add) = rydberg_hamiltonian_symmetric_blockaded(ϕ[1], 0, 0)
amp_error(time_step, ϕ, x_add, ϵ) = rydberg_hamiltonian_symmetric_blockaded(ϕ[1], ϵ, 0) - H0(time_step, ϕ, x_add)
freq_error(time_step, ϕ, x_add, δ) = rydberg_hamiltonian_symmetric_blockaded(ϕ[1], 0, δ) - H0(time_step, ϕ, x_add)
sources = ErrorSource[]
nerr >= 1 && push!(sources, ErrorSource(amp_error))
nerr >= 2 && push!(sources, ErrorSource(freq_error))
up = UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=5, H0=H0, nb_additional_param=1, error_sources=sources)
fp = FidelityRobustGRAPEProblem(unitary_problem=up, projector=collect(Diagonal([1, 2, 1, 0, 0])),
                                target_unitary=x_add -> cz_with_1q_phase_symmetric(x_add[1]))

csv = joinpath(@__DIR__, "..", "tests", "golden", "c4_pulses.csv")
X = if isfile(csv)
    M = [parse.(Float64, split(l, ',')) for l in eachline(csv)]
    hcat(M[1:min(npulses, length(M))]...)
else
    Random.seed!(43); 2π .* rand(ntimes + 1, npulses)
end
npulses = size(X, 2)
coeff = fill(1e-4, nerr)
function cost_and_grad(x)                       # calculate_common! without regularisation (:177-184)
    F, F_dx, F_d2err, F_d2err_dx = calculate_fidelity_and_derivatives(fp, x)
    cost = 1 - F + sum(coeff .* F_d2err .^ 2)
    grad = -F_dx
    for e in 1:nerr
        grad .+= 2 * coeff[e] * F_d2err[e] .* F_d2err_dx[:, e]
    end
    return cost, grad
end
cost_and_grad(X[:, 1])                          # compile
costs = zeros(npulses)
t = @elapsed Threads.@threads for b in 1:npulses
    costs[b] = cost_and_grad(X[:, b])[1]
end
@printf("%d pulses x %d steps, nerr = %d: %.3f evals/s on %d Julia thread(s)\n", npulses, ntimes, nerr, npulses / t, Threads.nthreads())
@printf("{\"impl\": \"reference\", \"metric\": \"GRAPE cost+grad evals/sec (CZ, batched pulses)\", \"value\": %.6g, \"unit\": \"evals/s\", \"cpu_baseline\": {\"value\": %.6g, \"unit\": \"evals/s\", \"cores\": %d, \"kind\": \"reference\", \"sample\": \"%d pulses, Julia %s\"}, \"first_costs\": [%.17g, %.17g]}\n",
        npulses / t, npulses / t, Threads.nthreads(), npulses, string(VERSION), costs[1], costs[min(2, npulses)])
