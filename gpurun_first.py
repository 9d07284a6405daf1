import numpy as np, time, sys
sys.path.insert(0, '.')
from robustgrape_b200 import *
from robustgrape_b200 import rydberg_tools as rt
from robustgrape_b200._lib import Problem, default_context
from oracle import reference_oracle as ro, exact_oracle as eo
ctx = default_context()
print("fp64 peak (dfma, dmma) TF:", ctx.measure_fp64_peak(0.5))
def mk(N, t0, errs, model="symmetric_blockaded"):
    d = 5 if model == "symmetric_blockaded" else 7
    proj = np.diag([1,2,1,0,0]) if d == 5 else np.diag([1,1,1,1,0,0,0])
    es = []
    for i, e in enumerate(errs):
        es.append(ErrorSource(rt.rydberg_amplitude_error(model, source=i) if e == 'amp' else rt.rydberg_frequency_error(model, source=i)))
    return FidelityRobustGRAPEProblem(UnitaryRobustGRAPEProblem(t0=t0, ntimes=N, ndim=d, H0=rt.rydberg_h0(model), nb_additional_param=1, error_sources=es), proj.astype(float), rt.cz_target(model))
rng = np.random.default_rng(1)
for (N, errs, model) in [(20, [], "symmetric_blockaded"), (23, ['amp'], "symmetric_blockaded"), (20, ['amp','freq'], "symmetric_blockaded"), (12, ['amp','freq'], "full_blockaded")]:
    fp = mk(N, 7.613*N/200, errs, model)
    B = 3
    X = np.concatenate([2*np.pi*rng.random((N, B)), 2*np.pi*rng.random((1, B))], axis=0)
    pr = Problem(fp)
    F, Fdx, F2, F2dx = pr.fidelity_and_derivatives_batch(X)
    for b in range(B):
        a = ro.calculate_fidelity_and_derivatives(fp, X[:, b])
        e = eo.calculate_fidelity_and_derivatives(fp, X[:, b])
        got = (F[b], Fdx[:, b], F2[:, b], F2dx[:, :, b])
        for name, g, aa, ee in zip(['F', 'F_dx', 'F_d2err', 'F_d2err_dx'], got, a, e):
            g, aa, ee = np.asarray(g), np.asarray(aa), np.asarray(ee)
            if ee.size == 0: continue
            sc = np.abs(ee).max()
            print(f"N={N} errs={errs} {model[:4]} b={b} {name:11s} |gpu-exact|/max={np.abs(g-ee).max()/sc:.2e}  |fp64oracle-exact|/max={np.abs(aa-ee).max()/sc:.2e}  max={sc:.3e}")
    c, g = pr.cost_and_grad_batch(X, [1e-4]*len(errs))
    b0 = ro.cost_and_gradient(fp, X[:, 0], [1e-4]*len(errs))
    print("cost/grad vs fp64 oracle:", abs(c[0]-b0[0]), np.abs(g[:,0]-b0[1:]).max())
