import sys, time
sys.path.insert(0, '/root/repo')
from bsreg_b200 import capi
from bsreg_b200.synthetic import make_problem
for sort in (False, True):
    for model in ("sem", "sar"):
        p = make_problem(model, n=1_000_000, k_reg=20, knn=10, sort=sort)
        code = capi.MODEL_SEM if model == "sem" else capi.MODEL_SAR
        with capi.Handle(p.y, p.X, model=code, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4, n_chains=64, seed=1) as h:
            h.launch_sweep(3); h.sync()
            t0 = time.perf_counter(); h.launch_sweep(20); h.sync(); dt = time.perf_counter() - t0
            print(f"{model} Morton-sorted input={sort}: {1e6*dt/20:.1f} us per sweep", flush=True)
