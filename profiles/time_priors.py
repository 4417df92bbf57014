"""cfg4-sized SAR / LM fits under the four coefficient priors: microseconds per Gibbs iteration (64 chains, stream and cached)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
models = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sar", "lm"]
for model in models:
    p = make_problem(model, n=n, k_reg=20, knn=10)
    code = {"sar": capi.MODEL_SAR, "lm": capi.MODEL_LM, "sem": capi.MODEL_SEM}[model]
    for prior, name in ((capi.PRIOR_INDEPENDENT, "Independent"), (capi.PRIOR_CONJUGATE, "Conjugate"),
                        (capi.PRIOR_HORSESHOE, "Horseshoe"), (capi.PRIOR_SHRINKAGE, "Shrinkage")):
        for mode, mname in ((capi.STATS_STREAM, "stream"), (capi.STATS_CACHED, "cached")):
            with capi.Handle(p.y, p.X, model=code, prior=prior, W=p.W if model != "lm" else None,
                             ldet=capi.LDET_CHEBYSHEV if model != "lm" else capi.LDET_NONE, cheb_order=50, cheb_probes=8,
                             n_chains=64, seed=1, stats_mode=mode, sng=dict(theta_scale=0.3) if prior == capi.PRIOR_SHRINKAGE else None) as h:
                h.run(240, 0)
                h.run(0, 480)                 # the draw buffers exist before the timed run
                h.sync()
                t0 = time.perf_counter()
                d = h.run(0, 480)
                h.sync()
                dt = time.perf_counter() - t0
                print(f"{model} {name:12s} {mname:6s}: {1e6 * dt / 480:7.2f} us per iteration, {64 * 480 / dt / 1e6:6.2f} M chain-iter/s, "
                      f"last column mean {d[:, -1, :].mean():.4f}", flush=True)
