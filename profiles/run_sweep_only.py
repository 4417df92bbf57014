"""cfg4-sized handle, then REPS fused sweeps only (for ncu / timing of the sweep kernel)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "sar"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
p = make_problem(model, n=n, k_reg=20, knn=10)
code = {"sar": capi.MODEL_SAR, "sem": capi.MODEL_SEM, "lm": capi.MODEL_LM}[model]
with capi.Handle(p.y, p.X, model=code, W=p.W if model != "lm" else None, ldet=capi.LDET_CHEBYSHEV if model != "lm" else capi.LDET_NONE, cheb_order=8, cheb_probes=4,
                 n_chains=64, seed=1) as h:
    h.launch_sweep(3)
    h.sync()
    t0 = time.perf_counter()
    h.launch_sweep(reps)
    h.sync()
    dt = time.perf_counter() - t0
    print(f"{model} n={n} sweep: {1e6 * dt / reps:.2f} us per launch (host timed, back to back), "
          f"{h.sweep_bytes() / (dt / reps) / 1e9:.0f} GB/s algorithmic")
