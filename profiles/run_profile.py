"""Short cfg4 run for ncu: create the handle (setup kernels) then ITERS Gibbs iterations."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "sar"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 60
p = make_problem(model, n=n, k_reg=20, knn=10)
code = capi.MODEL_SAR if model == "sar" else capi.MODEL_SEM
with capi.Handle(p.y, p.X, model=code, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=64,
                 n_chains=64, seed=1) as h:
    d = h.run(0, iters)
    print("ok", d.shape, float(d[:, -1, -1].mean()))
