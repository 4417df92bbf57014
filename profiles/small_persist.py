"""Small persistent-kernel run (d = 16, 6 chains, 2 launches) for compute-sanitizer."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=5000, k_reg=14, knn=7)
with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=10, cheb_probes=4,
                 n_chains=6, seed=1) as h:
    d = h.run(6, 6)
    d2 = h.run(0, 5)
    print("ok", d.shape, float(d[:, -1, -1].mean()), float(d2[:, -1, -1].mean()))
