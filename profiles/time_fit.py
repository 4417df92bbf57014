"""Wall-clock phases of one cfg4 fit through the C ABI (create / burn-in / sampling)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=100_000, k_reg=20, knn=10)
for rep in range(3):
    t0 = time.perf_counter()
    h = capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=64,
                    n_chains=64, seed=1 + rep)
    h.sync()
    t1 = time.perf_counter()
    h.run(500, 0)
    t2 = time.perf_counter()
    d = h.run(0, 1000)
    t3 = time.perf_counter()
    h.close()
    t4 = time.perf_counter()
    print(f"create {1e3 * (t1 - t0):.1f} ms, burn(500) {1e3 * (t2 - t1):.1f} ms, save(1000) {1e3 * (t3 - t2):.1f} ms, "
          f"close {1e3 * (t4 - t3):.1f} ms; lambda mean {d[:, -1, :].mean():.4f}")
