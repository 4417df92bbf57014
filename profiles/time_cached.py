"""Per-iteration time of the launch-per-step paths (cached statistics; graph of sweep + chain step)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=100_000, k_reg=20, knn=10)
for stats in (capi.STATS_CACHED, capi.STATS_STREAM):
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=16,
                     n_chains=64, seed=1, stats_mode=stats) as h:
        h.run_into(0, 480, None, None)
        h.sync()
        t0 = time.perf_counter()
        h.run_into(0, 960, None, None)
        h.sync()
        dt = time.perf_counter() - t0
        print(f"stats={'cached' if stats else 'stream'} persistent={'off' if os.environ.get('BSREG_NO_PERSISTENT') else 'on'}: "
              f"{1e6 * dt / 960:.2f} us per iteration, launches {h.kernel_launches()}")
