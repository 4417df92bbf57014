"""Per-iteration time of the persistent Gibbs kernel for a few (N, chains): separates the chain-step latency
(N tiny) from the streaming time (few chains)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

cases = [(2000, 64), (100_000, 4), (100_000, 64), (100_000, 128), (1_000_000, 64)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for n, chains in cases:
    p = make_problem("sar", n=n, k_reg=20, knn=10, sort=bool(os.environ.get("BSREG_SORT")))
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=16,
                     n_chains=chains, seed=1) as h:
        h.run(2000, 0)                      # warm-up long enough for the clocks to settle
        h.sync()
        dt = 1e9
        for _ in range(5):                  # best of five blocks of 1000 iterations
            t0 = time.perf_counter()
            h.run_into(0, 1000, None, None)
            h.sync()
            dt = min(dt, time.perf_counter() - t0)
        print(f"n={n:8d} chains={chains:4d}: {1e3 * dt:7.2f} us per iteration, {chains * 1000 / dt / 1e6:6.2f} M chain-iter/s, "
              f"sweep stream {h.sweep_bytes() * 1000 / dt / 1e9:7.0f} GB/s")
