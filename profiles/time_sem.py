"""cfg5 on one GPU's share: synthetic SEM, N = 1e6, kNN(10) W, K = 20, 64 chains (512 chains over 8 GPUs), stream and cached."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
p = make_problem("sem", n=n, k_reg=20, knn=10)
for mode, name in ((capi.STATS_STREAM, "stream"), (capi.STATS_CACHED, "cached")):
    with capi.Handle(p.y, p.X, model=capi.MODEL_SEM, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=8,
                     n_chains=64, seed=1, stats_mode=mode) as h:
        h.run(100, 0)
        h.sync()
        t0 = time.perf_counter()
        d = h.run(0, 200)
        h.sync()
        dt = time.perf_counter() - t0
        print(f"SEM n={n} 64 chains {name}: {1e6 * dt / 200:8.1f} us per iteration, {64 * 200 / dt / 1e6:6.3f} M chain-iter/s, "
              f"sweep stream {h.sweep_bytes() * 200 / dt / 1e9:6.0f} GB/s, lambda mean {d[:, -1, :].mean():.3f} (true {p.lam})")
