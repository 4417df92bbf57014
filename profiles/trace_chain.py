"""clock64 trace of one chain step (library built with BSREG_TRACE=1): cycles between trace points."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=100_000, k_reg=20, knn=10)
with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=64,
                 n_chains=64, seed=1) as h:
    h.run(100, 0)
    h.launch_chain_step(5)
    h.sync()
    out = np.zeros(32, dtype=np.int64)
    capi.load().bsreg_debug_trace(out.ctypes.data_as(C.c_void_p))
    names = ["start", "loads+sync", "rss", "gamma wait", "sigma2 sync", "build T", "chol", "syncthreads", "dots+beta",
             "sar sums", "mh decision", "store"]
    t = out[:12]
    for i in range(1, 12):
        print(f"{names[i]:>14s} {t[i] - t[i - 1]:7d} cycles")
    print(f"{'total':>14s} {t[11] - t[0]:7d} cycles = {(t[11] - t[0]) / 1.965e3:.2f} us at 1965 MHz")
