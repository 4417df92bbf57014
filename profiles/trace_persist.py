"""clock64 trace of the last chain step inside the persistent kernel (library built with BSREG_TRACE=1)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
p = make_problem("sar", n=n, k_reg=20, knn=10)
with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=16,
                 n_chains=64, seed=1) as h:
    h.run(300, 0)
    h.sync()
    out = np.zeros(32, dtype=np.int64)
    capi.load().bsreg_debug_trace_persist(out.ctypes.data_as(C.c_void_p))
    names = ["", "wait sweep + sync", "stage G + sync", "stage rows", "row load + sigma2 wait", "elimination", "Yd store + sync",
             "dots + sync", "beta, t0/t1, sums, MH", "store"]
    for i in range(1, 10):
        print(f"{names[i]:>24s} {out[i] - out[i - 1]:7d} cycles")
    print(f"{'total':>24s} {out[9] - out[0]:7d} cycles = {(out[9] - out[0]) / 1.965e3:.2f} us at 1965 MHz (n = {n})")
