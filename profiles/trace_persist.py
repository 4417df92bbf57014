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
    out = np.zeros(64, dtype=np.int64)
    capi.load().bsreg_debug_trace_persist(out.ctypes.data_as(C.c_void_p))
    t = out.reshape(4, 16)
    t0 = t[0, 0]
    pts = {0: "iteration start", 1: "rows scaled by sigma^2", 2: "elimination done", 7: "random numbers done", 8: "next Gram matrix fetched",
           9: "phase A done (before barrier)", 3: "barrier 1 passed", 4: "phase B done, barrier 2 passed", 10: "phase C done (before barrier)",
           5: "barrier 3 passed", 6: "phase D done"}
    order = [0, 1, 2, 7, 8, 9, 3, 4, 10, 5, 6]
    print(f"{'cycles since the start of the iteration':>40s}" + "".join(f"{w:>10s}" for w in ("E", "F", "R", "S")))
    for k in order:
        print(f"{pts[k]:>40s}" + "".join(f"{(t[w, k] - t0) if t[w, k] else 0:10d}" for w in range(4)))
    out = np.array([t0] + [0] * 5 + [t[:, 6].max()])
    print(f"{'total':>24s} {out[6] - out[0]:7d} cycles = {(out[6] - out[0]) / 1.965e3:.2f} us at 1965 MHz (n = {n})")
