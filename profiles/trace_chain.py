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
    out = np.zeros(64, dtype=np.int64)
    capi.load().bsreg_debug_trace(out.ctypes.data_as(C.c_void_p))
    # thread 0 (critical-path warp) passes points 0,1,3..9; point 2 belongs to warp 3 and is not recorded by tid 0
    names = {1: "loads+sync", 2: "stage rows", 3: "row load + sigma2 wait", 4: "LDL elimination", 5: "Yd store + sync", 6: "dots + beta",
             7: "sar sums", 8: "mh decision", 9: "store"}
    prev = out[0]
    for i in (1, 2, 3, 4, 5, 6, 7, 8, 9):
        print(f"{names[i]:>24s} {out[i] - prev:7d} cycles")
        prev = out[i]
    print(f"{'total':>24s} {out[9] - out[0]:7d} cycles = {(out[9] - out[0]) / 1.965e3:.2f} us at 1965 MHz")
