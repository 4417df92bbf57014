"""clock64 trace of one step of the general chain kernel (library built with BSREG_TRACE=1): cycles between trace points."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=100_000, k_reg=20, knn=10)
names = {17: "Gram matrix + state", 18: "rss + sigma^2", 19: "factorisation (beta)", 20: "normals, mu1, beta", 21: "shrinkage",
         22: "factorisation (lambda)", 23: "sar coefficients", 24: "proposal, log-det, decision", 25: "tuning, store, state"}
for prior, pname in ((capi.PRIOR_CONJUGATE, "Conjugate"), (capi.PRIOR_HORSESHOE, "Horseshoe"), (capi.PRIOR_SHRINKAGE, "Shrinkage")):
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, prior=prior, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=8,
                     n_chains=64, seed=1, sng=dict(theta_scale=0.3) if prior == capi.PRIOR_SHRINKAGE else None) as h:
        h.run(100, 0)
        h.launch_chain_step(5)
        h.sync()
        out = np.zeros(64, dtype=np.int64)
        capi.load().bsreg_debug_trace(out.ctypes.data_as(C.c_void_p))
        print(pname)
        prev = out[16]
        for i in range(17, 26):
            print(f"{names[i]:>30s} {out[i] - prev:7d} cycles")
            prev = out[i]
        print(f"{'total':>30s} {out[25] - out[16]:7d} cycles = {(out[25] - out[16]) / 1.965e3:.2f} us at 1965 MHz")
        print("   columns 0,5,10,15 start:", [int(out[28 + i] - out[27]) for i in range(4)], "column 6: rsqrt+scale", out[35] - out[34], "barrier", out[36] - out[35], "update", out[37] - out[36])
        print(f"   inside the first factorisation: fill {out[26] - out[18]}, identity + barrier {out[27] - out[26]}, Cholesky {out[19] - out[27]} cycles")
