"""clock64 stamps of the SEM ring sweep's roles over eight consecutive tiles of CTA 0 (library built with BSREG_TRACE=1:
BSREG_TRACE=1 python -m bsreg_b200.build --variant=trace; BSREG_CUDA_LIB=bsreg_b200/lib/libbsreg_cuda_trace.so)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
p = make_problem("sem", n=n, k_reg=20, knn=10)
with capi.Handle(p.y, p.X, model=capi.MODEL_SEM, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4, n_chains=64, seed=1) as h:
    h.launch_sweep(3)
    h.sync()
    out = np.zeros(128, dtype=np.int64)
    capi.load().bsreg_debug_trace_sem(out.ctypes.data_as(C.c_void_p))
t = out.reshape(8, 16)
t0 = t[0, 0]
names = {0: "g:start", 1: "g:gathers issued", 2: "g:own gathers landed", 3: "g:tile landed (full)", 4: "g:sync", 5: "g:lags formed", 6: "g:sync",
         8: "m:start", 9: "m:lags ready", 10: "m:done", 12: "p:start", 13: "p:stage free"}
for j in range(8):
    print(f"tile {40 + j}: " + "  ".join(f"{names[k]} {t[j, k] - t0}" for k in sorted(names) if t[j, k]))
