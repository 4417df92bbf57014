"""Wall clock of bsreg_create / bsreg_run / destroy for the cfg4 fit from PINNED host buffers (what bench.py's e2e times),
for several Chebyshev probe counts: how much of create is the trace moments?"""
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bsreg_b200 import capi            # noqa: E402
from bsreg_b200.synthetic import make_problem   # noqa: E402

p = make_problem("sar", n=100_000, k_reg=20, knn=10)


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


ty, tXt = pinned(p.y), pinned(np.ascontiguousarray(p.X.T))
trp, tci, tva = pinned(p.W.indptr.astype(np.int32)), pinned(p.W.indices.astype(np.int32)), pinned(p.W.data)
Wp = sp.csr_matrix((tva.numpy(), tci.numpy(), trp.numpy()), shape=p.W.shape)
yp, Xp = ty.numpy(), tXt.numpy().T
draws = torch.empty((64, 22, 1000), dtype=torch.float64, pin_memory=True)
for probes in (64, 64, 16, 4, 64):
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        h = capi.Handle(yp, Xp, model=capi.MODEL_SAR, W=Wp, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=probes,
                        n_chains=64, seed=1 + rep)
        t1 = time.perf_counter()
        h.run_into(500, 1000, None, draws.data_ptr())
        t2 = time.perf_counter()
        h.close()
        t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2))
    a = 1e3 * np.array(ts[1:]).mean(0)
    print(f"probes {probes:3d}: create {a[0]:6.2f} ms, run(500 + 1000) {a[1]:6.2f} ms, destroy {a[2]:5.2f} ms, fit {a.sum():6.2f} ms")
