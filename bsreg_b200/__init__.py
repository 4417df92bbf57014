"""bsreg_b200 -- B200-native implementation of the bsreg spatial Gibbs-sampler hot path.

`bsreg_b200.interface` mirrors bsreg's user interface (bm / blm / bsar / bsem / ...,
set_options / set_mh); `bsreg_b200.capi` is the ctypes binding of the C ABI
(include/bsreg_cuda.h) behind which the CUDA kernels live.  There is no CPU fallback.
"""
from .interface import (BM, PsiPower, blm, bm, bsar, bsdem, bsdm, bsem, bslx, bsv, set_HS, set_mh, set_NG, set_options,  # noqa: F401
                        set_SAR, set_SEM, set_SLX, set_SNG, set_SV)

__version__ = "0.1.0"
