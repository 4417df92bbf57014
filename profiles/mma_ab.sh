#!/bin/bash
# Gram update of the persistent sweep: vector FMA (BSREG_NO_MMA=1) against FP64 mma (default), sweep-bound and full
python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
for dbg in 1 0; do echo -n "vector dbg=$dbg "; BSREG_NO_MMA=1 BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64; done
for dbg in 1 0 2; do echo -n "mma    dbg=$dbg "; BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64; done
echo -n "mma n=1e6 "; timeout 120 python profiles/time_persistent.py 1000000,64
timeout 60 python profiles/run_sweep_only.py sar 100000 240
