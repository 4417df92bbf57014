#!/bin/bash
# quick check of the persistent kernel: parity tests, per-iteration time (full / sweep-bound / chain-bound), chain-bound trace
python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
for dbg in 0 1 2; do echo -n "dbg=$dbg "; BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64; done
BSREG_PERSIST_DBG=2 BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda_trace.so python profiles/trace_persist.py 100000
