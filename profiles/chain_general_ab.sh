#!/bin/bash
# general chain step (priors other than Independent): tests, microseconds per iteration at cfg4 size, clock64 trace of one step
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/time_priors.py 100000 sar,lm
BSREG_NO_CHAIN_LOOP=1 BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda_trace.so timeout 300 python profiles/trace_chain_general.py
