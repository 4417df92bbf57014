#!/bin/bash
# SEM ring sweep: tests + sweep-only timing at cfg5's share (lag formation: BSREG_SEM_LAG=0 lane per column, 1 quarter warp per row) + trace
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sem" 2>&1 | tail -3
for v in 0 1; do echo -n "BSREG_SEM_LAG=$v "; BSREG_SEM_LAG=$v timeout 300 python profiles/run_sweep_only.py sem 1000000 20; done
for d in 1 2 4 7; do echo -n "BSREG_SEM_DBG=$d "; BSREG_SEM_DBG=$d timeout 300 python profiles/run_sweep_only.py sem 1000000 20; done
BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda_trace.so timeout 300 python profiles/trace_sem.py
