#!/bin/bash
for dbg in 0 4; do
BSREG_PERSIST_DBG=$dbg BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda_trace.so python profiles/trace_persist.py 100000
done
