#!/bin/bash
# chains per chain-CTA in the persistent kernel: 4 (default build), 2, 1; full pipeline and chain-bound (dbg 2)
./profiles/micro/fp64_throughput
for v in "" _cpc2 _cpc1; do
  for dbg in 0 2; do
    echo -n "lib${v:-_cpc4} dbg=$dbg "
    BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda$v.so BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64
  done
done
for v in _trace; do BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda$v.so python profiles/trace_persist.py 100000; done
