#!/bin/bash
# chains per chain-CTA in the persistent kernel (side builds): full pipeline, sweep-bound (1), chain-bound (2)
for v in "" _cpc3 _cpc2; do
  for dbg in 0 1 2; do
    echo -n "lib${v:-_cpc4} dbg=$dbg "
    BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda$v.so BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64
  done
done
