#!/bin/bash
# quick A/B of the persistent kernel against the build of the last commit (lib variant "base"), same box
python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py -m gpu -x -q -k "persistent or statistic or tensor_core or slot_major" 2>&1 | tail -3
for v in _base ""; do
  for dbg in 0 2; do echo -n "lib${v:-_new} dbg=$dbg "; BSREG_CUDA_LIB=$PWD/bsreg_b200/lib/libbsreg_cuda$v.so BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64; done
done
