#!/bin/bash
# persistent Gibbs kernel at cfg4: 0 = full, 1 = chains skip the step (sweep-bound), 2 = sweeps stop after 3 iterations (chain-bound)
for dbg in 0 1 2; do
  echo -n "dbg=$dbg "; BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64 100000,4 2000,64
done
