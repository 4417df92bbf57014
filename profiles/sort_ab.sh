#!/bin/bash
# does the sweep speed up when the observations are in a locality-preserving (Morton) order?  (existing kernel, gathers through L2)
for sort in "" 1; do
  for dbg in 1 0; do echo -n "sort=${sort:-0} dbg=$dbg "; BSREG_SORT=$sort BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64 1000000,64; done
done
