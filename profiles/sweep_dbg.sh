#!/bin/bash
# ablation of the standalone ring sweep: 0 full, 2 no math, 4 no gather, 6 neither (TMA streaming only)
for n in 100000 1000000; do
  for dbg in 0 2 4 6; do
    echo -n "dbg=$dbg "; BSREG_DBG=$dbg timeout 60 python profiles/run_sweep_only.py sar $n 240 2>&1 | tail -2
  done
done
