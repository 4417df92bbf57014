#!/bin/bash
# timing matrix of the fused sweep: model x N x kernel variant (host-timed, back-to-back launches)
for n in 100000 1000000; do
  for m in lm sar; do
    python profiles/run_sweep_only.py $m $n 200
    BSREG_SWEEP_LEGACY=1 python profiles/run_sweep_only.py $m $n 200 | sed 's/^/legacy /'
  done
done
