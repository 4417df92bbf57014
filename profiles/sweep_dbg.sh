#!/bin/bash
for n in 2000 100000; do
  for m in lm sar; do
    for dbg in 0 1 2 4 8 9 15; do
      echo -n "dbg=$dbg "; BSREG_DBG=$dbg python profiles/run_sweep_only.py $m $n 300
    done
  done
done
