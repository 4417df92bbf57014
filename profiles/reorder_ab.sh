#!/bin/bash
# breadth-first relabelling at create + de-duplicated gathers: off / on, sweep-bound (dbg 1), full (0) and chain-bound (2)
for ro in 0 1; do
  for dbg in 1 0 2; do echo -n "reorder=$ro dbg=$dbg "; BSREG_REORDER=$ro BSREG_PERSIST_DBG=$dbg timeout 120 python profiles/time_persistent.py 100000,64 1000000,64; done
done
