#!/bin/bash
for st in 9 7 5; do for n in 100000 1000000; do echo -n "S=$st "; BSREG_RING_STAGES=$st timeout 40 python profiles/run_sweep_only.py sar $n 240 || echo "rc=$?"; done; done
for n in 2000 100000 1000000; do timeout 40 python profiles/run_sweep_only.py lm $n 240 || echo "rc=$?"; done
timeout 40 python profiles/run_sweep_only.py sar 2000 240
