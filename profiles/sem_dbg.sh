#!/bin/bash
# SEM ring sweep: which role bounds a tile?  BSREG_SEM_DBG bits: 1 = no mma, 2 = no lag formation, 4 = no foreign-row gathers
for d in 0 1 2 4 3 6 7; do echo -n "BSREG_SEM_DBG=$d "; BSREG_SEM_DBG=$d timeout 300 python profiles/run_sweep_only.py sem 1000000 20; done
