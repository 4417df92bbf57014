#!/bin/bash
# round-2 final evidence: GPU tests, bench (plain), launch list of the bench command, ncu --set full of the SEM ring sweep
T=${1:-r2f}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_sem -s 3 -c 1 -o gpurun_out/${T}_sem_sweep \
  python profiles/run_sweep_only.py sem 1000000 5 > gpurun_out/${T}_ncu_sem.log 2>&1
python profiles/ncu_summary.py gpurun_out/${T}_sem_sweep.ncu-rep sweep_sem > gpurun_out/${T}_sem_sweep_ncu_full.txt 2>&1; head -24 gpurun_out/${T}_sem_sweep_ncu_full.txt
