#!/bin/bash
# launch list of the bench command + ncu --set full of the SEM ring sweep (each after its command has run plainly)
T=${1:-r2g2}
python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_bench_short.json 2> gpurun_out/${T}_bench_short.err; tail -c 300 gpurun_out/${T}_bench_short.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/${T}_launches.csv
timeout 300 python profiles/run_sweep_only.py sem 1000000 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_sem -s 3 -c 1 -o gpurun_out/${T}_sem_sweep \
  python profiles/run_sweep_only.py sem 1000000 5 > gpurun_out/${T}_ncu_sem.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/${T}_ncu_sem.log
python profiles/ncu_summary.py gpurun_out/${T}_sem_sweep.ncu-rep sweep_sem > gpurun_out/${T}_sem_sweep_ncu_full.txt 2>&1; head -22 gpurun_out/${T}_sem_sweep_ncu_full.txt
