#!/bin/bash
# round-2 evidence run: GPU tests, bench (plain), launch list of the bench command, single-pass ncu metrics of the persistent kernel
T=${1:-r2s}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
timeout 600 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${T}_ncu_bench.log 2>&1
M=lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --replay-mode application --clock-control none -k regex:gibbs_persistent --metrics $M --csv --log-file gpurun_out/${T}_persist_metrics.csv \
  python profiles/run_profile.py sar 100000 500 > gpurun_out/${T}_ncu_persist.log 2>&1
tail -3 gpurun_out/${T}_ncu_persist.log
