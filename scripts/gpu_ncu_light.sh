#!/bin/bash
# A few metrics for every boot_* kernel launch of a short bench run.
TAG=${1:-light}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:boot_ -c 8 --csv --log-file gpurun_out/light_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_light_${TAG}.log 2>&1
echo rc=$?
