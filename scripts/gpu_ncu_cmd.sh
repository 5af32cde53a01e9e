#!/bin/bash
# light ncu metrics for an arbitrary command: bash scripts/gpu_ncu_cmd.sh <tag> <kernel-regex> <cmd...>
TAG=$1; PAT=$2; shift 2
mkdir -p gpurun_out
"$@" > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:${PAT} -c 12 --csv --log-file gpurun_out/light_${TAG}.csv "$@" > gpurun_out/ncu_light_${TAG}.log 2>&1
echo rc=$?
