#!/bin/bash
# ONE ncu pass per invocation (one profiler run per gpurun call), always after the same command has exited 0 without ncu.
#   bash scripts/gpu_ncu.sh launches <tag>                            launch list of `python bench.py --steps 2 --warmup 1`
#   bash scripts/gpu_ncu.sh full <tag> <kernel-regex> <skip> <count> <cmd...>   --set full capture (with source) of the matching launches of <cmd>
MODE=$1; TAG=${2:-r02}; shift 2
mkdir -p gpurun_out
if [ "$MODE" = "launches" ]; then
    python bench.py --steps 2 --warmup 1 > gpurun_out/plain_${TAG}.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
        python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_${TAG}.log 2>&1
    echo "ncu launches rc=$?"
else
    PAT=$1; SKIP=$2; CNT=$3; shift 3
    "$@" > gpurun_out/plain_${TAG}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:${PAT} -s ${SKIP} -c ${CNT} -f -o gpurun_out/prof_${TAG} "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
    echo "ncu full rc=$?"
    tail -3 gpurun_out/ncu_full_${TAG}.log
fi
