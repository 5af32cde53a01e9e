#!/bin/bash
# ncu launch list + full capture of the bench's kernels.  Usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex>
TAG=${1:-r02}
PAT=${2:-boot_}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${PAT} -s 6 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_${TAG}.log
