#!/bin/bash
# full ncu capture (with source) of kernels matching a regex in an arbitrary command
# usage: bash scripts/gpu_ncu_full_cmd.sh <tag> <kernel-regex> <skip> <count> <cmd...>
TAG=$1; PAT=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${PAT} -s ${SKIP} -c ${CNT} -f -o gpurun_out/prof_${TAG} "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
echo rc=$?
