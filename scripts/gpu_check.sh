#!/bin/bash
# On-box sequence without a profiler: parity tests -> smoke -> bench (both arms).  (ncu: scripts/gpu_ncu.sh, one pass per call.)
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; cat gpurun_out/bench_ref_${TAG}.json
