#!/bin/bash
# On-box sequence: parity tests -> smoke -> bench (both arms) -> ncu launch list -> ncu full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; cat gpurun_out/bench_ref_${TAG}.json
if [ "${NCU:-1}" = "1" ]; then
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'boot_rank_loop|boot_block_rows' -s 6 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
fi
ls -la gpurun_out
