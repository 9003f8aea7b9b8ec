#!/bin/bash
# Run on the GPU box (under gpurun).  $1 = tag for the output files.
# 1) launch list with per-launch device time, 2) full-set capture of the heaviest kernels.
TAG=${1:-r01}
BENCH="python bench.py --batch 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_$TAG.out 2>&1
for K in ${KERNELS:-k_trellis_ac k_forward_tile}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-1} -c 1 -f -o gpurun_out/prof_${TAG}_$K $BENCH > gpurun_out/prof_${TAG}_$K.out 2>&1
done
ls -la gpurun_out/ | tail -8
