#!/bin/bash
# Run on the GPU box (under gpurun): one ncu full-set capture of a kernel.  $1 = tag, $2 = kernel regex, $3 = launches to skip
TAG=$1; K=$2; SKIP=${3:-0}
export B200JPEG_BENCH_CACHE=/dev/shm
BENCH="python bench.py --batch 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG $BENCH > gpurun_out/prof_$TAG.out 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
