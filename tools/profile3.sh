#!/bin/bash
# Run on the GPU box (under gpurun).  $1 = tag.  One ncu pass over one bench step (16 images of 3840x2160): per launch
# the device time and the DRAM bytes read / written; then full-set captures of the heaviest kernels.
TAG=${1:-r02}
export B200JPEG_BENCH_CACHE=/dev/shm B200JPEG_CHUNK_IMAGES=16
BENCH="python bench.py --batch 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity-gate ${BENCH_EXTRA}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_$TAG.out 2>&1
for SPEC in ${KERNELS:-k_trellis_ac3:2 k_forward_tile:1}; do
  K=${SPEC%%:*}; SKIP=${SPEC#*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG}_$K $BENCH > gpurun_out/prof_${TAG}_$K.out 2>&1
done
ls -la gpurun_out/ | grep $TAG
