#!/bin/bash
# Run on the GPU box (under gpurun).  $1 = tag.  Like profile3.sh, but the full-set reports (30 MB each with the
# source pages) are digested ON THE BOX (tools/ncu_summary.py + tools/ncu_stalls.py) and only the text comes back:
# gpurun_out/ is limited to 64 MiB.   KERNELS="name:skip ..." selects the launches (default: the two heaviest kernels).
TAG=${1:-r03}
export B200JPEG_BENCH_CACHE=/dev/shm B200JPEG_CHUNK_IMAGES=16
BENCH="python bench.py --batch 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity-gate ${BENCH_EXTRA}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_$TAG.out 2>&1
for SPEC in ${KERNELS:-k_trellis_ac3:2 k_forward_tile:1}; do
  K=${SPEC%%:*}; SKIP=${SPEC#*:}
  REP=/tmp/prof_${TAG}_$K
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $REP $BENCH > gpurun_out/prof_${TAG}_$K.out 2>&1
  {
    echo "# ncu --set full --clock-control none --import-source on, kernel $K (launch $SKIP of the run), capture $TAG (batch 16 of 3840x2160)"
    echo "# headline metrics, hot SASS regions (instruction index range, #instr, executions per instr, share of warp-instructions, share of stall samples, active threads), stall attribution"
    python tools/ncu_summary.py $REP.ncu-rep
    echo; echo "# ---- stall samples by reason, top instructions"
    python tools/ncu_stalls.py $REP.ncu-rep 6
  } > gpurun_out/prof_${TAG}_$K.txt 2>&1
  rm -f $REP.ncu-rep
done
ls -la gpurun_out/ | grep $TAG
