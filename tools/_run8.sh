mkdir -p gpurun_out
export BENCH_ARGS="--workload cfg3 --batch 32 --steps 3 --warmup 3 --no-parity-gate"
tools/ab2.sh main dcearly old820a181 old41e446e main 2>&1 | tee gpurun_out/ab_r03_cfg3_bisect.txt
