mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t2.log
tools/ab2.sh main dc0 direct gs0 sg0 main,B200JPEG_KEEP_PLAIN=1 main,B200JPEG_SYMREC=0 2>&1 | tee gpurun_out/ab_r3b.txt
BENCH_ARGS="--workload cfg3 --batch 32 --steps 3 --warmup 3" tools/ab2.sh main gs0 2>&1 | tee gpurun_out/ab_r3b_cfg3.txt
