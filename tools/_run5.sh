mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t5.log
tools/ab2.sh main main,B200JPEG_FWD_STATS=0 main 2>&1 | tee gpurun_out/ab_r3e.txt
export B200JPEG_BENCH_CACHE=/dev/shm
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r3e.json 2> gpurun_out/bench_r3e.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r3e.json")); print("cfg2 b256", round(d["value"]), "MP/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print("bench FAILED", e, open("gpurun_out/bench_r3e.err").read()[-400:])
PY
