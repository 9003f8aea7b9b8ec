mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fullsize or config1 or batch or chunk" > gpurun_out/t4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t4.log
tools/ab2.sh main r7 dw4 fw8 main 2>&1 | tee gpurun_out/ab_r3d.txt
export B200JPEG_BENCH_CACHE=/dev/shm
for ns in 2 3 4; do B200JPEG_STREAMS=$ns timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ns$ns.json 2> gpurun_out/bench_ns$ns.err; python - $ns <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/bench_ns{sys.argv[1]}.json")); print("streams", sys.argv[1], round(d["value"]), "MP/s", round(d["ms_per_step"],2), "ms")
except Exception as e: print("streams", sys.argv[1], "FAILED", e, open(f"gpurun_out/bench_ns{sys.argv[1]}.err").read()[-300:])
PY
done
for ns in 4; do B200JPEG_STREAMS=$ns B200JPEG_CHUNK_IMAGES=32 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ns${ns}c32.json 2> gpurun_out/bench_ns${ns}c32.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_ns4c32.json")); print("streams 4 chunk 32", round(d["value"]), "MP/s", round(d["ms_per_step"],2), "ms")
except Exception as e: print("ns4c32 FAILED", e)
PY
done
