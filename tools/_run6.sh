mkdir -p gpurun_out
export B200JPEG_BENCH_CACHE=/dev/shm
for spec in "65 2" "128 2" "86 2" "43 2" "32 2" "86 3" "52 3"; do set -- $spec
  B200JPEG_CHUNK_IMAGES=$1 B200JPEG_STREAMS=$2 timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-gate > gpurun_out/bench_c$1_s$2.json 2> gpurun_out/bench_c$1_s$2.err
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/bench_c{sys.argv[1]}_s{sys.argv[2]}.json")); print("chunk", sys.argv[1], "streams", sys.argv[2], round(d["value"]), "MP/s", round(d["ms_per_step"],2), "ms")
except Exception as e: print("chunk", sys.argv[1], "FAILED", e, open(f"gpurun_out/bench_c{sys.argv[1]}_s{sys.argv[2]}.err").read()[-300:])
PY
done
