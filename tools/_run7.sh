mkdir -p gpurun_out
export B200JPEG_BENCH_CACHE=/dev/shm
run() { # tag workload env...
  TAG=$1; W=$2; shift 2
  env "$@" timeout 240 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-parity-gate > gpurun_out/ck_$TAG.json 2> gpurun_out/ck_$TAG.err
  python - $TAG <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ck_{t}.json")); print(f"{t:24s} resident {d['value']:8.0f} MP/s {d['ms_per_step']:7.2f} ms | e2e {d['e2e']['value']:8.0f} MP/s {d['e2e']['ms_per_step']:7.2f} ms | clocks {d['clocks']}")
except Exception as e: print(t, "FAILED", e, open(f"gpurun_out/ck_{t}.err").read()[-300:])
PY
}
run def_cur default A=1
run def_c16 default B200JPEG_CHUNK_IMAGES=16
run def_c8 default B200JPEG_CHUNK_IMAGES=8
run def_c8s4 default B200JPEG_CHUNK_IMAGES=8 B200JPEG_STREAMS=4
run def_c11s3 default B200JPEG_CHUNK_IMAGES=11 B200JPEG_STREAMS=3
run cfg3_cur cfg3 A=1
run cfg3_c32 cfg3 B200JPEG_CHUNK_IMAGES=32
run cfg3_c16 cfg3 B200JPEG_CHUNK_IMAGES=16
run cfg3_c32s4 cfg3 B200JPEG_CHUNK_IMAGES=32 B200JPEG_STREAMS=4
tools/ab2.sh main pu1 pu4 main 2>&1 | tee gpurun_out/ab_pu.txt
