mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t1.log
tools/ab2.sh main main,B200JPEG_SYMREC=0 g1,B200JPEG_SYMREC=0 g1 g4 occ12 2>&1 | tee gpurun_out/ab_r3a.txt
BENCH_ARGS="--workload cfg3 --batch 32 --steps 3 --warmup 3" tools/ab2.sh main g1 g4 2>&1 | tee gpurun_out/ab_r3a_cfg3.txt
export B200JPEG_BENCH_CACHE=/dev/shm
for v in "" 0; do B200JPEG_SYMREC=$v B200JPEG_CHUNK_IMAGES=256 B200JPEG_STREAMS=1 timeout 300 python bench.py --workload cfg4 --batch 256 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/cfg4_symrec_$v.json 2> gpurun_out/cfg4_symrec_$v.err; python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/cfg4_symrec_{v}.json"))
    for q,s in d["config"]["sweep"].items(): print("symrec",v or "on",q, f"{s['ms_per_step']:.3f} ms", " ".join(f"{k}={x:.2f}" for k,x in s["stage_ms"].items() if x>=0.05))
except Exception as e: print("cfg4",v,"FAILED",e, open(f"gpurun_out/cfg4_symrec_{v}.err").read()[-300:])
PY
done
