# trimmed end-of-round evidence for the final build: ncu launch list with DRAM bytes (-> traffic table), then the cfg2 bench line
TAG=r03
mkdir -p gpurun_out
export B200JPEG_BENCH_CACHE=/dev/shm
B200JPEG_CHUNK_IMAGES=16 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --batch 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity-gate > gpurun_out/launches_$TAG.out 2>&1
python tools/profile_digest3.py $TAG $TAG 16 > gpurun_out/digest_$TAG.out 2>&1
cp profiles/${TAG}_launches_$TAG.md profiles/${TAG}_launches_$TAG.csv profiles/dominant_kernel_traffic.json gpurun_out/ 2>/dev/null
tail -n 16 profiles/${TAG}_launches_$TAG.md
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg2_b256.json 2> gpurun_out/bench_${TAG}_cfg2_b256.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r03_cfg2_b256.json"))
print("cfg2", round(d["value"]), "MP/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]), "; roofline", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], d["config"]["parity_gate"], d["clocks"])
print({k: round(v, 2) for k, v in d["roofline"]["stage_ms"].items()})
PY
