mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t3.log
tools/ab2.sh main r6 r7 2>&1 | tee gpurun_out/ab_r3c.txt
export B200JPEG_BENCH_CACHE=/dev/shm
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg2_r3c.json 2> gpurun_out/bench_cfg2_r3c.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_cfg2_r3c.json"))
print("cfg2 b256:", round(d["value"]), "MP/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), {k:round(v,2) for k,v in d["roofline"]["stage_ms"].items()})
PY
KERNELS="k_trellis_ac3:2 k_forward_tile:1 k_encode_seq:0 k_trellis_dc_v2:0 k_gather_comp:0 k_block_bits_seq:0" tools/profile3.sh r03 > gpurun_out/profile_r03.out 2>&1; tail -12 gpurun_out/profile_r03.out
