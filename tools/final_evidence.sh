#!/bin/bash
# Run on the GPU box (under gpurun) at the end of a round: ncu launch list + full-set digests (tools/profile4.sh), the DRAM
# traffic table the bench reports (tools/profile_digest3.py, run here so that the bench lines below carry it), one bench
# line per BASELINE.json configuration, two control measurements, and the sanitizer logs.  $1 = tag (r03).  Everything
# lands in gpurun_out/; the most valuable artefacts come first (a call may be cut short by the GPU budget).
TAG=${1:-r03}
mkdir -p gpurun_out
tools/profile4.sh $TAG > gpurun_out/profile_$TAG.out 2>&1
python tools/profile_digest3.py $TAG $TAG 16 > gpurun_out/digest_$TAG.out 2>&1
cp profiles/${TAG}_launches_$TAG.md profiles/${TAG}_launches_$TAG.csv profiles/dominant_kernel_traffic.json gpurun_out/ 2>/dev/null
export B200JPEG_BENCH_CACHE=/dev/shm
timeout 420 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg2_b256.json 2> gpurun_out/bench_${TAG}_cfg2_b256.err
for W in cfg3 cfg4 cfg5 default; do
  timeout 420 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
done
python - $TAG <<'PY'
import json, sys
for w in ("cfg2_b256", "cfg3", "cfg4", "cfg5", "default"):
    try:
        d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_{w}.json"))
        print(w, round(d["value"]), "MP/s", round(d["ms_per_step"], 2), "ms; e2e", d["e2e"] and round(d["e2e"]["value"]), "; roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "traffic", d["roofline"].get("traffic"), "; clocks", d["clocks"])
    except Exception as e:
        print(w, "FAILED", e, open(f"gpurun_out/bench_{sys.argv[1]}_{w}.err").read()[-300:])
PY
# controls: the library default profile on ONE compute stream (what the two-chunk split of resident batches buys), and the
# AC trellis with its predecessor loop unrolled 1x / 4x (variants built by tools/build_variant.sh, if present)
B200JPEG_STREAMS=1 timeout 300 python bench.py --workload default --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-gate > gpurun_out/bench_${TAG}_default_1stream.json 2> gpurun_out/bench_${TAG}_default_1stream.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_default_1stream.json')); print('default, one stream:', round(d['value']), 'MP/s', round(d['ms_per_step'],2), 'ms')" 2>&1 | tail -1
VARS=""; for v in pu1 pu4; do [ -f mozjpeg_b200/variants/libb200jpeg_$v.so ] && VARS="$VARS $v"; done
[ -n "$VARS" ] && tools/ab2.sh main $VARS main 2>&1 | tee gpurun_out/ab_${TAG}_pu.txt
if [ "$SANITIZE" != "0" ]; then tools/sanitize.sh 2>&1 | tail -8; fi
du -sh gpurun_out
