#!/bin/bash
# Development aid, runs on the GPU box: stage times of the default pipeline for the main library and each variant named.
#   tools/ab.sh v1 v2   (variants built by tools/build_variant.sh v1 -DSOME_SWITCH=1 ...)
for V in "" "$@"; do
  B200JPEG_LIB_VARIANT=$V B200JPEG_CHUNK_IMAGES=64 B200JPEG_STREAMS=1 timeout 150 python bench.py --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_${V:-main}.json 2> gpurun_out/ab_${V:-main}.err
  python - "$V" <<'PY'
import json, sys
v = sys.argv[1] or "main"
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json")); s = d["roofline"]["stage_ms"]
    print(f"{v:8s} {d['ms_per_step']:7.3f} ms  fwd {s['forward']:.3f}  ac {s['trellis_ac']:.3f}  dc {s['trellis_dc']:.3f}  enc {s['encode']:.3f}  " + " ".join(f"{k}={x:.2f}" for k, x in s.items() if k not in ('forward','trellis_ac','trellis_dc','encode','dummy')))
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-300:])
PY
done
