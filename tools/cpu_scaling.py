#!/usr/bin/env python3
"""How the reference's CPU encoder scales with threads on this host (for cpu_baseline.cores)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from mozjpeg_b200.synth import synth_image
w, h = 3840, 2160
imgs = [synth_image(1000 + i, w, h) for i in range(2)]
sw = "-baseline -quality 75 -sample 2x2".split()
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cgroup cpu.max: n/a", e)
for t in (1, 4, 8, 16, 32, 64, 128):
    if t > (os.cpu_count() or 1): break
    v, kind, secs = bench.cpu_reference_run(imgs, sw, t, 1)
    print(f"threads {t:4d}: {v:8.1f} MP/s  ({secs:.2f} s, {kind})", flush=True)
