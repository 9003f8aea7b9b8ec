#!/usr/bin/env python3
"""bench.py -- megapixels/s of the JPEG encode hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            (ours)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU encoder)

A "step" is one pass of the hot path over one batch of synthetic images.
Default workload = BASELINE.json configs[1]: a batch of 256 synthetic
3840x2160 RGB images, q75, 4:2:0, trellis on, baseline (cjpeg -baseline
-quality 75 -sample 2x2).  Images shard across ranks (weak scaling: every rank
encodes its own full batch); the only collective is the final MAX/SUM reduce.

One JSON line on rank 0 (see the task contract): `value` = whole-job MP/s with
inputs resident in HBM; `e2e` = the same metric through the public C-ABI call
with HOST buffers (H2D of the pixels and D2H of the JPEG files inside the timed
region); `roofline` for the dominant kernel; `cpu_baseline` = the unmodified
reference (oracle/_ref) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--switches", default="-baseline -quality 75 -sample 2x2",
                    help="cjpeg switch set naming the profile (config 2 of BASELINE.json)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images tiled to fill the batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return f"batch of {a.batch} synthetic {a.width}x{a.height} RGB, cjpeg {a.switches} (BASELINE.json configs[1] shape)"


# ---------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# the reference's CPU encoder on the host cores (oracle/_ref via refshim)
# ---------------------------------------------------------------------------
def cpu_reference_run(images, switches, threads, reps):
    """Encode len(images)*reps... each thread encodes `reps` images with the
    UNMODIFIED reference library (ctypes releases the GIL).  Returns
    (MP/s, kind, seconds)."""
    from oracle import oracle as O
    import mozjpeg_b200 as mj
    use_ref = O.ref_available()
    h, w = images[0].shape[:2]
    if use_ref:
        O.ref()
        fn = lambda im: O.ref_encode(im, switches)
    else:
        p = mj.params_from_switches(switches, w, h)
        O.orc()
        fn = lambda im: O.oracle_encode(p, im).jpeg
    fn(images[0][:64, :64].copy()) if False else None
    done = [0] * threads

    def work(t):
        for r in range(reps):
            fn(images[(t + r) % len(images)]); done[t] += 1
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    mp = sum(done) * w * h / 1e6
    return mp / dt, ("reference" if use_ref else "port"), dt


def host_threads():
    """Host threads the reference arm can really use: the CPU affinity mask,
    capped by the cgroup CPU quota (the GPU boxes expose 128 logical CPUs but
    cap the container at 16 CPUs' worth of time; more threads only thrash)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return n


def run_reference_arm(a, rank, world):
    """--impl reference: the reference's own CPU implementation, all host threads."""
    if rank != 0:
        return
    from mozjpeg_b200.synth import synth_image
    sw = a.switches.split()
    threads = host_threads()
    imgs = [synth_image(1000 + i, a.width, a.height) for i in range(min(a.distinct, 4))]
    per = a.width * a.height / 1e6
    # size the step so K+W steps end within a few minutes: 1 image per thread per step
    reps = 1
    for _ in range(a.warmup):
        cpu_reference_run(imgs, sw, threads, reps)
    t0 = time.perf_counter(); kind = "reference"
    for _ in range(a.steps):
        _, kind, _ = cpu_reference_run(imgs, sw, threads, reps)
    dt = time.perf_counter() - t0
    val = a.steps * threads * reps * per / dt
    sample = f"{threads * reps} images {a.width}x{a.height} per step, one per host thread"
    line = {"impl": "reference", "metric": "megapixels/sec encode (4K RGB q75 4:2:0)", "value": val, "unit": "MP/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/fp32 (CPU)", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def rank_seeds(rank: int, distinct: int):
    """Synthetic-image seeds of one rank: the batch shards by rank, every rank encodes images of its own."""
    return [1000 * (rank + 1) + i for i in range(distinct)]


def max_over_ranks(value: float, world: int, device) -> float:
    """The contract's timing rule: a multi-rank number is the MAX over ranks (one all-reduce at the end; the data path
    itself has no collective).  `device` is where the process group lives (cuda:<local> for NCCL, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_mp_per_step(world: int, batch: int, width: int, height: int) -> float:
    """Megapixels one step encodes over ALL ranks (weak scaling: every rank has its own full batch)."""
    return world * batch * width * height / 1e6


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mozjpeg_b200 as mj
    from mozjpeg_b200.synth import synth_image

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sw = a.switches.split()
    W, H, B = a.width, a.height, a.batch
    p = mj.params_from_switches(sw, W, H)

    # ---- synthetic inputs: `distinct` images per rank, tiled to B (6.4 GB at the default size >> 126 MB L2)
    base = np.stack([synth_image(seed, W, H) for seed in rank_seeds(rank, a.distinct)])
    host = torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True)
    hb = torch.from_numpy(base)
    for i in range(B):
        host[i].copy_(hb[i % a.distinct])
    devbuf = host.to(dev, non_blocking=False)
    row_pitch, image_stride = W * 3, W * H * 3

    enc = mj.Encoder(local)
    stream = torch.cuda.current_stream()
    enc.set_stream(stream.cuda_stream)

    def step_resident():
        enc.encode_batch_ptr(p, devbuf.data_ptr(), True, row_pitch, image_stride, B, device_only=True)

    def step_e2e():
        enc.encode_batch_ptr(p, host.data_ptr(), False, row_pitch, image_stride, B)
        return sum(enc.output_size(i) for i in range(B))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident (kernel pipeline only) ----
    for _ in range(a.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local); clocks.start()
    l0 = enc.kernel_launches()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    ev0.record(stream)
    for _ in range(a.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = enc.kernel_launches() - l0
    # per-kernel times for the roofline: one extra, untimed pass with a single compute stream
    # (in the timed region consecutive chunks overlap on two streams, so stage intervals overlap too)
    enc.set_streams(1)
    step_resident()
    torch.cuda.synchronize()
    stage_acc = {k: v * a.steps for k, v in enc.stage_times().items()}
    resident_chunk = enc.chunk_images()
    enc.set_streams(2)
    clk = clocks.stop()
    ms_total = max_over_ranks(ms_total, world, dev)
    mp_per_step = whole_job_mp_per_step(world, B, W, H)
    value = mp_per_step * a.steps / (ms_total / 1e3)

    # ---- end to end through the public API: host pixels in, JPEG files out ----
    e2e = None
    if not a.no_e2e:
        jpeg_bytes = step_e2e()                       # warm (pinned output buffers get allocated)
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(a.steps):
            jpeg_bytes = step_e2e()
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3     # includes host-side file assembly, which events do not see
        e2e_ms = max_over_ranks(wall_ms, world, dev)
        e2e = {"value": mp_per_step * a.steps / (e2e_ms / 1e3), "unit": "MP/s", "h2d_bytes_per_step": B * W * H * 3,
               "d2h_bytes_per_step": int(jpeg_bytes), "ms_per_step": e2e_ms / a.steps, "timer": "host wall clock around the API calls, max over ranks"}
    else:
        step_e2e_bytes = 0

    # ---- roofline of the dominant kernel (CUDA events inside the library, averaged over the timed steps) ----
    stages = {k: v / a.steps for k, v in stage_acc.items() if k not in ("h2d", "h2d_wait")}
    dom = max(stages, key=stages.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    out_bytes = enc.last_scan_bytes() / B if not a.no_e2e else 0.0
    # one launch of the dominant kernel covers one chunk of the batch; stage times are summed over the chunks,
    # so bytes per launch / average launch duration == batch bytes / summed stage time
    imgs_per_launch = resident_chunk
    n_launch = (B + imgs_per_launch - 1) // imgs_per_launch
    alg_bytes = imgs_per_launch * (W * H * 3 + out_bytes)            # input bytes + JPEG bytes per image
    achieved = alg_bytes / (stages[dom] / n_launch / 1e3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        if prof.get("kernel") == dom:
            traffic = prof.get("dram_bytes_per_image") * imgs_per_launch
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": stages[dom] / n_launch, "launches_per_step": n_launch,
                "images_per_launch": imgs_per_launch,
                "algorithmic_bytes_per_launch": alg_bytes, "stage_ms": stages}

    # ---- the reference's CPU encoder on this box's host cores (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = host_threads()
        imgs = [base[i] for i in range(min(4, a.distinct))]
        reps = 3
        v, kind, secs = cpu_reference_run(imgs, sw, threads, reps)
        cpu = {"value": v, "unit": "MP/s", "cores": threads, "kind": kind,
               "sample": f"{threads * reps} images {W}x{H} ({reps} per host thread, {secs:.1f} s wall = {secs * threads:.0f} CPU-seconds), same switches"}

    if rank == 0:
        line = {"metric": "megapixels/sec encode (4K RGB q75 4:2:0)", "value": value, "unit": "MP/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8 in / int32 DCT / fp32 trellis costs / u8 out", "data": "synthetic",
                "config": {"workload": workload_name(a), "images_per_gpu": B, "global_images": B * world,
                           "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2" % (B * W * H * 3 / 1e9), "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    enc.close()


if __name__ == "__main__":
    main()
