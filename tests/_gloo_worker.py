"""Worker of tests/test_multirank_host.py: one of WORLD_SIZE processes (gloo, CPU) exercising the rank logic of
bench.py - the per-rank shard of the synthetic batch, the MAX-over-ranks reduction and the whole-job accounting -
with the CPU oracle standing in for the device encoder (test infrastructure)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
import mozjpeg_b200 as mj
from oracle import oracle as O


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    cpu = torch.device("cpu")
    # every rank encodes its own images (no data-path collective); only the elapsed time is reduced
    seeds = bench.rank_seeds(rank, 2)
    w, h = 48, 40
    p = mj.params_from_switches(["-baseline", "-quality", "75"], w, h)
    digests = [hashlib.md5(O.oracle_encode(p, O.synth_image(s, w, h)).jpeg).hexdigest() for s in seeds]
    elapsed = bench.max_over_ranks(10.0 + rank, world, cpu)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "seeds": seeds, "digests": digests, "elapsed": elapsed})
    if rank == 0:
        print(json.dumps({"world": world, "ranks": gathered, "mp_per_step": bench.whole_job_mp_per_step(world, 2, w, h)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
