"""A3R_TIMING=1: wall-clock of the phases of a host-side call on stderr (diagnostic; off by default)."""
import os
import sys
import time

import torch


class Phases:
    def __init__(self, name, device=None, sync=False):
        self.on = bool(os.environ.get("A3R_TIMING"))
        self.name, self.dev, self.sync, self.rows = name, device, sync, []
        if self.on:
            if self.sync:
                torch.cuda.synchronize(self.dev)
            self.t = time.perf_counter()

    def mark(self, label):
        if self.on:
            if self.sync:
                torch.cuda.synchronize(self.dev)
            now = time.perf_counter()
            self.rows.append(f"{label} {1e3 * (now - self.t):.1f}")
            self.t = now

    def done(self):
        if self.on:
            print(f"[a3r {self.name} ms] " + "  ".join(self.rows), file=sys.stderr, flush=True)


def host_workers(n_jobs, shard=None, total_jobs=None):
    """Size of a host thread pool (RANSAC-PnP solves): the box's cores are shared by the ranks of a frame-window sharded
    job - every rank starting cpu_count() workers oversubscribes the host 8x at 8 GPUs.  total_jobs: the jobs of ALL ranks
    when every rank knows the number (the frames the spanning tree leaves without a pose): a rank then takes the share of the
    cores its own jobs are of the total, instead of cpu_count // world whatever the others have to do (three solves on a rank
    with two threads are two rounds of 45 ms).  A3R_HOST_THREADS overrides."""
    import os
    env = os.environ.get("A3R_HOST_THREADS")
    if env:
        return max(1, min(int(n_jobs), int(env)))
    world = int(shard['world']) if shard is not None else 1
    cores = os.cpu_count() or 1
    share = cores // max(world, 1)
    if total_jobs:
        share = max(share, cores * int(n_jobs) // max(int(total_jobs), int(n_jobs), 1))
    return max(1, min(int(n_jobs), share))
