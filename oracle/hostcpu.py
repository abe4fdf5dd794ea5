"""Host-CPU helper for the checker and the CPU baseline.  *** TEST INFRASTRUCTURE ONLY *** (same rule as
oracle/navillm_oracle.py: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product).

``pick_cpu_threads`` chooses the intra-op thread count that gives torch's CPU GEMMs their best throughput on this
host.  os.cpu_count() can exceed what the container may use (affinity mask, cgroup quota: the GPU boxes show 128
CPUs and grant a quota of 16) and oversubscribed intra-op threads ran the oracle 31x slower in round 1."""
from __future__ import annotations

import os
import time
from pathlib import Path

import torch

_CPU_THREADS = None


def pick_cpu_threads(d_model: int = 4096) -> int:
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        torch.set_num_threads(_CPU_THREADS)
        return _CPU_THREADS
    cands = {os.cpu_count() or 1}
    if hasattr(os, "sched_getaffinity"):
        cands.add(len(os.sched_getaffinity(0)))
    for path, parse in (("/sys/fs/cgroup/cpu.max", lambda t: None if t.split()[0] == "max" else int(t.split()[0]) / int(t.split()[1])),
                        ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", lambda t: None if int(t) <= 0 else int(t) / 100000.0)):
        try:
            q = parse(Path(path).read_text())
            if q:
                cands.add(max(1, int(q)))
        except Exception:
            pass
    top = max(cands)
    cands |= {t for t in (8, 16, 32, 64, 128) if t <= top}
    a = torch.randn(640, d_model)
    w = torch.randn(d_model, d_model)
    best, best_t = None, None
    for t in sorted(cands):
        torch.set_num_threads(t)
        a @ w
        t0 = time.perf_counter()
        for _ in range(3):
            a @ w
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    torch.set_num_threads(best)
    _CPU_THREADS = best
    return best
