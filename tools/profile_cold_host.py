"""Host time of a cold cohort pass WITHOUT a GPU: the per-pass work of bench.py's ``e2e_cold`` leg (typer objects,
homozygosity decisions, work-item tables, staging of the small uploads, the call phase) on a backend whose
launches and large copies do nothing.  What it measures is the floor the host thread sets for a cohort stream on
THIS machine's CPU; the device results are all zero, so every call comes back as ``fail`` (the call phase still walks
every problem).  CPU only:

    python tools/profile_cold_host.py [samples=96] [passes=6] [profile=1]
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from kir_graph_b200 import cohort  # noqa: E402
from tests.fake_backend import FakeBackend  # noqa: E402


class NullBackend(FakeBackend):
    """Buffers are untouched virtual memory, kernels do not run, copies of more than 1 MB (the page-locked pools,
    DMA on the real backend) cost nothing; small uploads are copied once, like the staging arena does."""

    def zeros(self, n, dtype):
        return np.zeros(max(int(n), 1), dtype=dtype)           # calloc: no page is touched

    def empty(self, n, dtype):
        return np.empty(max(int(n), 1), dtype=dtype)

    def upload(self, array):
        array = np.ascontiguousarray(array)
        self.h2d_bytes += array.nbytes
        return array.reshape(-1).copy() if array.nbytes < (1 << 20) else array.reshape(-1)

    def copy_into(self, tensor, array):
        pass

    def zero_(self, tensor):
        pass

    def launch(self, name, *args, work=0.0):
        self.launches += 1

    def download_async(self, tensors, sizes=None):
        return [np.zeros(t.size if sizes is None else s, dtype=np.int32).reshape(-1)
                for t, s in zip(tensors, sizes or [0] * len(tensors))]


def main():
    n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    profiling = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
    import argparse
    args = argparse.Namespace(samples=n_samples, scale=1.0, top_n=300, steps=passes)
    sets = bench.build_cold_sets(args, 0, 1, 8, n_sets=2)
    be = NullBackend()
    hosts = []
    for packs, cns, _ in sets:
        probe = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=1, group_size=17)
        hosts.append([p.host for p in probe.parts])

    def one_pass(i):
        packs, cns, _ = sets[i % len(sets)]
        typer = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=1, group_size=17,
                                   host_batches=hosts[i % len(sets)])
        token = typer.start_pass(upload=True)
        return typer.finish_pass(token)

    for i in range(2):
        one_pass(i)
    if os.environ.get("GK_GC_FREEZE", "1") != "0":      # as bench.py: the packed cohorts leave the collector's view
        import gc
        gc.collect()
        gc.freeze()
    prof = cProfile.Profile()
    l0 = be.launches
    t0 = time.perf_counter()
    if profiling:
        prof.enable()
    each = []
    for i in range(passes):
        t1 = time.perf_counter()
        calls = one_pass(i)
        each.append(1e3 * (time.perf_counter() - t1))
    if profiling:
        prof.disable()
    ms = 1e3 * (time.perf_counter() - t0) / passes
    each.sort()
    print(f"host time per cold pass of {n_samples} samples: mean {ms:.1f} ms, median {each[len(each) // 2]:.1f}, "
          f"min {each[0]:.1f} ({(be.launches - l0) / passes:.0f} launches, {len(calls)} calls; "
          f"cProfile {'on' if profiling else 'off'})")
    if profiling:
        pstats.Stats(prof).sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
