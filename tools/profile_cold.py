"""cProfile of cold cohort passes (bench.py's e2e_cold: a CohortTyper constructed per pass from page-locked
pools, device buffers from an arena): where the host time of a pass that types a cohort for the first time
goes.  Needs a GPU.

    python tools/profile_cold.py [samples=96] [passes=6] [parts=2] [profile=1]

profile=0 only times the passes (cProfile inflates them by about a fifth).
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from kir_graph_b200 import engine  # noqa: E402

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 96
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
profiling = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0
args = argparse.Namespace(samples=n_samples, scale=1.0, top_n=300, steps=passes)
packed_sets = bench.build_cold_sets(args, 0, 1, 16, n_sets=2)          # before CUDA is initialised (the pool forks)
be = engine.CudaBackend(0)
prof = cProfile.Profile()


def timed(fn, steps, finalize=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if profiling:
        prof.enable()
    for _ in range(steps):
        fn()
    if finalize is not None:
        finalize()
    torch.cuda.synchronize()
    if profiling:
        prof.disable()
    return 1e3 * (time.perf_counter() - t0)


out = bench.cold_leg(args, be, 0, 1, timed, 17, parts, packed_sets)
print({k: v for k, v in out.items() if k != "timed_region"})
if profiling:
    pstats.Stats(prof).sort_stats("cumulative").print_stats(40)
