"""cProfile of BatchTyper.run on the real CUDA backend (host-side cost breakdown)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
packs, cns, truth = bench.build_cohort(list(range(100, 100 + n)), 1.0, 16)
import torch
from kir_graph_b200 import cohort, engine
be = engine.CudaBackend(0)
typer = cohort.BatchTyper(packs, cns, top_n=300, backend=be)
typer.upload()
for _ in range(3): typer.run()
torch.cuda.synchronize()
t0 = time.perf_counter(); typer.run(); torch.cuda.synchronize(); print("run wall ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); typer.run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
be.timing = {}
typer.run(); torch.cuda.synchronize()
tot = 0.0
for name, evs in be.timing.items():
    ms = sum(s.elapsed_time(e) for s, e, _ in evs)
    tot += ms
    print(f"  {name:20s} {len(evs):3d} launches {ms:8.3f} ms")
print("kernel sum ms", tot)
be.timing = None
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): typer.run()
torch.cuda.synchronize()
print("20 runs: wall ms per run", 1e3 * (time.perf_counter() - t0) / 20)
ct = cohort.CohortTyper(packs, cns, top_n=300, backend=be, n_parts=1, group_size=17)
ct.pin(); ct.upload()
for _ in range(3): ct.run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ct.run()
torch.cuda.synchronize()
print("CohortTyper 20 runs: wall ms per run", 1e3 * (time.perf_counter() - t0) / 20)
import gc
gc.collect(); gc.disable()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): typer.run()
torch.cuda.synchronize()
print("gc disabled, 20 runs: wall ms per run", 1e3 * (time.perf_counter() - t0) / 20)
gc.enable()
for np_ in (2, 3):
    ct = cohort.CohortTyper(packs, cns, top_n=300, backend=be, n_parts=np_, group_size=17)
    ct.pin(); ct.upload()
    for _ in range(3): ct.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): ct.run()
    torch.cuda.synchronize()
    print(f"CohortTyper parts={np_} 20 runs: wall ms per run", 1e3 * (time.perf_counter() - t0) / 20)
be.timing = {}
typer.run(); torch.cuda.synchronize()
for s, e, work in be.timing.get("gk_score", []):
    ms = s.elapsed_time(e)
    print(f"gk_score launch: {ms:.3f} ms, {work/1e9:.1f} G cells, {work/ms/1e9:.2f} TCells/s")
be.timing = None
g = typer.group
print("searches", g.n_search, "A mean", g.A.mean(), "A min/max", g.A.min(), g.A.max(), "R mean", g.R.mean())
import numpy as np
steps = np.where(typer.homo[typer.live], 1, typer.cns[typer.live])
print("steps histogram", np.bincount(steps))
