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
