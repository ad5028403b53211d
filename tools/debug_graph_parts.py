import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from kir_graph_b200 import cohort, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
packs, cns, truth = bench.build_cohort(list(range(100, 100 + n)), 1.0, 16)
be = engine.CudaBackend(0)
for parts in (1, 3):
    for graph in (True, False):
        ct = cohort.CohortTyper(packs, cns, top_n=300, backend=be, n_parts=parts, group_size=17)
        for p in ct.parts: p.use_graph = graph
        ct.pin(); ct.upload()
        ts = []
        for i in range(12):
            torch.cuda.synchronize(); t0 = time.perf_counter(); ct.run(); torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        print(f"parts={parts} graph={graph}: " + " ".join(f"{t:.1f}" for t in ts),
              [getattr(p, "graph_error", None) for p in ct.parts], [p._graph is not None for p in ct.parts])
