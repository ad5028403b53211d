"""Profile the host-side orchestration of BatchTyper.run with a backend whose kernels do nothing
but fill in plausible counts (no GPU needed)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.fake_backend import FakeBackend
from kir_graph_b200._cabi import STEP_INFO_DTYPE, SEARCH_DTYPE
from kir_graph_b200 import cohort
import bench

class NullBackend(FakeBackend):
    def empty(self, n, dtype):
        return np.zeros(max(int(n), 1), dtype=dtype)
    def launch(self, name, *args, work=0.0):
        self.launches += 1
        if name == "gk_first_step":
            stab, ns, top_n = args[1].view(SEARCH_DTYPE), args[2], args[3]
            info = args[10].view(STEP_INFO_DTYPE); kept = args[11]
            k = np.minimum(stab["n_cand"], top_n); info["n_kept"] = k; kept[:] = k
        elif name == "gk_select":
            stab, top_n = args[1].view(SEARCH_DTYPE), args[3]
            kept = args[7]; info = args[15].view(STEP_INFO_DTYPE)
            info["n_alive"] = np.minimum(kept * stab["n_cand"], top_n + 20)
        elif name == "gk_rank":
            info = args[17].view(STEP_INFO_DTYPE); kept = args[18]; top_n = args[3]
            k = np.minimum(info["n_alive"], top_n); info["n_kept"] = k; kept[:] = k

if __name__ != "__main__":
    raise SystemExit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
t0 = time.time()
packs, cns, truth = bench.build_cohort(list(range(100, 100 + n)), 1.0, 8)
print("build", time.time() - t0)
typer = cohort.BatchTyper(packs, cns, top_n=300, backend=NullBackend())
typer.upload(); typer.run()
t0 = time.time(); typer.run(); print("run host time", time.time() - t0)
pr = cProfile.Profile(); pr.enable(); typer.run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
