"""Where do the cycles of the packed scoring kernel go?  Needs a probe build:
    GK_NVCC_EXTRA=-DGK_SCORE_PROBE python -m kir_graph_b200.build --force   (then rebuild without it)
Prints, per stage, the cycles thread 0 spends waiting for a free stage and issuing the bulk
copies, and the cycles a consumer warp waits for a full stage / computes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import _cabi, engine, packing, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
A = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
TOP = int(sys.argv[3]) if len(sys.argv) > 3 else 300
gene = synthetic.make_gene([4, 0], "KIRDEEP*BACKBONE", A, 8 * A, 2, R, homo_prob=0.0)
pack = packing.pack_synthetic(gene)
be = engine.CudaBackend()
lib = _cabi.load()
out = (ctypes.c_ulonglong * 8)()
batch = engine.MatrixBatch([pack], backend=be, packed=True)
group = engine.SearchGroup(batch, [0], TOP)
be.timing = {}
group.step(need_next=[True])
be.sync()
assert lib.gk_score_probe(out) == 0
group.step(need_next=[False])
be.sync()
assert lib.gk_score_probe(out) == 0
for (s, e, work) in be.timing.get("gk_score", [])[-1:]:
    ms = s.elapsed_time(e)
    print(f"gk_score {ms:.3f} ms  {work / ms / 1e9:.2f} TCells/s")
v = np.array(list(out), dtype=np.float64)
st, ctas = v[5], v[6]
print(f"CTAs {ctas:.0f} stages {st:.0f}")
print(f"per stage: thread0 wait-free {v[0]/st:.0f}  issue {v[1]/st:.0f}  thread0 wait-full {v[2]/st:.0f} | "
      f"warp1 wait-full {v[3]/st:.0f}  compute {v[4]/st:.0f} cycles")
