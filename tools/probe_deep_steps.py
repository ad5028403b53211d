import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from kir_graph_b200 import engine, packing, synthetic
be = engine.CudaBackend()
pack = packing.pack_synthetic(synthetic.make_gene([900 + 1000, 0], "KIRP0*BACKBONE", 1000, 8000, 4, 131072, homo_prob=0.0))
batch = engine.MatrixBatch([pack], backend=be)
group = engine.SearchGroup(batch, [0], 300)
for rep in range(3):
    group.reset()
    be.timing = {}
    for step in range(4):
        group.step(need_next=[step + 1 < 4], collect=[False])
    be.sync()
    print(rep, " | ".join(f"{s.elapsed_time(e):7.3f} ms {w / s.elapsed_time(e) / 1e9:6.2f} T" for s, e, w in be.timing["gk_score"]), flush=True)
