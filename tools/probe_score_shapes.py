"""Scoring throughput per tile shape: batches of identical small genes (the shapes of a cfg3 sample), CUDA
events around gk_score, useful TCells/s per copy-number step.  Needs a GPU.

    python tools/probe_score_shapes.py [copies=96]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import engine, packing, synthetic

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 96
be = engine.CudaBackend()
SPECS = [(160, 2, 13333), (130, 2, 13333), (90, 2, 13333), (80, 3, 20000), (45, 2, 13333), (40, 2, 13333),
         (35, 4, 26667), (25, 2, 13333), (8, 2, 13333), (1000, 3, 131072)]
for a, cn, r in SPECS:
    n = copies if a < 1000 else 1
    packs = [packing.pack_synthetic(synthetic.make_gene([900 + a, i], f"KIRP{i}*BACKBONE", a, max(64, 8 * a), cn, r,
                                                        homo_prob=0.0)) for i in range(n)]
    batch = engine.MatrixBatch(packs, backend=be)
    group = engine.SearchGroup(batch, list(range(n)), 300)
    for rep in range(2):
        group.reset()
        be.timing = {}
        for step in range(cn):
            group.step(need_next=np.full(n, step + 1 < cn), collect=np.zeros(n, bool))
        be.sync()
    line = []
    for (s, e, work) in be.timing.get("gk_score", []):
        ms = s.elapsed_time(e)
        line.append(f"{ms:7.3f} ms {work / ms / 1e9:6.2f} TCells/s")
    be.timing = None
    print(f"A={a:5d} CN={cn} R={r:6d} x{n:3d}: " + " | ".join(line), flush=True)
    del group, batch
