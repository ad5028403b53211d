"""Which warp-split row pieces cover a remainder of g row groups (8 kept sets each) fastest?  For every g in
1..15 and two allele counts, batches of identical genes are scored at the third copy-number step with
top_n = 8 g (so the kept sets are exactly g groups) under each candidate cut of SearchGroup._W_CUT[g]; prints
useful TCells/s per candidate.  Needs a GPU.

    python tools/tune_row_cuts.py [copies=96]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import engine, packing, synthetic

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 96
be = engine.CudaBackend()
SG = engine.SearchGroup
base = dict(SG._W_CUT)


def candidates(g):
    out = {"now": base[g]}
    out["wk4"] = ((-(-g // 4), 2),)
    if g <= 8:
        out["wk2"] = ((-(-g // 2), 1),)
    if g <= 4:
        out["wk1"] = ((g, 0),)
    if g > 4 and g % 4:
        out["wk4+rest"] = ((g // 4, 2),) + base[g % 4] if g // 4 else None
    return {k: v for k, v in out.items() if v}


for a in (40, 80):
    r = 20000
    packs = [packing.pack_synthetic(synthetic.make_gene([700 + a, i], f"KIRT{i}*BACKBONE", a, 8 * a, 3, r, homo_prob=0.0))
             for i in range(copies)]
    batch = engine.MatrixBatch(packs, backend=be)
    for g in range(1, 16):
        line = []
        for name, cut in candidates(g).items():
            SG._W_CUT = dict(base); SG._W_CUT[g] = cut
            group = engine.SearchGroup(batch, list(range(copies)), 8 * g)
            best = None
            for rep in range(2):
                group.reset()
                be.timing = {}
                for step in range(3):
                    group.step(need_next=np.full(copies, step < 2), collect=np.zeros(copies, bool))
                be.sync()
                s, e, work = be.timing["gk_score"][-1]
                ms = s.elapsed_time(e)
                best = ms if best is None else min(best, ms)
            be.timing = None
            line.append(f"{name} {cut}: {best:6.3f} ms {work / best / 1e9:5.2f} T")
            del group
        print(f"A={a} g={g:2d} | " + " | ".join(line), flush=True)
    del batch
SG._W_CUT = base
