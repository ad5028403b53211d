"""Quick per-kernel timing probe (CUDA events on the launching stream)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import engine, packing, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
A = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
CN = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time()
gene = synthetic.make_gene([4, 0], "KIRDEEP*BACKBONE", A, 8 * A, CN, R, homo_prob=0.0)
pack = packing.pack_synthetic(gene)
print(f"gen+pack {time.time()-t0:.1f}s  R={pack.n_reads} A={pack.n_alleles} E={pack.n_entries}", flush=True)
be = engine.CudaBackend()
be.timing = {}
for rep in range(2):
    batch = engine.MatrixBatch([pack], backend=be)
    group = engine.SearchGroup(batch, [0], 300)
    for step in range(CN):
        out = group.step(need_next=[step + 1 < CN])[0]
be.sync()
for name, evs in be.timing.items():
    for (s, e, work) in evs[len(evs)//2:]:
        ms = s.elapsed_time(e)
        extra = ""
        if name == "gk_score":
            extra = f"{work/ms/1e9:.2f} TCells/s"
        if name == "gk_likelihood":
            extra = f"{work/ms/1e6:.1f} GCells/s  {batch.bytes_out/ms/1e6:.0f} GB/s written"
        print(f"{name:18s} {ms:9.3f} ms  {extra}")
print("kept", out.ids[:3].tolist(), "truth", sorted(gene.truth))
