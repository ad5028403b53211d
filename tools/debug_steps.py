import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kir_graph_b200 import engine, packing, synthetic
R, A, CN = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gene = synthetic.make_gene([4, 0], "KIRDEEP*BACKBONE", A, 8 * A, CN, R, homo_prob=0.0)
pack = packing.pack_synthetic(gene)
be = engine.CudaBackend()
orig = be.launch
def launch(name, *args, work=0.0):
    orig(name, *args, work=work)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED after", name, "n =", getattr(launch, "n", None), str(e)[:80]); sys.exit(1)
    print("ok", name, flush=True)
be.launch = launch
batch = engine.MatrixBatch([pack], backend=be)
group = engine.SearchGroup(batch, [0], 300)
for step in range(CN):
    launch.n = step + 1
    out = group.step(need_next=[step + 1 < CN])[0]
    print("step", step + 1, len(out.score), out.n_unique, out.n_alive, flush=True)
