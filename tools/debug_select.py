import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import engine, packing, synthetic
from tests.fake_backend import FakeBackend
A, CN, R, TOPN = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
gene = synthetic.make_gene([4, 0], "KIRDBG*BACKBONE", A, 8 * A, CN, R, homo_prob=0.0)
pack = packing.pack_synthetic(gene)
cuda, fake = engine.CudaBackend(), FakeBackend()
bg, bf = engine.MatrixBatch([pack], backend=cuda), engine.MatrixBatch([pack], backend=fake)
gg, gf = engine.SearchGroup(bg, [0], TOPN), engine.SearchGroup(bf, [0], TOPN)
for step in range(CN):
    og = gg.step(need_next=[step + 1 < CN])[0]
    of = gf.step(need_next=[step + 1 < CN])[0]
    same = np.array_equal(og.ids, of.ids) and np.array_equal(og.score, of.score)
    print("n", step + 1, "gpu", og.n_unique, og.n_alive, "fake", of.n_unique, of.n_alive, "same", same, flush=True)
    if step:
        N = int(gf.kept[0]) if False else None
        fg = cuda.download(gg.d_flag, np.uint8); ff = gf.d_flag
        K_prev_C = len(ff)
        diff = np.flatnonzero(fg[:K_prev_C] != ff[:K_prev_C])
        print("   flag diffs", len(diff), diff[:10])
