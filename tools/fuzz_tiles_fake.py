"""Random gene shapes (allele counts and top_n around the tile boundaries 8 / 32 / 64 / 128 / 256) through
SearchGroup on the NumPy fake backend against the oracle's integer search: every step's ids, scores, fraction
numerators and N_uniq.  Exercises the host-side tiling (row pieces, column tiles, work items) on CPU.

    python tools/fuzz_tiles_fake.py [seed=0] [seconds=120]
"""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kir_graph_b200 import engine, packing, synthetic
from tests.fake_backend import FakeBackend
from tests.test_engine_fake import oracle_for_pack, check_steps
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); n=0; bad=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 120:
    a=int(rng.choice([rng.integers(2,40), rng.integers(40,140), rng.integers(118,135), rng.integers(250,262), rng.integers(2,300)]))
    cn=int(rng.integers(2,4))
    top_n=int(rng.choice([rng.integers(5,64), rng.integers(100,140), rng.integers(240,260), rng.integers(5,400)]))
    half=bool(rng.integers(0,2))
    gene=synthetic.make_gene([int(rng.integers(1<<30)),0],"KIRF*BACKBONE",a,max(64,4*a),cn,int(rng.integers(20,90)))
    reads,variants=gene.to_objects()
    try:
        pack=packing.pack_gene(reads,variants,variant_correction=True)
    except Exception as e:
        continue
    if pack.n_alleles<2 or pack.n_reads<1: continue
    m,k,search=oracle_for_pack(pack,reads,variants,top_n)
    be=FakeBackend()
    batch=engine.MatrixBatch([pack],backend=be,packed=half)
    group=engine.SearchGroup(batch,[0],top_n)
    try:
        for step in range(cn):
            out=group.step(need_next=[step+1<cn])[0]
            ref=search.add_candidate()
            check_steps(out,ref)
    except AssertionError as e:
        bad+=1; print("MISMATCH a",pack.n_alleles,"cn",cn,"top_n",top_n,"half",half,"step",step, "kept", len(ref.score), flush=True)
    n+=1
print("cases",n,"bad",bad)
