"""Host preparation of one sample: .variant.json -> packed gene problems.
Object path (what the reference reader does: json.load, dataclasses, per-object packing) against the
C++ scanner + array-level packing (kir_graph_b200.fastjson).  CPU only."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import fastjson, packing, synthetic
from kir_graph_b200.hisat2 import loadReadsAndVariantsData, removeMultipleMapped, writeReadsAndVariantsData
from kir_graph_b200.kir_typing import groupReads, groupVariants

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
sam = "r%d\t99\tKIR*BACKBONE\t1234\t60\t150M\t=\t1500\t416\t" + "ACGT" * 37 + "AC\t" + "F" * 150 + \
      "\tNM:i:1\tMD:Z:75A74\tZs:Z:75|S|hv12\tNH:i:1"
genes = synthetic.make_wgs30x_sample(seed=3, scale=scale)
reads, variants = [], []
for g in genes:
    r, v = g.to_objects()
    reads += r
    variants += v
for i, r in enumerate(reads):                       # realistic record size: ~0.7 KB of SAM text per pair
    r.l_sam = sam % i
    r.r_sam = sam % i
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "s.variant.json")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, path)
    size = os.path.getsize(path)
    print(f"{len(reads)} read pairs, {len(variants)} variants, {size / 1e6:.1f} MB")
    t0 = time.perf_counter()
    data = removeMultipleMapped(loadReadsAndVariantsData(path))
    t1 = time.perf_counter()
    rg, vg = groupReads(data["reads"]), groupVariants(data["variants"])
    slow = {g: packing.pack_gene(rg.get(g, []), v, mutate_reads=False, gene=g) for g, v in vg.items()}
    t2 = time.perf_counter()
    fast = fastjson.load_packs(path)
    t3 = time.perf_counter()
    sc = fastjson.scan(path)
    t4 = time.perf_counter()
for g in slow:
    assert np.array_equal(slow[g].ent_pos, fast[g].ent_pos) and np.array_equal(slow[g].k_obs, fast[g].k_obs)
n = len(reads)
print(f"object path : load {t1 - t0:.2f} s + group/pack {t2 - t1:.2f} s = {t2 - t0:.2f} s ({1e6 * (t2 - t0) / n:.1f} us per pair)")
print(f"fast path   : {t3 - t2:.2f} s ({1e6 * (t3 - t2) / n:.2f} us per pair; scan alone {t4 - t3:.2f} s = {size / (t4 - t3) / 1e6:.0f} MB/s)"
      f"  -> {(t2 - t0) / (t3 - t2):.1f}x")
