"""Host extraction of one sample: name-sorted SAM text -> per-pair variant lists.
Object path (the reference's loop restated in kir_graph_b200.hisat2: pairRecords, filterRead,
extractVariant - a Variant per CIGAR segment, a PairRead per pair) against the native batch routine
gk_sam_extract (kir_graph_b200.fastsam).  CPU only.

    python tools/bench_sam_extract.py [pairs=20000]
"""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kir_graph_b200 import fastjson, fastsam, hisat2
from kir_graph_b200.msa2hisat import Variant
from tests import sam_sim

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(1)
table, pairs = [], []
genes = [f"KIR{g}*BACKBONE" for g in range(8)]
for g, name in enumerate(genes):
    seq, variants = sam_sim.make_table(rng, name, length=6000, n_single=120, n_del=12)
    for i, v in enumerate(variants):
        v.id = f"hv{10000 * g + i}"
    table += variants
    for i in range(n_pairs // len(genes)):
        s1 = int(rng.integers(0, len(seq) - 400))
        s2 = s1 + int(rng.integers(100, 200))
        pairs.append((sam_sim.simulate_record(rng, f"g{g}r{i}", 99, name, seq, variants, s1, n_ref=150, novel=0.001, soft=0.02),
                      sam_sim.simulate_record(rng, f"g{g}r{i}", 147, name, seq, variants, s2, n_ref=150, novel=0.001, soft=0.02)))
table.sort()
text = sam_sim.sam_text(pairs)
raw = text.encode()
print(f"{len(pairs)} pairs, {len(table)} index variants, {len(raw) / 1e6:.1f} MB of SAM text")

Variant.novel_id = 0
t0 = time.perf_counter()
ext = fastsam.extract(raw, table)
t1 = time.perf_counter()
packs = fastjson.packs_from_scan(ext.scan()) if all(v.allele for v in table) else None
t2 = time.perf_counter()
Variant.novel_id = 0
tab2 = copy.deepcopy(table)
t3 = time.perf_counter()
pr = hisat2.pairRecords(text.split("\n"))
pr = filter(lambda lr: hisat2.filterRead(lr[0]) and hisat2.filterRead(lr[1]), pr)
want = hisat2.extractVariant(pr, tab2)
t4 = time.perf_counter()
got = ext.reads_data()
assert [(r.lpv, r.lnv, r.rpv, r.rnv, r.multiple, r.backbone) for r in got["reads"]] == \
       [(r.lpv, r.lnv, r.rpv, r.rnv, r.multiple, r.backbone) for r in want["reads"]]
n = len(pairs)
import tempfile
with tempfile.TemporaryDirectory() as d:
    Variant.novel_id = 0
    t5 = time.perf_counter()
    fastsam.extract(raw, table, json_reads=True).write_json(os.path.join(d, "a.json"))
    t6 = time.perf_counter()
    hisat2.writeReadsAndVariantsData(want, os.path.join(d, "b.json"))
    t7 = time.perf_counter()
    assert open(os.path.join(d, "a.json"), "rb").read() == open(os.path.join(d, "b.json"), "rb").read()
    size = os.path.getsize(os.path.join(d, "a.json"))
print(f"SAM -> .json ({size / 1e6:.1f} MB, byte-identical): object path {t4 - t3 + t7 - t6:.2f} s "
      f"({1e6 * (t4 - t3 + t7 - t6) / n:.1f} us per pair), native {t6 - t5:.3f} s ({1e6 * (t6 - t5) / n:.2f} us per pair)"
      f"  -> {(t4 - t3 + t7 - t6) / (t6 - t5):.0f}x")
print(f"kept {ext.n_reads} pairs, {len(ext.novel)} novel variants")
print(f"object path : {t4 - t3:.2f} s ({1e6 * (t4 - t3) / n:.1f} us per pair)")
print(f"native path : {t1 - t0:.3f} s ({1e6 * (t1 - t0) / n:.2f} us per pair, {len(raw) / (t1 - t0) / 1e6:.0f} MB/s)"
      f"  -> {(t4 - t3) / (t1 - t0):.0f}x")
