"""Differential run of the batched cohort path against the UNMODIFIED reference (build container
only): random multi-gene samples, .json -> fastjson packs -> cohort.BatchTyper (NumPy test double of
the kernels) against selectKirTypingModel("full", top_n, variant_correction=True).typing(cn) of the
imported reference, gene by gene.  A difference counts as explained when the gene's GeneCall carries
tie flags, the called multiset is the same, or the object-API mirror reports a tie for that gene
(bit3: a fraction near the selectBest threshold, which the batched path does not read back).

    python tools/fuzz_cohort_vs_reference.py <seed> <seconds>
"""
import logging
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
from make_golden import import_reference  # noqa: E402

_, _, kt, _, _ = import_reference()
logging.getLogger("graphkir").setLevel(logging.ERROR)
from kir_graph_b200 import cohort, fastjson, kir_typing, synthetic  # noqa: E402
from kir_graph_b200.hisat2 import writeReadsAndVariantsData  # noqa: E402
from tests.fake_backend import FakeBackend  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60
tmp = tempfile.mkdtemp()
t0 = time.time()
n_genes = bad = explained = 0
while time.time() - t0 < seconds:
    seed = int(rng.integers(1 << 30))
    names = ["KIR2DL1S1", "KIR2DL5", "KIR3DL2", "KIR2DS4"]          # the first two are always typed as heterozygous
    genes = [synthetic.make_gene([seed, i], f"{names[i]}*BACKBONE", int(rng.integers(2, 25)), 64, int(rng.integers(1, 5)),
                                 int(rng.integers(20, 220)), hierarchical=bool(rng.integers(2)), variant_id_base=1000 * i)
             for i in range(int(rng.integers(1, 5)))]
    reads, variants = [], []
    for g in genes:
        rd, va = g.to_objects()
        reads += rd
        variants += va
    if reads and rng.random() < 0.3:
        reads[0].multiple = 2
    path = os.path.join(tmp, "s.json")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, path)
    cn = {g.gene: int(rng.integers(1, 5)) for g in genes}
    top_n = int(rng.choice([10, 40, 300]))
    ref = kt.selectKirTypingModel("full", path, top_n=top_n, variant_correction=True)
    packs = fastjson.load_packs(path, variant_correction=True)
    order = [g for g in cn if g in packs]
    calls = cohort.BatchTyper([packs[g] for g in order], [cn[g] for g in order], top_n=top_n, backend=FakeBackend()).run()
    mirror = None
    for gene, call in zip(order, calls):
        n_genes += 1
        try:
            want, want_reads = ref.typingPerGene(gene, cn[gene])
        except Exception as exc:                      # e.g. no usable reads and cn >= 2: AxisError in the reference
            if call.alleles == ["fail"] * cn[gene]:
                explained += 1
            else:
                bad += 1
                print("REF-EXC", type(exc).__name__, seed, gene, call.alleles)
            continue
        got = [a if a != "fail" else gene.split("*")[0] + "*" for a in call.alleles]
        if got == want and call.n_reads == want_reads:
            continue
        if call.n_reads == want_reads and (call.tie_flags or sorted(got) == sorted(want)):
            explained += 1
            continue
        if mirror is None:
            mirror = kir_typing.selectKirTypingModel("full", path, top_n=top_n, variant_correction=True, _backend=FakeBackend())
        mirror.typingPerGene(gene, cn[gene])
        if call.n_reads == want_reads and mirror.tie_report.get(gene):
            explained += 1
            continue
        bad += 1
        print("MISMATCH", seed, gene, cn[gene], top_n, got, want, call.n_reads, want_reads)
        if os.environ.get("FUZZ_KEEP"):
            import shutil
            shutil.copy(path, os.environ["FUZZ_KEEP"])
            print("kept", os.environ["FUZZ_KEEP"], "cn", cn, "top_n", top_n)
            sys.exit(1)
print("genes", n_genes, "bad", bad, "explained by ties", explained)
sys.exit(1 if bad else 0)
