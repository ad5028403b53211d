"""Differential run of AlleleTyping's intermediates against the UNMODIFIED reference (build container
only): random small genes (some reads emptied, variant correction and no_empty on or off); the reads
after correction, the allele order, log_probs, every step's sorted values, selectAllPossible and the
exception type must agree (ours on the NumPy test double of the kernels).

    python tools/fuzz_steps_vs_reference.py <seed> <seconds>
"""
import sys, os, time, copy, logging
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, ROOT)
from make_golden import import_reference, ref_objects
tma, tem, kt, h2, m2h = import_reference()
logging.getLogger("graphkir").setLevel(logging.ERROR)
from kir_graph_b200 import synthetic
from kir_graph_b200.typing_mulit_allele import AlleleTyping
from tests.fake_backend import FakeBackend
rng = np.random.default_rng(int(sys.argv[1])); T = float(sys.argv[2])
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < T:
    a = int(rng.integers(2, 30)); cn = int(rng.integers(1, 5)); r = int(rng.integers(3, 250)); top_n = int(rng.choice([5, 20, 60]))
    seed = int(rng.integers(1 << 30)); hier = bool(rng.integers(2)); vc = bool(rng.integers(2)); ne = bool(rng.integers(2))
    gene = synthetic.make_gene([seed, 0], "KIRQ*BACKBONE", a, max(64, 8 * a), cn, r, hierarchical=hier)
    reads, variants = gene.to_objects()
    if rng.random() < 0.3:
        for rd in reads[: int(rng.integers(1, 4))]: rd.lpv = rd.lnv = rd.rpv = rd.rnv = []; rd.lpv = []; rd.lnv = []; rd.rpv = []; rd.rnv = []
    rr, rv = ref_objects(h2, m2h, copy.deepcopy(reads), copy.deepcopy(variants))
    ref = tma.AlleleTyping(rr, rv, force_homo=None, top_n=top_n, variant_correction=vc, no_empty=ne)
    ours = AlleleTyping(reads, variants, force_homo=None, top_n=top_n, variant_correction=vc, no_empty=ne, _backend=FakeBackend())
    n += 1
    ok = True
    if [(x.lpv, x.lnv, x.rpv, x.rnv) for x in ref.reads] != [(x.lpv, x.lnv, x.rpv, x.rnv) for x in ours.reads]: ok = False; why = "reads"
    elif ref.id_to_allele != ours.id_to_allele: ok = False; why = "alleles"
    elif np.shape(ref.probs) != np.shape(ours.probs) or (np.size(ref.probs) and not np.allclose(ref.log_probs, ours.log_probs, rtol=1e-12, atol=1e-12)): ok = False; why = "log_probs"
    elif ref.getReadsNum() != ours.getReadsNum(): ok = False; why = "nreads"
    else:
        try:
            rres = ref.typing(cn)
        except Exception as e:
            
            try:
                ours.typing(cn); o = 'no exception'
            except Exception as e2:
                o = type(e2).__name__
            if o != type(e).__name__: print('REF-EXC', type(e).__name__, 'ours', o, seed); bad += 1
            continue
        ores = ours.typing(cn)
        if len(ref.result) != len(ours.result): ok = False; why = "steps"
        else:
            tied = False
            for x, y in zip(ref.result, ours.result):
                if tied:
                    break                 # a flagged tie at a cut: the kept sets may differ from here on
                tied = bool(getattr(y, 'tie_flags', 0))
                if x.n != y.n or len(x.value) != len(np.asarray(y.value)): ok = False; why = f"shape n={x.n} {len(x.value)} {len(np.asarray(y.value))}"; break
                if len(x.value) and not np.allclose(np.sort(x.value)[::-1], np.asarray(y.value), rtol=1e-11): ok = False; why = "values"; break
            if ok and not tied:
                pa = rres.selectAllPossible(.9) if len(rres.value) else []; pb = ores.selectAllPossible(.9) if len(np.asarray(ores.value)) else []
                if len(pa) != len(pb) or not np.allclose([v for v, _ in pa], [v for v, _ in pb], rtol=1e-11): ok = False; why = f"possible {len(pa)} {len(pb)}"
    if not ok:
        bad += 1; print("MISMATCH", why, seed, a, cn, r, top_n, hier, vc, ne)
print("cases", n, "bad", bad)
