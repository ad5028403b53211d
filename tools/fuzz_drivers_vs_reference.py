"""Differential run of the whole-sample drivers against the UNMODIFIED reference (build container only):
random multi-gene samples through selectKirTypingModel("em" | "full" | "exonfirst_1" | "exonfirst") of
both implementations (ours on the NumPy test double of the kernels); (alleles, warning genes) or the
exception type compared.  Differences with a tie report, or the same called set in another order,
are skipped.  The EM driver of the reference orders exactly tied abundances by set iteration order,
i.e. by PYTHONHASHSEED: run with PYTHONHASHSEED=1 for a stable comparison.

    PYTHONHASHSEED=1 python tools/fuzz_drivers_vs_reference.py <seed> <seconds>
"""
import sys, os, time, copy, logging, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, ROOT)
from make_golden import import_reference, ref_objects
tma, tem, kt, h2, m2h = import_reference()
logging.getLogger("graphkir").setLevel(logging.ERROR)
from kir_graph_b200 import synthetic, kir_typing
from kir_graph_b200.hisat2 import writeReadsAndVariantsData
from tests.fake_backend import FakeBackend
rng = np.random.default_rng(int(sys.argv[1])); T = float(sys.argv[2])
t0 = time.time(); n = 0; bad = 0; em_ties = 0
tmp = tempfile.mkdtemp()
while time.time() - t0 < T:
    seed = int(rng.integers(1 << 30))
    ng = int(rng.integers(1, 4))
    genes = [synthetic.make_gene([seed, i], f"KIRE{i}*BACKBONE", int(rng.integers(2, 25)), 64, int(rng.integers(1, 5)), int(rng.integers(3, 200)), hierarchical=bool(rng.integers(2)), variant_id_base=1000 * i) for i in range(ng)]
    reads, variants = [], []
    for g in genes:
        rd, va = g.to_objects(); reads += rd; variants += va
    if rng.random() < 0.3 and reads: reads[0].multiple = 2
    path = os.path.join(tmp, "s.json")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, path)
    cn = {g.gene: int(rng.integers(0, 5)) for g in genes}
    method = str(rng.choice(["em", "full", "exonfirst_1", "exonfirst"]))
    kw = {} if method == "em" else dict(top_n=int(rng.choice([10, 40])))
    def outcome(fn):
        try: return fn()
        except Exception as e: return "EXC:" + type(e).__name__
    want = outcome(lambda: kt.selectKirTypingModel(method, path, **kw).typing(cn))
    holder = {}
    def ours():
        holder["t"] = kir_typing.selectKirTypingModel(method, path, _backend=FakeBackend(), **kw); return holder["t"].typing(cn)
    got = outcome(ours)
    n += 1
    if got != want:
        same_sets = (not isinstance(got, str) and not isinstance(want, str) and sorted(got[0]) == sorted(want[0]) and got[1] == want[1])
        tie = getattr(holder.get("t"), "tie_report", None)
        if method == "em" and same_sets: continue
        if tie or same_sets: continue
        if method == "em" and not isinstance(got, str) and not isinstance(want, str) and got[1] == want[1]:
            # an exact tie of two abundances at the last copy of a gene: the reference orders those by a
            # set's iteration order, the mirror by name
            probs = {x.allele: x.prob for g in holder["t"]._result.values() for x in g}
            only_a = [a for a in got[0] if a not in want[0]]
            only_b = [a for a in want[0] if a not in got[0]]
            if len(only_a) == len(only_b) and all(abs(probs.get(x, -1) - probs.get(y, -2)) <= 1e-12 * max(probs.get(x, 1), 1e-300)
                                                 for x, y in zip(sorted(only_a), sorted(only_b))):
                em_ties += 1
                continue
        print("MISMATCH", method, seed, cn, kw, got, want); bad += 1
        if os.environ.get("FUZZ_KEEP"):
            import shutil
            shutil.copy(path, os.environ["FUZZ_KEEP"]); print("kept", os.environ["FUZZ_KEEP"]); sys.exit(1)
print("cases", n, "bad", bad, "exact EM ties", em_ties)
