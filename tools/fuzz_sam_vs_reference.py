"""Differential run of the native extraction loop (kir_graph_b200.fastsam, gk_sam_extract) against the
UNMODIFIED reference (build container only): random two-gene name-sorted SAM texts, some with records
damaged at random, through readPair -> filterRead -> extractVariant of the imported reference (only
readBam's samtools subprocess is replaced) and through fastsam.extract; the .json content (variants
and reads, every field) or the exception type must agree.

    python tools/fuzz_sam_vs_reference.py <seed> <seconds>
"""
import logging
import os
import sys
import time
from dataclasses import asdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
from make_golden import import_reference  # noqa: E402

_, _, _, h2, m2h = import_reference()
logging.getLogger("graphkir").setLevel(logging.ERROR)
from kir_graph_b200 import fastsam  # noqa: E402
from kir_graph_b200.msa2hisat import Variant  # noqa: E402
from tests import sam_sim  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60
ALPHABET = "0123456789MIDSNHX=^ACGT|,*Zs:\t"


def outcome(fn):
    try:
        data = fn()
        return [asdict(v) for v in data["variants"]], [asdict(r) for r in data["reads"]]
    except (NotImplementedError, AssertionError, IndexError, ValueError, KeyError) as exc:
        return type(exc).__name__


t0 = time.time()
n = bad = n_exc = 0
while time.time() - t0 < seconds:
    seed = int(rng.integers(1 << 30))
    nm = int(rng.choice([4, 9, 1000]))
    table, pairs = sam_sim.multi_gene(seed, n_pairs=int(rng.integers(5, 40)), novel=float(rng.choice([0.0, 0.002, 0.02])))
    lines = sam_sim.sam_text(pairs, header=bool(rng.integers(2))).rstrip("\n").split("\n")
    for _ in range(int(rng.integers(0, 3))):                # damage a few fields
        j = int(rng.integers(len(lines)))
        cols = lines[j].split("\t")
        if len(cols) < 12:
            continue
        c = int(rng.choice([1, 3, 5, 7, 9] + list(range(11, len(cols)))))
        s = cols[c]
        if not s:
            continue
        i = int(rng.integers(len(s)))
        ch = ALPHABET[int(rng.integers(len(ALPHABET)))]
        kind = int(rng.integers(3))
        cols[c] = s[:i] + ch + s[i + 1:] if kind == 0 else s[:i] + s[i + 1:] if kind == 1 else s[:i] + ch + s[i:]
        lines[j] = "\t".join(cols)
    text = "\n".join(lines) + "\n"
    rtable = [m2h.Variant(**asdict(v)) for v in table]
    h2.readBam = lambda _name, _text=text: iter(_text.split("\n"))

    def reference():
        m2h.Variant.novel_id = 0
        pr = h2.readPair("unused.bam")
        pr = filter(lambda lr: h2.filterRead(lr[0], nm) and h2.filterRead(lr[1], nm), pr)
        return h2.extractVariant(pr, rtable, pileup=None)

    def native():
        Variant.novel_id = 0
        return fastsam.extract(text, table, nm).reads_data()

    want, got = outcome(reference), outcome(native)
    n += 1
    n_exc += isinstance(want, str)
    if want != got or (not isinstance(want, str) and m2h.Variant.novel_id != Variant.novel_id):
        bad += 1
        print("MISMATCH seed", seed, "nm", nm, want if isinstance(want, str) else "data", got if isinstance(got, str) else "data")
        if os.environ.get("FUZZ_DEBUG"):
            import traceback
            for fn in (reference, native):
                try:
                    fn()
                except Exception:
                    traceback.print_exc(limit=-3)
            with open("/tmp/fuzz_sam_case.sam", "w") as f:
                f.write(text)
        if bad > (0 if os.environ.get("FUZZ_DEBUG") else 5):
            break
print("cases", n, "exceptions", n_exc, "bad", bad)
sys.exit(1 if bad else 0)
