"""Differential run of the record walk (gk_sam_walk behind hisat2.recordToRawVariant) against the
UNMODIFIED reference's recordToRawVariant (build container only): simulated records with up to three
random edits in CIGAR / MD / Zs / sequence / position; segments (typ, pos, length, val, id, ref) and
soft clips, or the exception type, must agree.

    python tools/fuzz_walk_vs_reference.py <seed> <seconds>
"""
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
from make_golden import import_reference  # noqa: E402

_, _, _, h2, m2h = import_reference()
logging.getLogger("graphkir").setLevel(logging.ERROR)
from kir_graph_b200 import hisat2  # noqa: E402
from tests import sam_sim  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60
ALPHABET = "0123456789MIDSNHX=^ACGT|,*Zs:"


def outcome(fn, rec):
    try:
        raw, clip = fn(rec)
        return [[v.typ, v.pos, v.length, v.val, v.id, v.ref] for v in raw], list(clip)
    except (NotImplementedError, AssertionError, IndexError, ValueError, KeyError) as exc:
        return type(exc).__name__


t0 = time.time()
n = bad = n_exc = 0
while time.time() - t0 < seconds:
    _, _, pairs = sam_sim.simulate_pairs(int(rng.integers(1 << 30)), n_pairs=10)
    for rec in (r for pair in pairs for r in pair):
        cols = rec.split("\t")
        for _ in range(int(rng.integers(0, 4))):
            c = int(rng.choice([3, 5, 9] + list(range(11, len(cols)))))
            s = cols[c]
            if not s:
                continue
            i = int(rng.integers(len(s)))
            ch = ALPHABET[int(rng.integers(len(ALPHABET)))]
            kind = int(rng.integers(3))
            cols[c] = s[:i] + ch + s[i + 1:] if kind == 0 else s[:i] + s[i + 1:] if kind == 1 else s[:i] + ch + s[i:]
        bad_rec = "\t".join(cols)
        want, got = outcome(h2.recordToRawVariant, bad_rec), outcome(hisat2.recordToRawVariant, bad_rec)
        n += 1
        n_exc += isinstance(want, str)
        if want != got:
            bad += 1
            print("MISMATCH", want if isinstance(want, str) else "data", got if isinstance(got, str) else "data", repr(bad_rec))
            if bad > 5:
                break
    if bad > 5:
        break
print("records", n, "exceptions", n_exc, "bad", bad)
sys.exit(1 if bad else 0)
