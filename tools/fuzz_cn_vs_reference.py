"""Differential run of the CN callers against the UNMODIFIED reference (imported from /root/reference with the
stubs of tests/golden/make_golden.py - build container only): random cohorts of gene depths through
``depthToCN`` of both implementations (ours on the NumPy test double of gk_cn_fit) - CN per gene per sample, the
fitted base and the bin count must be equal, the likelihood curve within 1e-10 relative.

    python tools/fuzz_cn_vs_reference.py <seed> <seconds>
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

import_reference()
import graphkir.kir_cn as ref_cn  # noqa: E402
from make_golden_cn import GENES, cohort  # noqa: E402

from kir_graph_b200 import kir_cn  # noqa: E402
from tests.fake_backend import FakeBackend  # noqa: E402

logging.getLogger("graphkir").setLevel(logging.ERROR)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60
t0 = time.time()
n = bad = errors = 0
while time.time() - t0 < seconds:
    n_samples = int(rng.choice([1, 1, 2, 5, 12, 40]))
    depths = cohort(rng, n_samples, per_copy=float(rng.uniform(3, 60)), noise=float(rng.uniform(0.02, 0.3)),
                    dl3_cn=int(rng.choice([1, 2, 2, 2, 3])))
    kw = {}
    if rng.integers(3) == 0:
        kw["cluster_method_kwargs"] = {"base_dev": float(rng.uniform(0.04, 0.15)), "start_base": int(rng.choice([1, 2]))}
    if rng.integers(3) == 0:
        kw["assume_3DL3_diploid"] = True

    def outcome(fn, **extra):
        try:
            with np.errstate(all="ignore"):
                cns, dist = fn([dict(d) for d in depths], **kw, **extra)
            return [{k: int(v) for k, v in c.items()} for c in cns], float(dist.base), int(dist.bin_num), np.asarray(dist.likelihood, float)
        except Exception as exc:                                  # the 3DL3 loop ends in AssertionError when it cannot make 3DL3 diploid
            return "EXC:" + type(exc).__name__

    want, got = outcome(ref_cn.depthToCN), outcome(kir_cn.depthToCN, _backend=FakeBackend())
    n += 1
    if isinstance(want, str) or isinstance(got, str):
        errors += want == got
        if want != got:
            print("MISMATCH", n, want if isinstance(want, str) else "result", got if isinstance(got, str) else "result", kw)
            bad += 1
        continue
    same = want[0] == got[0] and want[1] == got[1] and want[2] == got[2] and \
        np.allclose(want[3], got[3], rtol=1e-10, atol=0, equal_nan=True)
    if not same:
        print("MISMATCH", n, n_samples, kw, want[1], got[1], want[2], got[2])
        bad += 1
print("cases", n, "same exception on both sides", errors, "bad", bad)
