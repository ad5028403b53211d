#!/usr/bin/env python
"""
Golden fixtures of the CN model from the UNMODIFIED reference:
``graphkir.cn_model.CNgroup`` (cn_model.py:55-204) and ``graphkir.kir_cn.depthToCN`` (kir_cn.py:41-123).

Build container only (needs /root/reference):    python tests/golden/make_golden_cn.py
Writes tests/golden/cn_model.json.gz: per case the gene depths per sample, the arguments, and what the
reference computed (CN per gene per sample, base, x_max, base_dev, bin_num, the likelihood curve, the
CN-group probabilities at the fitted base).
"""
from __future__ import annotations

import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402  (installs the plotly / Bio / pyhlamsa stubs)

import_reference()
import graphkir.cn_model as cm  # noqa: E402
import graphkir.kir_cn as kc  # noqa: E402

GENES = ["KIR2DL1", "KIR2DL2", "KIR2DL3", "KIR2DL4", "KIR2DL5", "KIR2DP1", "KIR2DS1", "KIR2DS2", "KIR2DS3", "KIR2DS4",
         "KIR2DS5", "KIR3DL1", "KIR3DL2", "KIR3DL3", "KIR3DP1", "KIR3DS1"]


def cohort(rng, n_samples, per_copy=15.0, noise=0.08, dl3_cn=2):
    out = []
    for _ in range(n_samples):
        depth = per_copy * rng.uniform(0.9, 1.1)
        sample = {}
        for g in GENES:
            cn = dl3_cn if g == "KIR3DL3" else int(rng.choice([0, 1, 1, 2, 2, 2, 3]))
            sample[g + "*BACKBONE"] = float(max(0.0, cn * depth * (1 + noise * rng.standard_normal()) + 0.3 * abs(rng.standard_normal())))
        out.append(sample)
    return out


def loop_case(seed):
    """One sample whose KIR3DL3 depth looks like three copies: depthToCN refits with fewer bins around half
    of that depth until KIR3DL3 is called diploid (kir_cn.py:88-108)."""
    rng = np.random.default_rng(seed)
    d = cohort(rng, 1, per_copy=12.0, noise=0.25)
    d[0]["KIR3DL3*BACKBONE"] *= rng.uniform(1.2, 1.5)
    return d


def record(name, depths, **kw):
    diploid = kw.pop("diploid", None)
    with tempfile.TemporaryDirectory() as d:
        path = ""
        if diploid is not None:
            path = os.path.join(d, "dp")
            json.dump({"mean": diploid[0], "std": diploid[1]}, open(path + ".json", "w"))
        cns, dist = kc.depthToCN(depths, diploid_depth=path, **kw)
    return {"name": name, "depths": depths, "kwargs": kw, "diploid": diploid, "cns": [{k: int(v) for k, v in c.items()} for c in cns],
            "base": float(dist.base), "x_max": float(dist.x_max), "base_dev": float(dist.base_dev), "bin_num": int(dist.bin_num),
            "likelihood": np.asarray(dist.likelihood, dtype=float).tolist(),
            "group_prob": np.asarray(dist.calcCNGroupProb(dist.base), dtype=float).tolist()}


def main():
    rng = np.random.default_rng(20261019)
    cases = [
        record("cohort20", cohort(rng, 20)),
        record("cohort8_diploid_bounds", cohort(rng, 8, per_copy=22.0), diploid=(44.0, 5.0)),
        record("cohort12_start_base2", cohort(rng, 12, per_copy=9.0), cluster_method_kwargs={"start_base": 2}),
        record("single_sample", cohort(rng, 1, per_copy=30.0)),
        record("single_sample_3dl3_loop", loop_case(1), assume_3DL3_diploid=True),
        record("single_sample_3dl3_loop_b", loop_case(6), assume_3DL3_diploid=True),
        record("cohort5_3dl3", cohort(rng, 5, per_copy=18.0, noise=0.12), assume_3DL3_diploid=True),
        record("lcnd_alias_wide_dev", cohort(rng, 6, per_copy=40.0, noise=0.15), cluster_method="lcnd",
               cluster_method_kwargs={"base_dev": 0.1, "dev_decay": 1.0}),
        record("all_zero", [{g + "*BACKBONE": 0.0 for g in GENES}]),
    ]
    path = os.path.join(HERE, "cn_model.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump({"kind": "cn_model", "cases": cases}, f)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")
    for c in cases:
        print(c["name"], "base", round(c["base"], 4), "bins", c["bin_num"], "cn of sample 0", list(c["cns"][0].values()))


if __name__ == "__main__":
    main()
