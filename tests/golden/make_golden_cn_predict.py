#!/usr/bin/env python
"""
Golden fixtures of the CN callers (the callers of the model) from the UNMODIFIED
reference: ``graphkir.kir_cn.predictSamplesCN`` (kir_cn.py:148-231), ``filterDepth`` (:126-145),
``graphkir.cn_model.loadCNModel`` (cn_model.py:382-390) and ``graphkir.utils.mergeCN`` (utils.py:168-180) on synthetic ``samtools depth`` tables.

Build container only (needs /root/reference):    python tests/golden/make_golden_cn_predict.py
Writes tests/golden/cn_predict.json.gz: per case the depth tables (file texts), the arguments, and what the
reference wrote - every ``.cn.tsv`` byte for byte and the saved model files (parsed, the temporary directory replaced by
``@DIR@``, the likelihood curve reduced to its length).
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
import graphkir.utils as gu  # noqa: E402
from make_golden_cn import GENES  # noqa: E402


def depth_table(rng, per_copy, n_pos=40, dl3_cn=2, noise=0.1):
    """Text of a ``samtools depth -a`` table: gene, position, depth; a few positions of every gene dip (ends)."""
    lines = []
    for g in GENES:
        cn = dl3_cn if g == "KIR3DL3" else int(rng.choice([0, 1, 1, 2, 2, 2, 3]))
        level = cn * per_copy * (1 + noise * rng.standard_normal())
        for p in range(n_pos):
            edge = min(p, n_pos - 1 - p, 5) / 5
            lines.append(f"{g}*BACKBONE\t{1 + 25 * p}\t{int(rng.poisson(max(level, 0.05) * (0.4 + 0.6 * edge)))}")
    return "\n".join(lines) + "\n"


def slim(params):
    """A saved model without its likelihood curve (tests/golden/cn_model.json.gz pins the curves; here the
    files' structure, the fitted base and the recorded inputs are what is compared) - key order kept."""
    if isinstance(params, list):
        return [slim(p) for p in params]
    return {k: (len(v) if k == "likelihood" else v) for k, v in params.items()}


def run(name, tables, diploid=None, **kw):
    with tempfile.TemporaryDirectory(prefix="gkcn_") as d:
        assert "-" not in d                       # the per-gene keys of the reference are split at "-"
        depth_files, cn_files = [], []
        for i, text in enumerate(tables):
            depth_files.append(os.path.join(d, f"s{i}.depth.tsv"))
            cn_files.append(os.path.join(d, f"s{i}.cn.tsv"))
            open(depth_files[-1], "w").write(text)
        path = ""
        if diploid is not None:
            path = os.path.join(d, "dp")
            json.dump({"mean": diploid[0], "std": diploid[1]}, open(path + ".json", "w"))
        model_path = os.path.join(d, "model.json")
        kc.predictSamplesCN(depth_files, cn_files, diploid_depth=path, save_cn_model_path=model_path, **kw)
        loaded = None
        if not kw.get("per_gene"):
            dist = cm.loadCNModel(model_path)
            loaded = {"base": float(dist.base), "cn_of_first": [int(x) for x in dist.assignCN(
                [float(v) for v in json.loads(open(model_path).read())["data"][:8]])]}
        models = {f[len("model.json"):]: slim(json.loads(open(os.path.join(d, f)).read().replace(d, "@DIR@")))
                  for f in sorted(os.listdir(d)) if f.startswith("model.json")}
        merged = os.path.join(d, "cohort.cn.tsv")
        gu.mergeCN(cn_files, merged)                                   # utils.py:168-180 (main.py:592)
        return {"name": name, "tables": tables, "diploid": diploid, "kwargs": kw,
                "cn_tsv": [open(f).read() for f in cn_files], "models": models, "loaded": loaded,
                "merged_cn": open(merged).read().replace(d, "@DIR@")}


def run_filter(tables):
    regions = {"KIR2DL1*BACKBONE": [(26, 201), (501, 600)], "KIR3DL3*BACKBONE": [(1, 76)], "KIR9XX*BACKBONE": [(1, 10)]}
    with tempfile.TemporaryDirectory(prefix="gkcn_") as d:
        src, dst = os.path.join(d, "a.tsv"), os.path.join(d, "b.tsv")
        open(src, "w").write(tables[0])
        kc.filterDepth(src, dst, regions)
        return {"regions": {k: [list(r) for r in v] for k, v in regions.items()}, "table": tables[0],
                "filtered": open(dst).read()}


def main():
    rng = np.random.default_rng(20261020)
    five = [depth_table(rng, 15.0) for _ in range(5)]
    six = [depth_table(rng, 22.0, n_pos=30) for _ in range(6)]
    one = [depth_table(rng, 30.0)]
    cases = [
        run("cohort5_p75", five),
        run("cohort5_median_bounds", five, diploid=(30.0, 4.0), select_mode="median"),
        run("single_mean_3dl3", one, select_mode="mean", assume_3DL3_diploid=True),
        run("cohort6_per_gene", six, per_gene=True),
        run("cohort6_lcnd_kwargs", six, cluster_method="lcnd", cluster_method_kwargs={"base_dev": 0.1}),
    ]
    names = ["KIR3DP1*0010101", "KIR2DL1*0320102N", "KIR3DP1*BACKBONE", "KIR2DS4", "KIR2DL5A*0010101e2", "KIR2DL1*new7", ""]
    fields = [[n, r, gu.getAlleleField(n, r), gu.limitAlleleField(n, r), gu.getGeneName(n)] for n in names for r in (3, 5, 7)]
    out = {"kind": "cn_predict", "cases": cases, "filter": run_filter(five), "allele_fields": fields}
    path = os.path.join(HERE, "cn_predict.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")
    for c in cases:
        print(c["name"], c["cn_tsv"][0].splitlines()[1:4], sorted(c["models"])[:3])


if __name__ == "__main__":
    main()
