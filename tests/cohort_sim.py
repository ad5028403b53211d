"""A three-sample synthetic cohort on disk ({name}.json + {name}.depth.cn.tsv), shared by the golden
generator of the caller side (tests/golden/make_golden_main.py) and tests/test_main_callers.py."""
from __future__ import annotations

import os
from dataclasses import asdict

import pandas as pd

from kir_graph_b200 import synthetic
from kir_graph_b200.hisat2 import writeReadsAndVariantsData


def write_cohort(folder: str, n_samples: int = 3) -> tuple[list[str], list[str], list[dict]]:
    names, cn_files, inputs = [], [], []
    for s in range(n_samples):
        genes = [synthetic.make_gene([70 + s, i], f"KIRM{i}*BACKBONE", a, max(64, 8 * a), cn, r, hierarchical=True,
                                     variant_id_base=1000 * i)
                 for i, (a, cn, r) in enumerate([(12, 2, 260), (6, 1, 60 if s == 1 else 150), (9, 3, 300)])]
        reads, variants = [], []
        for g in genes:
            rd, va = g.to_objects()
            reads += rd
            variants += va
        name = os.path.join(folder, f"cohort.{s:02d}.index.variant") if folder != "." else f"cohort.{s:02d}.index.variant"
        writeReadsAndVariantsData({"variants": variants, "reads": reads}, name + ".json")
        cn = {g.gene: g.cn for g in genes}
        if s == 2:
            cn["KIRM1*BACKBONE"] = 0                 # a gene that is not typed
        cn["KIRNONE*BACKBONE"] = 1 if s == 0 else 0  # a gene without variants or reads
        cn_file = name + ".depth.cn.tsv"
        pd.DataFrame({"gene": list(cn), "cn": list(cn.values())}).to_csv(cn_file, sep="\t", index=False)
        names.append(name)
        cn_files.append(cn_file)
        inputs.append({"name": name, "cn": cn, "variants": [asdict(v) for v in variants],
                       "reads": [asdict(r) for r in reads]})
    return names, cn_files, inputs
