#!/usr/bin/env python
"""
Golden fixture of the index readers, from the UNMODIFIED reference: graphkir.hisat2.getVariants
(readVariants / readLink / readExons / isInExon, hisat2.py:121-225) over small synthetic
.snp / .link / .locus files.  Build container only:

    python tests/golden/make_golden_index.py
"""
from __future__ import annotations

import os
import sys
import tempfile
from dataclasses import asdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import dump, import_reference  # noqa: E402


def random_index(rng: np.random.Generator) -> dict[str, str]:
    genes = [f"KIR{g}*BACKBONE" for g in range(int(rng.integers(1, 4)))]
    snp, link, locus = [], [], []
    k = 0
    for g in genes:
        cuts = sorted(rng.choice(np.arange(1, 3000), size=2 * int(rng.integers(1, 5)), replace=False).tolist())
        exons = " ".join(f"{cuts[i]}-{cuts[i + 1]}" for i in range(0, len(cuts), 2))
        locus.append("\t".join([g, g, "0", "3000", "3000", exons, "+"]))
        for _ in range(int(rng.integers(5, 40))):
            typ = str(rng.choice(["single", "deletion", "insertion"]))
            pos = int(rng.integers(0, 3000))
            val = (str(int(rng.integers(1, 30))) if typ == "deletion"
                   else "".join(rng.choice(list("ACGT"), size=1 if typ == "single" else int(rng.integers(1, 4)))))
            vid = f"hv{k}"
            k += 1
            snp.append("\t".join([vid, typ, g, str(pos), val]))
            if rng.random() < 0.9:               # a variant without a .link line carries no allele
                link.append(vid + "\t" + " ".join(f"{g.split('*')[0]}*{int(a):03d}"
                                                   for a in rng.choice(20, size=int(rng.integers(1, 6)), replace=False)))
    return {"snp": "\n".join(snp) + "\n", "link": "\n".join(link) + "\n", "locus": "\n".join(locus) + "\n"}


def main() -> None:
    _, _, _, h2, _ = import_reference()
    rng = np.random.default_rng(8)
    cases = []
    for _ in range(4):
        files = random_index(rng)
        with tempfile.TemporaryDirectory() as tmp:
            index = os.path.join(tmp, "kir")
            for ext, text in files.items():
                with open(f"{index}.{ext}", "w") as handle:
                    handle.write(text)
            variants = h2.getVariants(index)
        cases.append({"files": files, "variants": [asdict(v) for v in variants]})
        print(len(variants), "variants,", sum(v.in_exon for v in variants), "in exons")
    dump("index_readers", {"kind": "index", "cases": cases})


if __name__ == "__main__":
    main()
