#!/usr/bin/env python
"""
Golden fixture of the whole extraction loop, from the UNMODIFIED reference:

    readPair -> filter(filterRead both mates) -> extractVariant(pileup=None)

(graphkir/hisat2.py:228-276, :541-578, :803-844, driven as extractVariantFromBam :923-932 does with
error_correction=False).  ``readBam`` - a samtools subprocess - is replaced by an iterator over the
simulated name-sorted SAM text; nothing else of the reference is touched.  Build container only:

    python tests/golden/make_golden_sam_extract.py
"""
from __future__ import annotations

import os
import sys
from dataclasses import asdict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import dump, import_reference  # noqa: E402

from tests import sam_sim  # noqa: E402


def main() -> None:
    _, _, _, h2, m2h = import_reference()
    cases = []
    for seed, n_pairs, nm, novel in ((51, 150, 4, 0.002), (52, 90, 9, 0.03), (53, 60, 1000, 0.01)):
        table, pairs = sam_sim.multi_gene(seed, n_pairs=n_pairs, novel=novel)
        text = sam_sim.sam_text(pairs)
        lines = text.rstrip("\n").split("\n")
        # records the pairing has to cope with: mate on another reference, a record without a mate,
        # two first-mate records under one name, an empty line
        first = lines[2].split("\t")
        extra = ["\t".join(["lonely", "99", first[2], "10", "60", "*", "chrX", "50"] + first[8:]),
                 "\t".join(["single", "73", first[2], "10", "60"] + first[5:]),
                 "\t".join(["odd", "99"] + first[2:]),
                 "\t".join(["odd", "99", first[2], first[7]] + first[4:7] + [first[3]] + first[8:]), ""]
        text = "\n".join(lines[:30] + extra + lines[30:]) + "\n"
        rtable = [m2h.Variant(**asdict(v)) for v in table]
        h2.readBam = lambda _name, _text=text: iter(_text.split("\n"))
        m2h.Variant.novel_id = 0
        pair_reads = h2.readPair("unused.bam")
        pair_reads = filter(lambda lr: h2.filterRead(lr[0], nm) and h2.filterRead(lr[1], nm), pair_reads)
        data = h2.extractVariant(pair_reads, rtable, pileup=None)
        cases.append({
            "num_editdist": nm, "sam": text, "table": [asdict(v) for v in table],
            "variants": [asdict(v) for v in data["variants"]], "reads": [asdict(r) for r in data["reads"]],
            "novel_id_after": m2h.Variant.novel_id,
        })
        full = sum(1 for r in data["reads"] if r.lpv or r.lnv or r.rpv or r.rnv)
        print(f"seed {seed}: {len(data['reads'])} pairs kept ({full} with observations), "
              f"{m2h.Variant.novel_id} novel variants")
    dump("sam_extract", {"kind": "sam_extract", "cases": cases})


if __name__ == "__main__":
    main()
