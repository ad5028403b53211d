"""
Read grouping of the novel-variant discovery step (reference: graphkir/novel_discover.py:48-70,
:267-277) on the device-resident likelihood - SURVEY.md section 8f rank 3, the next consumer of
``AlleleTyping.probs`` after typing itself.

Only the part that touches the likelihood is mirrored: ``groupReadByAllele`` (which reads pair
belongs to which of the called alleles) and its driver ``splitReadsByAlleles``.  The pile-up /
MSA side of novel discovery (samtools, pyhlamsa) stays in the reference.

Difference to the reference, on purpose: the reference compares float64 products
(``np.equal(probs, probs.max(1))``); two alleles with the same mismatch count can differ in the last
bit there, because the factors .999 / .001 are multiplied in variant order.  Here the comparison is
on the integer mismatch counts (``gk_group_reads``), i.e. exact ties are ties.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np

from .hisat2 import PairRead
from .msa2hisat import Variant
from .typing_mulit_allele import AlleleTyping

GroupPairRead = dict[tuple[str, ...], list[PairRead]]


def groupReadByAllele(typ: AlleleTyping, predict_alleles: list[str], reads: list[PairRead]) -> GroupPairRead:
    """Assign the reads to the called alleles they fit best (novel_discover.py:48-70).

    ``reads`` must be the list ``typ`` was built from with ``no_empty=False`` (one likelihood row per
    read), as in ``splitReadsByAlleles``."""
    allele_names = []
    allele_ids = []
    for i in predict_alleles:
        if i in typ.allele_to_id:
            allele_names.append(i)
            allele_ids.append(typ.allele_to_id[i])
    if not allele_names:
        return {}
    if len(reads) != typ.getReadsNum():
        raise ValueError("groupReadByAllele needs one likelihood row per read (build the model with no_empty=False)")
    pattern = typ.group_pattern(allele_ids)
    names = np.array(allele_names)
    bits = np.arange(len(allele_names), dtype=np.uint32)
    assign_reads: GroupPairRead = {}
    keys: dict[int, tuple[str, ...]] = {}
    for read, pat in zip(reads, pattern.tolist()):              # dict order = first occurrence, as in the reference
        key = keys.get(pat)
        if key is None:
            key = keys[pat] = tuple(sorted(names[(np.uint32(pat) >> bits) & 1 == 1].tolist()))
            assign_reads.setdefault(key, [])
        assign_reads[key].append(read)
    return assign_reads


def splitReadsByAlleles(pn_typing_model, predict_alleles: list[str]
                        ) -> Iterable[tuple[str, tuple[str, ...], list[PairRead], dict[str, Variant]]]:
    """Assign the reads of every gene to alleles (novel_discover.py:267-277)."""
    for gene, reads in pn_typing_model._gene_reads.items():
        typ = AlleleTyping(reads, pn_typing_model._gene_variants[gene], no_empty=False,
                           _backend=getattr(pn_typing_model, "_backend", None))
        assert typ.getReadsNum() == len(reads)
        assign_reads = groupReadByAllele(typ, predict_alleles, reads)
        for alleles, group in assign_reads.items():
            yield gene, alleles, group, typ.variants
