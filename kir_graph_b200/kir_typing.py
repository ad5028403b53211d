"""
Drop-in replacement of the reference's ``graphkir/kir_typing.py``: per-sample typing driver.

``selectKirTypingModel(method, json, **kwargs)`` returns an object with the reference's
interface (``typing(gene_cn) -> (alleles, warning_genes)``, ``getAllPossibleTyping()``,
``save()``); methods ``full``, ``exonfirst[_x]`` and ``em`` are known, anything else raises
NotImplementedError (reference: kir_typing.py:207-228).  The classes below only orchestrate;
all arithmetic is in :mod:`typing_mulit_allele` / :mod:`typing_em` (GPU).
"""
from __future__ import annotations

import dataclasses
import json
from collections import defaultdict
from typing import Any

import numpy as np

from .engine import capacity_violation
from .hisat2 import loadReadsAndVariantsData, removeMultipleMapped
from .packing import CapacityError
from .typing_em import hisat2TypingPerGene, preprocessHisatReads, printHisatTyping
from .typing_mulit_allele import (AlleleTyping, AlleleTypingExonFirst, LazyAlleleProb, isHetrozygous)
from .utils import logger


class NumpyEncoder(json.JSONEncoder):
    """JSON encoder for results holding arrays / dataclasses (reference: utils.py:119-127)."""

    def default(self, obj: Any) -> Any:
        if isinstance(obj, LazyAlleleProb):
            return obj.tolist()
        if dataclasses.is_dataclass(obj):
            return {f.name: getattr(obj, f.name) for f in dataclasses.fields(obj)}
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.generic):
            return obj.item()
        return json.JSONEncoder.default(self, obj)


def groupReads(reads) -> dict[str, list]:
    by_gene = defaultdict(list)
    for read in reads:
        by_gene[read.backbone].append(read)
    return by_gene


def groupVariants(variants) -> dict[str, list]:
    by_gene = defaultdict(list)
    for variant in variants:
        by_gene[variant.ref].append(variant)
    return by_gene


class Typing:
    """Common driver: loop over the genes of the CN table (kir_typing.py:31-74)."""

    def __init__(self) -> None:
        self._result: dict[str, Any] = {}
        self.capacity_failures: dict[str, str] = {}      # gene -> which limit of the device path it exceeded

    def typingPerGene(self, gene: str, cn: int) -> tuple[list[str], int]:
        raise NotImplementedError

    def typing(self, gene_cn: dict[str, int], min_reads_num: int = 100) -> tuple[list[str], list[str]]:
        alleles: list[str] = []
        warning_genes: list[str] = []
        for gene, cn in gene_cn.items():
            if not cn:
                continue
            try:
                called, n_reads = self.typingPerGene(gene, cn)
            except CapacityError as exc:
                # a capacity of the device path the reference does not have (255 observations per read
                # pair, copy number 8, ...): this gene is reported as failed, the sample goes on
                logger.warning(f"[Allele] {gene} (cn={cn}) not typed: {exc}")
                self.capacity_failures[gene] = str(exc)
                called, n_reads = [f"{gene.split('*')[0]}*"] * cn, 0
            alleles.extend(called)
            if n_reads < min_reads_num:
                warning_genes.append(gene)
        return alleles, warning_genes

    def save(self, filename: str) -> None:
        with open(filename, "w") as f:
            json.dump(self._result, f, cls=NumpyEncoder)

    def getAllPossibleTyping(self) -> list[dict[Any, Any]]:
        raise NotImplementedError


class TypingWithPosNegAllele(Typing):
    """Positive/negative-variant typing (kir_typing.py:77-150)."""

    def __init__(self, filename_variant_json: str, top_n: int = 300, multiple: bool = False,
                 exon_first: bool = False, exon_only: bool = False, exon_candidate_threshold: float = .9,
                 variant_correction: bool = False, _backend=None, _fast: bool = False, _scan=None):
        """``_fast`` (not in the reference): read the ``.variant.json`` with the C++ scanner and pack
        the genes from arrays (:mod:`kir_graph_b200.fastjson`) instead of building a ``PairRead`` per
        read pair; same calls, ~3.5x less host time per sample.  The read objects are then not kept
        (``_gene_reads`` is empty), so it applies to the full-variant strategy only.  ``_scan``: the
        reads as arrays already in memory (``fastsam.extract(sam, variants).scan()``: SAM text ->
        calls without writing or reading the JSON); ``filename_variant_json`` is then not opened."""
        super().__init__()
        self._packs = None
        if _scan is not None and (exon_first or exon_only):
            raise ValueError("_scan holds no read objects: full-variant strategy only")
        if (_fast or _scan is not None) and not exon_first and not exon_only:
            from . import fastjson
            sc = _scan if _scan is not None else fastjson.scan(filename_variant_json)
            self._packs = fastjson.packs_from_scan(sc, variant_correction=variant_correction,
                                                   single_mapped_only=not multiple)
            self._gene_reads = {}
            self._gene_variants = fastjson.group_variants(sc.variants)
        else:
            reads_data = loadReadsAndVariantsData(filename_variant_json)
            if not multiple:
                reads_data = removeMultipleMapped(reads_data)
            self._gene_reads = groupReads(reads_data["reads"])
            self._gene_variants = groupVariants(reads_data["variants"])
        self._top_n = top_n
        self._exon_first = exon_first
        self._exon_only = exon_only
        self._exon_candidate_threshold = exon_candidate_threshold
        self._variant_correction = variant_correction
        self._backend = _backend
        self.tie_report: dict[str, list] = {}

    def typingPerGene(self, gene: str, cn: int) -> tuple[list[str], int]:
        logger.debug(f"[Allele] {gene=} {cn=}")
        force_homo = False if isHetrozygous(gene) else None
        problem = capacity_violation(len({a for v in self._gene_variants.get(gene, []) for a in v.allele}), cn,
                                     self._top_n)
        if problem:
            raise CapacityError(f"gene {gene}: {problem}")
        if self._packs is not None:
            pack = self._packs.get(gene)
            if isinstance(pack, CapacityError):
                raise pack
            if pack is None:                     # no variants, no reads: the reference's defaultdicts give []
                from .packing import pack_gene
                pack = pack_gene([], [], gene=gene)
            typ = AlleleTyping(None, self._gene_variants.get(gene, []), force_homo=force_homo, top_n=self._top_n,
                               variant_correction=self._variant_correction, _backend=self._backend,
                               _pack=pack)
        elif not self._exon_first and not self._exon_only:
            typ = AlleleTyping(self._gene_reads[gene], self._gene_variants[gene], force_homo=force_homo,
                               top_n=self._top_n, variant_correction=self._variant_correction,
                               _backend=self._backend)
        else:
            # the reference does not forward variant_correction here (kir_typing.py:117-124)
            typ = AlleleTypingExonFirst(self._gene_reads[gene], self._gene_variants[gene], force_homo=force_homo,
                                        top_n=self._top_n, exon_only=self._exon_only,
                                        candidate_set_threshold=self._exon_candidate_threshold,
                                        _backend=self._backend)
        res = typ.typing(cn)
        self._result[gene] = typ.result
        # a tie at a cut of an EARLIER step decides which sets the later steps grow, even when the
        # last step itself is tie-free: report every flagged step
        flagged = [{"n": r.n, "tie_flags": r.tie_flags} for r in typ.result if r.tie_flags]
        if flagged:
            self.tie_report[gene] = flagged
        alleles = res.selectBest()
        pure_gene = gene.split("*")[0]
        return [a if a != "fail" else f"{pure_gene}*" for a in alleles], typ.getReadsNum()

    def getAllPossibleTyping(self) -> list[dict[Any, Any]]:
        rows = []
        for gene, result in self._result.items():
            for rank, (value, alleles) in enumerate(result[-1].selectAllPossible(.9)):
                row = {"gene": gene, "rank": rank, "value": value}
                for i, allele in enumerate(alleles):
                    row[str(i + 1)] = allele
                rows.append(row)
        return rows


class TypingWithReport(Typing):
    """Allele calls from EM abundances (kir_typing.py:153-204)."""

    def __init__(self, filename_variant_json: str, _backend=None):
        super().__init__()
        reads_data = removeMultipleMapped(loadReadsAndVariantsData(filename_variant_json))
        self._gene_reads = preprocessHisatReads(reads_data)
        self._backend = _backend

    def typingPerGene(self, gene: str, cn: int) -> tuple[list[str], int]:
        if gene not in self._gene_reads:
            self._result[gene] = []
            return [], 0
        report = hisat2TypingPerGene(self._gene_reads[gene], _backend=self._backend)
        report = sorted(report, key=lambda x: -x.prob)           # stable; ties keep name order
        share = 1 / cn
        called = []
        for allele in report:
            copies = max(1, round(allele.prob / share))
            called.extend([allele.allele] * min(cn, copies))
            allele.cn = copies
            cn -= copies
            if cn <= 0:
                break
        self._result[gene] = report
        return called, len(self._gene_reads[gene])

    def save(self, filename: str) -> None:
        super().save(filename)
        name = filename[:-5] if filename.endswith(".json") else filename
        with open(name + ".txt", "w") as f:
            printHisatTyping(self._result, file=f)


def selectKirTypingModel(method: str, filename_variant_json: str, **kwargs: Any) -> Typing:
    """Typing model by strategy name (kir_typing.py:207-228)."""
    if method == "full":
        return TypingWithPosNegAllele(filename_variant_json, **kwargs)
    if method.startswith("exonfirst"):
        fields = method.split("_")
        threshold = 0.0
        if len(fields) == 2:
            threshold = float(method[len("exonfirst_"):])
        return TypingWithPosNegAllele(filename_variant_json, exon_first=True,
                                      exon_candidate_threshold=threshold, **kwargs)
    if method == "em":
        return TypingWithReport(filename_variant_json, **{k: v for k, v in kwargs.items() if k == "_backend"})
    raise NotImplementedError
