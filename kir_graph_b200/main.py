"""
The callers of the typing path: per-sample allele typing with the reference's outputs, and a
cohort entry that types every sample's genes as one GPU batch.

``alleleTyping`` mirrors the reference function of that name (graphkir/main.py:171-220): per sample
``selectKirTypingModel(method, name + ".json", top_n=600, variant_correction=True)``, ``loadCN``,
``typing``, then ``{name}{suffix}.tsv`` (name / alleles / warnings) and ``.possible.tsv``; file names,
columns and formatting come from the same pandas calls, so the files are interchangeable.
``mergeAllele`` (utils.py:161-165) concatenates them into ``cohort.allele.tsv``; ``mergeCN`` (utils.py:168-180)
pivots the ``.cn.tsv`` files into ``cohort.cn.tsv``.

``cohortAlleleTyping`` is new: the (sample, gene) problems of all samples of
this rank go through ``cohort.CohortTyper`` in one pass - samples are independent, so ranks take
samples ``rank, rank + world, ...`` with no data-path communication and write their own ``.tsv``
files; the caller merges on rank 0 in input order.  It covers the ``full`` strategy (the only one
that needs no per-gene Python objects) and writes no ``.possible.tsv`` (only the called set is read
back from the device); other strategies go through ``alleleTyping``.  One deliberate difference: a
gene with copy number >= 2 and no usable reads is called ``<gene>*`` (fail) here, where the reference
- and the per-sample mirror - end in numpy's AxisError (``createHomoResult`` on the empty first-step
result, typing_mulit_allele.py:441) unless the gene is one of the always-heterozygous ones; likewise a
gene on which ``isHomozygous`` ends in IndexError (:853, a site with ten or more equally frequent values)
is reported and called fail instead of stopping every sample of the batch.

Mapping and BAM handling stay in the reference (copy-number estimation from depth tables:
:mod:`kir_graph_b200.kir_cn`); the CLI itself is not rebuilt.
"""
from __future__ import annotations

from typing import Any

import pandas as pd

from .kir_typing import selectKirTypingModel
from .utils import logger


def getCommonName(r1: str, r2: str) -> str:
    """Longest common dot-separated prefix of two file names (main.py:223-250)."""
    name = ""
    for s1, s2 in zip(r1.split("."), r2.split(".")):
        if s1 != s2:
            return name
        name = name + "." + s1 if name else s1
    return name


def loadCN(filename_cn: str) -> dict[str, int]:
    """gene -> copy number of a ``.cn.tsv`` (kir_cn.py:234-243)."""
    data = pd.read_csv(filename_cn, sep="\t", index_col=[0])
    return dict(data.to_dict()["cn"])


def mergeAllele(allele_result_files: list[str], final_result_file: str) -> pd.DataFrame:
    """Per-sample ``.tsv`` files -> ``cohort.allele.tsv`` (utils.py:161-165)."""
    df = pd.concat(pd.read_csv(f, sep="\t") for f in allele_result_files)
    df.to_csv(final_result_file, index=False, sep="\t")
    return df


def mergeCN(cn_result_files: list[str], final_result_file: str) -> pd.DataFrame:
    """Gene x sample table of the copy numbers of several ``.cn.tsv`` files, columns named by file, a gene a
    sample does not list counted as 0 (utils.py:168-180; what the reference's CLI writes as ``cohort.cn.tsv``)."""
    tables = []
    for f in cn_result_files:
        one = pd.read_csv(f, sep="\t")
        one["name"] = f
        tables.append(one)
    merged = pd.pivot_table(pd.concat(tables), values="cn", index="gene", columns=["name"]).fillna(0).astype(int)
    merged.to_csv(final_result_file, sep="\t")
    return merged


def _suffix(name: str, cn_file: str, method: str) -> str:
    """``.cn<rest of the CN file name>.<method>`` (main.py:182-192)."""
    return (".cn" + cn_file[len(getCommonName(name, cn_file)):].replace("/", "_").replace(".", "_") + "."
            + method)


def _write_sample(name: str, called_alleles: list[str], warning_genes: list[str]) -> str:
    df = pd.DataFrame({"name": [name], "alleles": ["_".join(called_alleles)],
                       "warnings": ["_".join(warning_genes)]})
    df.to_csv(name + ".tsv", sep="\t", index=False)
    return name + ".tsv"


def alleleTyping(processed_bam: list[str], cn_files: list[str], method: str = "full", **kwargs: Any) -> list[str]:
    """Allele typing sample by sample (main.py:171-220).  ``kwargs`` reach the typing model
    (``_backend``, ``_fast``)."""
    allele_files = []
    for name, cn_file in zip(processed_bam, cn_files):
        logger.debug(f"[Allele] Allele typing ({method}) with CN {cn_file} ({name})")
        if method == "exonfirst":
            method += "_1"
        suffix = _suffix(name, cn_file, method)
        t = selectKirTypingModel(method, name + ".json", top_n=600, variant_correction=True, **kwargs)
        called_alleles, warning_genes = t.typing(loadCN(cn_file))
        logger.info(f"[Allele] {called_alleles} ({name})")
        name += suffix
        allele_files.append(_write_sample(name, called_alleles, warning_genes))
        logger.info("[Allele] All possible allele set in Allele typing saved in [name].possible.tsv")
        pd.DataFrame(t.getAllPossibleTyping()).fillna("").to_csv(name + ".possible.tsv", index=False, sep="\t")
    return allele_files


def cohortAlleleTyping(processed_bam: list[str], cn_files: list[str], method: str = "full", top_n: int = 600,
                       min_reads_num: int = 100, rank: int = 0, world: int = 1, n_parts: int = 6,
                       workers: int = 0, batch_samples: int = 96, _backend=None) -> list[str]:
    """The samples ``rank, rank + world, ...`` typed as one batch on this rank's GPU; writes the
    ``{name}{suffix}.tsv`` of each (same bytes as ``alleleTyping``) and returns the names of ALL
    samples' files in input order, so that rank 0 can ``mergeAllele`` them once every rank is done.
    ``workers``: processes that scan and pack the samples' ``.json`` files in parallel (the host
    preparation is seconds per 200k-pair sample, the typing itself a fraction of a millisecond);
    0 packs them one after the other in this process.  ``batch_samples``: at most that many samples are
    loaded and typed per GPU pass (96 samples of 200 k read pairs take 58 GB of device memory and 0.46 G of the
    2^31 observation entries a batch can address), so a rank's share of a large cohort goes through in several
    passes instead of failing; the files do not depend on it."""
    if method != "full":
        raise NotImplementedError("cohortAlleleTyping covers the full-variant strategy; use alleleTyping")
    mine = list(range(len(processed_bam)))[rank::world]
    step = max(1, int(batch_samples))
    for at in range(0, len(mine), step):
        _type_samples(mine[at:at + step], processed_bam, cn_files, method, top_n, min_reads_num, n_parts, workers,
                      _backend)
    return [n + _suffix(n, c, method) + ".tsv" for n, c in zip(processed_bam, cn_files)]


def _type_samples(mine: list[int], processed_bam: list[str], cn_files: list[str], method: str, top_n: int,
                  min_reads_num: int, n_parts: int, workers: int, _backend) -> None:
    """One GPU pass over the samples ``mine`` (indices into ``processed_bam``) and their ``.tsv`` files."""
    import functools
    from . import cohort, engine, packio
    from .packing import CapacityError
    jsons = [processed_bam[i] + ".json" for i in mine]
    # one sample's packed genes: the .gkpack.npz sidecar when it is fresh, else the .json through the scanner
    load = functools.partial(packio.load_sample_packs, variant_correction=True)
    if workers > 1 and len(mine) > 1:
        import multiprocessing
        from concurrent.futures import ProcessPoolExecutor
        # spawn: the parent may hold a CUDA context, which a forked child must not inherit; the workers
        # import kir_graph_b200.packio / fastjson only (numpy + the C++ scanner, no pandas / torch)
        with ProcessPoolExecutor(max_workers=min(workers, len(mine)),
                                 mp_context=multiprocessing.get_context("spawn")) as pool:
            packed = list(pool.map(load, jsons))
    else:
        packed = [load(j) for j in jsons]
    packs, cns = [], []
    plans = {}
    for i, by_gene in zip(mine, packed):
        name, cn_file = processed_bam[i], cn_files[i]
        gene_cn = loadCN(cn_file)
        # CN-file order, as Typing.typing; the flag says whether the gene goes to the device.  A gene that
        # exceeds a capacity of the device path (255 observations in a read pair, copy number above 8,
        # ...) is reported and called "fail"; it does not take the other genes and samples down with it.
        plans[i] = []
        for g, c in gene_cn.items():
            if not c:
                continue
            known = g in by_gene
            if known:
                pack = by_gene[g]
                problem = str(pack) if isinstance(pack, CapacityError) else \
                    engine.capacity_violation(pack.n_alleles, int(c), top_n)
                if not problem and cohort.undecidable_homozygosity(pack, int(c)):
                    # the reference (and the per-sample mirror) end in IndexError inside isHomozygous here,
                    # which would take the whole batch down
                    problem = "isHomozygous cannot decide (a site without any value above the share filter)"
                if problem:
                    logger.warning(f"[Allele] {g} (cn={c}) of {name} not typed: {problem}")
                    known = False
            plans[i].append((g, int(c), known))
        for gene, cn, known in plans[i]:
            if known:
                packs.append(by_gene[gene])
                cns.append(cn)
    calls = []
    if packs:
        try:
            typer = cohort.CohortTyper(packs, cns, top_n=top_n, backend=_backend, n_parts=n_parts)
        except ValueError as exc:
            # more observation entries (or wire units) than a batch addresses - samples far deeper than the
            # 200 k pairs ``batch_samples`` is sized for: type the two halves on their own
            if "split the batch" not in str(exc) or len(mine) < 2:
                raise
            logger.info(f"[Allele] {len(mine)} samples do not fit one batch ({exc}): typing them in two")
            del packs, packed
            for half in (mine[: len(mine) // 2], mine[len(mine) // 2:]):
                _type_samples(half, processed_bam, cn_files, method, top_n, min_reads_num, n_parts, workers, _backend)
            return
        calls = typer.upload_and_run()
    at = 0
    for i in mine:
        alleles, warnings = [], []
        for gene, cn, known in plans[i]:
            pure_gene = gene.split("*")[0]
            if known:
                call = calls[at]
                at += 1
                alleles.extend(a if a != "fail" else f"{pure_gene}*" for a in call.alleles)
                n_reads = call.n_reads
            else:
                # a gene without variants or reads never reaches the device: the reference's
                # defaultdict grouping (kir_typing.py:15-28) gives it no reads -> ["fail"] * cn
                alleles.extend([f"{pure_gene}*"] * cn)
                n_reads = 0
            if n_reads < min_reads_num:
                warnings.append(gene)
        name = processed_bam[i] + _suffix(processed_bam[i], cn_files[i], method)
        logger.info(f"[Allele] {alleles} ({processed_bam[i]})")
        _write_sample(name, alleles, warnings)
