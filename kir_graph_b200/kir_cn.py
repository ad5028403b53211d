"""
The CN-prediction caller of the reference (``graphkir/kir_cn.py``) for the CNgroup model: ``depthToCN``
(kir_cn.py:41-123) with its KIR3DL3-is-diploid refit loop (:88-108), ``loadCN`` (:234-243), and the callers that turn
``samtools depth`` tables into ``.cn.tsv`` files: ``readSamtoolsDepth`` (samtools_utils.py:17-22),
``selectSamtoolsDepth`` / ``aggrDepths`` / ``filterDepth`` (kir_cn.py:15-38, :126-145) and ``predictSamplesCN``
(:148-231, cohort-wide and per-gene fits, the saved model files).  The ``samtools depth`` subprocess itself
(``bam2Depth``) and the KDE method stay in the reference.
"""
from __future__ import annotations

import json
from itertools import chain
from typing import Any

import pandas as pd

from .cn_model import CNgroup, Dist
from .kir_typing import NumpyEncoder
from .main import loadCN  # noqa: F401  (same function, kept importable from here as in the reference)
from .utils import logger


def depthToCN(sample_gene_depths: list[dict[str, float]], diploid_depth: str = "", cluster_method: str = "CNgroup",
              cluster_method_kwargs: dict[str, Any] = {}, assume_3DL3_diploid: bool = False,
              _backend=None) -> tuple[list[dict[str, int]], Dist]:
    """Depths of gene -> CN of gene, per sample, and the fitted model (kir_cn.py:41-123)."""
    values = list(chain.from_iterable(map(lambda i: i.values(), sample_gene_depths)))
    logger.info(f"[CN] Predict copy number by {cluster_method} with data size {len(values)}")
    if cluster_method == "CNgroup" or cluster_method.lower() == "lcnd":
        dist = CNgroup()
        if cluster_method_kwargs:
            dist = CNgroup.setParams(dist.getParams() | cluster_method_kwargs)
        dist._backend = _backend
        lower_bound = 0.0
        upper_bound = None
        if diploid_depth != "":
            with open(diploid_depth + ".json", "r") as f:
                dp_info = json.load(f)
                mean = float(dp_info["mean"])
                dev = float(dp_info["std"])
                lower_bound = (mean - dev) / 2
                upper_bound = (mean + dev) / 2
        else:
            dist.bin_num += 200
        dist.fit(values, lower_bound, upper_bound)
        if assume_3DL3_diploid:
            kir3dl3_depths = [float(gene_depths["KIR3DL3*BACKBONE"]) for gene_depths in sample_gene_depths]
            cn = dist.assignCN(kir3dl3_depths)
            decrease_perc = float(1)
            decrease_rate = 0.2
            original_bin_num = dist.bin_num
            while not all(i == 2 for i in cn):
                logger.debug("[CN] Assume 3DL3 cn=2")
                kir3dl3_depth = sum(kir3dl3_depths) / len(kir3dl3_depths)
                lower_3dl3 = (kir3dl3_depth - decrease_perc * 10) / 2
                upper_3dl3 = (kir3dl3_depth + decrease_perc * 10) / 2
                dist.bin_num = int(original_bin_num * decrease_perc)
                dist.fit(values, lower_3dl3, upper_3dl3)
                cn = dist.assignCN(kir3dl3_depths)
                decrease_perc = decrease_perc - decrease_rate
                if decrease_perc <= 0:
                    break
            assert all(i == 2 for i in cn)
        logger.info(f"[CN] {cluster_method} base = {dist.base}")
    elif cluster_method.lower() == "kde":
        raise NotImplementedError("the KDE method (scikit-learn KernelDensity) stays in the reference")
    else:
        raise NotImplementedError
    sample_gene_cns = []
    for gene_depths in sample_gene_depths:
        genes, depths = zip(*gene_depths.items())
        sample_gene_cns.append(dict(zip(genes, dist.assignCN(depths))))
    return sample_gene_cns, dist


def readSamtoolsDepth(depth_filename: str) -> pd.DataFrame:
    """The three columns of a ``samtools depth`` table: gene, pos, depth (samtools_utils.py:17-22)."""
    return pd.read_csv(depth_filename, sep="\t", header=None, names=["gene", "pos", "depth"])


def selectSamtoolsDepth(df: pd.DataFrame, ref_regions: dict[str, list[tuple[int, int]]]) -> pd.DataFrame:
    """Rows inside the given (inclusive) regions of each gene, region by region in the order given
    (kir_cn.py:15-26; used to drop intron depths).  No region at all is an error, as in the reference."""
    gene, pos = df["gene"], df["pos"]
    return pd.concat([df[(gene == name) & (pos >= lo) & (pos <= hi)]
                      for name, regions in ref_regions.items() for lo, hi in regions])


_AGGREGATE = {"median": lambda g: g.median(), "mean": lambda g: g.mean(), "p75": lambda g: g.quantile(0.75)}


def aggrDepths(depths: pd.DataFrame, select_mode: str = "p75") -> pd.DataFrame:
    """Depth per position -> one depth per gene (kir_cn.py:29-38)."""
    if select_mode not in _AGGREGATE:
        raise NotImplementedError
    return _AGGREGATE[select_mode](depths.groupby(by="gene", as_index=False)["depth"])


def filterDepth(depth_file: str, filtered_depth_file: str,
                bam_selected_regions: dict[str, list[tuple[int, int]]] = {}) -> None:
    """Depth table restricted to the selected regions, written in the same headerless format (kir_cn.py:126-145)."""
    depths = selectSamtoolsDepth(readSamtoolsDepth(depth_file), bam_selected_regions)
    depths.to_csv(filtered_depth_file, header=False, index=False, sep="\t")


def predictSamplesCN(samples_depth_tsv: list[str], samples_cn: list[str], diploid_depth: str = "",
                     save_cn_model_path: str | None = None, assume_3DL3_diploid: bool = False,
                     select_mode: str = "p75", per_gene: bool = False, cluster_method: str = "CNgroup",
                     cluster_method_kwargs: dict[str, Any] = {}, _backend=None) -> None:
    """Depth tables of the samples -> one ``gene / cn / depth`` table per sample (kir_cn.py:148-231).

    One model over every gene of every sample, or (``per_gene``) one model per gene over the samples - fitted
    without the diploid bounds and without the KIR3DL3 assumption, as the reference does.  The model
    parameters go to ``save_cn_model_path`` (per gene: a list, plus one ``.{gene}.json`` each).  Per-gene
    results are keyed ``{gene}-{file}``; the reference splits that key at every ``-`` and so fails on file
    names that contain one, here the key is split at the first ``-`` only."""
    assert len(samples_depth_tsv) == len(samples_cn)
    tables = []
    for depth_file in samples_depth_tsv:
        logger.info(f"[CN] Select {select_mode} of depths per gene ({depth_file})")
        table = aggrDepths(readSamtoolsDepth(depth_file), select_mode=select_mode)
        table["depth_file"] = depth_file
        tables.append(table)
    logger.info(f"[CN] Predict CN from {len(tables)} samples")
    depths_dict = [dict(zip(t["gene"], t["depth"])) for t in tables]
    if not per_gene:
        cns, model = depthToCN(depths_dict, diploid_depth, cluster_method=cluster_method,
                               cluster_method_kwargs=cluster_method_kwargs,
                               assume_3DL3_diploid=assume_3DL3_diploid, _backend=_backend)
        model.raw_df = [t.to_dict() for t in tables]
        if save_cn_model_path:
            model.save(save_cn_model_path)
    else:
        sample_of = {name: i for i, name in enumerate(samples_depth_tsv)}
        every = pd.concat(tables)
        every["gene_sampleid"] = every["gene"] + "-" + every["depth_file"]
        cns = [{} for _ in tables]
        params = []
        for gene in sorted(set(every["gene"])):
            logger.info(f"[CN] Predict per gene: {gene}")
            rows = every[every["gene"] == gene]
            gene_cns, gene_model = depthToCN([dict(zip(rows["gene_sampleid"], rows["depth"]))],
                                             cluster_method=cluster_method,
                                             cluster_method_kwargs=cluster_method_kwargs, _backend=_backend)
            gene_model.raw_df = [rows.to_dict()]
            params.append(gene_model.getParams() | {"gene": gene})
            for key, cn in gene_cns[0].items():
                cns[sample_of[key.split("-", 1)[1]]][gene] = cn
        if save_cn_model_path:
            for one in params:
                with open(save_cn_model_path + f".{one['gene']}.json", "w") as f:
                    json.dump(one, f, cls=NumpyEncoder)
            with open(save_cn_model_path, "w") as f:
                json.dump(params, f, cls=NumpyEncoder)
    for filename, cn, depths in zip(samples_cn, cns, depths_dict):
        calls = pd.DataFrame(list(cn.items()), columns=["gene", "cn"])
        seen = pd.DataFrame(list(depths.items()), columns=["gene", "depth"])
        calls.merge(seen, on="gene").to_csv(filename, index=False, sep="\t")
