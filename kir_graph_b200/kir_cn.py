"""
The CN-prediction caller of the reference (``graphkir/kir_cn.py``) for the CNgroup model: ``depthToCN``
(kir_cn.py:41-123) with its KIR3DL3-is-diploid refit loop (:88-108), and ``loadCN`` (:234-243).  Reading
samtools depth files and aggregating them per gene (``predictSamplesCN``, ``filterDepth``) stay in the
reference, as does the KDE method.
"""
from __future__ import annotations

import json
from itertools import chain
from typing import Any

from .cn_model import CNgroup, Dist
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
