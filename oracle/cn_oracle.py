"""
TEST INFRASTRUCTURE -- NumPy restatement of the reference's CN model:
``CNgroup.calcCNGroupProb`` / ``fit`` / ``assignCN`` (graphkir/cn_model.py:124-204) and the CNgroup branch of
``depthToCN`` (graphkir/kir_cn.py:41-123).  Pinned to tests/golden/cn_model.json.gz, which
tests/golden/make_golden_cn.py wrote with the unmodified reference.  Only tests/ may import it.
"""
from __future__ import annotations

from itertools import chain

import numpy as np

SQRT_2PI = np.sqrt(2 * np.pi)


def norm_pdf(x: np.ndarray, loc: float, scale: float) -> np.ndarray:
    """scipy.stats.norm.pdf: exp(-y^2 / 2) / sqrt(2 pi) / scale with y = (x - loc) / scale; nan for scale <= 0."""
    if not scale > 0:
        return np.full(np.shape(x), np.nan)
    y = (np.asarray(x, dtype=np.float64) - loc) / scale
    return np.exp(-y ** 2 / 2.0) / SQRT_2PI / scale


class CNParams:
    """The parameters of CNgroup (cn_model.py:69-88)."""

    def __init__(self, **kw):
        self.bin_num, self.max_cn = 300, 7
        self.x_max, self.base, self.base_dev, self.y0_dev = 1.0, None, 0.08, 1.5
        self.dev_decay, self.dev_decay_neg, self.start_base = 0.5, 0.3, 1
        for k, v in kw.items():
            setattr(self, k, v)


def group_prob(p: CNParams, base: float) -> np.ndarray:
    """calcCNGroupProb (cn_model.py:176-204): (CN x bins) probabilities of a depth bin under each CN."""
    x = np.linspace(0, p.x_max, p.bin_num)
    if p.start_base == 1:
        rows = [norm_pdf(x, 0, p.base_dev * p.y0_dev)]
        rows += [norm_pdf(x, base * n, p.base_dev * (p.dev_decay * (n - 1) + 1)) for n in range(1, p.max_cn)]
    elif p.start_base == 2:
        rows = []
        for n in range(0, p.max_cn):
            dev = p.base_dev * (p.dev_decay_neg * (p.start_base - n) + 1) if n < p.start_base \
                else p.base_dev * (p.dev_decay * (n - p.start_base) + 1)
            rows.append(norm_pdf(x, base * n, dev))
    else:
        raise NotImplementedError
    return np.array(rows) * (p.x_max / p.bin_num)


def fit(p: CNParams, values, lower_bound: float = 0, upper_bound=None) -> np.ndarray:
    """CNgroup.fit (cn_model.py:124-168); returns the likelihood curve and sets p.base."""
    if p.base is None:
        max_depth = max(values) * 1.2
        p.base_dev *= max_depth
        p.x_max = max(max_depth, 1e-6)
    if upper_bound is None:
        upper_bound = p.x_max
    density, _ = np.histogram(values, bins=p.bin_num, range=(0, p.x_max))
    curve = []
    for base in np.linspace(lower_bound, upper_bound, p.bin_num):
        max_prob = group_prob(p, base).max(axis=0)
        curve.append((base, np.sum(np.log(max_prob + 1e-9) * density)))
    curve = np.array(curve)
    p.base = curve[np.argmax(curve[:, 1]), 0]
    return curve


def assign_cn(p: CNParams, values) -> list[int]:
    """CNgroup.assignCN (cn_model.py:168-174)."""
    cn_max = group_prob(p, p.base).argmax(axis=0)
    space = p.x_max / p.bin_num
    return [int(cn_max[int(depth / space)]) for depth in values]


def depth_to_cn(sample_gene_depths, diploid=None, kwargs=None, assume_3dl3_diploid=False):
    """The CNgroup branch of depthToCN (kir_cn.py:41-123); ``diploid`` = (mean, std) of the diploid coverage."""
    values = list(chain.from_iterable(d.values() for d in sample_gene_depths))
    p = CNParams(**(kwargs or {}))
    lower, upper = 0.0, None
    if diploid is not None:
        lower, upper = (diploid[0] - diploid[1]) / 2, (diploid[0] + diploid[1]) / 2
    else:
        p.bin_num += 200
    curve = fit(p, values, lower, upper)
    if assume_3dl3_diploid:
        dl3 = [float(d["KIR3DL3*BACKBONE"]) for d in sample_gene_depths]
        cn = assign_cn(p, dl3)
        perc, original = 1.0, p.bin_num
        while not all(i == 2 for i in cn):
            mean = sum(dl3) / len(dl3)
            p.bin_num = int(original * perc)
            curve = fit(p, values, (mean - perc * 10) / 2, (mean + perc * 10) / 2)
            cn = assign_cn(p, dl3)
            perc -= 0.2
            if perc <= 0:
                break
        assert all(i == 2 for i in cn)
    cns = [dict(zip(d.keys(), assign_cn(p, list(d.values())))) for d in sample_gene_depths]
    return cns, p, curve
