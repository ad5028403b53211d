"""
Drop-in replacement of the reference's ``graphkir/cn_model.py`` for the proposed method: ``CNgroup``
(reference: cn_model.py:55-204).

Same class, attributes, parameters file (``getParams`` / ``setParams`` / ``save`` / ``load``) and methods
(``fit``, ``assignCN``, ``calcCNGroupProb``).  The likelihood curve of ``fit`` - for every candidate base the
sum over the depth bins of ``log(max over CN of the normal density) * histogram`` - and the CN-group
probabilities are evaluated on the GPU by ``gk_cn_fit`` in float64; the histogram, the grids
(``np.linspace``) and the argmax stay NumPy expressions, so the chosen base is a grid point computed exactly
as the reference computes it.  ``KDEcut`` (scikit-learn's KernelDensity) and the plots stay in the reference.
"""
from __future__ import annotations

import json
from typing import Any

import numpy as np

from . import engine


class Dist:
    """Abstract class of CN prediction model (cn_model.py:23-52)."""

    def __init__(self) -> None:
        self.raw_df: list[Any] = []

    def save(self, filename: str) -> None:
        from .kir_typing import NumpyEncoder
        with open(filename, "w") as f:
            json.dump(self.getParams(), f, cls=NumpyEncoder)

    @classmethod
    def load(cls, filename: str) -> "Dist":
        with open(filename) as f:
            data = json.load(f)
        return cls.setParams(data)

    def getParams(self) -> dict[str, Any]:
        raise NotImplementedError

    @classmethod
    def setParams(cls, data: dict[str, Any]) -> "Dist":
        raise NotImplementedError

    def plot(self, title: str = ""):
        raise NotImplementedError("plotting stays in the reference (plotly is not a dependency here)")


class CNgroup(Dist):
    """CN_group: linear copy-number distributions (cn_model.py:55-204)."""

    def __init__(self, _backend=None) -> None:
        super().__init__()
        self.bin_num: int = 300
        self.max_cn: int = 7
        self.x_max: float = 1
        self.base: float | None = None
        self.base_dev: float = 0.08
        self.y0_dev: float = 1.5
        self.dev_decay: float = 0.5
        self.dev_decay_neg: float = 0.3
        self.start_base: int = 1
        self.data: list[float] = []
        self.likelihood: np.ndarray = np.array([])
        self._backend = _backend

    def getParams(self) -> dict[str, Any]:
        return {
            "method": "CNgroup", "x_max": self.x_max, "base": self.base, "base_dev": self.base_dev,
            "y0_dev": self.y0_dev, "dev_decay": self.dev_decay, "dev_decay_neg": self.dev_decay_neg,
            "bin_num": self.bin_num, "max_cn": self.max_cn, "data": self.data, "likelihood": self.likelihood,
            "start_base": self.start_base, "raw_df": self.raw_df,
        }

    @classmethod
    def setParams(cls, data: dict[str, Any]) -> "CNgroup":
        assert data["method"] == "CNgroup"
        self = cls()
        self.base = data["base"]
        self.base_dev = data["base_dev"]
        self.x_max = data["x_max"]
        self.y0_dev = data["y0_dev"]
        self.dev_decay = data["dev_decay"]
        self.bin_num = data["bin_num"]
        self.max_cn = data["max_cn"]
        self.data = data["data"]
        self.raw_df = data.get("raw_df", [])
        self.likelihood = np.array(data["likelihood"])
        self.start_base = data.get("start_base", 1)
        self.dev_decay_neg = data.get("dev_decay_neg", self.dev_decay)
        return self

    # --- device evaluation ---------------------------------------------------------------
    def _evaluate(self, bases: np.ndarray, density: np.ndarray | None, want_prob: bool):
        """(likelihood [len(bases)], probabilities [len(bases), max_cn, bin_num] or None) from gk_cn_fit."""
        if self.start_base not in (1, 2):
            raise NotImplementedError
        be = self._backend if self._backend is not None else engine.default_backend()
        bases = np.ascontiguousarray(bases, dtype=np.float64)
        x = np.linspace(0, self.x_max, self.bin_num)
        dens = np.zeros(self.bin_num) if density is None else np.asarray(density, dtype=np.float64)
        d_like = be.empty(len(bases), np.float64)
        d_prob = be.empty(len(bases) * self.max_cn * self.bin_num, np.float64) if want_prob else None
        be.launch("gk_cn_fit", be.upload(x), be.upload(dens), be.upload(bases), len(bases), int(self.bin_num),
                  int(self.max_cn), int(self.start_base), float(self.base_dev), float(self.y0_dev),
                  float(self.dev_decay), float(self.dev_decay_neg), float(self.x_max / self.bin_num), d_like, d_prob)
        like = be.download(d_like, np.float64)[: len(bases)].copy()
        prob = None
        if want_prob:
            prob = be.download(d_prob, np.float64)[: len(bases) * self.max_cn * self.bin_num] \
                .reshape(len(bases), self.max_cn, self.bin_num).copy()
        return like, prob

    def fit(self, values: list[float], lower_bound: float = 0, upper_bound: float | None = None) -> None:
        """Find the base (mean depth of one copy) whose CN distributions fit the depths best (:124-168)."""
        if self.base is None:                      # normalise the first time
            max_depth = max(values) * 1.2
            self.base_dev *= max_depth
            self.x_max = max(max_depth, 1e-6)
            self.data = values
        if upper_bound is None:
            upper_bound = self.x_max
        density, _ = np.histogram(values, bins=self.bin_num, range=(0, self.x_max))
        bases = np.linspace(lower_bound, upper_bound, self.bin_num)
        like, _ = self._evaluate(bases, density, want_prob=False)
        self.likelihood = np.stack([bases, like], axis=1)          # n x 2 (base, likelihood of the base)
        self.base = self.likelihood[np.argmax(self.likelihood[:, 1]), :][0]

    def assignCN(self, values: list[float]) -> list[int]:
        """CN group of each depth (:168-174)."""
        assert self.base is not None
        cn_max = self.calcCNGroupProb(self.base).argmax(axis=0)
        space = self.x_max / self.bin_num
        return [cn_max[int(depth / space)] for depth in values]

    def calcCNGroupProb(self, base: float) -> np.ndarray:
        """(CN x bins) array: the probability that a normalised read depth belongs to the CN (:176-204).
        The last evaluation is kept: depthToCN calls assignCN once per sample with the same model."""
        key = (float(base), self.x_max, self.bin_num, self.max_cn, self.base_dev, self.y0_dev, self.dev_decay,
               self.dev_decay_neg, self.start_base)
        if getattr(self, "_prob_cache", (None, None))[0] != key:
            _, prob = self._evaluate(np.array([base], dtype=np.float64), None, want_prob=True)
            self._prob_cache = (key, prob[0])
        return self._prob_cache[1].copy()


def loadCNModel(filename: str) -> Dist:
    """Model of a saved parameter file (cn_model.py:382-390); only ``CNgroup`` files are read here."""
    with open(filename) as f:
        method = json.load(f)["method"]
    if method == "CNgroup":
        return CNgroup.load(filename)
    raise NotImplementedError("KDEcut models (scikit-learn) stay in the reference" if method == "KDEcut" else method)
