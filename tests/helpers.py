"""Shared test helpers: golden loaders and tie-aware comparisons."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

from kir_graph_b200.hisat2 import PairRead
from kir_graph_b200.msa2hisat import Variant

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str) -> dict:
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        return json.load(f)


def golden_names(kind: str) -> list[str]:
    out = []
    for fn in sorted(os.listdir(GOLDEN)):
        if fn.endswith(".json.gz"):
            name = fn[: -len(".json.gz")]
            if load_golden(name).get("kind") == kind:
                out.append(name)
    return out


def objects_from_input(inp: dict) -> tuple[list[PairRead], list[Variant]]:
    return [PairRead(**r) for r in inp["reads"]], [Variant(**v) for v in inp["variants"]]


def counts_from_log_probs(log_probs: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Invert log_probs = (K - m) c1 + m c0 to the integer mismatch counts."""
    from oracle.typing_oracle import C_HIT, C_MISS
    m = (log_probs - k[:, None] * C_HIT) / (C_MISS - C_HIT)
    mi = np.rint(m).astype(np.int64)
    assert np.abs(m - mi).max() < 1e-6
    return mi


def int_scores_from_values(values, k_total: int) -> np.ndarray:
    from oracle.typing_oracle import C_HIT, C_MISS
    s = (np.asarray(values, dtype=float) - k_total * C_HIT) / (C_MISS - C_HIT)
    si = np.rint(s).astype(np.int64)
    assert np.abs(s - si).max() < 1e-6, "reference value is not an integer mismatch score"
    return si


def assert_same_modulo_ties(ref_ids, ref_scores, our_ids, our_scores, kept_all: bool):
    """Kept sets agree except inside an exact-score tie group that straddles the cut.

    ``kept_all``: nothing was cut (fewer candidates than top_n), so every group must match.
    """
    ref_scores = np.asarray(ref_scores)
    our_scores = np.asarray(our_scores)
    assert len(ref_scores) == len(our_scores)
    assert np.array_equal(np.sort(ref_scores), np.sort(our_scores)), "score multisets differ"
    if not len(ref_scores):
        return
    last = ref_scores.max()
    for s in np.unique(ref_scores):
        a = {tuple(sorted(x)) for x in np.asarray(ref_ids)[ref_scores == s].tolist()}
        b = {tuple(sorted(x)) for x in np.asarray(our_ids)[our_scores == s].tolist()}
        if s == last and not kept_all:
            assert len(a) == len(b)
        else:
            assert a == b, f"kept sets differ at score {s}"
