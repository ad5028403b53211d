"""
Drop-in replacement of the reference's ``graphkir/typing_mulit_allele.py``.

Same classes, constructor arguments, attributes and error behaviour
(``TypingResult``, ``AlleleTyping``, ``AlleleTypingExonFirst``, ``rankScore``,
``isHomozygous`` ...; reference lines are cited per member), but the arithmetic
runs on the GPU through :mod:`kir_graph_b200.engine`:

* the likelihood matrix is built by ``gk_likelihood`` (popcount over packed
  variant bitsets) and kept on the device as integer mismatch counts
  ``m[r, a]``; ``probs`` / ``log_probs`` are unit conversions of it,
* ``addCandidate`` is one ``SearchGroup.step``: scoring, dedup, cut, rescoring,
  ranking all happen in CUDA kernels on exact integers.

Tie policy: the reference orders exactly-tied
candidates by float rounding noise and an unstable ``argsort``; here ties are
exact and broken by the documented secondary keys, then by candidate order.
Whenever a tie group touches a cut or the best rank, or read-level ties leave the
outcome of ``selectBest``'s fraction test open for a rank it examines,
``TypingResult.tie_flags`` is set and ``AlleleTyping.tie_report`` records it.
"""
from __future__ import annotations

import copy
import io
import math
from collections import defaultdict
from dataclasses import dataclass, field
from itertools import chain
from typing import Iterable, Optional

import numpy as np

from . import engine
from .packing import GenePack, csr_from_reads, error_correction_masks, pack_gene, site_tallies
from .utils import logger

C_HIT = float(np.log10(0.999))
C_MISS = float(np.log10(0.001))


INTRON_SEARCH_BUDGET_BYTES = 4 << 30     # device memory the tied exon candidates of exon-first may hold at once
TIE_FRACTION_NEAR_THRESHOLD = 8      # tie_flags bit3 (set by gk_rank as well; see GkStepInfo.tie_flags)


def _lcm_upto(n: int) -> int:
    out = 1
    for i in range(2, n + 1):
        out = out * i // math.gcd(out, i)
    return out


class LazyAlleleProb:
    """``allele_prob`` (reads x kept sets, log10) materialised from the GPU on first use."""

    def __init__(self, loader, n_rows: int, cols: np.ndarray):
        self._loader = loader            # () -> float64 [R, K_all]
        self._n_rows = n_rows
        self._cols = np.asarray(cols, dtype=np.int64)
        self._cache: np.ndarray | None = None

    @property
    def shape(self) -> tuple[int, int]:
        return (self._n_rows, len(self._cols))

    def __len__(self) -> int:
        return self._n_rows

    def select(self, cols) -> "LazyAlleleProb":
        return LazyAlleleProb(self._loader, self._n_rows, self._cols[np.asarray(cols, dtype=np.int64)])

    def materialize(self) -> np.ndarray:
        if self._cache is None:
            self._cache = np.ascontiguousarray(self._loader()[:, self._cols])
        return self._cache

    def __array__(self, dtype=None, copy=None):
        out = self.materialize()
        return out.astype(dtype) if dtype is not None else out

    def __getitem__(self, key):
        return self.materialize()[key]

    @property
    def T(self) -> np.ndarray:
        return self.materialize().T

    def tolist(self):
        return self.materialize().tolist()


@dataclass
class TypingResult:
    """Result of one CN step (reference: typing_mulit_allele.py:27-58)."""

    n: int
    value: np.ndarray               # top_n            log10 likelihood of each allele set
    value_sum_indv: np.ndarray      # top_n x n        CN=1 value of each member
    allele_id: np.ndarray           # top_n x n
    allele_name: list               # top_n x n
    allele_prob: object             # reads x top_n    (ndarray or LazyAlleleProb)
    fraction: np.ndarray            # top_n x n
    fraction_uniq: np.ndarray       # top_n x n
    allele_name_group: list = field(default_factory=list)
    # exact integer side (not in the reference): mismatch score, member column sums,
    # fraction numerators (fraction * reads * lcm(1..n)); tie_flags as in GkStepInfo
    score: np.ndarray | None = None
    member_colsum: np.ndarray | None = None
    frac_num: np.ndarray | None = None
    tie_flags: int = 0
    n_unique: int = 0
    p_colsum: np.ndarray | None = None   # sum over reads of allele_prob in mismatch counts (search state)

    def isFail(self) -> bool:
        return not len(self.value)

    def selectBest(self, filter_fraction: bool = True, filter_minor: bool = False) -> list[str]:
        """First rank whose every member has abundance >= 0.5 / CN, else rank 0 (:63-103)."""
        ranks: Iterable[int] = range(len(self.fraction))
        if filter_fraction:
            floor = (1 / self.n) / 2
            ranks = [i for i in ranks if all(f >= floor for f in self.fraction[i])]
        if filter_minor:
            ranks = [i for i in ranks
                     if np.abs(self.value_sum_indv[i]).min() / np.abs(self.value_sum_indv[i]).max() > 0.8]
        ranks = list(ranks) or [0]
        if self.isFail():
            logger.warning("[Allele] No candidates found. Return fail")
            return ["fail"] * self.n
        best = ranks[0]
        logger.debug(f"[Allele] Select best rank: {best}")
        assert len(self.allele_name[best]) == self.n
        return self.allele_name[best]

    def print(self, num: int = 100, top_threshold: float = 0.9) -> None:
        """Log the leading ranks at DEBUG level (:105-148)."""
        if not logger.isEnabledFor(10):
            return
        out = io.StringIO()
        print("Allele_num = ", self.n, file=out)
        if self.isFail():
            print("Fail Alleles:", ["fail"] * self.n)
            return
        for shown, rank in enumerate(self.topRank(top_threshold)):
            if shown > num:
                break
            print("Rank", rank, "probility", self.value[rank], "sum", self.value_sum_indv[rank].sum(), file=out)
            for i in range(self.n):
                print("  id", f"{self.allele_id[rank][i]:3}", "  name", f"{self.allele_name[rank][i]:20s}",
                      "  fraction", f"{self.fraction[rank][i]:.5f}",
                      "  sum", f"{self.value_sum_indv[rank][i]:8.3f}", end=" ", file=out)
                if self.allele_name_group:
                    print("  group", f"{self.allele_name_group[rank][i]}", end=" ", file=out)
                print(file=out)
        logger.debug(f"[Allele] {out.getvalue()}")

    def setNameGroup(self, allele_group_mapping: dict[str, list[str]]) -> None:
        self.allele_name_group = [[allele_group_mapping[j] for j in i] for i in self.allele_name]

    def _rank_index(self) -> list[int]:
        if self.score is not None and self.member_colsum is not None and self.frac_num is not None:
            n = self.n
            even = int(self.frac_num[0].sum()) // n if len(self.frac_num) else 0
            uneven = np.abs(self.frac_num - even).sum(axis=1)
            keys = list(zip(self.score.tolist(), self.member_colsum.sum(axis=1).tolist(), uneven.tolist()))
            return sorted(range(len(keys)), key=keys.__getitem__)
        return rankScore(self.value, self.value_sum_indv, self.fraction)

    def sortByScoreAndEveness(self, preserve_topn: int = -1) -> "TypingResult":
        """Stable re-rank on (score, member sums, evenness); keep ``preserve_topn`` (:156-171)."""
        if preserve_topn == -1:
            preserve_topn = self.value.shape[0]
        order = self._rank_index()[:preserve_topn]
        take = lambda x: None if x is None else x[order]
        prob = self.allele_prob
        prob = prob.select(order) if isinstance(prob, LazyAlleleProb) else prob[:, order]
        return TypingResult(
            n=self.n, value=self.value[order], value_sum_indv=self.value_sum_indv[order],
            allele_id=self.allele_id[order], allele_name=[self.allele_name[i] for i in order],
            allele_prob=prob, fraction=self.fraction[order], fraction_uniq=self.fraction_uniq[order],
            score=take(self.score), member_colsum=take(self.member_colsum), frac_num=take(self.frac_num),
            tie_flags=self.tie_flags, n_unique=self.n_unique, p_colsum=take(self.p_colsum))

    def topRank(self, threshold: float = 0.9) -> Iterable[int]:
        """Rank 0 and every rank with value * threshold >= best value (:173-184)."""
        assert not self.isFail()
        yield 0
        best = self.value[0]
        for i, v in enumerate(self.value):
            if i and v * threshold >= best:
                yield i

    def selectAllPossible(self, threshold: float = 0.9) -> list[tuple[float, list[str]]]:
        if self.isFail():
            return []
        return [(self.value[rank], self.allele_name[rank]) for rank in self.topRank(threshold)]


def argSortRow(data: np.ndarray) -> list[int]:
    """Stable argsort of rows as tuples (:197-199)."""
    return sorted(range(len(data)), key=lambda i: tuple(data[i]))


def rankScore(value: np.ndarray, value_sum_indv: np.ndarray, fraction: np.ndarray) -> list[int]:
    """Order by likelihood, then per-allele likelihood sum, then evenness (:202-214)."""
    uneven = np.abs(fraction - fraction.mean(axis=1, keepdims=True)).sum(axis=1)
    return argSortRow(np.array([-value, -value_sum_indv.sum(axis=1), uneven]).T)


def _empty_result(n: int) -> TypingResult:
    e = lambda: np.array([])
    return TypingResult(n=n, value=e(), value_sum_indv=e(), allele_id=e(), allele_name=[],
                        allele_prob=e(), fraction=e(), fraction_uniq=e())


class AlleleTyping:
    """Multi-allele typing of one gene (reference: typing_mulit_allele.py:217-619)."""

    def __init__(self, reads, variants, force_homo: bool | None = None, top_n: int = 300,
                 no_empty: bool = True, variant_correction: bool = True, *, _backend=None,
                 _pack: GenePack | None = None):
        self.top_n = top_n
        self._no_empty = no_empty
        self.force_homo: bool | None = force_homo
        variants = list(variants)
        self.variants = {str(v.id): v for v in variants}                           # (:253)
        pack = _pack if _pack is not None else pack_gene(
            reads, variants, variant_correction=variant_correction, no_empty=no_empty, mutate_reads=True)
        self._pack = pack
        self.id_to_allele: dict[int, str] = dict(enumerate(pack.allele_names))      # (:254)
        self.allele_to_id: dict[str, int] = {j: i for i, j in self.id_to_allele.items()}
        self.reads = [reads[i] for i in pack.kept_reads] if reads is not None else []
        self._reads_given = reads is not None
        self.result: list[TypingResult] = []
        self.tie_report: list[dict] = []
        # K_r as used by the likelihood: an empty read kept by no_empty=False counts one 0.999 (:372-374)
        self._k_eff = np.where(pack.k_obs == 0, 1, pack.k_obs).astype(np.int64)
        self._k_total = int(self._k_eff.sum())
        self._backend = _backend
        self._batch: engine.MatrixBatch | None = None
        self._search: engine.SearchGroup | None = None
        self._search_len = -1
        self._m_host: np.ndarray | None = None
        if pack.n_reads:
            self._batch = engine.MatrixBatch([pack], backend=_backend)
        else:
            logger.warning("[Allele] Error: Empty reads for typing (or Maybe read depth is too low)")

    # --- likelihood views -----------------------------------------------------------
    def mismatch_counts(self) -> np.ndarray:
        """m[r, a] (uint8) as computed by the likelihood kernel."""
        if self._m_host is None:
            self._m_host = (self._batch.mismatch_counts(0) if self._batch is not None
                            else np.zeros((0, len(self.id_to_allele)), np.uint8))
        return self._m_host

    @property
    def log_probs(self) -> np.ndarray:
        """log10 of ``probs`` (:263): (K_r - m) log10(.999) + m log10(.001)."""
        if self._batch is None:
            return np.array([])
        m = self.mismatch_counts().astype(np.float64)
        return (self._k_eff[:, None] - m) * C_HIT + m * C_MISS

    @property
    def probs(self) -> np.ndarray:
        if self._batch is None:
            return np.array([])
        return np.power(10.0, self.log_probs)

    def getReadsNum(self) -> int:
        return self._pack.n_reads

    def group_pattern(self, allele_ids) -> np.ndarray:
        """uint32 per read: bit t set where allele ``allele_ids[t]`` attains the row maximum of
        ``probs[:, allele_ids]`` (novel_discover.py:62-64), from the device-resident likelihood."""
        if self._batch is None:
            return np.zeros(0, dtype=np.uint32)
        return self._batch.group_pattern(0, allele_ids)

    # --- reference static helpers ------------------------------------------------------
    @staticmethod
    def removeEmptyReads(reads):
        return [read for read in reads if read.lpv + read.lnv + read.rpv + read.rnv]

    @staticmethod
    def collectAlleleNames(variants) -> set[str]:
        return set(chain.from_iterable(v.allele for v in variants))

    @staticmethod
    def uniqueAllele(data: np.ndarray) -> np.ndarray:
        """Host utility with the reference's contract (:456-476); the search itself dedups on the GPU."""
        seen, mask = set(), []
        for ids in data:
            key = tuple(sorted(ids))
            mask.append(key not in seen)
            seen.add(key)
        return np.array(mask)

    def mapAlleleIDs(self, list_ids) -> list[list[str]]:
        return [[self.id_to_allele[int(i)] for i in ids] for ids in list_ids]

    # The constructor does the work of the four methods below in packed form (packing.pack_gene +
    # the likelihood kernel); they are kept under the reference's names for callers that use them
    # on their own.
    def read2Onehot(self, variant) -> np.ndarray:
        """Alleles carrying ``variant`` as a boolean vector over the allele columns (:287-292)."""
        onehot = np.zeros(len(self.allele_to_id), dtype=bool)
        for allele in variant.allele:
            onehot[self.allele_to_id[allele]] = True
        return onehot

    @staticmethod
    def onehot2Prob(onehot: np.ndarray) -> np.ndarray:
        """0.999 where set, 0.001 elsewhere (:294-300)."""
        prob = np.ones(onehot.shape) * 0.001
        prob[onehot] = 0.999
        return prob

    def errorCorrection(self, reads):
        """Drop shallow (< 3 observations) variants and minor (< 0.2) polarities from the reads'
        id lists, in place (:302-338); returns ``reads``."""
        ids = list(self.variants)
        csr = csr_from_reads(reads, {vid: i for i, vid in enumerate(ids)})
        drop_pos, drop_neg = error_correction_masks(csr, len(ids))
        bad_pos = {ids[i] for i in np.flatnonzero(drop_pos)}
        bad_neg = {ids[i] for i in np.flatnonzero(drop_neg)}
        for read in reads:
            read.lpv = [v for v in read.lpv if v not in bad_pos]
            read.rpv = [v for v in read.rpv if v not in bad_pos]
            read.lnv = [v for v in read.lnv if v not in bad_neg]
            read.rnv = [v for v in read.rnv if v not in bad_neg]
        return reads

    def reads2AlleleProb(self, reads) -> np.ndarray:
        """probs[r, a] of arbitrary reads of this gene (:340-381) through the likelihood kernel:
        0.999^(K_r - m) * 0.001^m.  As in the reference: no reads -> warning and ``np.array([])``;
        a read without observations is a row of 0.999 when ``no_empty`` is off and a ValueError
        (``np.stack`` of nothing) when it is on."""
        if not reads:
            logger.warning("[Allele] Error: Empty reads for typing (or Maybe read depth is too low)")
            return np.array([])
        pack = pack_gene(reads, list(self.variants.values()), variant_correction=False, no_empty=False,
                         mutate_reads=False)
        if self._no_empty and bool((pack.k_obs == 0).any()):
            raise ValueError("need at least one array to stack")
        m = engine.MatrixBatch([pack], backend=self._backend).mismatch_counts(0).astype(np.float64)
        k_eff = np.where(pack.k_obs == 0, 1, pack.k_obs).astype(np.float64)
        return np.power(10.0, (k_eff[:, None] - m) * C_HIT + m * C_MISS)

    # --- typing ----------------------------------------------------------------------
    def typing(self, cn: int) -> TypingResult:
        """Top-n allele sets of size ``cn`` (:383-410)."""
        if cn < 1:
            raise ValueError(f"CN should be >= 1, got {cn}")
        if self.force_homo is not None:
            homo = self.force_homo
        elif self._reads_given:
            homo = isHomozygous(self.reads, self.variants, cn)
        else:          # built from a pack only (fast .variant.json path): the same tally from the packed lists
            homo = cn > 1 and _no_hetero_site(site_tallies(self._pack), cn)
        self.result = []
        if homo:
            self.addCandidate()
            self.addHomoResultForCn(cn)
        else:
            for _ in range(cn):
                self.addCandidate()
        self.result[-1].print()
        return self.result[-1]

    def addHomoResultForCn(self, cn: int) -> None:
        if cn > 1:
            self.result.append(self.createHomoResult(self.result[0], cn))

    @staticmethod
    def createHomoResult(cn1_result: TypingResult, cn: int) -> TypingResult:
        """CN copies of every CN=1 allele; likelihood scaled by CN (:423-454)."""
        if cn <= 1:
            raise ValueError(f"CN should be > 1, got {cn}")
        k = len(cn1_result.value)
        rep = lambda x: None if x is None else np.repeat(x, cn, axis=1)
        return TypingResult(
            n=cn, value=cn1_result.value * cn, value_sum_indv=np.repeat(cn1_result.value_sum_indv, cn, axis=1),
            allele_id=np.repeat(cn1_result.allele_id, cn, axis=1),
            allele_name=[[name[0]] * cn for name in cn1_result.allele_name],
            allele_prob=cn1_result.allele_prob, fraction=np.ones((k, cn)) / cn,
            fraction_uniq=np.ones((k, cn)) / cn,
            score=None if cn1_result.score is None else cn1_result.score * cn,
            member_colsum=rep(cn1_result.member_colsum), tie_flags=cn1_result.tie_flags,
            p_colsum=cn1_result.p_colsum)

    def _ensure_search(self) -> engine.SearchGroup:
        """Device search state is a cache of ``self.result``; rebuild it after deepcopy / reset."""
        if self._search is None or self._search_len != len(self.result) or self._search.n != len(self.result):
            self._search = engine.SearchGroup(self._batch, [0], self.top_n)
            if self.result:
                last = self.result[-1]
                self._search.restore(0, np.asarray(last.allele_id, dtype=np.int32), last.p_colsum)
        return self._search

    def _to_result(self, out: engine.StepOutput, search: engine.SearchGroup, s: int = 0) -> TypingResult:
        return step_to_result(out, self._batch.colsum(0), self._k_total, self._k_eff, self.id_to_allele,
                              lambda ids=out.ids: search.materialize_p(s, ids))

    def addCandidate(self, candidate_allele: Optional[list[str]] = None) -> TypingResult:
        """One greedy step: grow every kept allele set by one allele (:478-598)."""
        if self._batch is None:
            logger.warning("[Allele] Empty reads for typing. Skip")
            self.result.append(_empty_result(len(self.result) + 1))
            return self.result[-1]
        if self.top_n < 1:
            # top_n == 0 (the restricted model of exon-first with top_n < 5, :716): the reference's
            # first step keeps nothing, a further step indexes with the empty float id array (:540)
            if self.result:
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            self.result.append(_empty_result(1))
            return self.result[-1]
        cand = None
        if candidate_allele is not None:
            cand = np.array([self.allele_to_id[a] for a in candidate_allele], dtype=np.int32)
        search = self._ensure_search()
        out = search.step(cands=[cand])[0]
        res = self._to_result(out, search)
        self.result.append(res)
        self._search_len = len(self.result)
        if res.tie_flags:
            self.tie_report.append({"n": res.n, "tie_flags": res.tie_flags})
        return res

    def __deepcopy__(self, memo):
        """Share the immutable device likelihood; copy only the search results."""
        new = copy.copy(self)
        new.result = list(self.result)
        new.tie_report = list(self.tie_report)
        new._search = None
        new._search_len = -1
        memo[id(self)] = new
        return new

    def plot(self, title: str = ""):
        raise NotImplementedError("plotting stays in the reference (plotly is not a dependency here)")


def step_to_result(out: engine.StepOutput, colsum: np.ndarray, k_total: int, k_eff: np.ndarray,
                   id_to_allele: dict[int, str], p_loader) -> TypingResult:
    """Integer step output -> the reference's float64 ``TypingResult`` fields."""
    n, k = out.n, len(out.score)
    n_reads = len(k_eff)
    member_colsum = colsum[out.ids] if k else np.zeros((0, n), np.int64)
    weights = np.array([_lcm_upto(n) // q for q in range(1, n + 1)], dtype=np.int64)
    frac_num = (out.cnt * weights[None, None, :]).sum(axis=2)
    fraction = frac_num / float(n_reads * _lcm_upto(n))

    def load_log10():
        p = p_loader().astype(np.float64)
        return (k_eff[:, None] - p) * C_HIT + p * C_MISS

    # bit3: the outcome of selectBest's test "every member fraction >= 1 / (2n)" (:83-87) is not
    # decided by exact arithmetic alone for a rank it looks at.  A read tied between q members counts
    # 1/q for each here; in the reference the tied log-probabilities can differ in the last bit
    # (ordered float product), and then the read counts 1 for one member and 0 for the
    # others (:575-580).  Member t's fraction in the reference therefore lies anywhere between
    # (reads t wins alone) / R and (reads t wins or ties) / R; if the threshold falls inside that
    # range the reference may pass or fail the rank differently.
    tie_flags = int(out.tie_flags)
    if len(fraction):
        floor = 0.5 / n
        passing = np.flatnonzero((fraction >= floor).all(axis=1))
        last = int(passing[0]) + 1 if len(passing) else len(fraction)
        alone = out.cnt[:last, :, 0] / float(n_reads)
        at_most = out.cnt[:last].sum(axis=2) / float(n_reads)
        if bool(((alone < floor) & (at_most >= floor)).any()):
            tie_flags |= TIE_FRACTION_NEAR_THRESHOLD

    return TypingResult(
        n=n,
        value=k_total * C_HIT + out.score.astype(np.float64) * (C_MISS - C_HIT),
        value_sum_indv=k_total * C_HIT + member_colsum.astype(np.float64) * (C_MISS - C_HIT),
        allele_id=out.ids.astype(np.int64),
        allele_name=[[id_to_allele[int(i)] for i in ids] for ids in out.ids],
        allele_prob=LazyAlleleProb(load_log10, n_reads, np.arange(k)),
        fraction=fraction,
        fraction_uniq=np.ones(fraction.shape),                                      # "fake" in the reference (:584)
        score=out.score, member_colsum=member_colsum, frac_num=frac_num,
        tie_flags=tie_flags, n_unique=out.n_unique, p_colsum=out.score.copy())


class AlleleTypingExonFirst(AlleleTyping):
    """Type exon variants over allele groups first, then full variants per tied exon
    candidate (reference: typing_mulit_allele.py:622-797)."""

    def __init__(self, reads, variants, top_n: int = 300, exon_only: bool = False,
                 candidate_set_threshold: float = 1.0, variant_correction: bool = True,
                 force_homo: bool | None = None, *, _backend=None):
        variants = list(variants)
        exon_variants = [v for v in variants if v.in_exon]
        exon_reads = self.removeIntronVariant(reads, exon_variants)
        # the reference corrects the exon reads here and again inside the base constructor (:644-645, :664)
        if variant_correction:
            pack_gene(exon_reads, exon_variants, variant_correction=True, no_empty=False, mutate_reads=True)
        exon_reads = self.removeEmptyReads(exon_reads)

        variantset_to_allele = self.aggrVariantsByAllele(exon_variants)
        other = self.collectAlleleNames(variants) - self.collectAlleleNames(exon_variants)
        if other:
            variantset_to_allele[tuple()] = sorted(other)
        self.allele_group = {"|".join(alleles): alleles for alleles in variantset_to_allele.values()}
        exon_variants = self.removeDuplicateAllele(variants, self.createInverseMapping(self.allele_group))

        super().__init__(exon_reads, exon_variants, force_homo=force_homo, top_n=top_n, _backend=_backend)
        self.candidate_set_threshold = candidate_set_threshold
        self.full_model: AlleleTyping | None = None
        if not exon_only:
            self.full_model = AlleleTyping(reads, variants, force_homo=force_homo, top_n=top_n // 5,
                                           variant_correction=variant_correction, _backend=_backend)

    @staticmethod
    def aggrVariantsByAllele(variants) -> dict[tuple[str, ...], list[str]]:
        """variant -> alleles turned into (set of variant ids) -> alleles (:689-700)."""
        per_allele = defaultdict(list)
        for variant in variants:
            for allele in variant.allele:
                per_allele[allele].append(str(variant.id))
        by_set = defaultdict(list)
        for allele, vids in per_allele.items():
            by_set[tuple(sorted(set(vids)))].append(allele)
        return by_set

    @staticmethod
    def removeIntronVariant(reads, exon_variants):
        """Deep-copied reads whose lists only keep exon variant ids (:702-714)."""
        exon_ids = {v.id for v in exon_variants}
        new_reads = copy.deepcopy(reads)
        for read in new_reads:
            read.lpv = [v for v in read.lpv if v in exon_ids]
            read.lnv = [v for v in read.lnv if v in exon_ids]
            read.rpv = [v for v in read.rpv if v in exon_ids]
            read.rnv = [v for v in read.rnv if v in exon_ids]
        return new_reads

    @staticmethod
    def createInverseMapping(allele_group: dict[str, list[str]]) -> dict[str, str]:
        return {allele: group for group, alleles in allele_group.items() for allele in alleles}

    @staticmethod
    def removeDuplicateAllele(variants, allele_map: dict[str, str]):
        """Rewrite ``variant.allele`` to group names (:725-738)."""
        variants = copy.deepcopy(variants)
        for variant in variants:
            variant.allele = list(set(filter(None, [allele_map.get(v, "") for v in variant.allele])))
        return variants

    def typingIntron(self, exon_candidates: list[list[str]]) -> AlleleTyping:
        """Sequential form kept for API parity (:740-746); ``typing`` batches these on the GPU."""
        assert self.full_model
        model = copy.deepcopy(self.full_model)
        for cand in exon_candidates:
            model.addCandidate(cand)
        return model

    def typing(self, cn: int) -> TypingResult:
        result = super().typing(cn)
        result.setNameGroup(self.allele_group)
        logger.debug("[Allele] Typing exon:")
        result.print()
        if self.full_model is None:
            return result
        assert cn == result.n
        if not result.value.shape[0]:
            logger.warning("[Allele] Cannot typing with exon-only reads. Typing with exon+intron")
            return self.full_model.typing(cn)

        ranks = list(result.topRank(threshold=self.candidate_set_threshold))
        if self.full_model.top_n < 1:
            # top_n < 5 leaves the restricted model with top_n // 5 == 0 kept sets (:716): the reference's
            # first step then keeps nothing (the merged result is empty: "fail") and a second step
            # indexes with an empty float array
            if cn > 1:
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            self.result.extend(_empty_result(1) for _ in ranks)
            self.result.append(_empty_result(1))
            return self.result[-1]
        candidate_result = self._typing_intron_batched([result.allele_name_group[i] for i in ranks])
        logger.debug(f"[Allele] Intron Candidate {len(candidate_result)} Done")
        cat = lambda xs: None if any(x is None for x in xs) else np.concatenate(xs)
        merged = TypingResult(
            n=candidate_result[0].n,
            value=np.concatenate([r.value for r in candidate_result]),
            value_sum_indv=np.concatenate([r.value_sum_indv for r in candidate_result]),
            allele_id=np.concatenate([r.allele_id for r in candidate_result]),
            allele_name=list(chain.from_iterable(r.allele_name for r in candidate_result)),
            allele_prob=_ConcatProb([r.allele_prob for r in candidate_result]),
            fraction=np.concatenate([r.fraction for r in candidate_result]),
            fraction_uniq=np.concatenate([r.fraction for r in candidate_result]),      # (:791)
            score=cat([r.score for r in candidate_result]),
            member_colsum=cat([r.member_colsum for r in candidate_result]),
            frac_num=cat([r.frac_num for r in candidate_result]),
            tie_flags=int(np.bitwise_or.reduce([r.tie_flags for r in candidate_result])) | result.tie_flags)
        merged = merged.sortByScoreAndEveness()
        self.result.append(merged)
        logger.debug("[Allele] Typing intron + exon")
        merged.print()
        return merged

    def _typing_intron_batched(self, candidates: list[list[list[str]]]) -> list[TypingResult]:
        """All tied exon candidates advance together: one search per candidate over the
        shared full-variant likelihood (replaces the per-candidate deepcopy loop, :774-779)."""
        model = self.full_model
        n_search = len(candidates)
        cn = len(candidates[0])
        if model._batch is None:
            out = []
            for _ in candidates:
                steps = [_empty_result(i + 1) for i in range(cn)]
                self.result.extend(steps)
                out.append(steps[-1])
            return out
        # Every search holds its own P (reads x kept sets) and score buffers, so the tied exon candidates are
        # typed in chunks that fit a device-memory budget (the reference handles them one at a time, :774-779;
        # exact integer ties can make them many)
        r_pad = int(model._batch.table["r_pad"][0])
        n_kblk = -(-model.top_n // engine.GK_KB)
        per_search_bytes = r_pad * n_kblk * engine.GK_KB * 2 + n_kblk * engine.GK_KB * int(
            model._batch.table["n_ablk"][0]) * 32 * 4 + model.top_n * max(len(model.id_to_allele), 1) * 8
        chunk = int(max(1, min(n_search, INTRON_SEARCH_BUDGET_BYTES // max(per_search_bytes, 1))))
        per_search: list[list[TypingResult]] = [[] for _ in candidates]
        for lo in range(0, n_search, chunk):
            hi = min(lo + chunk, n_search)
            group = engine.SearchGroup(model._batch, [0] * (hi - lo), model.top_n)
            for step in range(cn):
                cands = [np.array([model.allele_to_id[a] for a in candidates[s][step]], dtype=np.int32)
                         for s in range(lo, hi)]
                outs = group.step(cands=cands, need_next=np.full(hi - lo, step + 1 < cn))
                for s in range(lo, hi):
                    res = step_to_result(outs[s - lo], model._batch.colsum(0), model._k_total, model._k_eff,
                                         model.id_to_allele,
                                         lambda g=group, j=s - lo, ids=outs[s - lo].ids: g.materialize_p(j, ids))
                    per_search[s].append(res)
        finals = []
        for steps in per_search:
            self.result.extend(steps)
            finals.append(steps[-1])
        return finals


class _ConcatProb(LazyAlleleProb):
    """Lazy column-wise concatenation of several ``allele_prob`` blocks."""

    def __init__(self, parts, cols=None):
        self._parts = parts
        n_rows = parts[0].shape[0] if parts else 0
        total = sum(p.shape[1] for p in parts)
        super().__init__(self._load, n_rows, np.arange(total) if cols is None else cols)

    def _load(self) -> np.ndarray:
        return np.concatenate([np.asarray(p) for p in self._parts], axis=1)

    def select(self, cols) -> "LazyAlleleProb":
        return _ConcatProb(self._parts, self._cols[np.asarray(cols, dtype=np.int64)])


def isHetrozygous(gene: str) -> bool:
    """Genes typed as heterozygous by name alone (:800-804)."""
    return "2DL1S1" in gene or "2DL5" in gene


def isHomozygous(reads, variants_map: dict, cn: int) -> bool:
    """No position shows convincing bi-allelic support (:807-857)."""
    if cn <= 1:
        return False
    tally: dict = defaultdict(lambda: defaultdict(int))
    for read in reads:
        for vid in chain(read.lpv, read.rpv):
            v = variants_map[vid]
            if v.typ != "deletion":
                tally[v.pos][str(v.val)] += 1
        for vid in chain(read.lnv, read.rnv):
            v = variants_map[vid]
            if v.typ != "deletion":
                tally[v.pos][f"*{v.val}"] += 1
    return _no_hetero_site(tally.values(), cn)


def _no_hetero_site(sites, cn: int) -> bool:
    hits = 0
    for site in sites:
        if len(site) <= 1 or all("*" in key for key in site):
            continue
        counts = sorted((c for c in site.values() if c > 3), reverse=True)
        depth = sum(counts)
        if depth < 20:
            continue
        shares = [c / depth for c in counts if c / depth > 0.1]
        if len(shares) == 1:
            continue
        if shares[1] > 1 / (cn * 2):      # IndexError for an empty list, as in the reference
            hits += 1
    return hits == 0
