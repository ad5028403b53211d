"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``kir_graph_b200``.

A CPU (NumPy) restatement of Graph-KIR's allele-typing core, written from the
algorithm description of the reference, with the reference location each
function follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.

Two flavours of the greedy search are provided:

``F64Search``   the reference's float64 expressions, operation for operation
                (``np.maximum`` broadcast + ``sum``, python-set dedup, ``argsort``,
                gathers, ``np.equal`` fractions, 3-key python sort), optionally
                chunked over reads so that it fits in RAM.  This is what the CPU
                baseline times, and what is pinned against the imported reference
                (tests/golden/make_golden.py).
``IntSearch``   the same search on exact integers: ``log_probs[r, a]`` is a
                function of K_r (observations in read r) and m[r, a] (how many
                disagree with allele a), so every score is an integer and ties are
                exact.  It defines the deterministic total order that the CUDA
                path must reproduce bit for bit:
                  cut to M = max(top_n, N_uniq // 5) by (S, flat index),
                  rank by (S, sum of member column sums, fraction unevenness, flat index).

Parity pin: ``tests/golden/*.json`` hold inputs and outputs of the reference
itself (imported from /root/reference with plotly/Bio/pyhlamsa stubbed);
``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

import math
from collections import Counter, defaultdict
from dataclasses import dataclass, field
from itertools import chain

import numpy as np

P_HIT, P_MISS = 0.999, 0.001
C_HIT = float(np.log10(P_HIT))      # log10 of an agreeing observation
C_MISS = float(np.log10(P_MISS))    # log10 of a disagreeing observation
LISTS = ("lpv", "rpv", "lnv", "rnv")  # multiplication order, typing_mulit_allele.py:363-368


# ---------------------------------------------------------------------------
# a7-a9: allele universe, variant error correction, empty reads
# ---------------------------------------------------------------------------
def collect_allele_names(variants) -> list[str]:
    """Sorted union of ``variant.allele`` (typing_mulit_allele.py:254, :283-285)."""
    return sorted(set(chain.from_iterable(v.allele for v in variants)))


def error_correction(reads):
    """typing_mulit_allele.py:302-338.  Mutates ``reads`` in place and returns it."""
    seen: dict[str, list[int]] = {}
    for read in reads:
        for vid in read.lpv + read.rpv:
            seen.setdefault(vid, [0, 0])[0] += 1
        for vid in read.lnv + read.rnv:
            seen.setdefault(vid, [0, 0])[1] += 1
    drop_pos, drop_neg = set(), set()
    for vid, (n_pos, n_neg) in seen.items():
        total = n_pos + n_neg
        if total < 3:                       # min depth (:316)
            drop_pos.add(vid)
            drop_neg.add(vid)
            continue
        if n_pos / total < 0.2:             # minor polarity (:319-325)
            drop_pos.add(vid)
        if n_neg / total < 0.2:
            drop_neg.add(vid)
    for read in reads:
        read.lpv = [v for v in read.lpv if v not in drop_pos]
        read.rpv = [v for v in read.rpv if v not in drop_pos]
        read.lnv = [v for v in read.lnv if v not in drop_neg]
        read.rnv = [v for v in read.rnv if v not in drop_neg]
    return reads


def remove_empty_reads(reads):
    """typing_mulit_allele.py:274-281."""
    return [r for r in reads if (r.lpv or r.lnv or r.rpv or r.rnv)]


# ---------------------------------------------------------------------------
# a10-a11: likelihood
# ---------------------------------------------------------------------------
def mismatch_counts(reads, variants_by_id, allele_to_id) -> tuple[np.ndarray, np.ndarray]:
    """m[r, a] and K_r by direct set logic (no bit tricks).

    A positive observation disagrees with the alleles that lack the variant, a
    negative one with the alleles that carry it (typing_mulit_allele.py:287-300,
    :363-368); duplicates count every time they are listed.
    """
    n_allele = len(allele_to_id)
    m = np.zeros((len(reads), n_allele), dtype=np.int64)
    k = np.zeros(len(reads), dtype=np.int64)
    carriers_cache: dict[str, np.ndarray] = {}

    def carriers(vid: str) -> np.ndarray:
        if vid not in carriers_cache:
            onehot = np.zeros(n_allele, dtype=bool)
            for name in variants_by_id[vid].allele:
                onehot[allele_to_id[name]] = True
            carriers_cache[vid] = onehot
        return carriers_cache[vid]

    for r, read in enumerate(reads):
        for vid in chain(read.lpv, read.rpv):
            m[r] += ~carriers(vid)
            k[r] += 1
        for vid in chain(read.lnv, read.rnv):
            m[r] += carriers(vid)
            k[r] += 1
    return m, k


def probs_ordered_product(reads, variants_by_id, allele_to_id, no_empty=True) -> np.ndarray:
    """``probs`` with the reference's rounding: one 0.999/0.001 vector per
    observation in list order lpv, rpv, lnv, rnv, multiplied along axis 0
    (typing_mulit_allele.py:340-381)."""
    n_allele = len(allele_to_id)
    rows = []
    for read in reads:
        factors = []
        for name in LISTS:
            positive = name in ("lpv", "rpv")
            for vid in getattr(read, name):
                onehot = np.zeros(n_allele, dtype=bool)
                for allele in variants_by_id[vid].allele:
                    onehot[allele_to_id[allele]] = True
                hit = onehot if positive else ~onehot
                factors.append(np.where(hit, P_HIT, P_MISS))
        if not factors and not no_empty:
            factors = [np.full(n_allele, P_HIT)]
        rows.append(np.stack(factors).prod(axis=0))
    return np.stack(rows) if rows else np.array([])


def log_probs_from_counts(m: np.ndarray, k: np.ndarray) -> np.ndarray:
    """log10 prob = (K_r - m) log10(.999) + m log10(.001)  (SURVEY.md headline fact 3)."""
    return (k[:, None] - m) * C_HIT + m * C_MISS


def value_from_score(score: np.ndarray, k_total: int) -> np.ndarray:
    """Sum over reads of the best log-probability, from the integer mismatch score."""
    return k_total * C_HIT + np.asarray(score, dtype=np.float64) * (C_MISS - C_HIT)


# ---------------------------------------------------------------------------
# a12: homozygosity heuristic
# ---------------------------------------------------------------------------
def is_homozygous(reads, variants_by_id, cn: int) -> bool:
    """typing_mulit_allele.py:807-857 (an IndexError of the reference for a site
    with no share above 0.1 is reproduced, not hidden)."""
    if cn <= 1:
        return False
    tally: dict[int, Counter] = defaultdict(Counter)
    for read in reads:
        for vid in chain(read.lpv, read.rpv):
            v = variants_by_id[vid]
            if v.typ != "deletion":
                tally[v.pos][str(v.val)] += 1
        for vid in chain(read.lnv, read.rnv):
            v = variants_by_id[vid]
            if v.typ != "deletion":
                tally[v.pos][f"*{v.val}"] += 1
    hetero_sites = 0
    for site in tally.values():
        if len(site) <= 1 or all("*" in key for key in site):
            continue
        counts = sorted((c for c in site.values() if c > 3), reverse=True)
        depth = sum(counts)
        if depth < 20:
            continue
        shares = [c / depth for c in counts if c / depth > 0.1]
        if len(shares) == 1:
            continue
        if shares[1] > 1 / (cn * 2):
            hetero_sites += 1
    return hetero_sites == 0


# ---------------------------------------------------------------------------
# results
# ---------------------------------------------------------------------------
@dataclass
class StepResult:
    """One CN step (fields of the reference's TypingResult, typing_mulit_allele.py:27-58)."""

    n: int
    value: np.ndarray
    value_sum_indv: np.ndarray
    allele_id: np.ndarray
    allele_prob: np.ndarray          # F64Search: log-probs [R, K]; IntSearch: mismatch counts [R, K]
    fraction: np.ndarray
    # integer side (IntSearch only)
    score: np.ndarray | None = None          # int64 [K]
    member_colsum: np.ndarray | None = None  # int64 [K, n]
    frac_num: np.ndarray | None = None       # int64 [K, n], fraction * R * lcm(1..n)
    n_unique: int = 0
    ties: list = field(default_factory=list)


def rank_rows(value, value_sum_indv, fraction) -> list[int]:
    """Stable sort on (-value, -sum(value_sum_indv), sum|frac-mean|)  (:197-214)."""
    uneven = np.abs(fraction - fraction.mean(axis=1, keepdims=True)).sum(axis=1)
    keys = np.array([-value, -value_sum_indv.sum(axis=1), uneven]).T
    return sorted(range(len(keys)), key=lambda i: tuple(keys[i]))


def first_occurrence_mask(id_rows: np.ndarray) -> np.ndarray:
    """True for the first row of every distinct sorted tuple (:456-476)."""
    seen, mask = set(), np.zeros(len(id_rows), dtype=bool)
    for i, row in enumerate(id_rows):
        key = tuple(sorted(row))
        if key not in seen:
            seen.add(key)
            mask[i] = True
    return mask


def select_best(result: StepResult, names: list[str]) -> list[str]:
    """First rank whose every fraction >= 1/(2n), else rank 0; 'fail' when empty (:63-103)."""
    if not len(result.value):
        return ["fail"] * result.n
    floor = (1 / result.n) / 2
    for i in range(len(result.fraction)):
        if all(f >= floor for f in result.fraction[i]):
            return [names[a] for a in result.allele_id[i]]
    return [names[a] for a in result.allele_id[0]]


def top_rank(value: np.ndarray, threshold: float) -> list[int]:
    """Rank 0 plus every rank i with value[i] * threshold >= value[0]  (:173-184)."""
    return [0] + [i for i in range(1, len(value)) if value[i] * threshold >= value[0]]


# ---------------------------------------------------------------------------
# a13-a18, float64 flavour (the reference's NumPy expressions)
# ---------------------------------------------------------------------------
class F64Search:
    """Greedy top-N search in float64, following typing_mulit_allele.py:478-598."""

    def __init__(self, log_probs: np.ndarray, top_n: int = 300, read_chunk: int | None = None):
        self.log_probs = log_probs
        self.top_n = top_n
        self.read_chunk = read_chunk
        self.result: list[StepResult] = []

    def _scores(self, cand: np.ndarray, prev_prob: np.ndarray) -> np.ndarray:
        lp = self.log_probs[:, cand]
        if self.read_chunk is None:
            # literal: (K, R, A) broadcast then reduce over reads (:540-542)
            return np.maximum(lp, prev_prob.T[:, :, None]).sum(axis=1).flatten()
        total = np.zeros((prev_prob.shape[1], len(cand)))
        for lo in range(0, lp.shape[0], self.read_chunk):
            hi = lo + self.read_chunk
            total += np.maximum(lp[lo:hi], prev_prob[lo:hi].T[:, :, None]).sum(axis=1)
        return total.flatten()

    def add_candidate(self, cand=None) -> StepResult:
        lp = self.log_probs
        cand = np.arange(lp.shape[1]) if cand is None else np.asarray(cand)
        if not self.result:                                          # CN = 1 (:512-532)
            col = lp[:, cand].sum(axis=0)
            order = np.argsort(col)[::-1][: self.top_n]
            ids = cand[:, None][order]
            res = StepResult(1, col[order], col[order][:, None], ids, lp[:, ids.flatten()],
                             np.ones(ids.shape))
            self.result.append(res)
            return res
        prev = self.result[-1]
        score = self._scores(cand, prev.allele_prob)                 # (:540-542)
        ids = np.hstack([np.repeat(prev.allele_id, len(cand), axis=0),
                         np.tile(cand, len(prev.allele_id))[:, None]])   # (:551-558)
        uniq = first_occurrence_mask(ids)                            # (:561-563)
        ids, score = ids[uniq], score[uniq]
        keep = max(self.top_n, score.shape[0] // 5)                  # (:567)
        top = np.argsort(score)[::-1][:keep]
        top_ids = ids[top]
        gathered = lp[:, top_ids]                                    # R x M x n
        best = gathered.max(axis=2)                                  # (:569)
        indv = gathered.sum(axis=0)                                  # (:571)
        belong = np.equal(gathered, best[:, :, None])                # (:575-580)
        share = belong / belong.sum(axis=2)[:, :, None]
        frac = share.sum(axis=0) / lp.shape[0]
        order = rank_rows(score[top], indv, frac)[: self.top_n]      # (:587-596)
        res = StepResult(len(self.result) + 1, score[top][order], indv[order], top_ids[order],
                         best[:, order], frac[order], n_unique=int(uniq.sum()))
        self.result.append(res)
        return res


# ---------------------------------------------------------------------------
# a13-a18, exact integer flavour (defines the deterministic order)
# ---------------------------------------------------------------------------
def lcm_upto(n: int) -> int:
    out = 1
    for i in range(2, n + 1):
        out = out * i // math.gcd(out, i)
    return out


class IntSearch:
    """Greedy top-N search on mismatch counts.  ``m`` int [R, A], ``k`` int [R]."""

    def __init__(self, m: np.ndarray, k: np.ndarray, top_n: int = 300, read_chunk: int = 2048):
        self.m = np.ascontiguousarray(m, dtype=np.int32)
        self.k_total = int(np.sum(k))
        self.n_reads = self.m.shape[0]
        self.colsum = self.m.sum(axis=0, dtype=np.int64)
        self.top_n = top_n
        self.read_chunk = read_chunk
        self.result: list[StepResult] = []

    # --- pieces ------------------------------------------------------------
    def scores(self, cand: np.ndarray, prev_min: np.ndarray) -> np.ndarray:
        """S[k, j] = sum_r min(m[r, cand[j]], prev_min[r, k]) as int64 [K, C]."""
        total = np.zeros((prev_min.shape[1], len(cand)), dtype=np.int64)
        sub = self.m[:, cand]
        for lo in range(0, self.n_reads, self.read_chunk):
            hi = lo + self.read_chunk
            total += np.minimum(sub[lo:hi][:, None, :], prev_min[lo:hi][:, :, None]).sum(axis=0)
        return total

    def rescore(self, ids: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Per set: min over members [R, M] and tie-split counts cnt[M, n, q-1]."""
        n = ids.shape[1]
        best = np.empty((self.n_reads, len(ids)), dtype=np.int32)
        cnt = np.zeros((len(ids), n, n), dtype=np.int64)
        for lo in range(0, len(ids), 64):
            block = ids[lo:lo + 64]
            g = self.m[:, block]                                  # R x b x n
            mn = g.min(axis=2)
            best[:, lo:lo + 64] = mn
            belong = g == mn[:, :, None]
            q = belong.sum(axis=2)                                # R x b
            for t in range(1, n + 1):
                cnt[lo:lo + 64, :, t - 1] = (belong & (q == t)[:, :, None]).sum(axis=0)
        return best, cnt

    @staticmethod
    def frac_numerators(cnt: np.ndarray) -> np.ndarray:
        """fraction * R * lcm(1..n) as exact integers, [M, n]."""
        n = cnt.shape[1]
        big = lcm_upto(n)
        weights = np.array([big // q for q in range(1, n + 1)], dtype=np.int64)
        return (cnt * weights[None, None, :]).sum(axis=2)

    def _finish(self, n, ids, score, member_colsum, best, frac_num, n_unique, ties):
        big = lcm_upto(n)
        fraction = frac_num / float(self.n_reads * big)
        res = StepResult(
            n=n, value=value_from_score(score, self.k_total),
            value_sum_indv=value_from_score(member_colsum, self.k_total),
            allele_id=ids, allele_prob=best, fraction=fraction, score=score,
            member_colsum=member_colsum, frac_num=frac_num, n_unique=n_unique, ties=ties)
        self.result.append(res)
        return res

    # --- steps ---------------------------------------------------------------
    def add_candidate(self, cand=None) -> StepResult:
        cand = np.arange(self.m.shape[1]) if cand is None else np.asarray(cand, dtype=np.int64)
        if not self.result:
            col = self.colsum[cand]
            order = np.lexsort((np.arange(len(cand)), col))[: self.top_n]
            ids = cand[order][:, None]
            ties = []
            if len(cand) > self.top_n and col[order[-1]] == np.sort(col)[self.top_n]:
                ties.append(("cut", 1, int(col[order[-1]])))
            return self._finish(1, ids, col[order], col[order][:, None], self.m[:, ids[:, 0]],
                                np.full((len(order), 1), self.n_reads, dtype=np.int64),
                                len(cand), ties)
        prev = self.result[-1]
        n = prev.n + 1
        n_prev, n_cand = len(prev.allele_id), len(cand)
        score = self.scores(cand, prev.allele_prob).reshape(-1)
        ids = np.hstack([np.repeat(prev.allele_id, n_cand, axis=0),
                         np.tile(cand, n_prev)[:, None]])
        flat = np.arange(n_prev * n_cand)
        uniq = first_occurrence_mask(ids)
        ids, score, flat = ids[uniq], score[uniq], flat[uniq]
        n_unique = len(ids)
        keep = max(self.top_n, n_unique // 5)
        sel = np.lexsort((flat, score))[:keep]
        ties = []
        if n_unique > keep and score[sel[-1]] == np.sort(score)[keep]:
            ties.append(("cut", n, int(score[sel[-1]])))
        ids, score, flat = ids[sel], score[sel], flat[sel]
        # only sets that can still reach the final top_n need rescoring:
        # everything strictly better than the top_n-th score plus its tie group
        if len(score) > self.top_n:
            bar = np.sort(score)[self.top_n - 1]
            alive = score <= bar
            ids, score, flat = ids[alive], score[alive], flat[alive]
        best, cnt = self.rescore(ids)
        member_colsum = self.colsum[ids]
        frac_num = self.frac_numerators(cnt)
        even = self.n_reads * lcm_upto(n) // n
        uneven = np.abs(frac_num - even).sum(axis=1)
        order = np.lexsort((flat, uneven, member_colsum.sum(axis=1), score))
        if len(order) > self.top_n:
            a, b = order[self.top_n - 1], order[self.top_n]
            if (score[a], member_colsum[a].sum(), uneven[a]) == (score[b], member_colsum[b].sum(), uneven[b]):
                ties.append(("rank-cut", n, int(score[a])))
        order = order[: self.top_n]
        if len(order) > 1 and score[order[0]] == score[order[1]]:
            ties.append(("best", n, int(score[order[0]])))
        return self._finish(n, ids[order], score[order], member_colsum[order], best[:, order],
                            frac_num[order], n_unique, ties)

    def typing(self, cn: int, homo: bool = False) -> StepResult:
        self.result = []
        if homo:
            self.add_candidate()
            if cn > 1:
                self.result.append(homo_result(self.result[0], cn))
        else:
            for _ in range(cn):
                self.add_candidate()
        return self.result[-1]


def homo_result(first: StepResult, cn: int) -> StepResult:
    """CN copies of each CN=1 allele (typing_mulit_allele.py:423-454)."""
    if cn <= 1:
        raise ValueError(f"CN should be > 1, got {cn}")
    k = len(first.value)
    return StepResult(
        n=cn, value=first.value * cn, value_sum_indv=np.repeat(first.value_sum_indv, cn, axis=1),
        allele_id=np.repeat(first.allele_id, cn, axis=1), allele_prob=first.allele_prob,
        fraction=np.ones((k, cn)) / cn,
        score=None if first.score is None else first.score * cn,
        member_colsum=None if first.member_colsum is None else np.repeat(first.member_colsum, cn, axis=1))


# ---------------------------------------------------------------------------
# a20: exon-first grouping helpers
# ---------------------------------------------------------------------------
def exon_allele_groups(variants) -> dict[str, list[str]]:
    """Alleles with identical exon-variant sets form a group named by '|'.join(members);
    alleles without any exon variant form one extra group (:649-656, :689-700)."""
    exon = [v for v in variants if v.in_exon]
    per_allele: dict[str, list[str]] = defaultdict(list)
    for v in exon:
        for allele in v.allele:
            per_allele[allele].append(str(v.id))
    by_set: dict[tuple, list[str]] = defaultdict(list)
    for allele, vids in per_allele.items():
        by_set[tuple(sorted(set(vids)))].append(allele)
    rest = set(collect_allele_names(variants)) - set(collect_allele_names(exon))
    if rest:
        by_set[tuple()] = sorted(rest)
    return {"|".join(members): members for members in by_set.values()}


# ---------------------------------------------------------------------------
# a21-a23: EM path
# ---------------------------------------------------------------------------
def candidate_alleles_per_mate(positive: list[list[str]], negative: list[list[str]]) -> list[str]:
    """Intersection of the positive allele sets minus every negative set (typing_em.py:68-87)."""
    if not positive:
        return []
    keep = set(positive[0])
    for alleles in positive[1:]:
        keep &= set(alleles)
    for alleles in negative:
        keep -= set(alleles)
    return list(keep)


def most_frequent(candidates: list[str]) -> list[str]:
    """Alleles with the maximal multiplicity (typing_em.py:90-104)."""
    tally = Counter(candidates)
    if not tally:
        return []
    top = max(tally.values())
    return [name for name, c in tally.items() if c == top]


def em_abundance(allele_per_read: list[list[str]], iter_max: int = 300,
                 diff_threshold: float = 1e-4) -> dict[str, float]:
    """SQUAREM-accelerated EM over a 0/1 read x allele matrix (typing_em.py:107-188)."""
    names = sorted(set(chain.from_iterable(allele_per_read)))
    col = {name: i for i, name in enumerate(names)}
    compat = np.zeros((len(allele_per_read), len(names)))
    for r, alleles in enumerate(allele_per_read):
        for name in alleles:
            compat[r, col[name]] = 1

    def step(p):
        w = p * compat
        tot = w.sum(axis=1)[:, None]
        w = np.divide(w, tot, out=np.zeros(w.shape), where=tot != 0)
        s = w.sum(axis=0)
        return s / s.sum()

    p = step(np.ones(len(names)))
    for _ in range(iter_max):
        p1 = step(p)
        p2 = step(p1)
        r = p1 - p
        v = p2 - p1 - r
        rr, vv = (r ** 2).sum(), (v ** 2).sum()
        if vv > 0.0:
            g = -np.sqrt(rr / vv)
            p1 = step(np.maximum(p - r * g * 2 + v * g ** 2, 0))
        if np.abs(p - p1).sum() <= diff_threshold:
            break
        p = p1
    return dict(zip(names, p))


def em_call(report: list[tuple[str, float]], cn: int) -> list[str]:
    """Allocate CN copies by abundance (kir_typing.py:163-195); report = [(allele, prob)]."""
    share = 1 / cn
    called = []
    for allele, prob in sorted(report, key=lambda item: -item[1]):
        copies = max(1, round(prob / share))
        called.extend([allele] * min(cn, copies))
        cn -= copies
        if cn <= 0:
            break
    return called


# ---------------------------------------------------------------------------
# read grouping of novel discovery (graphkir/novel_discover.py:48-70)
# ---------------------------------------------------------------------------
def group_reads_float(probs_called: np.ndarray, names: list[str]) -> dict[tuple[str, ...], list[int]]:
    """Literal restatement: ``is_max = np.equal(probs[:, ids], probs[:, ids].max(axis=1)[:, None])``
    (:62-64), then every read goes to the sorted tuple of the alleles at the row maximum (:66-69).
    Returns read indices per key, keys in order of first occurrence (the reference's dict order)."""
    probs_called = np.asarray(probs_called, dtype=np.float64)
    is_max = np.equal(probs_called, probs_called.max(axis=1)[:, None])
    groups: dict[tuple[str, ...], list[int]] = {}
    arr = np.array(names)
    for i, row in enumerate(is_max):
        groups.setdefault(tuple(sorted(arr[row].tolist())), []).append(i)
    return groups


def group_reads_int(m_called: np.ndarray, names: list[str]) -> dict[tuple[str, ...], list[int]]:
    """The same grouping decided on the integer mismatch counts: probs is strictly decreasing in m for a
    fixed read, so the row maximum of probs is the row minimum of m; exact ties stay ties (the float
    comparison above can split them by the rounding of the ordered product)."""
    m_called = np.asarray(m_called, dtype=np.int64)
    is_min = m_called == m_called.min(axis=1)[:, None]
    groups: dict[tuple[str, ...], list[int]] = {}
    arr = np.array(names)
    for i, row in enumerate(is_min):
        groups.setdefault(tuple(sorted(arr[row].tolist())), []).append(i)
    return groups
