"""
Drop-in replacement of the reference's ``graphkir/typing_em.py`` (HISAT-genotype style EM).

Same function names and return types; the set algebra per read pair and the SQUAREM EM
run on the GPU (``gk_em_compat`` / ``gk_em_squarem`` in csrc/gk_em.cu):

* ``getCandidateAllelePerRead`` / ``getMostFreqAllele`` (typing_em.py:68-104) become bitset
  AND / ANDN / OR over the variant -> allele membership table, one warp per read pair;
* ``hisatEMnp`` (:107-188) iterates on the distinct compatibility rows with multiplicities,
  in float64 with a fixed summation order.

Floating point: the EM sums over reads in a different order than NumPy's pairwise sum, so
abundances agree with the reference to ~1e-12, not bit for bit (tests use 1e-9).
"""
from __future__ import annotations

from collections import Counter, defaultdict
from dataclasses import dataclass
from itertools import chain

import numpy as np

from . import engine
from ._cabi import EM_PROBLEM_DTYPE
from .hisat2 import ReadsAndVariantsData, loadReadsAndVariantsData, removeMultipleMapped


@dataclass
class Hisat2AlleleResult:
    """Abundance of one allele (reference: typing_em.py:22-29)."""

    allele: str
    count: int
    prob: float
    cn: int = 0


@dataclass
class GeneEmReads:
    """Reads of one gene for the EM path: allele sets of the variants + four CSR lists."""

    allele_names: list[str]          # sorted universe
    sets: list[tuple[str, ...]]      # distinct allele sets referenced by the lists ("variants")
    lists: dict[str, tuple[np.ndarray, np.ndarray]]   # lp/ln/rp/rn -> (offsets int32[R+1], set index int32[nnz])
    n_reads: int

    def __len__(self) -> int:
        return self.n_reads


def readAlleleLength(file_fasta: str) -> dict[str, int]:
    """Sequence length per record id of a fasta file (typing_em.py:32-34; the reference goes through
    Bio.SeqIO, whose record id is the header up to the first blank); feeds ``hisatEMnp(seq_len=...)``."""
    lengths: dict[str, int] = {}
    name = None
    with open(file_fasta) as handle:
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                fields = line[1:].split(None, 1)
                name = fields[0] if fields else ""
                lengths[name] = 0
            elif name is not None:
                lengths[name] += len(line.strip())
    return lengths


def _gene_em_reads(per_read: list[dict[str, list[list[str]]]]) -> GeneEmReads:
    """Reference format (list of {'lp','ln','rp','rn': list of allele-name lists}) -> packed."""
    set_id: dict[tuple[str, ...], int] = {}
    lists = {}
    for key in ("lp", "ln", "rp", "rn"):
        off = np.zeros(len(per_read) + 1, dtype=np.int32)
        idx: list[int] = []
        for r, read in enumerate(per_read):
            for alleles in read[key]:
                t = tuple(alleles)
                idx.append(set_id.setdefault(t, len(set_id)))
            off[r + 1] = len(idx)
        lists[key] = (off, np.asarray(idx, dtype=np.int32))
    sets = list(set_id.keys())
    names = sorted(set(chain.from_iterable(sets)))
    return GeneEmReads(names, sets, lists, len(per_read))


def preprocessHisatReads(reads_data: ReadsAndVariantsData) -> dict[str, GeneEmReads]:
    """Group reads by backbone and attach the allele set of every variant id (:37-65).

    Returns the packed per-gene form; ``hisat2TypingPerGene`` also accepts the reference's
    list-of-dicts form."""
    allele_of = {v.id: tuple(v.allele) for v in reads_data["variants"]}
    reads = reads_data["reads"]
    assert all(r.multiple == 1 for r in reads)
    per_gene: dict[str, list] = defaultdict(list)
    for read in reads:
        per_gene[read.backbone].append(read)
    out = {}
    for gene, gene_reads in per_gene.items():
        set_id: dict[str, int] = {}
        sets: list[tuple[str, ...]] = []
        lists = {}
        for key, attr in (("lp", "lpv"), ("ln", "lnv"), ("rp", "rpv"), ("rn", "rnv")):
            off = np.zeros(len(gene_reads) + 1, dtype=np.int32)
            idx: list[int] = []
            for r, read in enumerate(gene_reads):
                for vid in getattr(read, attr):
                    if vid not in set_id:
                        set_id[vid] = len(sets)
                        sets.append(allele_of[vid])
                    idx.append(set_id[vid])
                off[r + 1] = len(idx)
            lists[key] = (off, np.asarray(idx, dtype=np.int32))
        names = sorted(set(chain.from_iterable(sets)))
        out[gene] = GeneEmReads(names, sets, lists, len(gene_reads))
    return out


def _membership_words(names: list[str], sets: list[tuple[str, ...]]) -> tuple[np.ndarray, int]:
    col = {n: i for i, n in enumerate(names)}
    n_aw = max(1, (len(names) + 31) // 32)
    table = np.zeros((max(len(sets), 1), n_aw), dtype=np.uint32)
    for v, alleles in enumerate(sets):
        for a in alleles:
            c = col[a]
            table[v, c >> 5] |= np.uint32(1) << np.uint32(c & 31)
    return table, n_aw


def compatible_alleles(gene: GeneEmReads, backend=None) -> np.ndarray:
    """uint32 [R, ceil(A/32)]: per read pair the alleles of getMostFreqAllele(left + right)."""
    be = backend if backend is not None else engine.default_backend()
    table, n_aw = _membership_words(gene.allele_names, gene.sets)
    if gene.n_reads == 0:
        return np.zeros((0, n_aw), dtype=np.uint32)
    d_tab = be.upload(table)
    args = []
    for key in ("lp", "ln", "rp", "rn"):
        off, idx = gene.lists[key]
        args += [be.upload(off), be.upload(idx)]
    d_out = be.zeros(2 * gene.n_reads * n_aw, np.uint32)
    be.launch("gk_em_compat", d_tab, n_aw, len(gene.allele_names), *args, gene.n_reads, d_out)
    return be.download(d_out, np.uint32)[: gene.n_reads * n_aw].reshape(gene.n_reads, n_aw).copy()


def _rows_to_names(rows: np.ndarray, names: list[str]) -> list[list[str]]:
    out = []
    for row in rows:
        ids = [32 * w + b for w, word in enumerate(row) for b in range(32) if (int(word) >> b) & 1]
        out.append([names[i] for i in ids])
    return out


def em_from_rows(rows: np.ndarray, names: list[str], seq_len: dict[str, int] | None = None,
                 iter_max: int = 300, diff_threshold: float = 1e-4, backend=None) -> dict[str, float]:
    """Abundances of the alleles that occur in ``rows`` (compatibility bitsets, one per read)."""
    be = backend if backend is not None else engine.default_backend()
    n_aw = rows.shape[1] if rows.ndim == 2 else 1
    present = np.zeros(n_aw, dtype=np.uint32)
    for w in range(n_aw):
        present[w] = np.bitwise_or.reduce(rows[:, w]) if len(rows) else 0
    ids = [32 * w + b for w in range(n_aw) for b in range(32) if (int(present[w]) >> b) & 1]
    if not ids:
        return {}
    # re-index the columns to the alleles that occur (the reference's EM universe, :137-138)
    a_new = len(ids)
    w_new = (a_new + 31) // 32
    bits = ((rows[:, [i >> 5 for i in ids]] >> np.array([i & 31 for i in ids], dtype=np.uint32)) & 1).astype(np.uint32)
    packed = np.zeros((len(rows), w_new), dtype=np.uint32)
    for j in range(a_new):
        packed[:, j >> 5] |= bits[:, j] << np.uint32(j & 31)
    uniq, counts = np.unique(packed, axis=0, return_counts=True)
    keep = uniq.any(axis=1)
    uniq, counts = uniq[keep], counts[keep]
    sel_names = [names[i] for i in ids]
    lengths = np.array([float(seq_len[n]) for n in sel_names]) if seq_len else np.ones(a_new)
    prob = be.zeros(0, np.float64)
    tab = np.zeros(1, dtype=EM_PROBLEM_DTYPE)
    tab[0] = (0, 0, 0, 0, len(uniq), a_new, w_new, 0)
    d_out = be.zeros(5 * a_new + len(uniq), np.float64)
    d_iters = be.zeros(1, np.int32)
    be.launch("gk_em_squarem", be.upload(tab), 1, be.upload(uniq.astype(np.uint32)),
              be.upload(counts.astype(np.uint32)), be.upload(lengths.astype(np.float64)), d_out, d_iters,
              int(iter_max), float(diff_threshold))
    prob = be.download(d_out, np.float64)[:a_new]
    return dict(zip(sel_names, (float(x) for x in prob)))


def getCandidateAllelePerRead(positive_allele: list[list[str]], negative_allele: list[list[str]],
                              _backend=None) -> list[str]:
    """Single-mate form of the set algebra, kept for API parity (:68-87)."""
    gene = _gene_em_reads([{"lp": positive_allele, "ln": negative_allele, "rp": [], "rn": []}])
    rows = compatible_alleles(gene, _backend)
    return _rows_to_names(rows, gene.allele_names)[0]


def getMostFreqAllele(candidates: list[str]) -> list[str]:
    """Alleles with the maximal multiplicity (:90-104); host utility for API parity."""
    tally = Counter(candidates)
    if not tally:
        return []
    top = max(tally.values())
    return [name for name, c in tally.items() if c == top]


def hisatEMnp(allele_per_read: list[list[str]], seq_len: dict[str, int] = {}, iter_max: int = 300,
              diff_threshold: float = 0.0001, _backend=None) -> dict[str, float]:
    """EM abundance from per-read compatible alleles (:107-188)."""
    names = sorted(set(chain.from_iterable(allele_per_read)))
    col = {n: i for i, n in enumerate(names)}
    n_aw = max(1, (len(names) + 31) // 32)
    rows = np.zeros((len(allele_per_read), n_aw), dtype=np.uint32)
    for r, alleles in enumerate(allele_per_read):
        for a in alleles:
            c = col[a]
            rows[r, c >> 5] |= np.uint32(1) << np.uint32(c & 31)
    return em_from_rows(rows, names, seq_len or None, iter_max, diff_threshold, _backend)


def hisat2TypingPerGene(reads_alleles, _backend=None) -> list[Hisat2AlleleResult]:
    """Compatible alleles per read pair, then EM (:191-215).  Accepts ``GeneEmReads`` or the
    reference's list of {'lp','ln','rp','rn'} dicts."""
    gene = reads_alleles if isinstance(reads_alleles, GeneEmReads) else _gene_em_reads(reads_alleles)
    rows = compatible_alleles(gene, _backend)
    prob = em_from_rows(rows, gene.allele_names, None, 300, 1e-4, _backend)
    counts = {}
    for w in range(rows.shape[1]):
        for b in range(32):
            i = 32 * w + b
            if i < len(gene.allele_names):
                c = int(((rows[:, w] >> np.uint32(b)) & 1).sum())
                if c:
                    counts[gene.allele_names[i]] = c
    return [Hisat2AlleleResult(allele=a, count=counts.get(a, 0), prob=prob.get(a, 0.0))
            for a in sorted(set(prob) | set(counts))]


def hisat2Typing(read_and_variant_json: str, output_prefix: str) -> None:
    """EM report of every gene of a sample (:218-241)."""
    import json
    from dataclasses import asdict
    reads_data = removeMultipleMapped(loadReadsAndVariantsData(read_and_variant_json))
    result = {gene: hisat2TypingPerGene(reads) for gene, reads in preprocessHisatReads(reads_data).items()}
    with open(output_prefix + ".txt", "w") as f:
        printHisatTyping(result, file=f)
    with open(output_prefix + ".json", "w") as f:
        json.dump({g: [asdict(x) for x in r] for g, r in result.items()}, f)


def printHisatTyping(hisat_result, first_n: int = 10, file=None) -> None:
    import sys
    file = file or sys.stdout
    for backbone, result in hisat_result.items():
        print(backbone, file=file)
        for i, allele in enumerate(sorted(result, key=lambda x: x.count, reverse=True)[:first_n]):
            print(f"  {i+1:2d} {allele.allele:18s} (count: {allele.count})", file=file)
        for i, allele in enumerate(sorted(result, key=lambda x: x.prob, reverse=True)[:first_n]):
            print(f"  Rank {i+1:2d} {allele.allele:18s} (abundance: {allele.prob:.2f}, cn: {allele.cn})", file=file)
