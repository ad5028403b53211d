"""
Seeded synthetic typing workloads.

A gene problem is generated directly in CSR form (variant indices per read for
the four lists lpv/rpv/lnv/rnv) so that the 200k-read and 2M-read shapes of
BASELINE.json can be built in seconds; ``to_objects`` materialises the same
problem as ``PairRead``/``Variant`` objects for the object-level API, the
oracle and the reference.

Generator (per gene): V single-base variants at pos 25*v, ``in_exon`` for every
4th; membership Bernoulli(f_v), f_v ~ Beta(0.3, 1); CN truth alleles (with
probability 0.2 one of them is duplicated); a read pair picks a truth allele and
a window start uniformly, the left mate observes ``w`` consecutive variants and
the right mate the next ``w`` overlapping by two; each observation is flipped
with probability 0.01 and recorded as positive if the (flipped) allele carries
the variant, else negative.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .hisat2 import PairRead
from .msa2hisat import Variant

LIST_NAMES = ("lpv", "rpv", "lnv", "rnv")  # product order of the likelihood


@dataclass
class ReadCSR:
    """Four ragged lists per read, as variant indices into the gene's table."""

    n_reads: int
    offsets: dict[str, np.ndarray]  # name -> int64[R+1]
    indices: dict[str, np.ndarray]  # name -> int32[nnz]

    def row(self, name: str, r: int) -> np.ndarray:
        off = self.offsets[name]
        return self.indices[name][off[r]:off[r + 1]]

    def take(self, keep: np.ndarray) -> "ReadCSR":
        """Sub-select reads (boolean mask or index array), preserving order."""
        keep = np.asarray(keep)
        if keep.dtype == bool:
            keep = np.flatnonzero(keep)
        offsets, indices = {}, {}
        for name in LIST_NAMES:
            off = self.offsets[name]
            lens = (off[1:] - off[:-1])[keep]
            new_off = np.zeros(len(keep) + 1, dtype=np.int64)
            np.cumsum(lens, out=new_off[1:])
            # gather ranges
            starts = off[:-1][keep]
            total = int(new_off[-1])
            pos = np.arange(total, dtype=np.int64)
            row = np.repeat(np.arange(len(keep)), lens)
            src = starts[row] + (pos - new_off[:-1][row])
            offsets[name] = new_off
            indices[name] = self.indices[name][src]
        return ReadCSR(len(keep), offsets, indices)


@dataclass
class SyntheticGene:
    gene: str                       # backbone name, e.g. "KIRS03*BACKBONE"
    allele_names: list[str]         # index -> name (already sorted)
    member: np.ndarray              # bool [V, A]
    in_exon: np.ndarray             # bool [V]
    variant_ids: list[str]          # index -> "hv<n>"
    cn: int
    truth: list[int]                # allele indices, len == cn
    reads: ReadCSR
    groups: list[list[int]] = field(default_factory=list)  # CDS groups (hierarchical mode)

    @property
    def n_alleles(self) -> int:
        return len(self.allele_names)

    @property
    def n_variants(self) -> int:
        return len(self.variant_ids)

    def variants(self) -> list[Variant]:
        out = []
        for v, vid in enumerate(self.variant_ids):
            out.append(Variant(
                pos=25 * v, typ="single", ref=self.gene, val="ACGT"[v % 4], id=vid,
                allele=[self.allele_names[a] for a in np.flatnonzero(self.member[v])],
                in_exon=bool(self.in_exon[v]),
            ))
        return out

    def pair_reads(self) -> list[PairRead]:
        ids = np.array(self.variant_ids, dtype=object)
        out = []
        for r in range(self.reads.n_reads):
            out.append(PairRead(
                backbone=self.gene, multiple=1,
                lpv=list(ids[self.reads.row("lpv", r)]),
                rpv=list(ids[self.reads.row("rpv", r)]),
                lnv=list(ids[self.reads.row("lnv", r)]),
                rnv=list(ids[self.reads.row("rnv", r)]),
            ))
        return out

    def to_objects(self) -> tuple[list[PairRead], list[Variant]]:
        return self.pair_reads(), self.variants()


def _membership(rng: np.random.Generator, n_var: int, n_allele: int,
                hierarchical: bool, in_exon: np.ndarray) -> tuple[np.ndarray, list[list[int]]]:
    freq = rng.beta(0.3, 1.0, size=n_var)
    if not hierarchical:
        return rng.random((n_var, n_allele)) < freq[:, None], []
    # CDS groups: geometric sizes (mean 4, max 12); exon variants shared by a group
    groups: list[list[int]] = []
    a = 0
    while a < n_allele:
        size = int(min(12, rng.geometric(0.25), n_allele - a))
        groups.append(list(range(a, a + size)))
        a += size
    group_of = np.zeros(n_allele, dtype=np.int64)
    for g, members in enumerate(groups):
        group_of[members] = g
    per_allele = rng.random((n_var, n_allele)) < freq[:, None]
    per_group = rng.random((n_var, len(groups))) < freq[:, None]
    member = np.where(in_exon[:, None], per_group[:, group_of], per_allele)
    return member, groups


def make_gene(
    seed: int | list[int],
    gene: str,
    n_allele: int,
    n_var: int,
    cn: int,
    n_reads: int,
    w: int = 12,
    flip: float = 0.01,
    hierarchical: bool = False,
    variant_id_base: int = 0,
    homo_prob: float = 0.2,
) -> SyntheticGene:
    rng = np.random.default_rng(seed)
    in_exon = (np.arange(n_var) % 4) == 0
    member, groups = _membership(rng, n_var, n_allele, hierarchical, in_exon)
    width = len(str(max(n_allele - 1, 1)))
    width = max(width, 5)
    names = [f"{gene.split('*')[0]}*{i:0{width}d}" for i in range(n_allele)]

    truth = list(rng.choice(n_allele, size=min(cn, n_allele), replace=False))
    while len(truth) < cn:
        truth.append(truth[0])
    if cn > 1 and rng.random() < homo_prob:
        truth[1] = truth[0]
    truth = [int(t) for t in truth]

    span = 2 * w - 2
    assert n_var >= span, "variant table shorter than one read pair window"
    which = rng.integers(0, cn, size=n_reads)
    allele_of_read = np.asarray(truth, dtype=np.int64)[which]
    start = rng.integers(0, n_var - span + 1, size=n_reads)
    # column j < w: left mate, variant start+j ; j >= w: right mate, start + (w-2) + (j-w)
    col = np.concatenate([np.arange(w), np.arange(w) + (w - 2)])
    vi = (start[:, None] + col[None, :]).astype(np.int32)           # [R, 2w]
    has = member[vi, allele_of_read[:, None]]
    flipped = rng.random(vi.shape) < flip
    positive = has ^ flipped
    is_left = np.broadcast_to(np.arange(2 * w) < w, vi.shape)

    offsets, indices = {}, {}
    for name, mate_left, pol in (("lpv", True, True), ("rpv", False, True),
                                 ("lnv", True, False), ("rnv", False, False)):
        mask = (is_left == mate_left) & (positive == pol)
        off = np.zeros(n_reads + 1, dtype=np.int64)
        np.cumsum(mask.sum(axis=1), out=off[1:])
        offsets[name] = off
        indices[name] = vi[mask].astype(np.int32)
    reads = ReadCSR(n_reads, offsets, indices)
    ids = [f"hv{variant_id_base + v}" for v in range(n_var)]
    return SyntheticGene(gene, names, member, in_exon, ids, cn, truth, reads, groups)


# ---------------------------------------------------------------------------
# BASELINE.json shapes
# ---------------------------------------------------------------------------
WGS30X_ALLELES = [160, 130, 110, 90, 80, 60, 50, 45, 40, 35, 30, 25, 15, 12, 8, 6, 4]
WGS30X_CN = [2, 2, 1, 2, 3, 1, 1, 2, 2, 4, 1, 2, 1, 1, 2, 2, 1]


def make_wgs30x_sample(seed: int = 3, total_reads: int = 200_000,
                       hierarchical: bool = False, scale: float = 1.0) -> list[SyntheticGene]:
    """cfg3: 17 genes, 900 alleles, 200k read pairs, CN <= 4 (``scale`` shrinks reads for tests)."""
    genes = []
    base = 0
    cn_total = sum(WGS30X_CN)
    for g, (n_allele, cn) in enumerate(zip(WGS30X_ALLELES, WGS30X_CN)):
        n_var = max(64, 8 * n_allele)
        n_reads = max(1, int(round(total_reads * scale * cn / cn_total)))
        genes.append(make_gene([seed, g], f"KIRS{g:02d}*BACKBONE", n_allele, n_var, cn,
                               n_reads, w=12, hierarchical=hierarchical,
                               variant_id_base=base))
        base += n_var
    return genes


def make_deep_sample(seed: int = 4, n_reads: int = 2_000_000, n_allele: int = 1000,
                     n_var: int = 8000, cn: int = 6) -> SyntheticGene:
    """cfg4: one gene, 2M read pairs x 1000 alleles, CN 6 with distinct truth alleles."""
    return make_gene([seed, 0], "KIRDEEP*BACKBONE", n_allele, n_var, cn, n_reads,
                     w=12, homo_prob=0.0)


def make_cohort(n_samples: int = 96, first_seed: int = 100, **kw) -> list[list[SyntheticGene]]:
    """cfg5: ``n_samples`` independent cfg3 samples with seeds first_seed.."""
    return [make_wgs30x_sample(seed=first_seed + i, **kw) for i in range(n_samples)]
