"""Host packing (kir_graph_b200/packing.py) against the oracle's set logic."""
import copy

import numpy as np
import pytest

from kir_graph_b200 import packing, synthetic
from oracle import typing_oracle as orc
from tests.helpers import golden_names, load_golden, objects_from_input


def counts_from_pack(pack: packing.GenePack) -> np.ndarray:
    """NumPy emulation of the likelihood kernel's formula on the packed arrays."""
    m = np.zeros((pack.n_reads, pack.n_alleles), dtype=np.int64)
    row = np.repeat(np.arange(pack.n_reads), np.diff(pack.ent_off))
    mem = pack.mem_words[pack.ent_word]                      # [E, A]
    x = (pack.ent_pos[:, None] & ~mem) | (pack.ent_neg[:, None] & mem)
    pop = np.zeros(x.shape, dtype=np.int64)
    for b in range(32):
        pop += (x >> np.uint32(b)) & np.uint32(1)
    np.add.at(m, row, pop)
    return m


@pytest.mark.parametrize("name", golden_names("typing"))
def test_pack_matches_oracle_and_reference(name):
    case = load_golden(name)
    reads, variants = objects_from_input(case["input"])
    reads_o = copy.deepcopy(reads)
    pack = packing.pack_gene(reads, variants, variant_correction=case["variant_correction"])
    assert pack.allele_names == case["allele_names"]
    assert pack.n_reads == case["n_reads"]
    # in-place side effect of errorCorrection on the caller's reads
    kept = [reads[i] for i in pack.kept_reads]
    after = [{"lpv": r.lpv, "rpv": r.rpv, "lnv": r.lnv, "rnv": r.rnv} for r in kept]
    assert after == case["reads_after"]
    # oracle path
    if case["variant_correction"]:
        reads_o = orc.error_correction(reads_o)
    reads_o = orc.remove_empty_reads(reads_o)
    by_id = {str(v.id): v for v in variants}
    col = {n: i for i, n in enumerate(pack.allele_names)}
    m, k = orc.mismatch_counts(reads_o, by_id, col)
    assert np.array_equal(pack.k_obs, k)
    assert np.array_equal(counts_from_pack(pack), m)


def test_pack_synthetic_equals_object_path():
    gene = synthetic.make_gene([5, 1], "KIRQ*BACKBONE", 33, 264, 3, 500, hierarchical=True)
    a = packing.pack_synthetic(gene, variant_correction=True)
    reads, variants = gene.to_objects()
    b = packing.pack_gene(reads, variants, variant_correction=True)
    assert a.allele_names == b.allele_names
    for f in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "kept_reads"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_duplicates_and_conflicts():
    from kir_graph_b200.hisat2 import PairRead
    from kir_graph_b200.msa2hisat import Variant
    g = "X*BACKBONE"
    variants = [Variant(pos=i, typ="single", ref=g, val="A", id=f"hv{i}",
                        allele=[f"X*{j}" for j in range(3) if (i + j) % 2]) for i in range(40)]
    variants.append(Variant(pos=99, typ="single", ref=g, val="C", id="nv0", allele=[]))
    reads = [PairRead(backbone=g, lpv=["hv1", "hv1", "hv33"], rpv=["hv1", "nv0"], lnv=["hv1", "hv2"], rnv=["hv2", "hv2", "hv39"]),
             PairRead(backbone=g),
             PairRead(backbone=g, rnv=["hv0"])]
    pack = packing.pack_gene(reads, variants, variant_correction=False)
    assert pack.kept_reads.tolist() == [0, 2]
    by_id = {v.id: v for v in variants}
    col = {n: i for i, n in enumerate(pack.allele_names)}
    m, k = orc.mismatch_counts([reads[0], reads[2]], by_id, col)
    assert np.array_equal(pack.k_obs, k) and k.tolist() == [10, 1]
    assert np.array_equal(counts_from_pack(pack), m)
    full = packing.pack_gene(reads, variants, variant_correction=False, no_empty=False)
    assert full.n_reads == 3 and full.k_obs.tolist() == [10, 0, 1]


def test_too_many_observations_is_loud():
    gene = synthetic.make_gene([5, 2], "KIRW*BACKBONE", 8, 600, 1, 4, w=140)
    with pytest.raises(ValueError):
        packing.pack_synthetic(gene, variant_correction=False)


def test_batched_homozygosity_decisions_equal_the_per_gene_rule():
    """cohort.HomozygosityIndex (candidate sites of all problems in one array) against decide_homozygous per
    problem (isHomozygous, typing_mulit_allele.py:807-857), for every copy number 1..4."""
    from kir_graph_b200 import cohort
    packs = []
    for i in range(24):
        gene = synthetic.make_gene([61, i], f"KIRHZ{i}*BACKBONE", 6 + 3 * (i % 7), 96, 2 + i % 3, 600 + 40 * i,
                                   homo_prob=0.5)
        packs.append(packing.pack_synthetic(gene))
    index = cohort.HomozygosityIndex(packs)
    seen = set()
    for cn in (1, 2, 3, 4):
        got = index.decide(np.full(len(packs), cn, dtype=np.int64))
        want = np.array([cohort.decide_homozygous(p, cn) for p in packs])
        assert np.array_equal(got, want), cn
        seen |= set(want.tolist())
    assert seen == {True, False}
    mixed = np.array([1 + i % 4 for i in range(len(packs))], dtype=np.int64)
    assert np.array_equal(index.decide(mixed), np.array([cohort.decide_homozygous(p, int(c)) for p, c in zip(packs, mixed)]))


def _gene_with_an_undecidable_site(name: str):
    """Ten insertions at one position, each seen four times: the site is looked at by isHomozygous (depth 40,
    a positive value) but no value passes share > 0.1, and the reference indexes an empty list (:853)."""
    from kir_graph_b200.hisat2 import PairRead
    from kir_graph_b200.msa2hisat import Variant
    vals = ["A", "C", "G", "T", "AA", "CC", "GG", "TT", "AC", "AG"]
    variants = [Variant(pos=100, typ="insertion", ref=name, val=v, id=f"hv{i}", allele=[f"{name[:-9]}*{i % 3:03d}"])
                for i, v in enumerate(vals)]
    variants.append(Variant(pos=300, typ="single", ref=name, val="A", id="hv90", allele=[f"{name[:-9]}*000"]))
    reads = [PairRead(backbone=name, lpv=[f"hv{i % 10}"], rnv=["hv90"]) for i in range(40)]
    return reads, variants


def test_undecidable_site_fails_like_the_reference_only_where_it_is_looked_at():
    """The batched homozygosity index raises IndexError for a problem with such a site exactly when the per-gene
    rule does: copy number above 1 and not a gene that is heterozygous by name."""
    import pytest
    from kir_graph_b200 import cohort
    from kir_graph_b200.typing_mulit_allele import isHomozygous
    reads, variants = _gene_with_an_undecidable_site("KIR3DL9*BACKBONE")
    with pytest.raises(IndexError):                                             # the mirror of the reference's rule
        isHomozygous(reads, {v.id: v for v in variants}, 2)
    bad = packing.pack_gene(reads, variants, variant_correction=False, gene="KIR3DL9*BACKBONE")
    reads2, variants2 = _gene_with_an_undecidable_site("KIR2DL5*BACKBONE")       # heterozygous by name: never asked
    named = packing.pack_gene(reads2, variants2, variant_correction=False, gene="KIR2DL5*BACKBONE")
    good = packing.pack_synthetic(synthetic.make_gene([61, 3], "KIRHZ3*BACKBONE", 9, 96, 2, 600))
    index = cohort.HomozygosityIndex([good, bad, named])
    assert index.broken.tolist() == [False, True, True]
    for cn in (2, 3):
        with pytest.raises(IndexError, match="KIR3DL9"):
            index.decide(np.array([2, cn, 2]))
        with pytest.raises(IndexError):
            cohort.decide_homozygous(bad, cn)
    for cns in ([2, 1, 2], [2, 0, 3]):                                          # not asked for that problem: no failure
        got = index.decide(np.array(cns))
        assert got.tolist() == [cohort.decide_homozygous(p, c) for p, c in zip([good, bad, named], cns)]
