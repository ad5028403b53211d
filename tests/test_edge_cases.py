"""Edge cases of the typing path: tiny allele universes, empty inputs, top_n = 1, ragged batches.
Run on the NumPy test double here and on the GPU in test_gpu_kernels.py::test_edge_cases_on_gpu."""
import numpy as np
import pytest

from kir_graph_b200 import cohort, engine, packing, synthetic
from kir_graph_b200.hisat2 import PairRead
from kir_graph_b200.msa2hisat import Variant
from kir_graph_b200.typing_mulit_allele import AlleleTyping
from oracle import typing_oracle as orc
from tests.fake_backend import FakeBackend


def edge_problems():
    g = "KIRE*BACKBONE"
    one = [Variant(pos=10 * i, typ="single", ref=g, val="A", id=f"hv{i}", allele=["KIRE*001"]) for i in range(5)]
    reads_one = [PairRead(backbone=g, lpv=[f"hv{i % 5}"], rnv=[f"hv{(i + 1) % 5}"]) for i in range(40)]
    two = [Variant(pos=10 * i, typ="single", ref=g, val="A", id=f"hv{i}",
                   allele=["KIRE*001"] if i % 2 else ["KIRE*002"]) for i in range(6)]
    reads_two = [PairRead(backbone=g, lpv=[f"hv{i % 6}"], lnv=[f"hv{(i + 3) % 6}"]) for i in range(60)]
    return {"one_allele": (reads_one, one, 3), "two_alleles": (reads_two, two, 4)}


@pytest.mark.parametrize("name", ["one_allele", "two_alleles"])
@pytest.mark.parametrize("top_n", [1, 5])
def test_tiny_universe(name, top_n, backend=None):
    reads, variants, cn = edge_problems()[name]
    be = backend or FakeBackend()
    typ = AlleleTyping(reads, variants, force_homo=False, top_n=top_n, variant_correction=False, _backend=be)
    res = typ.typing(cn)
    by_id = {v.id: v for v in variants}
    m, k = orc.mismatch_counts(typ.reads, by_id, typ.allele_to_id)
    ref = orc.IntSearch(m, k, top_n=top_n)
    want = ref.typing(cn)
    assert np.array_equal(res.allele_id, want.allele_id)
    assert np.array_equal(res.score, want.score)
    assert res.selectBest() == orc.select_best(want, [typ.id_to_allele[i] for i in range(len(typ.id_to_allele))])


def test_ragged_batch_with_empty_and_zero_cn(backend=None):
    be = backend or FakeBackend()
    genes = [synthetic.make_gene([31, i], f"KIRR{i}*BACKBONE", a, 64, c, r)
             for i, (a, c, r) in enumerate([(3, 2, 50), (17, 1, 1), (33, 3, 129), (16, 2, 64), (5, 2, 200)])]
    packs = [packing.pack_synthetic(g) for g in genes]
    empty = packing.pack_gene([], genes[0].variants(), gene="KIREMPTY*BACKBONE")
    packs.insert(2, empty)
    cns = [2, 1, 2, 3, 0, 2]
    calls = cohort.BatchTyper(packs, cns, top_n=7, backend=be).run()
    assert [c.gene for c in calls] == [p.gene for p, c in zip(packs, cns) if c]
    assert calls[2].alleles == ["fail", "fail"] and calls[2].n_reads == 0
    for pack, cn, call in zip([p for p, c in zip(packs, cns) if c], [c for c in cns if c], calls):
        if pack.n_reads == 0:
            continue
        batch = engine.MatrixBatch([pack], backend=FakeBackend())
        ref = orc.IntSearch(batch.mismatch_counts(0).astype(np.int64), pack.k_obs, top_n=7)
        want = ref.typing(cn, homo=call.homozygous)
        names = pack.allele_names
        assert call.alleles == orc.select_best(want, names), pack.gene
