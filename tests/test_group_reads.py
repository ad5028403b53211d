"""Read grouping by called alleles (reference graphkir/novel_discover.py:48-70).

The golden fixture holds what the UNMODIFIED reference returned (tests/golden/make_golden.py).  The
oracle's literal float restatement must reproduce it exactly; the integer statement - what the CUDA
path computes - must agree with it wherever the reference's float comparison did not split an exact
tie (checked against the stored probabilities with a relative tolerance of 1e-12)."""
import numpy as np
import pytest

from kir_graph_b200 import novel_discover
from kir_graph_b200.typing_mulit_allele import AlleleTyping
from oracle import typing_oracle as orc
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden, objects_from_input

CASES = load_golden("group_reads")["cases"]


def _golden_groups(case):
    return {tuple(k): v for k, v in case["groups"]}


def _names(case, typ):
    return [n for n in case["predict_alleles"] if n in typ.allele_to_id]


def _tolerant_groups(case, names):
    """Grouping of the stored reference probabilities with ties up to 1e-12 relative."""
    p = np.array(case["probs_called"], dtype=np.float64)
    is_max = p >= p.max(axis=1)[:, None] * (1 - 1e-12)
    groups = {}
    arr = np.array(names)
    for i, row in enumerate(is_max):
        groups.setdefault(tuple(sorted(arr[row].tolist())), []).append(i)
    return groups


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_restatement_equals_reference(case):
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTyping(reads, variants, no_empty=False, _backend=FakeBackend())
    names = _names(case, typ)
    got = orc.group_reads_float(case["probs_called"], names)
    assert got == _golden_groups(case) and list(got) == [tuple(k) for k, _ in case["groups"]]


def _check_model(case, backend):
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTyping(reads, variants, no_empty=False, _backend=backend)
    assert typ.getReadsNum() == len(reads)                    # the empty read is kept
    names = _names(case, typ)
    ids = [typ.allele_to_id[n] for n in names]
    got = novel_discover.groupReadByAllele(typ, case["predict_alleles"], reads)
    index = {id(r): i for i, r in enumerate(reads)}
    got_idx = {k: [index[id(r)] for r in v] for k, v in got.items()}
    want = orc.group_reads_int(typ.mismatch_counts()[:, ids], names)
    assert got_idx == want and list(got_idx) == list(want)    # same groups, same key order
    assert got_idx == _tolerant_groups(case, names)           # = the reference modulo split float ties
    ref = _golden_groups(case)
    moved = sum(len(set(v) ^ set(ref.get(k, []))) for k, v in got_idx.items())
    assert moved <= 0.02 * len(reads), "the reference's float noise should touch few reads"
    # the read without observations ties over every called allele
    assert 3 in got_idx[tuple(sorted(names))]
    return typ, ids


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_grouping_on_numpy_backend(case):
    _check_model(case, FakeBackend())


def test_error_behaviour():
    case = CASES[0]
    reads, variants = objects_from_input(case["input"])
    fake = FakeBackend()
    typ = AlleleTyping(reads, variants, no_empty=False, _backend=fake)
    assert novel_discover.groupReadByAllele(typ, ["KIRNOT*00001"], reads) == {}     # no known allele (:58-59)
    with pytest.raises(ValueError):
        novel_discover.groupReadByAllele(typ, case["predict_alleles"], reads[:-1])
    with pytest.raises(ValueError):
        typ.group_pattern(list(range(33)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_grouping_on_gpu(case):
    from kir_graph_b200 import engine
    cuda = engine.CudaBackend()
    typ, ids = _check_model(case, cuda)
    reads, variants = objects_from_input(case["input"])
    ref = AlleleTyping(reads, variants, no_empty=False, _backend=FakeBackend())
    assert np.array_equal(typ.group_pattern(ids), ref.group_pattern(ids))
    assert np.array_equal(typ.group_pattern(ids[::-1] + ids), ref.group_pattern(ids[::-1] + ids))


def test_split_reads_by_alleles_driver(tmp_path):
    """splitReadsByAlleles (novel_discover.py:267-277) over a whole sample."""
    from kir_graph_b200 import synthetic
    from kir_graph_b200.hisat2 import writeReadsAndVariantsData
    from kir_graph_b200.kir_typing import TypingWithPosNegAllele
    genes = synthetic.make_wgs30x_sample(seed=9, total_reads=3000)[:4]
    reads, variants, predict = [], [], []
    for g in genes:
        r, v = g.to_objects()
        reads += r
        variants += v
        predict += [g.allele_names[t] for t in g.truth]
    path = str(tmp_path / "s.json")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, path)
    model = TypingWithPosNegAllele(path, _backend=FakeBackend())
    seen = {}
    for gene, alleles, group, var in novel_discover.splitReadsByAlleles(model, predict):
        assert all(a.split("*")[0] == gene.split("*")[0] for a in alleles) and len(group) > 0
        seen[gene] = seen.get(gene, 0) + len(group)
    assert seen == {g.gene: len(model._gene_reads[g.gene]) for g in genes}
