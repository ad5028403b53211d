"""Host orchestration (kir_graph_b200/engine.py + typing classes) on the NumPy test double.

These tests cover everything around the kernels -- pool layout, work-item lists, step
sequencing, result conversion -- without a GPU; the kernels themselves are checked by the
-m gpu tests against the same oracle.
"""
import copy

import numpy as np
import pytest

from kir_graph_b200 import engine, packing, synthetic
from kir_graph_b200.typing_mulit_allele import AlleleTyping, AlleleTypingExonFirst
from oracle import typing_oracle as orc
from tests.fake_backend import FakeBackend
from tests.helpers import (assert_same_modulo_ties, golden_names, int_scores_from_values, load_golden,
                           objects_from_input)


def oracle_for_pack(pack, reads, variants, top_n):
    by_id = {str(v.id): v for v in variants}
    col = {n: i for i, n in enumerate(pack.allele_names)}
    kept = [reads[i] for i in pack.kept_reads]
    m, k = orc.mismatch_counts(kept, by_id, col)
    return m, k, orc.IntSearch(m, k, top_n=top_n)


def check_steps(group_out, ref):
    assert np.array_equal(group_out.ids, ref.allele_id)
    assert np.array_equal(group_out.score, ref.score)
    n = group_out.n
    w = np.array([orc.lcm_upto(n) // q for q in range(1, n + 1)])
    assert np.array_equal((group_out.cnt * w[None, None, :]).sum(axis=2), ref.frac_num)
    assert group_out.n_unique == ref.n_unique


@pytest.mark.parametrize("spec", [
    dict(seed=[7, 0], n_allele=20, n_var=96, cn=3, n_reads=260, top_n=25),
    dict(seed=[7, 1], n_allele=5, n_var=64, cn=4, n_reads=150, top_n=300),
    dict(seed=[7, 2], n_allele=70, n_var=560, cn=2, n_reads=300, top_n=40),
    dict(seed=[7, 3], n_allele=150, n_var=600, cn=2, n_reads=140, top_n=130),
    dict(seed=[7, 4], n_allele=14, n_var=64, cn=7, n_reads=120, top_n=20),
    # 125 kept sets at the second step: a remainder of 16 row groups (121..127 rows) had no tile at all
    dict(seed=[7, 5], n_allele=125, n_var=500, cn=2, n_reads=120, top_n=300),
    dict(seed=[7, 6], n_allele=40, n_var=160, cn=3, n_reads=100, top_n=250),
])
@pytest.mark.parametrize("half", [False, True])
def test_search_group_equals_int_oracle(spec, half):
    spec = dict(spec)
    top_n = spec.pop("top_n")
    cn = spec["cn"]
    gene = synthetic.make_gene(gene="KIRT*BACKBONE", **spec)
    reads, variants = gene.to_objects()
    pack = packing.pack_gene(reads, variants, variant_correction=True)
    m, k, search = oracle_for_pack(pack, reads, variants, top_n)
    be = FakeBackend()
    batch = engine.MatrixBatch([pack], backend=be, packed=half)
    assert np.array_equal(batch.mismatch_counts(0), m)
    assert np.array_equal(batch.blocked_counts(0), m)
    assert np.array_equal(batch.colsum(0), m.sum(axis=0))
    group = engine.SearchGroup(batch, [0], top_n)
    for step in range(cn):
        out = group.step(need_next=[step + 1 < cn])[0]
        ref = search.add_candidate()
        check_steps(out, ref)
    p = group.materialize_p(0, out.ids)
    assert np.array_equal(p, ref.allele_prob)


def test_batch_of_genes_with_different_cn_and_restricted_candidates():
    genes = [synthetic.make_gene([9, i], f"KIRU{i}*BACKBONE", a, max(64, 8 * a), c, r)
             for i, (a, c, r) in enumerate([(12, 2, 130), (40, 3, 200), (6, 1, 90), (33, 2, 170)])]
    packs, oracles = [], []
    for g in genes:
        reads, variants = g.to_objects()
        p = packing.pack_gene(reads, variants)
        packs.append(p)
        oracles.append(oracle_for_pack(p, reads, variants, 30)[2])
    be = FakeBackend()
    batch = engine.MatrixBatch(packs, backend=be)
    group = engine.SearchGroup(batch, [0, 1, 2, 3, 1], 30)      # search 4 shares matrix 1, restricted candidates
    oracles.append(orc.IntSearch(oracles[1].m, np.array([oracles[1].k_total]), top_n=30))
    cns = [2, 3, 1, 2, 3]
    rng = np.random.default_rng(0)
    for step in range(3):
        active = np.array([c > step for c in cns])
        need = np.array([c > step + 1 for c in cns])
        cands = [None, None, None, None, rng.choice(packs[1].n_alleles, size=7, replace=False)]
        outs = group.step(cands=cands, active=active, need_next=need)
        for s in np.flatnonzero(active):
            ref = oracles[s].add_candidate(cands[s])
            check_steps(outs[s], ref)
    assert set(be.log) >= {"gk_likelihood", "gk_first_step", "gk_score", "gk_select", "gk_rescore_count",
                           "gk_rank", "gk_write_p"}


@pytest.mark.parametrize("name", golden_names("typing"))
def test_allele_typing_class_against_reference(name):
    case = load_golden(name)
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTyping(reads, variants, force_homo=case["force_homo"], top_n=case["top_n"],
                       variant_correction=case["variant_correction"], _backend=FakeBackend())
    assert [typ.id_to_allele[i] for i in range(len(typ.id_to_allele))] == case["allele_names"]
    assert typ.getReadsNum() == case["n_reads"]
    after = [{"lpv": r.lpv, "rpv": r.rpv, "lnv": r.lnv, "rnv": r.rnv} for r in typ.reads]
    assert after == case["reads_after"]
    np.testing.assert_allclose(typ.log_probs, np.array(case["log_probs"]), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(typ.probs, np.array(case["probs"]), rtol=1e-9)
    res = typ.typing(case["cn"])
    assert len(typ.result) == len(case["steps"])
    for got, ref in zip(typ.result, case["steps"]):
        assert got.n == ref["n"]
        np.testing.assert_allclose(got.value, np.sort(ref["value"])[::-1], rtol=1e-11)
        if got.frac_num is None:        # homozygous shortcut
            assert got.allele_id.tolist() == ref["allele_id"] or got.tie_flags
            continue
        ref_scores = int_scores_from_values(ref["value"], typ._k_total)
        assert_same_modulo_ties(ref["allele_id"], ref_scores, got.allele_id, got.score,
                                kept_all=got.n_unique <= case["top_n"])
    best = res.selectBest()
    assert best == case["best"] or res.tie_flags, (best, case["best"])
    # lazy allele_prob equals max over members of log_probs
    last = typ.result[-1]
    if last.frac_num is not None:
        lp = typ.log_probs
        want = lp[:, last.allele_id].max(axis=2)
        np.testing.assert_allclose(np.asarray(last.allele_prob), want, rtol=1e-12)


@pytest.mark.parametrize("name", golden_names("exonfirst"))
def test_exon_first_class_against_reference(name):
    case = load_golden(name)
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTypingExonFirst(reads, variants, force_homo=False, top_n=case["top_n"],
                                candidate_set_threshold=case["threshold"], _backend=FakeBackend())
    assert {k: sorted(v) for k, v in typ.allele_group.items()} == \
           {k: sorted(v) for k, v in case["allele_group"].items()}
    assert [typ.id_to_allele[i] for i in range(len(typ.id_to_allele))] == case["exon_allele_names"]
    assert typ.getReadsNum() == case["exon_n_reads"]
    res = typ.typing(case["cn"])
    for got, ref in zip(typ.result[: case["cn"]], case["exon_steps"]):
        np.testing.assert_allclose(got.value, np.sort(ref["value"])[::-1], rtol=1e-11)
    assert len(typ.result) == case["n_results"] or any(r.tie_flags for r in typ.result)
    np.testing.assert_allclose(res.value[:5], np.sort(case["final"]["value"])[::-1][:5], rtol=1e-11)
    best = res.selectBest()
    assert best == case["best"] or res.tie_flags, (best, case["best"])


def test_deepcopy_continues_search():
    gene = synthetic.make_gene([3, 3], "KIRV*BACKBONE", 18, 144, 3, 220)
    reads, variants = gene.to_objects()
    be = FakeBackend()
    a = AlleleTyping(copy.deepcopy(reads), variants, force_homo=False, top_n=20, _backend=be)
    a.addCandidate()
    b = copy.deepcopy(a)
    assert b._batch is a._batch
    a.addCandidate()
    b.addCandidate()
    assert np.array_equal(a.result[-1].allele_id, b.result[-1].allele_id)
    assert np.array_equal(a.result[-1].score, b.result[-1].score)
    assert len(a.result) == len(b.result) == 2


@pytest.mark.parametrize("name", ["syn_a6_cn2", "syn_a12_cn4_nocorr", "worked_example_corr"])
def test_reference_named_likelihood_methods(name):
    """read2Onehot / onehot2Prob / errorCorrection / reads2AlleleProb under the reference's names
    (typing_mulit_allele.py:287-381): same lists after correction and the reference's probs, also
    when called on their own; the reference's ordered product restated from the one-hot helpers."""
    import copy
    from kir_graph_b200.hisat2 import PairRead
    case = load_golden(name)
    reads, variants = objects_from_input(case["input"])
    raw = copy.deepcopy(reads)
    typ = AlleleTyping(reads, variants, force_homo=case["force_homo"], top_n=case["top_n"],
                       variant_correction=case["variant_correction"], _backend=FakeBackend())
    if case["variant_correction"]:
        fixed = typ.errorCorrection(raw)
        assert fixed is raw
        kept = AlleleTyping.removeEmptyReads(raw)
        assert [{"lpv": r.lpv, "rpv": r.rpv, "lnv": r.lnv, "rnv": r.rnv} for r in kept] == case["reads_after"]
    probs = typ.reads2AlleleProb(typ.reads)
    np.testing.assert_allclose(probs, np.array(case["probs"]), rtol=1e-9)
    for r, read in list(enumerate(typ.reads))[:25]:
        factors = [typ.onehot2Prob(typ.read2Onehot(typ.variants[i])) for i in read.lpv + read.rpv] + \
                  [typ.onehot2Prob(np.logical_not(typ.read2Onehot(typ.variants[i]))) for i in read.lnv + read.rnv]
        np.testing.assert_allclose(probs[r], np.stack(factors).prod(axis=0), rtol=1e-12)
    assert typ.reads2AlleleProb([]).shape == (0,)
    blank = PairRead(backbone=str(variants[0].ref))
    with pytest.raises(ValueError):
        typ.reads2AlleleProb([typ.reads[0], blank])              # no_empty: np.stack of nothing
    loose = AlleleTyping(copy.deepcopy(typ.reads), variants, no_empty=False, variant_correction=False,
                         _backend=FakeBackend())
    rows = loose.reads2AlleleProb([typ.reads[0], blank])
    np.testing.assert_allclose(rows[1], np.full(len(typ.allele_to_id), 0.999), rtol=1e-12)
    np.testing.assert_allclose(rows[0], probs[0], rtol=1e-12)


def test_errors_and_empty():
    gene = synthetic.make_gene([3, 4], "KIRY*BACKBONE", 8, 64, 2, 50)
    reads, variants = gene.to_objects()
    typ = AlleleTyping(reads, variants, force_homo=False, top_n=10, _backend=FakeBackend())
    with pytest.raises(ValueError):
        typ.typing(0)
    with pytest.raises(ValueError):
        AlleleTyping.createHomoResult(typ.addCandidate(), 1)
    empty = AlleleTyping([], variants, force_homo=False, top_n=10, _backend=FakeBackend())
    res = empty.typing(2)
    assert res.isFail() and res.selectBest() == ["fail", "fail"] and empty.getReadsNum() == 0
    assert empty.probs.shape == (0,)


def test_alive_overflow_falls_back_to_exact_grids(monkeypatch):
    """More alive sets than the pre-sized rescoring grid (large exact tie): the step is redone exactly."""
    monkeypatch.setattr(engine, "ALIVE_SLACK", 0)
    gene = synthetic.make_gene([7, 9], "KIRTIE*BACKBONE", 30, 240, 3, 40)      # few reads -> many exact ties
    reads, variants = gene.to_objects()
    pack = packing.pack_gene(reads, variants, variant_correction=True)
    m, k, search = oracle_for_pack(pack, reads, variants, 8)
    be = FakeBackend()
    group = engine.SearchGroup(engine.MatrixBatch([pack], backend=be), [0], 8)
    seen_overflow = False
    for step in range(3):
        before = be.log.count("gk_rank")
        out = group.step(need_next=[step < 2])[0]
        seen_overflow |= be.log.count("gk_rank") - before == 2
        check_steps(out, search.add_candidate())
    assert seen_overflow


def test_pipelined_batch_equals_stepwise_and_falls_back(monkeypatch):
    from kir_graph_b200 import cohort
    genes = synthetic.make_wgs30x_sample(seed=12, total_reads=2500)
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    ref = cohort.BatchTyper(packs, cns, top_n=25, backend=FakeBackend())
    ref.pipelined = False
    want = ref.run()
    fast = cohort.BatchTyper(packs, cns, top_n=25, backend=FakeBackend())
    got = fast.run()
    assert [(c.alleles, c.score, c.best_rank, c.tie_flags) for c in got] == \
           [(c.alleles, c.score, c.best_rank, c.tie_flags) for c in want]
    assert fast.score_cells == ref.score_cells
    # a pre-sized grid that is too small is detected at the single read-back and the run is redone stepwise
    monkeypatch.setattr(engine, "ALIVE_SLACK", 0)
    be = FakeBackend()
    slow = cohort.BatchTyper(packs, cns, top_n=25, backend=be)
    got = slow.run()
    assert [(c.alleles, c.score) for c in got] == [(c.alleles, c.score) for c in want]


def test_pass_pipeline_equals_serial_passes():
    """PassPipeline (passes in flight over replicas that share the host pools) returns, pass by pass
    and in order, what typer.run() returns; a single typer used with depth 2 as well."""
    from kir_graph_b200 import cohort
    genes = (synthetic.make_wgs30x_sample(seed=14, total_reads=900) + synthetic.make_wgs30x_sample(seed=15, total_reads=900))
    genes = [g for g in genes if g.n_alleles <= 40][:8]
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    key = lambda calls: [(c.gene, c.alleles, c.score, c.best_rank, c.tie_flags) for c in calls]
    typer = cohort.CohortTyper(packs, cns, top_n=25, backend=FakeBackend(), n_parts=2, group_size=4)
    want = key(typer.run())
    assert len(want) == len(packs)
    twin = typer.replica(packs, cns, top_n=25, group_size=4)
    assert all(a.host is b.host for a, b in zip(typer.parts, twin.parts))
    for typers, depth, upload in (([typer, twin], None, False), ([typer, twin], None, True), ([typer], 2, False),
                                  ([typer], 1, True)):
        pipe = cohort.PassPipeline(typers, upload=upload, depth=depth)
        out = []
        for i in range(3):
            done = pipe.submit()
            assert (done is None) == (i < pipe.depth)
            if done is not None:
                out.append(done)
        out += pipe.drain()
        assert len(out) == 3 and all(key(calls) == want for calls in out)
        assert pipe.drain() == []
    # a typer whose passes are not pipelined (read-back per step) still gives the same calls
    for part in typer.parts:
        part.pipelined = False
    pipe = cohort.PassPipeline([typer], depth=2)
    pipe.submit(), pipe.submit()
    assert [key(c) for c in pipe.drain()] == [want, want]


def test_host_batch_of_other_problems_is_refused():
    """A prepared host batch carries state derived from its packs (pools, homozygosity index): handing it
    to a typer of other problems - even of the same number - is an error, not a silent wrong call."""
    import pytest
    from kir_graph_b200 import cohort
    genes = [g for g in synthetic.make_wgs30x_sample(seed=21, total_reads=600) if g.n_alleles <= 40][:4]
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    first = cohort.BatchTyper(packs, cns, top_n=25, backend=FakeBackend())
    again = cohort.BatchTyper(packs, cns, top_n=25, backend=FakeBackend(), host_batch=first.host)
    assert again.homo_index is first.homo_index
    with pytest.raises(ValueError, match="other problems"):
        cohort.BatchTyper(packs[::-1], cns[::-1], top_n=25, backend=FakeBackend(), host_batch=first.host)
    with pytest.raises(ValueError, match="other problems"):
        cohort.BatchTyper(packs[:3], cns[:3], top_n=25, backend=FakeBackend(), host_batch=first.host)


def test_exon_first_with_top_n_below_five_follows_the_reference():
    """top_n < 5 leaves the restricted model of exon-first with top_n // 5 == 0 kept sets
    (typing_mulit_allele.py:716): the reference then answers "fail" for one step and raises IndexError
    (an empty float array used as an index, :540) for more - checked against the imported reference
    on random genes; the mirror does the same instead of its own ValueError."""
    gene = synthetic.make_gene([380855173, 0], "KIRQ*BACKBONE", 22, 176, 3, 143, hierarchical=True)
    reads, variants = gene.to_objects()
    typ = AlleleTypingExonFirst(reads, variants, force_homo=None, top_n=3, candidate_set_threshold=0.0,
                                _backend=FakeBackend())
    with pytest.raises(IndexError):
        typ.typing(3)
    gene = synthetic.make_gene([237585442, 0], "KIRQ*BACKBONE", 10, 80, 1, 24, hierarchical=True)
    reads, variants = gene.to_objects()
    typ = AlleleTypingExonFirst(reads, variants, force_homo=None, top_n=3, candidate_set_threshold=0.0,
                                _backend=FakeBackend())
    res = typ.typing(1)
    assert res.isFail() and res.selectBest() == ["fail"]


def test_fraction_near_the_select_best_threshold_is_reported():
    """tie_flags bit3: read-level ties leave the outcome of selectBest's test "every member fraction
    >= 1 / (2n)" open for a rank it looks at - a read tied between members counts 1/q for each in exact
    arithmetic, but 1 for one of them in the reference when the tied log-probabilities differ in the
    last bit (typing_mulit_allele.py:575-580).  This gene is the case a differential run against the
    imported reference found (the reference computes 31/122 = 0.254 for the second member of rank 0 and
    keeps it, exact arithmetic gives 30/122 = 0.246 and moves on to rank 1)."""
    from kir_graph_b200.typing_mulit_allele import TIE_FRACTION_NEAR_THRESHOLD
    gene = synthetic.make_gene([257500406, 1], "KIRE1*BACKBONE", 3, 64, 4, 122, hierarchical=False,
                               variant_id_base=1000)
    reads, variants = gene.to_objects()
    typ = AlleleTyping(reads, variants, force_homo=None, top_n=40, _backend=FakeBackend())
    res = typ.typing(2)
    assert res.tie_flags & TIE_FRACTION_NEAR_THRESHOLD
    assert abs(res.fraction[0].min() - 0.25) < 0.01 and res.selectBest() == res.allele_name[1]
    assert any(r["tie_flags"] & TIE_FRACTION_NEAR_THRESHOLD for r in typ.tie_report)
    # the batched path reads back only the called set: gk_rank reports the same bit in GkStepInfo
    from kir_graph_b200 import cohort
    call = cohort.BatchTyper([packing.pack_synthetic(gene)], [2], top_n=40, backend=FakeBackend()).run()[0]
    assert call.tie_flags & TIE_FRACTION_NEAR_THRESHOLD and call.alleles == res.selectBest()
    # far from the threshold: not set
    gene = synthetic.make_gene([3, 4], "KIRY*BACKBONE", 8, 64, 2, 300)
    reads, variants = gene.to_objects()
    typ = AlleleTyping(reads, variants, force_homo=False, top_n=10, _backend=FakeBackend())
    res = typ.typing(2)
    alone = res.frac_num is not None and typ.result[-1].fraction[0].min() > 0.3
    assert not (alone and res.tie_flags & TIE_FRACTION_NEAR_THRESHOLD and not res.fraction[0].min() < 0.35)


def test_exon_first_candidates_typed_in_chunks_give_the_same_result(monkeypatch):
    """The tied exon candidates of exon-first are typed in chunks bounded by a device-memory budget; one at a
    time (budget of one search) or all together, the merged result is the same."""
    import kir_graph_b200.typing_mulit_allele as tma
    name = golden_names("exonfirst")[0]
    case = load_golden(name)

    def run():
        reads, variants = objects_from_input(case["input"])
        typ = AlleleTypingExonFirst(reads, variants, force_homo=False, top_n=case["top_n"],
                                    candidate_set_threshold=0.0, _backend=FakeBackend())
        res = typ.typing(case["cn"])
        return res.allele_id.tolist(), res.value.tolist(), len(typ.result)

    whole = run()
    monkeypatch.setattr(tma, "INTRON_SEARCH_BUDGET_BYTES", 1)
    assert run() == whole and whole[2] > case["cn"] + 1


def test_row_pieces_cover_every_kept_set_once_inside_the_allocated_blocks():
    """Row tiles of the packed scoring path (SearchGroup._row_pieces): for every kept-set count the pieces
    under a warp-split column tile ("W") and under a full-width tile plus its halves ("F" + "H") cover rows
    0 .. k-1 exactly once, start at multiples of 8, may pad only beyond k, stay inside the 64-row blocks P
    and S are allocated in for top_n = k, and span at most two k-blocks each (what a stage copies)."""
    from kir_graph_b200.engine import SHAPE_WARP_SPLIT, SearchGroup
    for k in range(1, 700):
        alloc = 64 * -(-k // 64)
        for kinds in (("W",), ("F", "H")):
            cover = np.zeros(alloc, dtype=np.int64)
            for kind in kinds:
                for start, code, rows in SearchGroup._row_pieces(k, kind):
                    assert start % 8 == 0 and rows > 0 and start + rows <= alloc, (k, kind, start, rows)
                    assert (start % 64) + rows <= 128, (k, kind, start, rows)
                    if code & SHAPE_WARP_SPLIT:
                        gp, wk = code & 0xf, (code >> 4) & 0xf
                        assert rows == (8 * gp) << wk and 1 <= gp <= 4 and wk <= 2
                    cover[start:start + rows] += 1
            assert (cover[:k] == 1).all(), (k, kinds)
            assert (cover[k:] <= 1).all(), (k, kinds)
