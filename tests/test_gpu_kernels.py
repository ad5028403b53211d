"""CUDA kernels against (1) the NumPy launch-by-launch statement in tests/fake_backend.py,
(2) the oracle's exact integer search, (3) golden outputs of the reference itself.
All calls go through the C ABI (ctypes -> libgk_typing.so)."""
import numpy as np
import pytest

from kir_graph_b200 import engine, packing, synthetic
from kir_graph_b200._cabi import GK_KB, GK_RT, STEP_INFO_DTYPE
from oracle import typing_oracle as orc
from tests.fake_backend import FakeBackend
from tests.helpers import (assert_same_modulo_ties, golden_names, int_scores_from_values, load_golden,
                           objects_from_input)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    return engine.CudaBackend()


def _packs(specs, seed):
    packs = []
    for i, (a, c, r) in enumerate(specs):
        g = synthetic.make_gene([seed, i], f"KIRK{i}*BACKBONE", a, max(64, 8 * a), c, r)
        packs.append(packing.pack_synthetic(g))
    return packs


def _compare_outputs(a, b):
    assert set(a) == set(b)
    for s in a:
        x, y = a[s], b[s]
        assert np.array_equal(x.ids, y.ids), f"search {s}: ids"
        assert np.array_equal(x.score, y.score), f"search {s}: score"
        assert np.array_equal(x.cnt, y.cnt), f"search {s}: cnt"
        assert np.array_equal(x.flat, y.flat), f"search {s}: flat"
        assert (x.n_unique, x.n_alive, x.cut, x.tie_flags) == (y.n_unique, y.n_alive, y.cut, y.tie_flags)


@pytest.mark.parametrize("specs,top_n,cns", [
    ([(12, 2, 130), (40, 3, 700), (6, 1, 90), (150, 2, 400), (3, 4, 260)], 30, [2, 3, 1, 2, 4]),
    ([(70, 3, 1500), (33, 2, 90)], 300, [3, 2]),
    ([(40, 6, 500), (12, 5, 300), (90, 8, 260)], 30, [6, 5, 8]),
    # ragged allele / kept-set counts around every tile boundary of the packed path (full-width tiles of
    # 128 alleles, warp-split tiles of 8..64 alleles x 8..128 kept sets, two of them after a 128 tile)
    ([(4, 2, 200), (9, 2, 300), (17, 3, 330), (33, 2, 410), (65, 2, 280), (100, 2, 300), (129, 2, 310),
      (161, 2, 290), (200, 2, 270)], 300, [2, 2, 3, 2, 2, 2, 2, 2, 2]),
    ([(57, 3, 900), (136, 3, 520), (250, 2, 300)], 100, [3, 3, 2]),
])
@pytest.mark.parametrize("half", [False, True])
def test_every_launch_matches_numpy_statement(cuda, specs, top_n, cns, half):
    packs = _packs(specs, 101)
    fake = FakeBackend()
    bg, bf = engine.MatrixBatch(packs, backend=cuda, packed=half), engine.MatrixBatch(packs, backend=fake, packed=half)
    assert np.array_equal(cuda.download(bg.d_LT, np.uint8), bf.d_LT), "LT"
    assert np.array_equal(cuda.download(bg.d_L, np.float32), bf.d_L), "L"
    assert np.array_equal(cuda.download(bg.d_col, np.uint64), bf.d_col), "colsum"
    ids = list(range(len(packs))) + [1]
    cns = cns + [cns[1]]
    gg, gf = engine.SearchGroup(bg, ids, top_n), engine.SearchGroup(bf, ids, top_n)
    rng = np.random.default_rng(5)
    for step in range(max(cns)):
        active = np.array([c > step for c in cns])
        need = np.array([c > step + 1 for c in cns])
        cands = [None] * len(packs) + [rng.choice(packs[1].n_alleles, size=9, replace=False)]
        og = gg.step(cands=cands, active=active, need_next=need)
        of = gf.step(cands=cands, active=active, need_next=need)
        if step:
            assert np.array_equal(cuda.download(gg.d_S, np.uint32), gf.d_S), f"S at step {step + 1}"
        _compare_outputs(og, of)
        if need.any():
            pg = cuda.download(gg.d_P, np.uint16 if half else np.float32)
            for s in np.flatnonzero(need):
                rp, nkb = int(gg.mt["r_pad"][s]), int(gg.tab["n_kblk"][s])
                kb = -(-int(gg.kept[s]) // GK_KB)      # written kept-set blocks
                o = int(gg.tab["P_off"][s])
                shape = (rp // GK_RT, nkb, GK_RT, GK_KB)
                assert np.array_equal(pg[o:o + nkb * rp * GK_KB].reshape(shape)[:, :kb],
                                      gf.d_P[o:o + nkb * rp * GK_KB].reshape(shape)[:, :kb]), f"P of search {s}"


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("a,cn,r,top_n", [(120, 3, 6000, 300), (200, 2, 12000, 300), (45, 4, 3000, 64),
                                          (125, 2, 4000, 300), (40, 3, 3000, 250)])
def test_search_equals_oracle(cuda, a, cn, r, top_n, half):
    gene = synthetic.make_gene([77, a], "KIRO*BACKBONE", a, 8 * a, cn, r)
    pack = packing.pack_synthetic(gene)
    batch = engine.MatrixBatch([pack], backend=cuda, packed=half)
    m = batch.mismatch_counts(0)
    # independent m from the CSR lists by set logic
    want = np.zeros_like(m, dtype=np.int64)
    member = gene.member[:, [gene.allele_names.index(n) for n in pack.allele_names]]
    for name in ("lpv", "rpv", "lnv", "rnv"):
        off, idx = pack.csr.offsets[name], pack.csr.indices[name]
        row = np.repeat(np.arange(pack.n_reads), np.diff(off))
        contrib = ~member[idx] if name in ("lpv", "rpv") else member[idx]
        np.add.at(want, row, contrib.astype(np.int64))
    assert np.array_equal(m, want)
    search = orc.IntSearch(m.astype(np.int64), pack.k_obs, top_n=top_n)
    group = engine.SearchGroup(batch, [0], top_n)
    for step in range(cn):
        out = group.step(need_next=[step + 1 < cn])[0]
        ref = search.add_candidate()
        assert np.array_equal(out.ids, ref.allele_id)
        assert np.array_equal(out.score, ref.score)
        w = np.array([orc.lcm_upto(out.n) // q for q in range(1, out.n + 1)])
        assert np.array_equal((out.cnt * w[None, None, :]).sum(axis=2), ref.frac_num)
    assert np.array_equal(group.materialize_p(0, out.ids), ref.allele_prob)


@pytest.mark.parametrize("name", golden_names("typing"))
def test_reference_golden(cuda, name):
    from kir_graph_b200.typing_mulit_allele import AlleleTyping
    case = load_golden(name)
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTyping(reads, variants, force_homo=case["force_homo"], top_n=case["top_n"],
                       variant_correction=case["variant_correction"], _backend=cuda)
    np.testing.assert_allclose(typ.log_probs, np.array(case["log_probs"]), rtol=1e-5, atol=1e-12)
    res = typ.typing(case["cn"])
    for got, ref in zip(typ.result, case["steps"]):
        np.testing.assert_allclose(got.value, np.sort(ref["value"])[::-1], rtol=1e-5)
        if got.frac_num is None:
            continue
        ref_scores = int_scores_from_values(ref["value"], typ._k_total)
        assert_same_modulo_ties(ref["allele_id"], ref_scores, got.allele_id, got.score,
                                kept_all=got.n_unique <= case["top_n"])
    best = res.selectBest()
    assert best == case["best"] or res.tie_flags, (best, case["best"])


def test_split_invariance_and_idempotence(cuda, monkeypatch):
    """Scores do not depend on how reads are chunked into work items, and reruns are bit-identical."""
    gene = synthetic.make_gene([88, 0], "KIRP*BACKBONE", 130, 1040, 3, 20000)
    pack = packing.pack_synthetic(gene)
    outs = []
    for chunk in (8192, 2048, 8192):
        monkeypatch.setattr(engine, "SCORE_READ_CHUNK", chunk)
        batch = engine.MatrixBatch([pack], backend=cuda)
        group = engine.SearchGroup(batch, [0], 300)
        steps = [group.step(need_next=[i < 2])[0] for i in range(3)]
        outs.append(steps)
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a.ids, b.ids) and np.array_equal(a.score, b.score) and np.array_equal(a.cnt, b.cnt)
    for a, b in zip(outs[0], outs[2]):
        assert np.array_equal(a.ids, b.ids) and np.array_equal(a.score, b.score) and np.array_equal(a.cnt, b.cnt)
    last = outs[0][-1]
    assert np.all(np.diff(last.score) >= 0)
    assert np.all(last.cnt.sum(axis=(1, 2)) >= pack.n_reads)          # every read counted for >= 1 member
    w = np.array([orc.lcm_upto(3) // q for q in (1, 2, 3)])
    assert np.all((last.cnt * w).sum(axis=(1, 2)) == pack.n_reads * orc.lcm_upto(3))   # fractions sum to 1


def test_em_path_matches_reference_golden(cuda):
    from kir_graph_b200 import typing_em
    cases = load_golden("em_cases")["cases"]
    kat = cases["kat_candidate"]
    assert sorted(typing_em.getCandidateAllelePerRead(kat["positive"], kat["negative"], _backend=cuda)) == sorted(kat["out"])
    for key in ("kat_simple", "syn_a16"):
        prob = typing_em.hisatEMnp(cases[key]["allele_per_read"], _backend=cuda)
        assert set(prob) == set(cases[key]["prob"])
        for name, p in cases[key]["prob"].items():
            assert abs(prob[name] - p) < 1e-9          # float64, different summation order than NumPy
    reads, variants = objects_from_input(cases["syn_a16"]["input"])
    gene = typing_em.preprocessHisatReads({"reads": reads, "variants": variants})["KIRI*BACKBONE"]
    rows_gpu = typing_em.compatible_alleles(gene, cuda)
    rows_np = typing_em.compatible_alleles(gene, FakeBackend())
    assert np.array_equal(rows_gpu, rows_np)
    got = [sorted(x) for x in typing_em._rows_to_names(rows_gpu, gene.allele_names)]
    assert got == cases["syn_a16"]["allele_per_read"]
    # run-to-run bit reproducibility of the EM (fixed summation order)
    a = typing_em.hisatEMnp(cases["syn_a16"]["allele_per_read"], _backend=cuda)
    b = typing_em.hisatEMnp(cases["syn_a16"]["allele_per_read"], _backend=cuda)
    assert a == b


def test_em_larger_against_oracle(cuda):
    from kir_graph_b200 import typing_em
    gene = synthetic.make_gene([55, 0], "KIREM*BACKBONE", 180, 1440, 3, 20000)
    reads, variants = gene.to_objects()
    packed = typing_em.preprocessHisatReads({"reads": reads, "variants": variants})["KIREM*BACKBONE"]
    rows = typing_em.compatible_alleles(packed, cuda)
    per_read = typing_em._rows_to_names(rows, packed.allele_names)
    # oracle set algebra on a sample of reads, EM on everything
    by_id = {v.id: v.allele for v in variants}
    for r in range(0, len(reads), 97):
        want = orc.most_frequent(
            orc.candidate_alleles_per_mate([by_id[v] for v in reads[r].lpv], [by_id[v] for v in reads[r].lnv])
            + orc.candidate_alleles_per_mate([by_id[v] for v in reads[r].rpv], [by_id[v] for v in reads[r].rnv]))
        assert sorted(want) == sorted(per_read[r])
    ref = orc.em_abundance(per_read)
    got = typing_em.em_from_rows(rows, packed.allele_names, backend=cuda)
    assert set(ref) == set(got)
    for name in ref:
        assert abs(ref[name] - got[name]) < 1e-9


@pytest.mark.parametrize("method", ["full", "exonfirst_1", "em"])
def test_sample_driver_matches_reference(cuda, tmp_path, method):
    import os
    from kir_graph_b200 import kir_typing
    from kir_graph_b200.hisat2 import writeReadsAndVariantsData
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "sample.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    kw = {"full": dict(top_n=60, variant_correction=True), "exonfirst_1": dict(top_n=60), "em": {}}[method]
    t = kir_typing.selectKirTypingModel(method, path, _backend=cuda, **kw)
    alleles, warn = t.typing(sample["gene_cn"])
    ref = sample["calls"][method]
    assert warn == ref["warnings"]
    if method == "em":
        assert sorted(alleles) == sorted(ref["alleles"])
    else:
        assert alleles == ref["alleles"] or t.tie_report


def test_cohort_batch_equals_per_gene_class(cuda):
    from kir_graph_b200 import cohort
    from kir_graph_b200.typing_mulit_allele import AlleleTyping
    genes = synthetic.make_wgs30x_sample(seed=5, total_reads=30000)
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    calls = cohort.CohortTyper(packs, cns, top_n=300, backend=cuda, n_parts=2, group_size=6).run()
    for g, c in zip(genes, calls):
        reads, variants = g.to_objects()
        t = AlleleTyping(reads, variants, force_homo=None, top_n=300, _backend=cuda)
        r = t.typing(g.cn)
        assert r.selectBest() == c.alleles
        assert abs(c.value - r.value[c.best_rank]) <= 1e-9 * abs(c.value)


def test_repeated_passes_replay_a_cuda_graph(cuda):
    """A batch typed again and again is replayed as one CUDA graph from the third pass on; the calls
    must not depend on whether a pass was launched eagerly, captured, or replayed, nor on a re-upload."""
    from kir_graph_b200 import cohort
    genes = synthetic.make_wgs30x_sample(seed=11, total_reads=24000)
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    typer = cohort.BatchTyper(packs, cns, top_n=300, backend=cuda)
    key = lambda calls: [(c.gene, tuple(c.alleles), c.best_rank, c.score, c.tie_flags) for c in calls]
    first = key(typer.run())
    for _ in range(4):
        assert key(typer.run()) == first
    assert getattr(typer, "graph_error", None) is None and typer._graph is not None
    cuda.zero_(typer.batch.d_stream)                     # wreck an input on the device ...
    cuda.zero_(typer.batch.d_mem)
    assert key(typer.upload_and_run()) == first          # ... the re-upload restores it; same buffers,
    assert typer._graph is not None                      # so the recorded graph is still valid
    for _ in range(2):
        assert key(typer.run()) == first
    eager = cohort.BatchTyper(packs, cns, top_n=300, backend=cuda)
    eager.use_graph = False
    assert key(eager.run()) == first and key(eager.run()) == first


def test_two_batches_in_flight_on_one_stream(cuda):
    """start() of two typers before finish() of the first: the read-backs must not share a buffer."""
    from kir_graph_b200 import cohort
    key = lambda calls: [(c.gene, tuple(c.alleles), c.best_rank, c.score, c.tie_flags) for c in calls]
    typers, want = [], []
    for seed in (12, 13):
        genes = synthetic.make_wgs30x_sample(seed=seed, total_reads=20000)
        packs, cns = [packing.pack_synthetic(g) for g in genes], [g.cn for g in genes]
        typers.append(cohort.BatchTyper(packs, cns, top_n=300, backend=cuda))
        want.append(key(cohort.BatchTyper(packs, cns, top_n=300, backend=cuda).run()))
    for _ in range(4):                                   # eager passes, then graph replays
        typers[0].start()
        typers[1].start()
        assert key(typers[0].finish()) == want[0]
        assert key(typers[1].finish()) == want[1]


def test_pass_pipeline_over_replicas(cuda):
    """Consecutive passes with two in flight (cohort.PassPipeline): replicas with device buffers and
    streams of their own, resident and with the host->device copies inside every pass; and two passes
    of ONE typer in flight on its stream.  Every pass must give the calls of a plain run()."""
    from kir_graph_b200 import cohort
    key = lambda calls: [(c.gene, tuple(c.alleles), c.best_rank, c.score, c.tie_flags) for c in calls]
    packs, cns = [], []
    for seed in (21, 22, 23):
        genes = synthetic.make_wgs30x_sample(seed=seed, total_reads=15000)
        packs += [packing.pack_synthetic(g) for g in genes]
        cns += [g.cn for g in genes]
    want = key(cohort.BatchTyper(packs, cns, top_n=300, backend=cuda).run())
    for n_parts in (1, 3):
        typer = cohort.CohortTyper(packs, cns, top_n=300, backend=cuda, n_parts=n_parts, group_size=17, own_stream=True)
        assert typer.streams is not None and len(typer.streams) == n_parts
        twin = typer.replica(packs, cns, top_n=300, group_size=17)
        assert all(a.host is b.host for a, b in zip(typer.parts, twin.parts))
        for typers, depth, upload in (([typer, twin], None, False), ([typer, twin], None, True), ([typer], 2, False)):
            pipe = cohort.PassPipeline(typers, upload=upload, depth=depth)
            out = []
            for _ in range(7):                           # eager passes, graph capture, replays
                done = pipe.submit()
                if done is not None:
                    out.append(done)
            out += pipe.drain()
            assert len(out) == 7 and all(key(calls) == want for calls in out), (n_parts, depth, upload)
        assert all(getattr(p, "graph_error", None) is None for t in (typer, twin) for p in t.parts)
        assert all(p._graph is not None for t in (typer, twin) for p in t.parts)


@pytest.mark.parametrize("packed", [False, True])
def test_many_observations_per_read(cuda, packed):
    """Up to 255 observations per read pair stay exact on both scoring paths."""
    gene = synthetic.make_gene([91, 0], "KIRWIDE*BACKBONE", 30, 400, 2, 1500, w=100)
    pack = packing.pack_synthetic(gene)
    assert pack.k_obs.max() > 128
    batch = engine.MatrixBatch([pack], backend=cuda, packed=packed)
    m = batch.mismatch_counts(0)
    search = orc.IntSearch(m.astype(np.int64), pack.k_obs, top_n=40)
    group = engine.SearchGroup(batch, [0], 40)
    for step in range(2):
        out = group.step(need_next=[step < 1])[0]
        ref = search.add_candidate()
        assert np.array_equal(out.ids, ref.allele_id) and np.array_equal(out.score, ref.score)


def _cap_observations(csr, cap):
    """The reads with their lists cut from the end (rnv, lnv, rpv, lpv) to at most ``cap`` observations."""
    names = ("lpv", "rpv", "lnv", "rnv")
    lens = {n: np.diff(csr.offsets[n]) for n in names}
    excess = np.maximum(sum(lens.values()) - cap, 0)
    for n in reversed(names):
        cut = np.minimum(excess, lens[n])
        lens[n] = lens[n] - cut
        excess = excess - cut
    offsets, indices = {}, {}
    for n in names:
        off = np.zeros(csr.n_reads + 1, dtype=np.int64)
        np.cumsum(lens[n], out=off[1:])
        row = np.repeat(np.arange(csr.n_reads), lens[n])
        indices[n] = csr.indices[n][csr.offsets[n][:-1][row] + (np.arange(off[-1]) - off[:-1][row])]
        offsets[n] = off
    return type(csr)(csr.n_reads, offsets, indices)


@pytest.mark.parametrize("cap", [127, 128, 255])
def test_tie_counts_and_p_on_both_byte_predicate_paths(cuda, cap):
    """The tie-counting and P kernels take cheaper byte predicates when every count of the matrix is below
    128 (GkMatrix.m_max): reads with at most 127 observations take that path with the largest value it
    allows, 128 and 255 take the general one; two, three and four members, ids / scores / fraction
    numerators / P against the oracle's integer search."""
    import dataclasses
    gene = synthetic.make_gene([92, 0], "KIRWIDE*BACKBONE", 30, 400, 4, 2500, w=100)
    pack = packing.pack_synthetic(dataclasses.replace(gene, reads=_cap_observations(gene.reads, cap)))
    assert pack.k_obs.max() == min(cap, 200)          # the generator's reads carry about 200 observations
    batch = engine.MatrixBatch([pack], backend=cuda)
    assert int(batch.host.table["m_max"][0]) == int(pack.k_obs.max())
    m = batch.mismatch_counts(0)
    search = orc.IntSearch(m.astype(np.int64), pack.k_obs, top_n=40)
    group = engine.SearchGroup(batch, [0], 40)
    for step in range(4):
        out = group.step(need_next=[step < 3])[0]
        ref = search.add_candidate()
        assert np.array_equal(out.ids, ref.allele_id) and np.array_equal(out.score, ref.score)
        w = np.array([orc.lcm_upto(out.n) // q for q in range(1, out.n + 1)])
        assert np.array_equal((out.cnt * w[None, None, :]).sum(axis=2), ref.frac_num)
    assert np.array_equal(group.materialize_p(0, out.ids), ref.allele_prob)


def test_scattered_observations_overflow_the_staged_entries(cuda):
    """Reads whose observations are scattered over the variant table produce more than 1024 entries per
    64-read tile, which takes the likelihood kernel's unstaged path."""
    from kir_graph_b200.hisat2 import PairRead
    from kir_graph_b200.msa2hisat import Variant
    rng = np.random.default_rng(3)
    g = "KIRSCAT*BACKBONE"
    n_var, n_allele = 3000, 45
    variants = [Variant(pos=7 * i, typ="single", ref=g, val="ACGT"[i % 4], id=f"hv{i}",
                        allele=[f"KIRSCAT*{a:03d}" for a in np.flatnonzero(rng.random(n_allele) < 0.3)])
                for i in range(n_var)]
    reads = []
    for _ in range(300):
        ids = [f"hv{i}" for i in rng.choice(n_var, size=48, replace=False)]
        reads.append(PairRead(backbone=g, lpv=ids[:12], lnv=ids[12:24], rpv=ids[24:36], rnv=ids[36:]))
    pack = packing.pack_gene(reads, variants, variant_correction=False)
    assert pack.n_entries / pack.n_reads > 16
    batch = engine.MatrixBatch([pack], backend=cuda)
    by_id = {v.id: v for v in variants}
    col = {n: i for i, n in enumerate(pack.allele_names)}
    m, k = orc.mismatch_counts(reads, by_id, col)
    assert np.array_equal(batch.mismatch_counts(0), m)
    assert np.array_equal(batch.blocked_counts(0), m)
    assert np.array_equal(batch.colsum(0), m.sum(axis=0))


def test_edge_cases_on_gpu(cuda):
    from tests import test_edge_cases as edge
    for name in ("one_allele", "two_alleles"):
        for top_n in (1, 5):
            edge.test_tiny_universe(name, top_n, backend=cuda)
    edge.test_ragged_batch_with_empty_and_zero_cn(backend=cuda)


def test_full_size_wgs30x_sample_properties(cuda):
    """cfg3 at full size (200k read pairs, 17 genes): calls recover the generator's truth, reruns are
    bit-identical, the pipelined and the step-by-step drivers agree, and the called set's score equals
    an independent recomputation from the mismatch matrix read back from the device."""
    from kir_graph_b200 import cohort
    genes = synthetic.make_wgs30x_sample(seed=3)
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = [g.cn for g in genes]
    typer = cohort.BatchTyper(packs, cns, top_n=300, backend=cuda)
    a = typer.run()
    b = typer.run()
    assert [(c.alleles, c.score, c.best_rank) for c in a] == [(c.alleles, c.score, c.best_rank) for c in b]
    typer.pipelined = False
    c_step = typer.run()
    assert [(c.alleles, c.score, c.tie_flags) for c in a] == [(c.alleles, c.score, c.tie_flags) for c in c_step]
    for gene, call in zip(genes, a):
        assert sorted(call.alleles) == sorted(gene.allele_names[t] for t in gene.truth)
    for i in (0, 4, 9):                                   # CN 2, 3, 4
        m = typer.batch.mismatch_counts(i).astype(np.int64)
        want = m[:, a[i].ids].min(axis=1).sum() * (cns[i] if a[i].homozygous else 1)
        if a[i].homozygous:
            want = m[:, a[i].ids[0]].sum() * cns[i]
        assert a[i].score == want


def test_caller_side_files_equal_the_references(cuda, tmp_path, monkeypatch):
    """kir_graph_b200.main on the GPU: per-sample alleleTyping and the batched cohortAlleleTyping write
    the .tsv files of the reference's graphkir.main.alleleTyping (tests/golden/main_tsv.json.gz)."""
    import pandas as pd
    from kir_graph_b200 import main
    from kir_graph_b200.hisat2 import PairRead, writeReadsAndVariantsData
    from kir_graph_b200.msa2hisat import Variant
    data = load_golden("main_tsv")
    monkeypatch.chdir(tmp_path)
    names, cn_files = [], []
    for inp in data["inputs"]:
        writeReadsAndVariantsData({"variants": [Variant(**v) for v in inp["variants"]],
                                   "reads": [PairRead(**r) for r in inp["reads"]]}, inp["name"] + ".json")
        pd.DataFrame({"gene": list(inp["cn"]), "cn": list(inp["cn"].values())}).to_csv(
            inp["name"] + ".depth.cn.tsv", sep="\t", index=False)
        names.append(inp["name"])
        cn_files.append(inp["name"] + ".depth.cn.tsv")
    want = data["methods"]["full"]
    files = main.cohortAlleleTyping(names, cn_files, "full", n_parts=2, _backend=cuda)
    assert files == want["files"] and [open(f).read() for f in files] == want["tsv"]
    files = main.alleleTyping(names, cn_files, "full", _backend=cuda, _fast=True)
    assert files == want["files"] and [open(f).read() for f in files] == want["tsv"]
    files = main.alleleTyping(names, cn_files, "exonfirst", _backend=cuda)
    assert [open(f).read() for f in files] == data["methods"]["exonfirst"]["tsv"]


def test_wire_expansion_on_the_device_equals_the_numpy_statement(cuda):
    """gk_expand_reads: entry offsets and 16-byte entries rebuilt on the GPU from the wire records, regular
    reads and raw records, against tests/fake_backend.py; and the likelihood built from them against the
    likelihood of host-built entries."""
    from tests.test_wire_format import _counts_by_set_logic, _random_pack
    rng = np.random.default_rng(12)
    weird, member = _random_pack(rng, n_reads=900, n_var=700, n_allele=70, weird=True)
    packs = _packs([(40, 2, 1500), (7, 1, 300), (150, 3, 700)], 55) + [weird]
    fake = FakeBackend()
    hb = engine.HostBatch(packs, wire=True)
    assert int(((hb.hdr >> 8) == 0).sum()) > 0                       # raw records are exercised
    bg, bf = engine.MatrixBatch(hb, backend=cuda), engine.MatrixBatch(hb, backend=fake)
    assert np.array_equal(cuda.download(bg.d_entoff, np.int32)[: hb.n_offsets], bf.d_entoff[: hb.n_offsets])
    assert np.array_equal(cuda.download(bg.d_ent, np.uint32)[: 4 * hb.n_entries], bf.d_ent[: 4 * hb.n_entries])
    assert np.array_equal(cuda.download(bg.d_LT, np.uint8), bf.d_LT)
    assert np.array_equal(bg.mismatch_counts(3), _counts_by_set_logic(weird, member))
    legacy = engine.MatrixBatch(engine.HostBatch(packs, wire=False), backend=cuda)
    for name in ("d_LT", "d_L", "d_col"):
        assert np.array_equal(cuda.download(getattr(bg, name)), cuda.download(getattr(legacy, name))), name


@pytest.mark.parametrize("a", [36, 52, 100])
def test_every_row_remainder_of_the_packed_tiles(cuda, a):
    """Kept-set counts 8 g - 3 for g = 1 .. 16 row groups (every entry of SearchGroup._W_CUT, padded pieces
    included) under one (36, 52 alleles) or two (100) warp-split column tiles: scores and kept sets of the
    third step against the oracle's integer search."""
    gene = synthetic.make_gene([93, a], "KIRROW*BACKBONE", a, 4 * a, 3, 1500)
    pack = packing.pack_synthetic(gene)
    batch = engine.MatrixBatch([pack], backend=cuda)
    m = batch.mismatch_counts(0).astype(np.int64)
    for g in range(1, 17):
        top_n = 8 * g - 3
        search = orc.IntSearch(m, pack.k_obs, top_n=top_n)
        group = engine.SearchGroup(batch, [0], top_n)
        for step in range(3):
            out = group.step(need_next=[step < 2])[0]
            ref = search.add_candidate()
            assert np.array_equal(out.ids, ref.allele_id), (a, g, step)
            assert np.array_equal(out.score, ref.score), (a, g, step)
        assert len(ref.score) == top_n


@pytest.mark.parametrize("seed", range(8))
def test_random_batches_match_numpy_statement(cuda, seed):
    """Randomised shapes through every launch: allele counts 1..260 (all lane layouts of the likelihood kernel,
    full-width and warp-split scoring tiles with every row offset), 1..900 reads, copy numbers 1..5, ragged
    top_n, one gene of arbitrary reads (raw wire records) - CUDA against the NumPy statement, pool by pool."""
    from tests.test_wire_format import _random_pack
    rng = np.random.default_rng([2024, seed])
    n_genes = int(rng.integers(2, 6))
    specs = [(int(rng.choice([1, 2, 5, 8, 9, 16, 17, 31, 33, 47, 64, 65, 90, 127, 129, 200, 260])),
              int(rng.integers(1, 6)), int(rng.integers(1, 900))) for _ in range(n_genes)]
    packs = _packs(specs, 300 + seed)
    cns = [c for _, c, _ in specs]
    if seed % 2:
        weird, _ = _random_pack(rng, n_reads=int(rng.integers(50, 400)), n_var=500, n_allele=int(rng.integers(3, 80)),
                                weird=True)
        packs.append(weird)
        cns.append(int(rng.integers(1, 4)))
    top_n = int(rng.choice([5, 17, 64, 300]))
    fake = FakeBackend()
    bg, bf = engine.MatrixBatch(packs, backend=cuda), engine.MatrixBatch(packs, backend=fake)
    assert np.array_equal(cuda.download(bg.d_LT, np.uint8), bf.d_LT), "LT"
    assert np.array_equal(cuda.download(bg.d_L, np.float32).view(np.uint32), bf.d_L.view(np.uint32)), "L"
    assert np.array_equal(cuda.download(bg.d_col, np.uint64), bf.d_col), "colsum"
    ids = list(range(len(packs)))
    gg, gf = engine.SearchGroup(bg, ids, top_n), engine.SearchGroup(bf, ids, top_n)
    for step in range(max(cns)):
        active = np.array([c > step for c in cns])
        need = np.array([c > step + 1 for c in cns])
        og = gg.step(active=active, need_next=need)
        of = gf.step(active=active, need_next=need)
        if step:
            assert np.array_equal(cuda.download(gg.d_S, np.uint32), gf.d_S), f"S at step {step + 1}"
        _compare_outputs(og, of)
