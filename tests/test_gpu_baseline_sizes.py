"""Parity at BASELINE.json's sizes, on CUDA, against the oracle's exact integer search (`-m gpu`):

* every gene of a full-size cfg3 sample (200k read pairs, 900 alleles / 17 genes, top_n 300): kept allele
  ids, scores, fraction numerators, N_uniq and tie flags of every copy-number step;
* a cfg4-shaped problem (50k read pairs x 1000 alleles, CN 6, top_n 300), the same quantities;
* the capacity limits of the device path, on both sides of each limit.

The oracle needs about a minute for the first and two for the second on the GPU box's host."""
import numpy as np
import pytest

from kir_graph_b200 import cohort, engine, packing, synthetic
from oracle import typing_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    return engine.CudaBackend()


def _check_step(out, ref, label):
    n = out.n
    assert np.array_equal(out.ids, ref.allele_id), f"{label}: kept allele ids"
    assert np.array_equal(out.score, ref.score), f"{label}: scores"
    w = np.array([orc.lcm_upto(n) // q for q in range(1, n + 1)])
    assert np.array_equal((out.cnt * w[None, None, :]).sum(axis=2), ref.frac_num), f"{label}: fraction numerators"
    assert out.n_unique == ref.n_unique, f"{label}: N_uniq"
    kinds = {k for k, _, _ in ref.ties}
    if n >= 2:
        # bit0: the tie group at the M = max(top_n, N_uniq // 5) cut matters only when it is the group of the
        # top_n-th score (sets that can still reach the final top_n); a tie between worse scores at the cut is
        # reported by the oracle but cannot change the result
        cut_scores = {sc for k, _, sc in ref.ties if k == "cut"}
        relevant = bool(cut_scores) and int(ref.score[-1]) in cut_scores
        assert bool(out.tie_flags & 1) == relevant, f"{label}: tie at the M cut"
        assert bool(out.tie_flags & 2) >= ("rank-cut" in kinds), f"{label}: tie at the top_n cut"
    assert bool(out.tie_flags & 4) == ("best" in kinds), f"{label}: rank 0 tied with rank 1"


@pytest.mark.timeout(1500)
def test_full_size_wgs30x_every_gene_equals_int_search(cuda):
    genes = synthetic.make_wgs30x_sample(seed=3)
    packs = [packing.pack_synthetic(g) for g in genes]
    cns = np.array([g.cn for g in genes])
    batch = engine.MatrixBatch(packs, backend=cuda)
    group = engine.SearchGroup(batch, list(range(len(packs))), 300)
    searches = [orc.IntSearch(batch.mismatch_counts(i).astype(np.int64), p.k_obs, top_n=300)
                for i, p in enumerate(packs)]
    for step in range(int(cns.max())):
        outs = group.step(active=cns > step, need_next=cns > step + 1)
        for i in np.flatnonzero(cns > step):
            _check_step(outs[int(i)], searches[i].add_candidate(), f"{genes[i].gene} step {step + 1}")
    # and the batched typer (pipelined, selectBest on the device) calls what the oracle's last step says
    calls = cohort.BatchTyper(packs, list(cns), top_n=300, backend=cuda).run()
    for g, p, c, s in zip(genes, packs, calls, searches):
        if c.homozygous:
            continue
        last = s.result[-1]
        ok = np.flatnonzero((last.fraction >= 0.5 / last.n).all(axis=1))
        best = int(ok[0]) if len(ok) else 0
        assert c.best_rank == best and c.ids == last.allele_id[best].tolist() and c.score == int(last.score[best])


@pytest.mark.timeout(2400)
def test_deep_shaped_problem_equals_int_search(cuda):
    gene = synthetic.make_deep_sample(n_reads=50_000, n_allele=1000, n_var=8000, cn=6)
    pack = packing.pack_synthetic(gene)
    batch = engine.MatrixBatch([pack], backend=cuda)
    search = orc.IntSearch(batch.mismatch_counts(0).astype(np.int64), pack.k_obs, top_n=300, read_chunk=256)
    group = engine.SearchGroup(batch, [0], 300)
    for step in range(6):
        _check_step(group.step(need_next=[step < 5])[0], search.add_candidate(), f"deep step {step + 1}")
    call = cohort.BatchTyper([pack], [6], top_n=300, backend=cuda).run()[0]
    assert sorted(call.alleles) == sorted(gene.allele_names[t] for t in gene.truth)


def _wide_read_pack(k_obs, n_allele=6, n_var=600):
    """A gene with one read pair of exactly ``k_obs`` observations among ordinary ones."""
    from kir_graph_b200.synthetic import LIST_NAMES, ReadCSR
    rng = np.random.default_rng(k_obs)
    member = rng.random((n_var, n_allele)) < 0.4
    lists = {n: [] for n in LIST_NAMES}
    for r in range(40):
        lo = int(rng.integers(0, n_var - 30))
        obs = np.arange(lo, lo + 24)
        is_pos = member[obs, r % n_allele]
        lists["lpv"].append(list(obs[:12][is_pos[:12]])); lists["lnv"].append(list(obs[:12][~is_pos[:12]]))
        lists["rpv"].append(list(obs[12:][is_pos[12:]])); lists["rnv"].append(list(obs[12:][~is_pos[12:]]))
    half = k_obs // 2
    lists["lpv"].append([]); lists["lnv"].append(list(range(0, half)))
    lists["rpv"].append([]); lists["rnv"].append(list(range(half, k_obs)))
    offsets, indices = {}, {}
    for name in LIST_NAMES:
        lens = np.array([len(x) for x in lists[name]], dtype=np.int64)
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        offsets[name], indices[name] = off, np.array([v for x in lists[name] for v in x], dtype=np.int32)
    csr = ReadCSR(41, offsets, indices)
    pack, _ = packing._finish("KIRWIDE*BACKBONE", [f"KIRWIDE*{i:03d}" for i in range(n_allele)],
                              [f"hv{v}" for v in range(n_var)], member, csr, variant_correction=False, no_empty=True)
    return pack, member


def test_limit_255_observations_per_read_pair(cuda):
    pack, member = _wide_read_pack(255)
    m = engine.MatrixBatch([pack], backend=cuda).mismatch_counts(0)
    want = member[:255].sum(axis=0)                                   # the wide pair: 255 negatives
    assert np.array_equal(m[-1].astype(np.int64), want) and int(pack.k_obs[-1]) == 255
    ref = orc.IntSearch(m.astype(np.int64), pack.k_obs, top_n=10)
    group = engine.SearchGroup(engine.MatrixBatch([pack], backend=cuda), [0], 10)
    for step in range(2):
        _check_step(group.step(need_next=[step < 1])[0], ref.add_candidate(), f"K_r 255 step {step + 1}")
    with pytest.raises(packing.CapacityError):
        _wide_read_pack(256)


def test_limit_copy_number_eight_and_wide_dedup_keys(cuda):
    """CN 8 with 1000 alleles: 7 previous ids x 10 bits = 70 bits of canonical key (a 64-bit key refused this
    in round 1); CN 9 is refused per gene with CapacityError."""
    gene = synthetic.make_gene([5, 8], "KIRCN8*BACKBONE", 1000, 2000, 8, 900, homo_prob=0.0)
    pack = packing.pack_synthetic(gene)
    batch = engine.MatrixBatch([pack], backend=cuda)
    ref = orc.IntSearch(batch.mismatch_counts(0).astype(np.int64), pack.k_obs, top_n=40, read_chunk=64)
    group = engine.SearchGroup(batch, [0], 40)
    for step in range(8):
        _check_step(group.step(need_next=[step < 7])[0], ref.add_candidate(), f"CN 8 step {step + 1}")
    with pytest.raises(packing.CapacityError):
        group.step()
    assert engine.capacity_violation(1000, 9, 300) and engine.capacity_violation(1000, 8, 300) is None
    assert engine.capacity_violation(10, 2, 2049) and engine.capacity_violation(10, 2, 2048) is None
