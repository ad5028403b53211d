"""Wire format of the reads (csrc/gk_wire.cu): gk_wire_encode (host) -> the NumPy statement of
gk_expand_reads (tests/fake_backend.py) gives the mismatch counts of the canonical observation entries, on
the synthetic workloads and on reads the window form cannot express (raw records)."""
import numpy as np
import pytest

from kir_graph_b200 import engine, packing, synthetic
from kir_graph_b200.synthetic import LIST_NAMES, ReadCSR
from tests.fake_backend import FakeBackend


def _random_pack(rng, n_reads, n_var, n_allele, weird):
    """A gene whose reads are arbitrary lists: far-apart mates, positives beyond the window, and (``weird``)
    a variant twice in a list, positive and negative in one mate, windows of hundreds of variants."""
    member = rng.random((n_var, n_allele)) < 0.3
    lists = {name: [] for name in LIST_NAMES}
    for _ in range(n_reads):
        for pos_name, neg_name in (("lpv", "lnv"), ("rpv", "rnv")):
            kind = rng.integers(0, 10 if weird else 6)
            lo = int(rng.integers(0, n_var - 40))
            width = int(rng.integers(1, 40))
            window = np.arange(lo, min(lo + width, n_var))
            obs = window[rng.random(len(window)) < 0.85]              # some holes
            is_pos = rng.random(len(obs)) < 0.25
            pos, neg = list(obs[is_pos]), list(obs[~is_pos])
            if kind == 0:
                pos, neg = [], []                                     # empty mate
            elif kind == 1:
                pos = pos + [int(v) for v in rng.integers(0, n_var, size=2)]      # positives anywhere (novel-like)
                pos = [v for v in dict.fromkeys(pos) if v not in neg]
            elif kind == 2:
                neg = []                                              # positives only
            elif kind == 6 and pos:
                pos = pos + [pos[0]]                                  # a positive twice
            elif kind == 7 and neg:
                neg = neg + [neg[-1], neg[-1]]                        # a negative three times
            elif kind == 8 and neg:
                pos = pos + [neg[0]]                                  # positive and negative in one mate
            elif kind == 9:
                neg = sorted(set(neg) | {int(v) for v in rng.integers(0, n_var, size=3)})   # huge window
                pos = [v for v in pos if v not in neg]
            lists[pos_name].append(pos)
            lists[neg_name].append(neg)
    offsets, indices = {}, {}
    for name in LIST_NAMES:
        lens = np.array([len(x) for x in lists[name]], dtype=np.int64)
        off = np.zeros(n_reads + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        offsets[name] = off
        indices[name] = np.array([v for x in lists[name] for v in x], dtype=np.int32)
    csr = ReadCSR(n_reads, offsets, indices)
    names = [f"KIRW*{i:03d}" for i in range(n_allele)]
    pack, _ = packing._finish("KIRW*BACKBONE", names, [f"hv{v}" for v in range(n_var)], member, csr,
                              variant_correction=False, no_empty=False)
    return pack, member


def _counts_by_set_logic(pack, member):
    m = np.zeros((pack.n_reads, pack.n_alleles), dtype=np.int64)
    for name in LIST_NAMES:
        off, idx = pack.csr.offsets[name], pack.csr.indices[name]
        row = np.repeat(np.arange(pack.n_reads), np.diff(off))
        np.add.at(m, row, (~member[idx] if name in ("lpv", "rpv") else member[idx]).astype(np.int64))
    return m


@pytest.mark.parametrize("weird", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_wire_round_trip_gives_the_canonical_counts(seed, weird):
    rng = np.random.default_rng([77, seed, int(weird)])
    pack, member = _random_pack(rng, n_reads=300 + 37 * seed, n_var=700, n_allele=21 + 40 * seed, weird=weird)
    want = _counts_by_set_logic(pack, member)
    wire = packing.wire_encode(pack)
    raw = int(((wire.hdr >> 8) == 0).sum())
    assert (raw > 0) == weird or not weird              # raw records appear only for what windows cannot express
    for use_wire in (True, False):
        batch = engine.MatrixBatch(engine.HostBatch([pack], wire=use_wire), backend=FakeBackend())
        assert np.array_equal(batch.mismatch_counts(0), want), f"wire={use_wire}"
        assert np.array_equal(batch.colsum(0), want.sum(axis=0))
    if weird:
        assert raw > 0


def test_wire_is_compact_on_the_wgs_workload_and_keeps_the_entry_count():
    genes = synthetic.make_wgs30x_sample(seed=5, scale=0.02)
    packs = [packing.pack_synthetic(g) for g in genes]
    wires = [packing.wire_encode(p) for p in packs]
    reads = sum(p.n_reads for p in packs)
    assert sum(w.nbytes for w in wires) / reads < 16.0             # against 36 B of entries + offsets per read pair
    canon = sum(p.n_entries for p in packs)
    assert canon <= sum(w.n_entries for w in wires) <= 1.01 * canon
    assert all(int(((w.hdr >> 8) == 0).sum()) == 0 for w in wires)  # no raw records on regular reads
    a = engine.MatrixBatch(engine.HostBatch(packs, wire=True), backend=FakeBackend())
    b = engine.MatrixBatch(engine.HostBatch(packs, wire=False), backend=FakeBackend())
    assert np.array_equal(a.d_LT, b.d_LT) and np.array_equal(a.d_L, b.d_L) and np.array_equal(a.d_col, b.d_col)
    reads_bytes = lambda h: h.nbytes - h.mem.nbytes - h.table.nbytes          # what grows with the reads
    assert reads_bytes(a.host) < 0.45 * reads_bytes(b.host)


def test_wire_of_a_read_shard_and_of_an_empty_problem():
    gene = synthetic.make_gene([9, 1], "KIRS*BACKBONE", 30, 240, 2, 700)
    pack = packing.pack_synthetic(gene)
    whole = engine.MatrixBatch(engine.HostBatch([pack], wire=True), backend=FakeBackend()).mismatch_counts(0)
    parts = [packing.shard_reads(pack, r, 3) for r in range(3)]
    got = np.concatenate([engine.MatrixBatch(engine.HostBatch([p], wire=True), backend=FakeBackend()).mismatch_counts(0)
                          for p in parts])
    assert np.array_equal(got, whole)
    empty = packing.shard_reads(pack, 0, 10 ** 6)                  # a shard without reads still takes part
    assert empty.n_reads == 0
    batch = engine.MatrixBatch(engine.HostBatch([empty, pack], wire=True), backend=FakeBackend())
    assert np.array_equal(batch.mismatch_counts(1), whole) and batch.colsum(0).sum() == 0


def test_reads_without_observations():
    """no_empty=False keeps read pairs without any observation (a row of 0.999 in the reference,
    typing_mulit_allele.py:372-374): no entries at all, counts of zero, on both input paths."""
    member = np.random.default_rng(3).random((40, 5)) < 0.5
    zero = np.zeros(6, dtype=np.int64)
    csr = ReadCSR(5, {n: zero.copy() for n in LIST_NAMES}, {n: np.zeros(0, np.int32) for n in LIST_NAMES})
    pack, _ = packing._finish("KIRE*BACKBONE", [f"KIRE*{i}" for i in range(5)], [f"hv{v}" for v in range(40)], member, csr,
                              variant_correction=False, no_empty=False)
    for use_wire in (True, False):
        batch = engine.MatrixBatch(engine.HostBatch([pack], wire=use_wire), backend=FakeBackend())
        assert batch.mismatch_counts(0).shape == (5, 5) and not batch.mismatch_counts(0).any()
