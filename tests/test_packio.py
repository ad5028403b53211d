import os

import numpy as np

from kir_graph_b200 import cohort, packing, packio
from kir_graph_b200.hisat2 import writeReadsAndVariantsData
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden, objects_from_input


def test_sidecar_roundtrip_and_cohort_calls(tmp_path):
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "s.variant.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    packs = packio.pack_variant_json(path, variant_correction=True)
    side = os.path.join(tmp_path, "s.gkpack.npz")
    packio.save_packs(side, packs, {"variant_correction": True})
    back, meta = packio.load_packs(side)
    assert meta == {"variant_correction": True} and list(back) == list(packs)
    for g in packs:
        for name in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "obs_pos", "obs_neg"):
            assert np.array_equal(getattr(packs[g], name), getattr(back[g], name))
        assert packs[g].allele_names == back[g].allele_names and packs[g].var_val == back[g].var_val
        # the wire form of the reads travels with the sidecar: equal to a fresh encoding of the same lists
        stored, packs[g].wire = back[g].wire, None
        fresh = packing.wire_encode(packs[g])
        for name in ("hdr", "stream", "neg_keep", "tile_stream", "tile_entry"):
            assert np.array_equal(getattr(stored, name), getattr(fresh, name)), (g, name)
        assert stored.n_entries == fresh.n_entries
        for name in ("lpv", "rpv", "lnv", "rnv"):
            assert np.array_equal(packs[g].csr.indices[name], back[g].csr.indices[name])
            assert np.array_equal(packs[g].csr.offsets[name], back[g].csr.offsets[name])
    genes = [g for g, cn in sample["gene_cn"].items() if cn and g in back]
    typer = cohort.BatchTyper([back[g] for g in genes], [sample["gene_cn"][g] for g in genes], top_n=60,
                              backend=FakeBackend())
    calls = typer.run()
    got = [a for c in calls for a in c.alleles]
    assert got == sample["calls"]["full"]["alleles"] or any(c.tie_flags for c in calls)


def test_sidecar_of_another_format_or_stale_json_is_not_used(tmp_path):
    """A sidecar without the format-3 index (older layout), or one whose meta no longer matches the .json,
    is ignored: the packs come from the .json."""
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "s.variant.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    want = packio.pack_variant_json(path, variant_correction=True)
    side = packio.sidecar_path(path)
    np.savez_compressed(side, __genes__=np.array("[]"), __meta__=np.array("{}"))     # format-2 style file
    got = packio.load_sample_packs(path)
    assert list(got) == list(want) and all(np.array_equal(got[g].ent_pos, want[g].ent_pos) for g in want)
    packio.save_packs(side, want, packio.sidecar_meta(path))
    assert all(p.wire is not None for p in packio.load_sample_packs(path).values())       # fresh: used
    os.utime(path, ns=(1, 1))                                                               # .json touched
    assert all(p.wire is None for p in packio.load_sample_packs(path).values())           # stale: rebuilt from .json


def test_truncated_or_garbled_sidecar_falls_back_to_json(tmp_path):
    """A sidecar cut short (a crashed writer, a full disk) or overwritten with other bytes is not an error:
    the packs come from the .json."""
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "s.variant.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    want = packio.pack_variant_json(path, variant_correction=True)
    side = packio.sidecar_path(path)
    packio.save_packs(side, want, packio.sidecar_meta(path))
    whole = open(side, "rb").read()
    for damaged in (whole[: len(whole) // 2], whole[:100], b"", b"PK\x03\x04" + b"\0" * 64, b"not a zip file"):
        with open(side, "wb") as f:
            f.write(damaged)
        got = packio.load_sample_packs(path)
        assert list(got) == list(want) and all(p.wire is None for p in got.values())
        assert all(np.array_equal(got[g].ent_pos, want[g].ent_pos) for g in want)


def test_sidecar_round_trip_over_ragged_and_empty_genes(tmp_path):
    """Format 3 stores the arrays of all genes back to back per array name: genes without reads, with a single
    read, with reads that observe nothing after correction, and of very different sizes come back array for
    array, and type to the same calls."""
    from kir_graph_b200 import synthetic
    from kir_graph_b200.hisat2 import PairRead
    specs = [(4, 64, 1, 0), (9, 72, 2, 1), (30, 240, 3, 400), (2, 64, 1, 3), (17, 136, 2, 90), (6, 64, 2, 0)]
    packs = {}
    for i, (a, v, cn, r) in enumerate(specs):
        gene = synthetic.make_gene([81, i], f"KIRR{i}*BACKBONE", a, v, cn, max(r, 1), hierarchical=bool(i % 2))
        reads, variants = gene.to_objects()
        if r == 0:
            reads = []
        if i == 3:                                               # reads that carry nothing
            reads = [PairRead(backbone=gene.gene) for _ in range(3)]
        packs[gene.gene] = packing.pack_gene(reads, variants, variant_correction=bool(i % 2), gene=gene.gene,
                                             no_empty=i != 3)
    side = os.path.join(tmp_path, "r.gkpack.npz")
    packio.save_packs(side, packs, {"k": 1})
    back, meta = packio.load_packs(side)
    assert meta == {"k": 1} and list(back) == list(packs)
    for g, p in packs.items():
        q = back[g]
        assert (q.n_reads, q.n_alleles, q.n_variants, q.n_words) == (p.n_reads, p.n_alleles, p.n_variants, p.n_words)
        for name in packio._ARRAYS:
            a, b = np.asarray(getattr(p, name)), np.asarray(getattr(q, name))
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (g, name)
        for name in packio._WIRE_ARRAYS:
            a, b = np.asarray(getattr(p.wire, name)), np.asarray(getattr(q.wire, name))
            assert a.dtype == b.dtype and np.array_equal(a, b), (g, name)
        assert q.wire.n_entries == p.wire.n_entries and q.allele_names == p.allele_names and q.variant_ids == p.variant_ids
    cns = [c for (_, _, c, _) in specs]
    key = lambda calls: [(c.gene, c.alleles, c.score, c.tie_flags, c.n_reads) for c in calls]
    want = key(cohort.BatchTyper(list(packs.values()), cns, top_n=20, backend=FakeBackend()).run())
    got = key(cohort.BatchTyper(list(back.values()), cns, top_n=20, backend=FakeBackend()).run())
    assert got == want and any(c[1][0] == "fail" for c in got) and any(c[1][0] != "fail" for c in got)
