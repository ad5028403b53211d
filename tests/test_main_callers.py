"""The callers of the typing path (kir_graph_b200/main.py): alleleTyping writes the reference's
per-sample files, cohortAlleleTyping types all samples as one batch and writes the same bytes;
sharded over ranks the merged cohort.allele.tsv is the same."""
import io
import os

import numpy as np
import pandas as pd
import pytest

from kir_graph_b200 import main
from kir_graph_b200.hisat2 import PairRead, writeReadsAndVariantsData
from kir_graph_b200.msa2hisat import Variant
from tests.cohort_sim import write_cohort
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden


def _cohort(tmp_path, n_samples=3):
    names, cn_files, _ = write_cohort(str(tmp_path), n_samples)
    return names, cn_files


@pytest.mark.parametrize("method", ["full", "exonfirst"])
def test_allele_typing_files_equal_the_references(tmp_path, monkeypatch, method):
    """tests/golden/main_tsv.json.gz: graphkir.main.alleleTyping + mergeAllele of the reference on a
    three-sample cohort (make_golden_main.py).  Same file names and bytes for {name}.tsv and the merged
    cohort file; the .possible.tsv rows equal up to the order of exactly tied sets."""
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
    want = data["methods"][method]
    files = main.alleleTyping(names, cn_files, method, _backend=FakeBackend())
    assert files == want["files"]
    assert [open(f).read() for f in files] == want["tsv"]
    main.mergeAllele(files, "cohort.allele.tsv")
    assert open("cohort.allele.tsv").read() == want["merged"]
    for f, ref in zip(files, want["possible"]):
        got = pd.read_csv(f[:-4] + ".possible.tsv", sep="\t").fillna("")
        exp = pd.read_csv(io.StringIO(ref), sep="\t").fillna("")
        assert list(got.columns) == list(exp.columns) and len(got) == len(exp)
        np.testing.assert_allclose(got["value"], exp["value"], rtol=1e-11)
        cols = [c for c in got.columns if c.isdigit()]
        rows = lambda df: sorted((g, round(v, 6), tuple(sorted(map(str, r)))) for g, v, r in
                                 zip(df["gene"], df["value"], df[cols].values.tolist()))
        assert rows(got) == rows(exp)
    if method == "full":
        for f in files:
            os.remove(f)
        assert main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend()) == want["files"]
        assert [open(f).read() for f in files] == want["tsv"]


def test_helpers():
    assert main.getCommonName("data/x_30x.00.read.r1.fq", "data/x_30x.00.read.r2.fq") == "data/x_30x.00.read"
    assert main.getCommonName("a.b.c", "a.b.c") == "a.b.c" and main.getCommonName("a.b", "c.b") == ""
    assert main._suffix("d/s.00.variant", "d/s.00.variant.depth.cn.tsv", "full") == ".cn_depth_cn_tsv.full"


def test_cohort_entry_writes_the_same_files(tmp_path):
    names, cn_files = _cohort(tmp_path)
    assert main.loadCN(cn_files[0])["KIRM0*BACKBONE"] == 2
    per_sample = main.alleleTyping(names, cn_files, "full", _backend=FakeBackend())
    want = [open(f, "rb").read() for f in per_sample]
    assert all(os.path.exists(f[:-4] + ".possible.tsv") for f in per_sample)
    rows = [pd.read_csv(f, sep="\t").fillna("") for f in per_sample]
    assert rows[0]["alleles"][0].endswith("KIRNONE*") and "KIRNONE*BACKBONE" in rows[0]["warnings"][0]
    assert "KIRM1*BACKBONE" in rows[1]["warnings"][0] and rows[2]["warnings"][0] == ""
    assert len(rows[2]["alleles"][0].split("_")) == 5 and len(rows[1]["alleles"][0].split("_")) == 6
    merged = main.mergeAllele(per_sample, str(tmp_path / "cohort.allele.tsv"))
    assert list(merged.columns) == ["name", "alleles", "warnings"] and len(merged) == 3
    want_merged = open(tmp_path / "cohort.allele.tsv", "rb").read()
    for f in per_sample:
        os.remove(f)
    # one rank
    files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend(), n_parts=2)
    assert files == per_sample and [open(f, "rb").read() for f in files] == want
    for f in per_sample:
        os.remove(f)
    # a rank's samples in several GPU passes of bounded size: same files
    for step in (1, 2):
        files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend(), batch_samples=step)
        assert files == per_sample and [open(f, "rb").read() for f in files] == want
        for f in per_sample:
            os.remove(f)
    # a batch that exceeds the entry range of the device tables is typed in halves
    from kir_graph_b200 import cohort as cohort_mod
    real_typer, built = cohort_mod.CohortTyper, []

    def small_batches_only(packs, cns, **kw):
        built.append(len(packs))
        if len(packs) > 3:
            raise ValueError("entry pool exceeds 2^31 entries; split the batch")
        return real_typer(packs, cns, **kw)

    cohort_mod.CohortTyper = small_batches_only
    try:
        files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    finally:
        cohort_mod.CohortTyper = real_typer
    assert files == per_sample and [open(f, "rb").read() for f in files] == want
    assert built[0] > 3 and len(built) >= 3 and all(n <= 3 for n in built[-2:])
    for f in per_sample:
        os.remove(f)
    # host preparation in worker processes
    files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend(), workers=2)
    assert files == per_sample and [open(f, "rb").read() for f in files] == want
    for f in per_sample:
        os.remove(f)
    # two ranks: each writes its own samples, rank 0 merges in input order
    for rank in (1, 0):
        files = main.cohortAlleleTyping(names, cn_files, "full", rank=rank, world=2, _backend=FakeBackend())
        assert files == per_sample
    assert [open(f, "rb").read() for f in files] == want
    main.mergeAllele(files, str(tmp_path / "cohort2.allele.tsv"))
    assert open(tmp_path / "cohort2.allele.tsv", "rb").read() == want_merged
    # the fast per-sample driver gives the same files too (a gene without variants included)
    for f in per_sample:
        os.remove(f)
    assert main.alleleTyping(names, cn_files, "full", _backend=FakeBackend(), _fast=True) == per_sample
    assert [open(f, "rb").read() for f in per_sample] == want
    with pytest.raises(NotImplementedError):
        main.cohortAlleleTyping(names, cn_files, "exonfirst_1", _backend=FakeBackend())
    assert main.cohortAlleleTyping([], [], "full", _backend=FakeBackend()) == []


def test_sidecar_written_at_extraction_feeds_the_cohort_entry(tmp_path):
    """extractVariantFromSam(write_pack=True) leaves {prefix}.gkpack.npz next to the .json; the cohort
    entry types from it (the .json is not parsed: here it is blanked, same size), falls back to the
    .json when the sidecar does not belong to it, and the calls are those of the .json route."""
    from kir_graph_b200 import hisat2, packio
    from tests import sam_sim
    names, cn_files = [], []
    for s in range(2):
        table, pairs = sam_sim.multi_gene(90 + s, n_pairs=260, novel=0.002)
        alleles = {g: [f"{g.split('*')[0]}*{i:03d}" for i in range(6)] for g in ("KIRA*BACKBONE", "KIRB*BACKBONE")}
        rng = np.random.default_rng(7 + s)
        for v in table:
            v.allele = [a for a in alleles[v.ref] if rng.random() < 0.4] or [alleles[v.ref][0]]
        name = str(tmp_path / f"c.{s:02d}.variant")
        with open(name + ".sam", "w") as f:
            f.write(sam_sim.sam_text(pairs))
        Variant.novel_id = 0
        hisat2.extractVariantFromSam(table, name + ".sam", name, num_editdist=9, write_pack=True)
        assert os.path.exists(name + ".gkpack.npz")
        pd.DataFrame({"gene": ["KIRA*BACKBONE", "KIRB*BACKBONE"], "cn": [2, 1]}).to_csv(
            name + ".cn.tsv", sep="\t", index=False)
        names.append(name)
        cn_files.append(name + ".cn.tsv")
    want_files = main.alleleTyping(names, cn_files, "full", _backend=FakeBackend())          # the .json route
    want = [open(f, "rb").read() for f in want_files]
    assert all(len(pd.read_csv(f, sep="\t")["alleles"][0].split("_")) == 3 for f in want_files)
    for f in want_files:
        os.remove(f)
    sizes = [os.path.getsize(n + ".json") for n in names]
    stamps = [os.stat(n + ".json").st_mtime_ns for n in names]
    for n, size, stamp in zip(names, sizes, stamps):        # blank the .json (size and time stamp kept): only
        with open(n + ".json", "wb") as f:                  # the sidecar can answer now
            f.write(b" " * size)
        os.utime(n + ".json", ns=(stamp, stamp))
    files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    assert files == want_files and [open(f, "rb").read() for f in files] == want
    # a .json rewritten to the same size at another time does not belong to the sidecar any more
    os.utime(names[1] + ".json", ns=(stamps[1] + 10 ** 9, stamps[1] + 10 ** 9))
    with pytest.raises(ValueError):                         # the blanked .json is what gets parsed now
        packio.load_sample_packs(names[1] + ".json")
    os.utime(names[1] + ".json", ns=(stamps[1], stamps[1]))
    # a .json of another size does not belong to the sidecar: the scanner is asked (and finds no reads)
    with open(names[0] + ".json", "wb") as f:
        f.write(b"{}")
    assert packio.load_sample_packs(names[0] + ".json") == {}
    side, meta = packio.load_packs(packio.sidecar_path(names[1] + ".json"))
    assert meta == {"variant_correction": True, "multiple": False, "json_size": sizes[1],
                    "json_mtime_ns": stamps[1], "format": packio.PACK_FORMAT} and len(side) == 2
    assert list(packio.load_sample_packs(names[1] + ".json", variant_correction=True)) == list(side)


def test_cohort_entry_survives_a_gene_the_homozygosity_rule_cannot_decide(tmp_path, caplog):
    """One gene of one sample with a site on which isHomozygous indexes an empty list (the reference's
    IndexError): the per-sample path raises like the reference, the cohort entry reports the gene, calls it
    fail and types everything else as before."""
    import json
    from tests.test_packing import _gene_with_an_undecidable_site
    names, cn_files = _cohort(tmp_path)
    want = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    rows_before = [pd.read_csv(f, sep="\t").fillna("") for f in want]
    reads, variants = _gene_with_an_undecidable_site("KIR3DL9*BACKBONE")
    data = json.load(open(names[1] + ".json"))
    from dataclasses import asdict
    data["reads"] += [asdict(r) for r in reads]
    data["variants"] += [asdict(v) for v in variants]
    json.dump(data, open(names[1] + ".json", "w"))
    cn = pd.read_csv(cn_files[1], sep="\t")
    cn = pd.concat([cn, pd.DataFrame({"gene": ["KIR3DL9*BACKBONE"], "cn": [2]})])
    cn.to_csv(cn_files[1], sep="\t", index=False)
    with pytest.raises(IndexError):
        main.alleleTyping(names[1:2], cn_files[1:2], "full", _backend=FakeBackend())
    with caplog.at_level("WARNING", logger="graphkir"):
        files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    assert any("KIR3DL9" in r.getMessage() and "isHomozygous" in r.getMessage() for r in caplog.records)
    rows = [pd.read_csv(f, sep="\t").fillna("") for f in files]
    assert rows[0].equals(rows_before[0]) and rows[2].equals(rows_before[2])
    assert rows[1]["alleles"][0] == rows_before[1]["alleles"][0] + "_KIR3DL9*_KIR3DL9*"
    assert "KIR3DL9*BACKBONE" in rows[1]["warnings"][0]


def _gene_with_a_read_pair_beyond_the_capacity(name: str, n_obs: int = 256):
    """A gene whose last read pairs carry ``n_obs`` variant observations each (the device path counts mismatches
    per read pair in a byte: at most 255)."""
    variants = [Variant(pos=10 + 3 * i, typ="single", ref=name, val="A", id=f"hw{i}",
                        allele=[f"{name[:-9]}*{j:03d}" for j in range(4) if (i + j) % 3])
                for i in range(n_obs + 4)]
    reads = [PairRead(backbone=name, lpv=[f"hw{i % 7}"], rnv=[f"hw{7 + i % 5}"]) for i in range(40)]
    for _ in range(3):                                    # three of them: the observations survive errorCorrection
        reads.append(PairRead(backbone=name, lnv=[f"hw{i}" for i in range(n_obs // 2)],
                              rnv=[f"hw{i}" for i in range(n_obs // 2, n_obs)]))
    return reads, variants


@pytest.mark.parametrize("fast", [False, True])
def test_a_gene_beyond_a_capacity_fails_alone(tmp_path, caplog, fast):
    """A read pair with 256 observations exceeds a capacity the reference does not have: every entry point
    reports that gene and calls it fail, the other genes and samples are typed as before (the per-sample
    driver on the object and the fast path, and the cohort entry)."""
    import json
    from dataclasses import asdict
    from kir_graph_b200 import kir_typing
    names, cn_files = _cohort(tmp_path)
    before = [pd.read_csv(f, sep="\t").fillna("") for f in
              main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())]
    reads, variants = _gene_with_a_read_pair_beyond_the_capacity("KIRWIDE*BACKBONE")
    data = json.load(open(names[0] + ".json"))
    data["reads"] += [asdict(r) for r in reads]
    data["variants"] += [asdict(v) for v in variants]
    json.dump(data, open(names[0] + ".json", "w"))
    cn = pd.concat([pd.read_csv(cn_files[0], sep="\t"), pd.DataFrame({"gene": ["KIRWIDE*BACKBONE"], "cn": [2]})])
    cn.to_csv(cn_files[0], sep="\t", index=False)
    with caplog.at_level("WARNING", logger="graphkir"):
        t = kir_typing.selectKirTypingModel("full", names[0] + ".json", top_n=600, variant_correction=True,
                                            _backend=FakeBackend(), **({"_fast": True} if fast else {}))
        alleles, warnings = t.typing(main.loadCN(cn_files[0]))
    assert alleles[-2:] == ["KIRWIDE*", "KIRWIDE*"] and "KIRWIDE*BACKBONE" in warnings
    assert "KIRWIDE*BACKBONE" in t.capacity_failures and "255" in t.capacity_failures["KIRWIDE*BACKBONE"]
    assert any("KIRWIDE" in r.getMessage() and "not typed" in r.getMessage() for r in caplog.records)
    files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    rows = [pd.read_csv(f, sep="\t").fillna("") for f in files]
    assert rows[1].equals(before[1]) and rows[2].equals(before[2])
    assert rows[0]["alleles"][0] == before[0]["alleles"][0] + "_KIRWIDE*_KIRWIDE*"
    # one observation fewer fits: the gene is typed
    reads, variants = _gene_with_a_read_pair_beyond_the_capacity("KIRWIDE*BACKBONE", 254)
    data["reads"] = [r for r in data["reads"] if r["backbone"] != "KIRWIDE*BACKBONE"] + [asdict(r) for r in reads]
    json.dump(data, open(names[0] + ".json", "w"))
    files = main.cohortAlleleTyping(names, cn_files, "full", _backend=FakeBackend())
    assert "KIRWIDE*0" in pd.read_csv(files[0], sep="\t")["alleles"][0]


def test_extraction_goes_on_when_no_sidecar_can_be_written(tmp_path, monkeypatch, caplog):
    """The sidecar is a cache: when a gene exceeds a capacity of the device path no sidecar is written (and an
    older one is removed), the extraction still returns with its .json in place."""
    from kir_graph_b200 import fastjson, hisat2, packio
    from kir_graph_b200.packing import CapacityError
    from tests import sam_sim
    table, pairs = sam_sim.multi_gene(95, n_pairs=60, novel=0.0)
    name = str(tmp_path / "c.variant")
    with open(name + ".sam", "w") as f:
        f.write(sam_sim.sam_text(pairs))
    Variant.novel_id = 0
    hisat2.extractVariantFromSam(table, name + ".sam", name, num_editdist=9, write_pack=True)
    assert os.path.exists(packio.sidecar_path(name))
    real = fastjson.packs_from_scan

    def one_gene_too_wide(scan, **kw):
        packs = real(scan, **kw)
        packs[next(iter(packs))] = CapacityError("a read pair carries 256 variant observations (limit 255)")
        return packs

    monkeypatch.setattr(fastjson, "packs_from_scan", one_gene_too_wide)
    Variant.novel_id = 0
    with caplog.at_level("WARNING", logger="graphkir"):
        ext = hisat2.extractVariantFromSam(table, name + ".sam", name, num_editdist=9, write_pack=True)
    assert ext.n_reads > 0 and os.path.exists(name + ".json") and not os.path.exists(packio.sidecar_path(name))
    assert any("No packed sidecar" in r.getMessage() for r in caplog.records)
