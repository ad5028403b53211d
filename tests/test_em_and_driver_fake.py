"""EM path and the per-sample driver (kir_typing) on the NumPy test double, against
outputs of the reference itself (tests/golden)."""
import json
import os

import numpy as np
import pytest

from kir_graph_b200 import kir_typing, typing_em
from kir_graph_b200.hisat2 import writeReadsAndVariantsData
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden, objects_from_input


def test_em_kats_and_synthetic():
    cases = load_golden("em_cases")["cases"]
    be = FakeBackend()
    kat = cases["kat_candidate"]
    assert sorted(typing_em.getCandidateAllelePerRead(kat["positive"], kat["negative"], _backend=be)) == sorted(kat["out"])
    for key in ("kat_simple", "syn_a16"):
        prob = typing_em.hisatEMnp(cases[key]["allele_per_read"], _backend=be)
        assert set(prob) == set(cases[key]["prob"])
        for name, p in cases[key]["prob"].items():
            assert abs(prob[name] - p) < 1e-9
    # from raw reads: compat rows equal the reference's per-read candidate lists
    reads, variants = objects_from_input(cases["syn_a16"]["input"])
    gene = typing_em.preprocessHisatReads({"reads": reads, "variants": variants})["KIRI*BACKBONE"]
    rows = typing_em.compatible_alleles(gene, be)
    got = [sorted(x) for x in typing_em._rows_to_names(rows, gene.allele_names)]
    assert got == cases["syn_a16"]["allele_per_read"]
    report = typing_em.hisat2TypingPerGene(gene, _backend=be)
    for item in report:
        assert abs(item.prob - cases["syn_a16"]["prob"][item.allele]) < 1e-9


@pytest.mark.parametrize("method", ["full", "exonfirst_1", "exonfirst", "em"])
def test_select_kir_typing_model_matches_reference(tmp_path, method):
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "sample.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    kw = {"full": dict(top_n=60, variant_correction=True), "exonfirst_1": dict(top_n=60),
          "exonfirst": dict(top_n=60), "em": {}}[method]
    t = kir_typing.selectKirTypingModel(method, path, _backend=FakeBackend(), **kw)
    alleles, warn = t.typing(sample["gene_cn"])
    ref = sample["calls"][method]
    assert warn == ref["warnings"]
    if method == "em":
        assert sorted(alleles) == sorted(ref["alleles"])
    else:
        assert alleles == ref["alleles"] or t.tie_report, (alleles, ref["alleles"])
        possible = t.getAllPossibleTyping()
        assert [p["gene"] for p in possible][:1] == [p["gene"] for p in ref["possible"]][:1]
        np.testing.assert_allclose(possible[0]["value"], ref["possible"][0]["value"], rtol=1e-11)
    out = os.path.join(tmp_path, "dump.json")
    t.save(out)
    assert json.load(open(out))


def test_unknown_method():
    with pytest.raises(NotImplementedError):
        kir_typing.selectKirTypingModel("report", "nothing.json")


def test_read_allele_length(tmp_path):
    from kir_graph_b200.typing_em import readAlleleLength
    path = tmp_path / "a.fa"
    path.write_text(">KIR2DL1*001 some description\nACGT\nAC\n\n>KIR2DL1*002\nAAAA\r\n>empty\n")
    assert readAlleleLength(str(path)) == {"KIR2DL1*001": 6, "KIR2DL1*002": 4, "empty": 0}
