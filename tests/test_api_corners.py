"""Corners of the mirrored API that the other suites do not reach: host utilities kept for API parity
(``uniqueAllele``, the sequential ``typingIntron``), the DEBUG print of a result, the EM report writer."""
import io
import json
import logging
import os

import numpy as np

from kir_graph_b200 import synthetic, typing_em
from kir_graph_b200.hisat2 import writeReadsAndVariantsData
from kir_graph_b200.typing_mulit_allele import AlleleTyping, AlleleTypingExonFirst
from tests.fake_backend import FakeBackend
from tests.helpers import golden_names, load_golden, objects_from_input


def test_unique_allele_known_answer():
    kat = load_golden("kats")["unique_allele"]                 # the docstring example of typing_mulit_allele.py:463-465
    assert AlleleTyping.uniqueAllele(np.array(kat["in"])).tolist() == kat["out"]


def test_sequential_typing_intron_equals_the_batched_form():
    """``typingIntron`` (deepcopy of the full model + ``addCandidate`` per exon candidate list,
    typing_mulit_allele.py:740-746) gives the steps the batched ``typing`` merges."""
    case = load_golden(golden_names("exonfirst")[0])
    reads, variants = objects_from_input(case["input"])
    typ = AlleleTypingExonFirst(reads, variants, force_homo=False, top_n=case["top_n"],
                                candidate_set_threshold=case["threshold"], _backend=FakeBackend())
    typ.typing(case["cn"])
    exon_best = typ.result[case["cn"] - 1]
    groups = [[sorted(typ.allele_group[g]) for g in names] for names in exon_best.allele_name[:1]]
    model = typ.typingIntron([sum(g, []) for g in zip(*groups)] if case["cn"] > 1 else [sum(groups[0], [])])
    assert len(model.result) >= 1 and model.result[-1].value.shape[0] >= 1
    # the sequential model searched the same restricted candidates: its best set is made of members of the groups
    members = set(sum(groups[0], []))
    assert set(model.result[-1].allele_name[0]) <= members


def test_debug_print_of_results(caplog):
    gene = synthetic.make_gene([5, 1], "KIRP*BACKBONE", 14, 112, 2, 150, hierarchical=True)
    reads, variants = gene.to_objects()
    logger = logging.getLogger("graphkir")
    with caplog.at_level(logging.DEBUG, logger="graphkir"):
        typ = AlleleTypingExonFirst(reads, variants, force_homo=False, top_n=20, _backend=FakeBackend())
        res = typ.typing(2)
        res.print(num=3)
    text = "\n".join(r.getMessage() for r in caplog.records)
    assert "Allele_num =  2" in text and "Rank 0 probility" in text and "fraction" in text and "group" in text
    assert logger.name == "graphkir"


def test_em_report_writer(tmp_path, monkeypatch):
    """``hisat2Typing`` (typing_em.py:218-241): the report of every gene as text (alleles by count, then by
    abundance) and as JSON.  (The reference's ``json.dump`` of its dataclass objects raises TypeError after the
    text file is written; the mirror writes the fields.)"""
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "sample.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    be = FakeBackend()
    orig = typing_em.hisat2TypingPerGene
    monkeypatch.setattr(typing_em, "hisat2TypingPerGene", lambda reads, **kw: orig(reads, _backend=be))
    typing_em.hisat2Typing(path, os.path.join(tmp_path, "report"))
    data = json.load(open(os.path.join(tmp_path, "report.json")))
    text = open(os.path.join(tmp_path, "report.txt")).read().splitlines()
    assert set(data) and all(set(item) >= {"allele", "count", "prob"} for items in data.values() for item in items)
    for gene, items in data.items():
        assert gene in text
        at = text.index(gene)
        counts = sorted((i["count"] for i in items), reverse=True)[:10]
        shown = [int(line.split("count: ")[1].rstrip(")")) for line in text[at + 1: at + 1 + len(counts)]]
        assert shown == counts
        assert abs(sum(i["prob"] for i in items) - 1) < 1e-6 or not items
    out = io.StringIO()
    typing_em.printHisatTyping({"KIRX*BACKBONE": []}, file=out)
    assert out.getvalue() == "KIRX*BACKBONE\n"


def test_most_frequent_allele_and_empty_em():
    assert typing_em.getMostFreqAllele(["a", "b", "b", "c", "c"]) == ["b", "c"]
    assert typing_em.getMostFreqAllele([]) == []
    assert typing_em.hisatEMnp([], _backend=FakeBackend()) == {}
