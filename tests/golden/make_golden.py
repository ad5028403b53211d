#!/usr/bin/env python
"""
Generate golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):    python tests/golden/make_golden.py

The reference imports plotly / Bio / pyhlamsa at module top; none of them takes
part in the typing arithmetic (plots, fasta length reader, MSA writers), so they
are stubbed with empty modules.  Fixtures are self-contained: each holds the
input (reads + variants in the reference's ``.variant.json`` layout) and what
the reference computed from it.  Floats are stored with ``repr`` round-trip
precision by ``json``.
"""
from __future__ import annotations

import copy
import gzip
import json
import os
import sys
import types
from dataclasses import asdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for name in ("plotly", "plotly.express", "plotly.graph_objects", "plotly.subplots",
                 "Bio", "Bio.SeqIO", "pyhlamsa"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plotly"].express = sys.modules["plotly.express"]
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    sys.modules["plotly.graph_objects"].Figure = object
    sys.modules["plotly.subplots"].make_subplots = lambda *a, **k: None
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["pyhlamsa"].Genemsa = object
    sys.path.insert(0, "/root/reference")
    import graphkir.typing_mulit_allele as tma
    import graphkir.typing_em as tem
    import graphkir.kir_typing as kt
    import graphkir.hisat2 as h2
    import graphkir.msa2hisat as m2h
    return tma, tem, kt, h2, m2h


def dump(name: str, payload: dict) -> None:
    path = os.path.join(HERE, name + ".json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(payload, f)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def result_to_dict(res) -> dict:
    return {
        "n": int(res.n),
        "value": np.asarray(res.value, dtype=float).tolist(),
        "value_sum_indv": np.asarray(res.value_sum_indv, dtype=float).tolist(),
        "allele_id": np.asarray(res.allele_id).tolist(),
        "allele_name": [list(x) for x in res.allele_name],
        "fraction": np.asarray(res.fraction, dtype=float).tolist(),
        "allele_name_group": res.allele_name_group,
    }


def ref_objects(h2, m2h, reads, variants):
    """our dataclasses -> reference dataclasses (same fields)."""
    rv = [m2h.Variant(**asdict(v)) for v in variants]
    rr = [h2.PairRead(**asdict(r)) for r in reads]
    return rr, rv


def typing_case(tma, h2, m2h, name, reads, variants, cn, top_n, variant_correction,
                force_homo=False, store_probs=True):
    rr, rv = ref_objects(h2, m2h, reads, variants)
    inputs = {"variants": [asdict(v) for v in variants], "reads": [asdict(r) for r in reads]}
    typ = tma.AlleleTyping(rr, rv, force_homo=force_homo, top_n=top_n,
                           variant_correction=variant_correction)
    homo_auto = bool(tma.isHomozygous(typ.reads, typ.variants, cn))
    res = typ.typing(cn)
    payload = {
        "kind": "typing", "name": name, "cn": cn, "top_n": top_n,
        "variant_correction": variant_correction, "force_homo": force_homo,
        "input": inputs,
        "allele_names": [typ.id_to_allele[i] for i in range(len(typ.id_to_allele))],
        "n_reads": typ.getReadsNum(),
        "reads_after": [{"lpv": r.lpv, "rpv": r.rpv, "lnv": r.lnv, "rnv": r.rnv} for r in typ.reads],
        "is_homozygous": homo_auto,
        "steps": [result_to_dict(r) for r in typ.result],
        "best": res.selectBest(),
        "possible": [[float(v), list(a)] for v, a in res.selectAllPossible(0.9)],
    }
    if store_probs:
        payload["probs"] = np.asarray(typ.probs).tolist()
        payload["log_probs"] = np.asarray(typ.log_probs).tolist()
    dump(name, payload)


def exonfirst_case(tma, h2, m2h, name, reads, variants, cn, top_n, threshold):
    rr, rv = ref_objects(h2, m2h, reads, variants)
    inputs = {"variants": [asdict(v) for v in variants], "reads": [asdict(r) for r in reads]}
    typ = tma.AlleleTypingExonFirst(rr, rv, force_homo=False, top_n=top_n,
                                    candidate_set_threshold=threshold)
    res = typ.typing(cn)
    exon_steps = typ.result[:cn]
    dump(name, {
        "kind": "exonfirst", "name": name, "cn": cn, "top_n": top_n, "threshold": threshold,
        "input": inputs,
        "allele_group": typ.allele_group,
        "exon_allele_names": [typ.id_to_allele[i] for i in range(len(typ.id_to_allele))],
        "exon_n_reads": typ.getReadsNum(),
        "exon_steps": [result_to_dict(r) for r in exon_steps],
        "n_results": len(typ.result),
        "final": result_to_dict(res),
        "best": res.selectBest(),
    })


def main() -> None:
    tma, tem, kt, h2, m2h = import_reference()
    from kir_graph_b200 import synthetic as syn
    from kir_graph_b200.hisat2 import PairRead, writeReadsAndVariantsData
    from kir_graph_b200.msa2hisat import Variant

    # -- the worked 6-read example --------------------
    g = "G*BACKBONE"
    variants = [
        Variant(pos=10, typ="single", ref=g, val="A", id="hv0", allele=["G*001", "G*002"]),
        Variant(pos=20, typ="single", ref=g, val="C", id="hv1", allele=["G*002", "G*003"]),
        Variant(pos=30, typ="single", ref=g, val="G", id="hv2", allele=["G*003"]),
        Variant(pos=40, typ="single", ref=g, val="T", id="hv3", allele=["G*001", "G*004"]),
    ]
    spec = [(["hv0"], ["hv1"], ["hv3"], ["hv2"]), (["hv0"], ["hv1"], ["hv3"], []),
            (["hv0", "hv1"], [], [], ["hv2", "hv3"]), (["hv1"], ["hv2"], ["hv0"], ["hv3"]),
            (["hv1", "hv2"], ["hv0"], [], ["hv3"]), (["hv0"], [], ["hv0"], ["hv2"])]
    reads = [PairRead(backbone=g, lpv=a, lnv=b, rpv=c, rnv=d) for a, b, c, d in spec]
    typing_case(tma, h2, m2h, "worked_example_nocorr", copy.deepcopy(reads), variants, 2, 300, False)
    typing_case(tma, h2, m2h, "worked_example_corr", copy.deepcopy(reads), variants, 2, 300, True)

    # -- seeded synthetic genes ------------------------------------------------
    cases = [
        ("syn_a24_cn3", dict(seed=[11, 0], gene="KIRA*BACKBONE", n_allele=24, n_var=96, cn=3, n_reads=300), 3, 40, True),
        ("syn_a60_cn2", dict(seed=[12, 0], gene="KIRB*BACKBONE", n_allele=60, n_var=480, cn=2, n_reads=400), 2, 300, True),
        ("syn_a12_cn4_nocorr", dict(seed=[13, 0], gene="KIRC*BACKBONE", n_allele=12, n_var=64, cn=4, n_reads=250), 4, 25, False),
        ("syn_a6_cn2", dict(seed=[14, 0], gene="KIRD*BACKBONE", n_allele=6, n_var=64, cn=2, n_reads=120), 2, 300, True),
        ("syn_a40_cn1", dict(seed=[15, 0], gene="KIRE*BACKBONE", n_allele=40, n_var=320, cn=1, n_reads=200), 1, 10, True),
    ]
    for name, kw, cn, top_n, corr in cases:
        gene = syn.make_gene(**kw)
        reads, variants = gene.to_objects()
        typing_case(tma, h2, m2h, name, reads, variants, cn, top_n, corr)
    # homozygous decision exercised through force_homo=None
    gene = syn.make_gene(seed=[16, 0], gene="KIRF*BACKBONE", n_allele=20, n_var=160, cn=2, n_reads=600,
                         homo_prob=1.0)
    reads, variants = gene.to_objects()
    typing_case(tma, h2, m2h, "syn_homo_auto", reads, variants, 2, 50, True, force_homo=None)
    gene = syn.make_gene(seed=[17, 0], gene="KIRG*BACKBONE", n_allele=20, n_var=160, cn=2, n_reads=600,
                         homo_prob=0.0)
    reads, variants = gene.to_objects()
    typing_case(tma, h2, m2h, "syn_hetero_auto", reads, variants, 2, 50, True, force_homo=None)

    # -- exon-first (hierarchical generator) ----------------------------------
    for name, seed, cn, thr in (("exon_a30_cn2", 21, 2, 1.0), ("exon_a18_cn3_thr0", 22, 3, 0.0)):
        gene = syn.make_gene(seed=[seed, 0], gene="KIRH*BACKBONE", n_allele=30 if cn == 2 else 18,
                             n_var=240 if cn == 2 else 144, cn=cn, n_reads=350, hierarchical=True)
        reads, variants = gene.to_objects()
        exonfirst_case(tma, h2, m2h, name, reads, variants, cn, 50, thr)

    # -- EM path -----------------------------------------------------------------
    em_cases = {}
    em_cases["kat_simple"] = {
        "allele_per_read": [["a"], ["a", "b"], ["b"], ["a"]],
    }
    gene = syn.make_gene(seed=[31, 0], gene="KIRI*BACKBONE", n_allele=16, n_var=128, cn=2, n_reads=300)
    reads, variants = gene.to_objects()
    by_id = {v.id: v.allele for v in variants}
    per_read = []
    for r in reads:
        per_read.append(tem.getMostFreqAllele(
            tem.getCandidateAllelePerRead([by_id[v] for v in r.lpv], [by_id[v] for v in r.lnv])
            + tem.getCandidateAllelePerRead([by_id[v] for v in r.rpv], [by_id[v] for v in r.rnv])))
    em_cases["syn_a16"] = {"allele_per_read": [sorted(x) for x in per_read],
                           "input": {"variants": [asdict(v) for v in variants],
                                     "reads": [asdict(r) for r in reads]}}
    for case in em_cases.values():
        prob = tem.hisatEMnp(case["allele_per_read"])
        case["prob"] = {k: float(v) for k, v in prob.items()}
    em_cases["kat_candidate"] = {
        "positive": [["a", "b", "c"], ["b", "c"]], "negative": [["c"]],
        "out": tem.getCandidateAllelePerRead([["a", "b", "c"], ["b", "c"]], [["c"]]),
    }
    dump("em_cases", {"kind": "em", "cases": em_cases})

    # -- small doc-string KATs ------------------------------------------------
    uniq = tma.AlleleTyping.uniqueAllele(np.array([[0, 1], [1, 0], [2, 0], [2, 2], [2, 0], [1, 0]]))
    sel = tma.TypingResult(
        n=2, value=np.array([-1., -2., -3., -4.]), value_sum_indv=np.zeros((4, 2)),
        allele_id=np.arange(8).reshape(4, 2), allele_name=[[f"x{i}", f"y{i}"] for i in range(4)],
        allele_prob=np.zeros((1, 4)),
        fraction=np.array([[.1, .9], [.05, .95], [.2, .8], [.4, .6]]),
        fraction_uniq=np.ones((4, 2)))
    dump("kats", {
        "kind": "kats",
        "unique_allele": {"in": [[0, 1], [1, 0], [2, 0], [2, 2], [2, 0], [1, 0]],
                          "out": [bool(x) for x in uniq]},
        "select_best": {"out": sel.selectBest()},
        "log10": {"hit": float(np.log10(0.999)), "miss": float(np.log10(0.001))},
    })

    # -- SAM record walk and read-variant extraction (hisat2.py:279-844) --------------------------
    from tests import sam_sim
    sam_cases = []
    for seed in (1, 2):
        _, table, pairs = sam_sim.simulate_pairs(seed, 45)
        rtable = [m2h.Variant(**asdict(v)) for v in table]
        records = []
        for l, r in pairs:
            for rec in (l, r):
                raw, clip = h2.recordToRawVariant(rec)
                records.append({"raw": [[v.typ, v.pos, v.length, v.val, v.id] for v in raw], "clip": clip,
                                "filter": bool(h2.filterRead(rec))})
        m2h.Variant.novel_id = 0
        data = h2.extractVariant(pairs, rtable)
        sam_cases.append({
            "variants": [asdict(v) for v in table], "pairs": pairs, "records": records,
            "reads": [{"lpv": r.lpv, "lnv": r.lnv, "rpv": r.rpv, "rnv": r.rnv, "multiple": r.multiple,
                       "backbone": r.backbone} for r in data["reads"]],
            "variant_ids_after": [v.id for v in data["variants"]],
        })
    # Appendix C of SURVEY.md: hand-built records
    g = "KIRX*BACKBONE"
    ctab = [m2h.Variant(pos=10, typ="single", ref=g, val="T", id="hv0", allele=["a1", "a2"]),
            m2h.Variant(pos=20, typ="single", ref=g, val="C", id="hv2", allele=["a3"]),
            m2h.Variant(pos=20, typ="single", ref=g, val="G", id="hv1", allele=["a2"]),
            m2h.Variant(pos=30, typ="deletion", ref=g, val=2, id="hv3", allele=["a1"]),
            m2h.Variant(pos=45, typ="insertion", ref=g, val="AC", id="hv4", allele=["a3"]),
            m2h.Variant(pos=120, typ="deletion", ref=g, val=3, id="hv5", allele=["a3"]),
            m2h.Variant(pos=140, typ="single", ref=g, val="A", id="hv6", allele=["a1"])]
    ctab = sorted(ctab)

    def rec(pos, cigar, md, zs, seq, flag=99, nm=1):
        f = ["r", str(flag), g, str(pos), "60", cigar, "=", "300", "350", seq, "I" * len(seq), f"NM:i:{nm}", f"MD:Z:{md}"]
        if zs:
            f.append(f"Zs:Z:{zs}")
        f.append("NH:i:1")
        return "\t".join(f)

    kats = {
        "k0": rec(1, "30M2D18M", "10C9C9^CC18", "10|S|hv0,9|S|hv1,9|D|hv3", "A" * 10 + "T" + "A" * 9 + "G" + "A" * 27),
        "k1": rec(1, "60M", "60", "", "A" * 60),
        "k3": rec(1, "40M", "15G24", "", "A" * 15 + "C" + "A" * 24),
        "k4": rec(1, "5S35M", "35", "", "A" * 40),
        "k5": rec(1, "40M", "10G0G28", "10|S|hv0", "A" * 10 + "TC" + "A" * 28),
        "k6": rec(1, "12M3D28M", "12^GGG28", "", "A" * 40),
        "k2p": rec(1, "5M2I20M", "25", "", "A" * 5 + "AC" + "A" * 20),
        "k7": rec(101, "36M", "36", "", "A" * 36),
        "k8": rec(101, "29M", "29", "", "A" * 29),
        "k9": rec(22, "9M2D11M", "9^CC0G10", "9|D|hv3", "A" * 9 + "T" + "A" * 10),
    }
    kat_out = {}
    for key, line in kats.items():
        m2h.Variant.novel_id = 0
        vmap = {v: v for v in copy.deepcopy(ctab)}
        rv = h2.recordToVariants(line, vmap)
        pos_v, neg_v = h2.getPNFromVariantList(rv, ctab)
        kat_out[key] = {"line": line, "positive": [v.id for v in pos_v], "negative": [v.id for v in neg_v]}
    flt = {"99_4": h2.filterRead(rec(1, "60M", "60", "", "A" * 60, 99, 4)),
           "99_5": h2.filterRead(rec(1, "60M", "60", "", "A" * 60, 99, 5)),
           "97_1": h2.filterRead(rec(1, "60M", "60", "", "A" * 60, 97, 1)),
           "355_1": h2.filterRead(rec(1, "60M", "60", "", "A" * 60, 355, 1))}
    dump("sam_walk", {"kind": "sam", "cases": sam_cases, "kat_table": [asdict(v) for v in ctab],
                      "kats": kat_out, "filter": {k: bool(v) for k, v in flt.items()}})

    # -- whole-sample: selectKirTypingModel over a small multi-gene JSON --------
    genes = [syn.make_gene(seed=[41, i], gene=f"KIRZ{i}*BACKBONE", n_allele=a, n_var=max(64, 8 * a),
                           cn=c, n_reads=r, hierarchical=True, variant_id_base=1000 * i)
             for i, (a, c, r) in enumerate([(20, 2, 260), (9, 1, 150), (14, 3, 330)])]
    all_reads, all_variants = [], []
    for gene in genes:
        rd, va = gene.to_objects()
        all_reads += rd
        all_variants += va
    all_reads[3].multiple = 2          # dropped by removeMultipleMapped
    path = os.path.join(HERE, "_tmp_sample.json")
    writeReadsAndVariantsData({"variants": all_variants, "reads": all_reads}, path)
    gene_cn = {g.gene: g.cn for g in genes}
    gene_cn["KIRNONE*BACKBONE"] = 0
    sample = {"kind": "sample", "gene_cn": gene_cn,
              "input": {"variants": [asdict(v) for v in all_variants],
                        "reads": [asdict(r) for r in all_reads]}, "calls": {}}
    for method, kw in (("full", dict(top_n=60, variant_correction=True)),
                       ("exonfirst_1", dict(top_n=60)), ("exonfirst", dict(top_n=60)),
                       ("em", {})):
        t = kt.selectKirTypingModel(method, path, **kw)
        alleles, warn = t.typing(gene_cn)
        sample["calls"][method] = {"alleles": alleles, "warnings": warn}
        if method != "em":
            sample["calls"][method]["possible"] = t.getAllPossibleTyping()
    os.remove(path)
    dump("sample_small", sample)

    # -- read grouping of novel discovery: reference groupReadByAllele -------------
    for extra in ("Bio.Seq", "Bio.SeqRecord", "pysam"):
        sys.modules.setdefault(extra, types.ModuleType(extra))
    sys.modules["Bio.Seq"].Seq = object
    sys.modules["Bio.SeqRecord"].SeqRecord = object
    sys.modules["pysam"].AlignmentFile = object
    import graphkir.novel_discover as nd
    cases = []
    for seed, a, cn, r in ((21, 40, 2, 600), (22, 25, 3, 900), (23, 60, 4, 700), (24, 12, 1, 300)):
        gene = syn.make_gene([seed, 0], f"KIRG{seed}*BACKBONE", a, 8 * a, cn, r)
        reads, variants = gene.to_objects()
        # one read without any observation: kept by no_empty=False, ties over every allele
        reads.insert(3, PairRead(l_sam="empty\t", r_sam="empty\t", backbone=gene.gene))
        rr, rv = ref_objects(h2, m2h, reads, variants)
        typ = tma.AlleleTyping(rr, rv, no_empty=False)
        predict = [gene.allele_names[t] for t in gene.truth] + ["KIRNOT*00001"]
        if seed == 23:
            predict = predict[::-1]
        groups = nd.groupReadByAllele(typ, predict, rr)
        index = {id(x): i for i, x in enumerate(rr)}
        # probs of the called alleles are stored too: the test also checks the grouping on rounded
        # values (the reference's float comparison can split exact ties by the rounding of the ordered
        # product)
        names = [n for n in predict if n in typ.allele_to_id]
        ids = [typ.allele_to_id[n] for n in names]
        cases.append({
            "name": f"group_{seed}", "predict_alleles": predict,
            "input": {"variants": [asdict(v) for v in variants], "reads": [asdict(x) for x in reads]},
            "groups": [[list(k), [index[id(x)] for x in v]] for k, v in groups.items()],
            "probs_called": np.asarray(typ.probs)[:, ids].tolist(),
        })
    dump("group_reads", {"kind": "group_reads", "cases": cases})


if __name__ == "__main__":
    main()
