"""SAM record -> per-read variant ids (kir_graph_b200/hisat2.py) against outputs of the reference
(tests/golden/sam_walk.json.gz: simulated CIGAR/MD/Zs records and SURVEY.md Appendix C cases)."""
import copy

import pytest

from kir_graph_b200 import hisat2
from kir_graph_b200.msa2hisat import Variant
from tests.helpers import load_golden


@pytest.mark.parametrize("case", [0, 1])
def test_simulated_records_match_reference(case):
    data = load_golden("sam_walk")["cases"][case]
    table = [Variant(**v) for v in data["variants"]]
    pairs = [tuple(p) for p in data["pairs"]]
    records = [rec for pair in pairs for rec in pair]
    for rec, want in zip(records, data["records"]):
        raw, clip = hisat2.recordToRawVariant(rec)
        assert [[v.typ, v.pos, v.length, v.val, v.id] for v in raw] == want["raw"]
        assert clip == want["clip"]
        assert hisat2.filterRead(rec) == want["filter"]
    Variant.novel_id = 0
    out = hisat2.extractVariant(pairs, table)
    got = [{"lpv": r.lpv, "lnv": r.lnv, "rpv": r.rpv, "rnv": r.rnv, "multiple": r.multiple,
            "backbone": r.backbone} for r in out["reads"]]
    assert got == data["reads"]
    assert [v.id for v in out["variants"]] == data["variant_ids_after"]
    # pairing of name-sorted records
    lines = ["@HD\tVN:1.0"]
    for left, right in pairs:
        lf, rf = left.split("\t"), right.split("\t")
        lf[7], rf[7] = rf[3], lf[3]                      # PNEXT = mate position
        lines += ["\t".join(lf), "\t".join(rf)]
    paired = list(hisat2.pairRecords(lines))
    assert len(paired) == len(pairs)
    assert all(a.split("\t")[0] == b.split("\t")[0] for a, b in paired)


def test_appendix_c_known_answers():
    data = load_golden("sam_walk")
    table = sorted(Variant(**v) for v in data["kat_table"])
    for key, want in data["kats"].items():
        Variant.novel_id = 0
        vmap = {v: v for v in copy.deepcopy(table)}
        rv = hisat2.recordToVariants(want["line"], vmap)
        pos, neg = hisat2.getPNFromVariantList(rv, table)
        assert [v.id for v in pos] == want["positive"], key
        assert [v.id for v in neg] == want["negative"], key
    line = data["kats"]["k1"]["line"]
    def with_flag_nm(flag, nm):
        f = line.split("\t"); f[1] = str(flag); f[11] = f"NM:i:{nm}"; return "\t".join(f)
    assert hisat2.filterRead(with_flag_nm(99, 4)) is data["filter"]["99_4"] is True
    assert hisat2.filterRead(with_flag_nm(99, 5)) is data["filter"]["99_5"] is False
    assert hisat2.filterRead(with_flag_nm(97, 1)) is data["filter"]["97_1"] is False
    assert hisat2.filterRead(with_flag_nm(355, 1)) is data["filter"]["355_1"] is True
    assert hisat2.filterRead("\t".join(line.split("\t")[:11])) is False       # no NM tag
    assert hisat2.getNH(line) == 1 and hisat2.getNH(line.replace("NH:i:1", "NH:i:3")) == 3


def test_splicing_and_pileup_are_refused():
    line = load_golden("sam_walk")["kats"]["k1"]["line"].replace("60M", "30M5N30M")
    with pytest.raises(NotImplementedError):
        hisat2.recordToRawVariant(line)
    with pytest.raises(NotImplementedError):
        hisat2.recordToVariants(load_golden("sam_walk")["kats"]["k1"]["line"], {}, pileup={"x": 1})


def _outcome(fn, rec):
    try:
        raw, clip = fn(rec)
        return [[v.typ, v.pos, v.length, v.val, v.id, v.ref] for v in raw], clip
    except (NotImplementedError, AssertionError, IndexError, ValueError) as exc:
        return type(exc).__name__


def test_native_walk_equals_python_statement_on_mutated_records():
    """gk_sam_walk (host C++) against the Python walker: simulated records, then the same records
    with CIGAR / MD / Zs / sequence damaged at random - same segments or the same exception type."""
    import numpy as np
    from tests import sam_sim
    rng = np.random.default_rng(17)
    records = []
    for seed in (3, 4, 5):
        _, _, pairs = sam_sim.simulate_pairs(seed, n_pairs=40)
        records += [rec for pair in pairs for rec in pair]
    assert all(_outcome(hisat2.recordToRawVariant, r) == _outcome(hisat2.recordToRawVariantPy, r) for r in records)
    alphabet = "0123456789MIDSNHX=^ACGT|,*Zs:"
    n_err = 0
    for rec in records:
        for _ in range(12):
            cols = rec.split("\t")
            c = int(rng.choice([3, 5, 9] + list(range(11, len(cols)))))
            s = cols[c]
            if not s:
                continue
            i = int(rng.integers(len(s)))
            kind = int(rng.integers(3))
            ch = alphabet[int(rng.integers(len(alphabet)))]
            cols[c] = s[:i] + ch + s[i + 1:] if kind == 0 else s[:i] + s[i + 1:] if kind == 1 else s[:i] + ch + s[i:]
            bad = "\t".join(cols)
            a, b = _outcome(hisat2.recordToRawVariant, bad), _outcome(hisat2.recordToRawVariantPy, bad)
            assert a == b, (bad, a, b)
            n_err += isinstance(a, str)
    assert n_err > 100                                   # the damage does exercise the error paths
    for rec in ("r\t99\tG\t5\t60\t*\t=\t1\t0\tACGT\tFFFF", "r\t99\tG\t5\t60\t4M\t=\t1\t0\tACGT\tFFFF\tMD:Z:4",
                "r\t99\tG\t5\t60\t2S2M\t=\t1\t0\tACGT\tFFFF\tMD:Z:2\tZs:Z:", "r\t99\tG\tx\t60\t4M", ""):
        assert _outcome(hisat2.recordToRawVariant, rec) == _outcome(hisat2.recordToRawVariantPy, rec), rec


def test_zs_items_follow_the_references_tuple_indexing():
    """readZs builds (int(f[0]), f[1], f[2]) from item.split("|") (hisat2.py:518-527): the gap is parsed
    first (ValueError), a missing field is an IndexError, fields beyond the third are ignored - found
    by tools/fuzz_sam_vs_reference.py against the imported reference."""
    line = load_golden("sam_walk")["kats"]["k5"]["line"]          # ... Zs:Z:10|S|hv0
    assert "Zs:Z:10|S|hv0" in line
    for walk in (hisat2.recordToRawVariant, hisat2.recordToRawVariantPy):
        good = _outcome(walk, line)
        assert not isinstance(good, str)
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:10|S|hv0|more|fields")) == good
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:10|Shv0")) == "IndexError"
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:10")) == "IndexError"
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:x|S")) == "ValueError"
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:")) == "ValueError"
        assert _outcome(walk, line.replace("Zs:Z:10|S|hv0", "Zs:Z:10|S|hv0,")) == "ValueError"


def test_index_readers_match_reference(tmp_path):
    """getVariants (readVariants / readLink / readExons / isInExon) over .snp / .link / .locus files:
    tests/golden/index_readers.json.gz holds what the reference's getVariants returned
    (make_golden_index.py) - order, allele lists and exon flags included."""
    from dataclasses import asdict
    for i, case in enumerate(load_golden("index_readers")["cases"]):
        index = str(tmp_path / f"kir{i}")
        for ext, text in case["files"].items():
            with open(f"{index}.{ext}", "w") as handle:
                handle.write(text)
        got = hisat2.getVariants(index)
        assert [asdict(v) for v in got] == case["variants"]
        assert got == sorted(got) and any(v.in_exon for v in got) and any(not v.allele for v in got)


def test_record_with_more_segments_than_the_reusable_buffer():
    """A long read with 90 mismatches and 12 indels walks into 200+ segments: the native walk reports that its
    reusable 64-segment buffer is too small and is called again with a buffer sized from the record."""
    md, cigar, seq = [], [], []
    for i in range(12):                                   # 12 x (20M 1I 20M 2D) with 7 mismatches per block
        cigar.append("20M1I20M2D")
        md.append("2C2C2C2C2C2C2C5" + "^GG")             # 26 + ... : first 20M + second 20M = 40 aligned bases
        seq.append("A" * 41)
    # MD above covers 7 * 3 + 5 = 26 bases per block; pad the blocks to 40 aligned bases
    md = [m.replace("5^GG", "19^GG") for m in md]
    cigar.append("30M")
    md.append("2T" * 6 + "12")
    seq.append("A" * 30)
    line = "\t".join(["long", "99", "KIRX*BACKBONE", "101", "60", "".join(cigar), "=", "400", "700", "".join(seq),
                      "F" * sum(len(s) for s in seq), "NM:i:126", "MD:Z:" + "".join(md)])
    native, python = _outcome(hisat2.recordToRawVariant, line), _outcome(hisat2.recordToRawVariantPy, line)
    assert native == python and not isinstance(native, str)
    assert len(native[0]) > 64
    short = load_golden("sam_walk")["kats"]["k1"]["line"]  # the reusable buffer still serves the next record
    assert _outcome(hisat2.recordToRawVariant, short) == _outcome(hisat2.recordToRawVariantPy, short)
