"""Native batch extraction SAM text -> CSR variant lists (kir_graph_b200/fastsam.py, gk_sam_extract)
against the reference's outputs (tests/golden/sam_walk.json.gz) and against the Python statement of
the same loop (hisat2.pairRecords / filterRead / extractVariant) on simulated and damaged SAM text."""
import copy
from dataclasses import asdict

import numpy as np
import pytest

from kir_graph_b200 import fastjson, fastsam, hisat2
from kir_graph_b200.msa2hisat import Variant
from tests import sam_sim
from tests.helpers import load_golden


_sam_text = sam_sim.sam_text
_multi_gene = sam_sim.multi_gene


def _python_path(sam: str, table, nm: int = 4):
    pairs = hisat2.pairRecords(sam.split("\n"))
    pairs = filter(lambda lr: hisat2.filterRead(lr[0], nm) and hisat2.filterRead(lr[1], nm), pairs)
    return hisat2.extractVariant(pairs, table)


def _as_dicts(data):
    return [asdict(v) for v in data["variants"]], [asdict(r) for r in data["reads"]]


@pytest.mark.parametrize("case", [0, 1])
def test_golden_pairs_match_reference(case):
    data = load_golden("sam_walk")["cases"][case]
    table = [Variant(**v) for v in data["variants"]]
    pairs = [tuple(p) for p in data["pairs"]]
    sam = _sam_text(pairs)
    Variant.novel_id = 0
    ext = fastsam.extract(sam, table)
    # the reference's extractVariant saw every pair; the batch path also applies filterRead to both mates.
    # Dropping a pair changes the numbering of the novel variants after it, so compare with the
    # reference's rows only when nothing was filtered, else with the Python statement (below)
    keep = [data["records"][2 * i]["filter"] and data["records"][2 * i + 1]["filter"] for i in range(len(pairs))]
    got = ext.reads_data()
    rows = [{"lpv": r.lpv, "lnv": r.lnv, "rpv": r.rpv, "rnv": r.rnv, "multiple": r.multiple, "backbone": r.backbone}
            for r in got["reads"]]
    assert len(rows) == sum(keep)
    if all(keep):
        assert rows == data["reads"]
        assert [v.id for v in got["variants"]] == data["variant_ids_after"]
    else:
        strip = lambda row: {k: ([i for i in v if not i.startswith("nv")] if isinstance(v, list) else v)
                             for k, v in row.items()}
        assert [strip(r) for r in rows] == [strip(r) for r, k in zip(data["reads"], keep) if k]
    n_novel = Variant.novel_id
    Variant.novel_id = 0
    want = _python_path(sam, copy.deepcopy(table))
    assert Variant.novel_id == n_novel
    assert _as_dicts(got) == _as_dicts(want)


def test_unfiltered_golden_rows():
    """With the edit-distance bound lifted and every record flagged as a proper pair nothing is
    filtered, so every row of the reference's extractVariant output must be reproduced."""
    for case in (0, 1):
        data = load_golden("sam_walk")["cases"][case]
        table = [Variant(**v) for v in data["variants"]]
        if not all(int(rec.split("\t")[1]) & 2 and "NM:i:" in rec for pair in data["pairs"] for rec in pair):
            continue
        Variant.novel_id = 0
        ext = fastsam.extract(_sam_text([tuple(p) for p in data["pairs"]]), table, num_editdist=10 ** 6)
        got = ext.reads_data()
        rows = [{"lpv": r.lpv, "lnv": r.lnv, "rpv": r.rpv, "rnv": r.rnv, "multiple": r.multiple,
                 "backbone": r.backbone} for r in got["reads"]]
        assert rows == data["reads"]
        assert [v.id for v in got["variants"]] == data["variant_ids_after"]


@pytest.mark.parametrize("case", [0, 1, 2])
def test_whole_loop_matches_reference(case):
    """tests/golden/sam_extract.json.gz: the reference's readPair -> filterRead -> extractVariant over
    a two-gene name-sorted SAM text (make_golden_sam_extract.py); every field of the .json it would
    write must come out the same, and so must the Python statement of the loop."""
    data = load_golden("sam_extract")["cases"][case]
    table = [Variant(**v) for v in data["table"]]
    Variant.novel_id = 0
    got = fastsam.extract(data["sam"], table, data["num_editdist"]).reads_data()
    assert Variant.novel_id == data["novel_id_after"]
    assert _as_dicts(got) == (data["variants"], data["reads"])
    Variant.novel_id = 0
    want = _python_path(data["sam"], copy.deepcopy(table), data["num_editdist"])
    assert _as_dicts(want) == (data["variants"], data["reads"])


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_multi_gene_equals_python_statement(seed):
    table, pairs = _multi_gene(seed)
    sam = _sam_text(pairs)
    lines = sam.rstrip("\n").split("\n")
    first = lines[2].split("\t")
    extra = [
        "\t".join(["lonely", "99", first[2], "10", "60", "*", "chrX", "50"] + first[8:]),       # mate elsewhere
        "\t".join(["single", "73", first[2], "10", "60"] + first[5:]),                          # never paired
        "\t".join(["odd", "99"] + first[2:]), "\t".join(["odd", "99", first[2], first[7]] + first[4:7] + [first[3]] + first[8:]),
        "",
    ]
    # secondary alignments pair among themselves (flag & 256 is part of the key)
    sec_l, sec_r = lines[4].split("\t"), lines[5].split("\t")
    sec_l[1], sec_r[1] = str(int(sec_l[1]) | 256), str(int(sec_r[1]) | 256)
    extra += ["\t".join(sec_l), "\t".join(sec_r)]
    sam = "\n".join(lines[:40] + extra + lines[40:]) + "\n"
    Variant.novel_id = 7
    ext = fastsam.extract(sam, table, num_editdist=9)
    got = ext.reads_data()
    n_after = Variant.novel_id
    Variant.novel_id = 7
    want = _python_path(sam, copy.deepcopy(table), 9)
    assert Variant.novel_id == n_after > 7
    assert _as_dicts(got) == _as_dicts(want)
    assert ext.n_strange == 1                                 # "odd": both records carry the first-mate flag
    assert ext.n_reads > 40 and set(ext.refs) == {"KIRA*BACKBONE", "KIRB*BACKBONE"}
    assert len({r.backbone for r in got["reads"]}) == 2
    assert any(r.multiple == 3 for r in got["reads"]) and any(v.id.startswith("nv") for v in got["variants"])


def test_scan_feeds_the_packing_path(tmp_path):
    """SAM text -> GenePack without JSON equals SAM text -> reference .json -> fastjson -> GenePack."""
    table, pairs = _multi_gene(21, n_pairs=120)
    alleles = [f"KIR*{i:03d}" for i in range(9)]
    rng = np.random.default_rng(5)
    for v in table:
        v.allele = [a for a in alleles if rng.random() < 0.4] or [alleles[0]]
    sam = _sam_text(pairs)
    Variant.novel_id = 0
    ext = fastsam.extract(sam, table, num_editdist=9)
    direct = fastjson.packs_from_scan(ext.scan())
    path = str(tmp_path / "s.variant.json")
    hisat2.writeReadsAndVariantsData(ext.reads_data(), path)
    via_json = fastjson.load_packs(path)
    assert list(direct) == list(via_json) and len(direct) == 2
    from tests.test_fastjson import _assert_same_pack
    for gene in direct:
        assert direct[gene].csr.n_reads > 10
        _assert_same_pack(direct[gene], via_json[gene])


def _outcome(fn):
    start = Variant.novel_id = 0
    try:
        data = fn()
        return _as_dicts(data), Variant.novel_id - start
    except (NotImplementedError, AssertionError, IndexError, ValueError) as exc:
        return type(exc).__name__, Variant.novel_id - start


def test_damaged_sam_text_same_result_or_same_exception():
    table, pairs = _multi_gene(31, n_pairs=12)
    rng = np.random.default_rng(32)
    sam = _sam_text(pairs, header=False)
    lines = sam.rstrip("\n").split("\n")
    alphabet = "0123456789MIDSNHX=^ACGT|,*Zs:\t"
    n_err = n_ok = 0
    for trial in range(400):
        bad = list(lines)
        for _ in range(int(rng.integers(1, 3))):
            j = int(rng.integers(len(bad)))
            cols = bad[j].split("\t")
            c = int(rng.choice([1, 3, 5, 7, 9] + list(range(11, len(cols)))))
            s = cols[c]
            if not s:
                continue
            i = int(rng.integers(len(s)))
            kind = int(rng.integers(3))
            ch = alphabet[int(rng.integers(len(alphabet)))]
            cols[c] = s[:i] + ch + s[i + 1:] if kind == 0 else s[:i] + s[i + 1:] if kind == 1 else s[:i] + ch + s[i:]
            bad[j] = "\t".join(cols)
        text = "\n".join(bad) + "\n"
        a = _outcome(lambda: fastsam.extract(text, table, 9).reads_data())
        b = _outcome(lambda: _python_path(text, copy.deepcopy(table), 9))
        assert a == b, (trial, a if isinstance(a[0], str) else "data", b if isinstance(b[0], str) else "data")
        n_err += isinstance(a[0], str)
        n_ok += not isinstance(a[0], str)
    assert n_err > 30 and n_ok > 30


def test_edges():
    table, pairs = _multi_gene(41, n_pairs=4)
    for text in ("", "\n", "@HD\tVN:1.0\n", "@HD\tVN:1.0"):
        ext = fastsam.extract(text, table)
        assert ext.n_reads == 0 and ext.novel == [] and ext.reads_data()["reads"] == []
        assert all(len(ext.offsets[k]) == 1 for k in fastjson.SCAN_LISTS)
    Variant.novel_id = 0
    got = _as_dicts(fastsam.extract(_sam_text(pairs), []).reads_data())   # empty table: every positive is novel
    Variant.novel_id = 0
    assert got == _as_dicts(_python_path(_sam_text(pairs), []))
    with pytest.raises(ValueError):
        fastsam.extract("", list(reversed(table)))                # unsorted table
    with pytest.raises(ValueError):
        fastsam.extract("a\tb\tc\n", table)                       # fewer than 8 columns
    with pytest.raises(ValueError):
        fastsam.extract("a\tx\tc\t1\t60\t4M\t=\t5\n", table)      # flag is not a number
    # no trailing newline, CRLF line ends are kept in the record as the reference keeps them
    text = _sam_text(pairs).rstrip("\n")
    Variant.novel_id = 0
    a = _as_dicts(fastsam.extract(text, table).reads_data())
    Variant.novel_id = 0
    assert a == _as_dicts(_python_path(text, copy.deepcopy(table)))


def test_sam_to_calls_without_json(tmp_path):
    """extractVariantFromSam + selectKirTypingModel(_scan=...) call the same alleles as the reference's
    route: write the .json, load it into objects, type from the objects."""
    from kir_graph_b200.kir_typing import selectKirTypingModel
    from tests.fake_backend import FakeBackend
    table, pairs = _multi_gene(61, n_pairs=250, novel=0.002)
    alleles = {g: [f"{g.split('*')[0]}*{i:03d}" for i in range(7)] for g in ("KIRA*BACKBONE", "KIRB*BACKBONE")}
    rng = np.random.default_rng(6)
    for v in table:
        v.allele = [a for a in alleles[v.ref] if rng.random() < 0.4] or [alleles[v.ref][0]]
    sam_path = str(tmp_path / "s.sam")
    with open(sam_path, "w") as f:
        f.write(_sam_text(pairs))
    Variant.novel_id = 0
    ext = hisat2.extractVariantFromSam(table, sam_path, str(tmp_path / "s.variant"), num_editdist=9)
    gene_cn = {"KIRA*BACKBONE": 2, "KIRB*BACKBONE": 1}
    slow = selectKirTypingModel("full", str(tmp_path / "s.variant.json"), top_n=30, variant_correction=True,
                                _backend=FakeBackend())
    fast = selectKirTypingModel("full", "unused", top_n=30, variant_correction=True, _backend=FakeBackend(),
                                _scan=ext.scan())
    a_slow, w_slow = slow.typing(gene_cn)
    a_fast, w_fast = fast.typing(gene_cn)
    assert a_fast == a_slow and w_fast == w_slow and len(a_fast) == 3
    assert fast.getAllPossibleTyping() == slow.getAllPossibleTyping()
    with pytest.raises(ValueError):
        selectKirTypingModel("exonfirst_1", "unused", _scan=ext.scan())
    with pytest.raises(NotImplementedError):
        hisat2.extractVariantFromSam(table, sam_path, None, error_correction=True)
    assert hisat2.extractVariantFromSam(table, sam_path, None, num_editdist=9).n_reads == ext.n_reads


def test_native_json_is_byte_identical(tmp_path):
    """SamExtract.write_json == writeReadsAndVariantsData(reads_data()) byte for byte, including
    the escapes json.dump applies (quotes and backslashes in quality strings, control characters,
    non-ASCII read names as \\uXXXX / surrogate pairs), and the file loads back into equal objects."""
    table, pairs = _multi_gene(71, n_pairs=60, novel=0.004)
    table[0].id = 'hv"0\\x'                                  # an id that needs escaping
    lines = _sam_text(pairs).rstrip("\n").split("\n")
    out = []
    for i, line in enumerate(lines):
        f = line.split("\t")
        if len(f) > 10 and i % 3 == 0:
            f[10] = ('"\\/' + "\x7f\x01" + f[10])[: len(f[10])] if len(f[10]) > 6 else f[10]
        if len(f) > 10 and i % 5 < 2:
            # both mates of a pair must keep one name: rename by the name itself
            f[0] = f[0] + "é€😀"
        out.append("\t".join(f))
    sam = "\n".join(out) + "\n"
    Variant.novel_id = 3
    ext = fastsam.extract(sam, table, num_editdist=9, json_reads=True)
    assert ext.n_reads > 20
    a, b = str(tmp_path / "a.json"), str(tmp_path / "b.json")
    ext.write_json(a)
    hisat2.writeReadsAndVariantsData(ext.reads_data(), b)
    raw_a, raw_b = open(a, "rb").read(), open(b, "rb").read()
    assert raw_a == raw_b
    assert b"\\u00e9\\u20ac\\ud83d\\ude00" in raw_a and b'\\"\\\\/\\u007f\\u0001' in raw_a
    back = hisat2.loadReadsAndVariantsData(a)
    assert _as_dicts(back) == _as_dicts(ext.reads_data())
    # the scanner of the fast typing path reads the natively written file
    sc = fastjson.scan(a)
    assert sc.n_reads == ext.n_reads and sc.multiple.tolist() == ext.multiple.tolist()
    with pytest.raises(ValueError):
        fastsam.extract(sam, table, num_editdist=9).write_json(a)          # json_reads not requested
    # empty result and invalid UTF-8
    empty = fastsam.extract("", table, json_reads=True)
    empty.write_json(a)
    hisat2.writeReadsAndVariantsData(empty.reads_data(), b)
    assert open(a, "rb").read() == open(b, "rb").read()
    bad = sam.encode("utf-8").replace("é".encode("utf-8"), b"\xff\xfe")
    with pytest.raises(UnicodeDecodeError):
        fastsam.extract(bad, table, num_editdist=9, json_reads=True)


def test_extract_variant_from_bam_with_a_samtools_shim(tmp_path, monkeypatch):
    """extractVariantFromBam (the reference's name and arguments) end to end; samtools is replaced by
    a shim on PATH that serves the name-sorted SAM text ("sort -n X -O SAM"), the header ("view -H")
    and copies for "sort X.sam -o X.bam" / "index"."""
    import os
    import stat
    table, pairs = _multi_gene(81, n_pairs=50, novel=0.002)
    sam = _sam_text(pairs)
    (tmp_path / "in.bam").write_text(sam)                     # the shim treats the "bam" as SAM text
    shim = tmp_path / "bin" / "samtools"
    shim.parent.mkdir()
    shim.write_text("""#!/bin/sh
if [ "$1" = sort ] && [ "$2" = -n ]; then cat "$3"; exit 0; fi
if [ "$1" = view ] && [ "$2" = -H ]; then grep '^@' "$3"; exit 0; fi
if [ "$1" = sort ]; then cp "$2" "$4"; exit 0; fi
if [ "$1" = index ]; then : > "$2.bai"; exit 0; fi
exit 1
""")
    shim.chmod(shim.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{shim.parent}{os.pathsep}{os.environ['PATH']}")
    with pytest.raises(NotImplementedError):
        hisat2.extractVariantFromBam(table, str(tmp_path / "in.bam"), str(tmp_path / "out"))   # default: pileup
    Variant.novel_id = 0
    ext = hisat2.extractVariantFromBam(table, str(tmp_path / "in.bam"), str(tmp_path / "out"), error_correction=False,
                                       num_editdist=9)
    Variant.novel_id = 0
    want = _python_path(sam, copy.deepcopy(table), 9)
    got = hisat2.loadReadsAndVariantsData(str(tmp_path / "out.json"))
    assert _as_dicts(got) == _as_dicts(want) and ext.n_reads == len(want["reads"]) > 10
    # the two "bam" files: header + the kept records, all pairs / single-mapped pairs only (:937-940)
    body = lambda path: [l for l in open(path).read().split("\n") if l and not l.startswith("@")]
    assert body(tmp_path / "out.bam") == [x for r in want["reads"] for x in (r.l_sam, r.r_sam)]
    assert body(tmp_path / "out.no_multi.bam") == [x for r in want["reads"] if r.multiple == 1 for x in (r.l_sam, r.r_sam)]
    assert open(tmp_path / "out.bam").read().startswith("@HD") and os.path.exists(tmp_path / "out.bam.bai")
    assert not os.path.exists(tmp_path / "out.sam")
    # the object-level functions of the reference over the same shim
    pr = list(hisat2.readPair(str(tmp_path / "in.bam")))
    assert len(pr) == len(pairs) and hisat2.readBamHeader(str(tmp_path / "in.bam")).startswith("@HD")
    hisat2.saveReadsToBam(want, str(tmp_path / "obj"), str(tmp_path / "in.bam"), filter_multi_mapped=True)
    assert open(tmp_path / "obj.bam").read() == open(tmp_path / "out.no_multi.bam").read()
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    with pytest.raises(FileNotFoundError):
        hisat2.extractVariantFromBam(table, str(tmp_path / "in.bam"), str(tmp_path / "out"), error_correction=False)


def _mask_first_mismatch(record: str):
    """The read base of the first MD mismatch replaced by N (records aligned without indels only)."""
    import re
    cols = record.split("\t")
    if len(cols) < 12 or not re.fullmatch(r"\d+M", cols[5]):
        return None
    md = next((c[5:] for c in cols[11:] if c.startswith("MD:Z:")), None)
    m = re.match(r"(\d+)[ACGT]", md or "")
    if m is None:
        return None
    k = int(m.group(1))
    cols[9] = cols[9][:k] + "N" + cols[9][k + 1:]
    return "\t".join(cols)


@pytest.mark.parametrize("seed", [21, 22])
def test_bases_masked_with_n_exclude_every_alternative(seed):
    """A read base masked as N (what the reference's pileup error correction writes) becomes a novel variant
    with val N, and none of A / T / C / G at that position counts as a negative (hisat2.py:758-768): native
    loop against the Python statement."""
    table, pairs = _multi_gene(seed)
    masked, n_masked = [], 0
    for left, right in pairs:
        new = [_mask_first_mismatch(r) for r in (left, right)]
        n_masked += sum(x is not None for x in new)
        masked.append(tuple(x if x is not None else r for x, r in zip(new, (left, right))))
    assert n_masked > 10
    sam = _sam_text(masked)
    Variant.novel_id = 0
    got = fastsam.extract(sam, table, num_editdist=9).reads_data()
    Variant.novel_id = 0
    want = _python_path(sam, copy.deepcopy(table), 9)
    assert _as_dicts(got) == _as_dicts(want)
    n_variants = [v for v in got["variants"] if v.val == "N"]
    assert n_variants and all(v.id.startswith("nv") for v in n_variants)
    # the masking changes the lists: the unmasked text gives other negatives for some pair
    Variant.novel_id = 0
    plain = fastsam.extract(_sam_text(pairs), table, num_editdist=9).reads_data()
    assert _as_dicts(plain)[1] != _as_dicts(got)[1]
