"""
Read records of the typing path: ``PairRead`` and the ``.variant.json`` contract.

Field names and JSON layout follow the reference (graphkir/hisat2.py:24-52 for
``PairRead``, :847-866 for the JSON reader/writer, :943-948 for the
multi-mapping filter) so files are interchangeable.  Running HISAT2/samtools is
out of scope and stays in the reference.
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from typing import TypedDict

from .msa2hisat import Variant


@dataclass
class PairRead:
    """A read pair with the variant ids it supports (p) or contradicts (n)."""

    l_sam: str = ""
    r_sam: str = ""
    multiple: int = 1
    backbone: str = ""
    lpv: list[str] = field(default_factory=list)
    lnv: list[str] = field(default_factory=list)
    rpv: list[str] = field(default_factory=list)
    rnv: list[str] = field(default_factory=list)


class ReadsAndVariantsData(TypedDict):
    variants: list[Variant]
    reads: list[PairRead]


def writeReadsAndVariantsData(reads_data: ReadsAndVariantsData, filename: str) -> None:
    """Serialise to the reference's ``{prefix}.json`` layout."""
    payload = {
        "variants": [asdict(v) for v in reads_data["variants"]],
        "reads": [asdict(r) for r in reads_data["reads"]],
    }
    with open(filename, "w") as handle:
        json.dump(payload, handle)


def loadReadsAndVariantsData(filename: str) -> ReadsAndVariantsData:
    """Inverse of :func:`writeReadsAndVariantsData`."""
    with open(filename) as handle:
        raw = json.load(handle)
    return {
        "variants": [Variant(**item) for item in raw["variants"]],
        "reads": [PairRead(**item) for item in raw["reads"]],
    }


def removeMultipleMapped(reads_data: ReadsAndVariantsData) -> ReadsAndVariantsData:
    """Keep pairs whose NH tag is exactly one."""
    return {
        "variants": reads_data["variants"],
        "reads": [r for r in reads_data["reads"] if r.multiple == 1],
    }


# ---------------------------------------------------------------------------
# SAM record -> per-read variant ids.
# Host-side text/integer logic; the subprocess wrappers (hisat2, samtools) stay in
# the reference.  Each function cites the reference code whose behaviour it keeps.
# ---------------------------------------------------------------------------
import bisect
import copy
import os
import re
from typing import Iterable

from .utils import logger

_CIGAR_RE = re.compile(r"(\d+)(\w)")
_MD_RE = re.compile(r"\d+|.")


def getNH(sam_info: str) -> int:
    """NH tag of a record, 1 when absent (reference: hisat2.py:95-100)."""
    found = re.search(r"NH:i:(\d+)", sam_info)
    return int(found.group(1)) if found else 1


def readZs(cols: list[str]) -> list[tuple[int, str, str]]:
    """``Zs:Z:43|D|hv862,19|D|hv868`` -> [(43, 'D', 'hv862'), (19, 'D', 'hv868')] (:518-527)."""
    for col in cols:
        if col.startswith("Zs"):
            out = []
            for item in col[5:].split(","):
                # fields beyond the third are ignored and a missing one is an IndexError after the gap
                # has been parsed, as in the reference's (int(i[0]), i[1], i[2])
                fields = item.split("|")
                out.append((int(fields[0]), fields[1], fields[2]))
            return out
    return []


def readMd(cols: list[str]) -> list[int | str]:
    """``MD:Z:43^G19`` -> [43, '^', 'G', 19]: numbers as int, everything else per character (:530-538)."""
    for col in cols:
        if col.startswith("MD"):
            return [int(tok) if tok.isdigit() else tok for tok in _MD_RE.findall(col[5:])]
    return []


def filterRead(line: str, num_editdist: int = 4) -> bool:
    """Keep properly paired records (flag & 2) whose NM tag exists and is <= num_editdist (:541-578)."""
    fields = line.strip().split("\t")
    if int(fields[1]) & 2 == 0:
        return False
    nm = None
    for col in fields[11:]:
        if col.startswith("NM"):
            nm = int(col[5:])
    return nm is not None and nm <= num_editdist


class _RecordWalker:
    """Walks CIGAR x MD x Zs of one record.

    State (same meaning as in the reference walk, hisat2.py:342-355):
      pos     reference position of the current CIGAR operation
      read_i  read offset of the current CIGAR operation
      md_i    next MD token;   md_len  reference bases of the MD match run already read but not
              yet consumed by CIGAR (it carries across an insertion, whose bases MD does not list)
      zs_i    next Zs entry;   zs_pos  read offset up to which Zs gaps have been consumed
    """

    def __init__(self, line: str):
        cols = line.strip().split("\t")
        self.backbone = cols[2]
        self.pos = int(cols[3]) - 1
        self.cigar = [(op, int(n)) for n, op in _CIGAR_RE.findall(cols[5])]
        self.seq = cols[9]
        self.zs = readZs(cols[11:])
        self.md = readMd(cols[11:])
        self.read_i = 0
        self.md_i = 0
        self.md_len = 0
        self.zs_i = 0
        self.zs_pos = 0

    def _known_id(self, kind: str) -> str:
        """Id of the Zs entry sitting exactly at the current read offset, else 'unknown' (:357-371)."""
        if self.zs_i < len(self.zs):
            gap, zs_kind, vid = self.zs[self.zs_i]
            if zs_kind == kind and self.read_i + self.md_len == self.zs_pos + gap:
                self.zs_pos += gap + (1 if kind == "S" else 0)
                self.zs_i += 1
                return vid
        return "unknown"

    def _skip_zero(self) -> None:
        if self.md_i < len(self.md) and self.md[self.md_i] == 0:
            self.md_i += 1

    def _match(self, length: int) -> list[Variant]:
        """An M operation: match runs split by mismatches (:373-446)."""
        out = []
        done = 0                      # reference bases of this operation already emitted
        while True:
            if self.md_len <= done and self.md_i < len(self.md) and type(self.md[self.md_i]) is int:
                self.md_len += self.md[self.md_i]
                self.md_i += 1
            if self.md_len >= length:                       # the run reaches the end of the operation
                self.md_len -= length
                out.append(Variant(typ="match", ref=self.backbone, pos=self.pos + done, length=length - done))
                return out
            base = self.seq[self.read_i + self.md_len]      # mismatching read base
            if self.md[self.md_i] == 0:
                self.md_i += 1
            assert str(self.md[self.md_i]) in "ACGT"
            assert str(self.md[self.md_i]) != base
            self.md_i += 1
            if self.md_len > done:
                out.append(Variant(typ="match", ref=self.backbone, pos=self.pos + done,
                                   length=self.md_len - done))
            out.append(Variant(typ="single", ref=self.backbone, pos=self.pos + self.md_len, length=1,
                               val=base, id=self._known_id("S")))
            self.md_len += 1
            done = self.md_len
            if self.md_len == length:
                self.md_len = 0
                return out

    def walk(self) -> tuple[list[Variant], list[int]]:
        segments: list[Variant] = []
        soft_clip = [0, 0]
        for i, (op, length) in enumerate(self.cigar):
            self._skip_zero()
            if op == "M":
                segments.extend(self._match(length))
            elif op == "I":
                segments.append(Variant(typ="insertion", ref=self.backbone, pos=self.pos,
                                        val=self.seq[self.read_i:self.read_i + length], length=length,
                                        id=self._known_id("I")))
            elif op == "D":
                assert self.md[self.md_i] == "^"
                self.md_i += 1
                while (self.md_i < len(self.md) and type(self.md[self.md_i]) is not int
                       and str(self.md[self.md_i]) in "ACGT"):
                    self.md_i += 1
                segments.append(Variant(typ="deletion", ref=self.backbone, pos=self.pos, val=length,
                                        length=length, id=self._known_id("D")))
            elif op == "S":
                soft_clip[0 if i == 0 else 1] = length
                self.zs_pos += length
            elif op == "N":
                raise NotImplementedError("Cannot typing with splicing")
            else:
                raise NotImplementedError
            if op in "MND":
                self.pos += length
            if op in "MIS":
                self.read_i += length
        self._skip_zero()
        assert self.zs_i == len(self.zs)
        assert self.md_i == len(self.md)
        assert self.read_i == len(self.seq)
        return segments, soft_clip


def recordToRawVariantPy(line: str) -> tuple[list[Variant], list[int]]:
    """Python statement of ``recordToRawVariant`` (the tests compare the C++ walk with it)."""
    return _RecordWalker(line).walk()


_WALK_ERRORS = {-3: lambda: NotImplementedError("Cannot typing with splicing"), -4: NotImplementedError,
                -5: AssertionError, -6: IndexError, -7: ValueError}
_walk_fn = None
_walk_seg = None
_walk_meta = None


def recordToRawVariant(line: str) -> tuple[list[Variant], list[int]]:
    """One SAM record -> (match / single / insertion / deletion segments with 0-based backbone
    positions, [head soft clip, tail soft clip])  (reference: hisat2.py:279-515), through the host
    routine ``gk_sam_walk`` of libgk_typing.so.  (Most of the time per record is the construction of
    the ``Variant`` objects the reference's interface asks for, not the walk.)"""
    global _walk_fn, _walk_seg, _walk_meta
    import ctypes
    if _walk_fn is None:
        from . import _cabi
        fn = _cabi.load().gk_sam_walk
        fn.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        _walk_fn, _walk_seg, _walk_meta = fn, (ctypes.c_int32 * (7 * 64))(), (ctypes.c_int32 * 4)()
    buf = line.encode("utf-8")
    seg_buf, meta, max_seg = _walk_seg, _walk_meta, 64
    n = _walk_fn(buf, len(buf), seg_buf, max_seg, meta)
    if n == -8:                                            # more segments than the reusable buffer holds
        max_seg = len(buf) + 1
        seg_buf = (ctypes.c_int32 * (7 * max_seg))()
        n = _walk_fn(buf, len(buf), seg_buf, max_seg, meta)
    if n < 0:
        raise _WALK_ERRORS[n]()
    flat = seg_buf[:7 * n]
    seg = [flat[i:i + 7] for i in range(0, 7 * n, 7)]
    backbone = buf[meta[2]:meta[2] + meta[3]].decode("utf-8")
    out = []
    for typ, pos, length, v_off, v_len, i_off, i_len in seg:
        vid = "unknown" if i_len == -1 else buf[i_off:i_off + i_len].decode("utf-8")
        if typ == 0:
            out.append(Variant(typ="match", ref=backbone, pos=pos, length=length))
        elif typ == 1:
            out.append(Variant(typ="single", ref=backbone, pos=pos, length=1,
                               val=buf[v_off:v_off + v_len].decode("utf-8"), id=vid))
        elif typ == 2:
            out.append(Variant(typ="insertion", ref=backbone, pos=pos, val=buf[v_off:v_off + v_len].decode("utf-8"),
                               length=length, id=vid))
        else:
            out.append(Variant(typ="deletion", ref=backbone, pos=pos, val=length, length=length, id=vid))
    return out, [int(meta[0]), int(meta[1])]


def findVariantId(variant: Variant, variants_map: dict[Variant, Variant]) -> Variant:
    """Known variant with the same (pos, ref, typ, val), else a new ``nv<k>`` which is also added
    to the map; match segments pass through (:581-606)."""
    if variant in variants_map:
        return variants_map[variant]
    if variant.typ in ("single", "insertion", "deletion"):
        variant.id = f"nv{Variant.novel_id}"
        Variant.novel_id += 1
        variants_map[variant] = variant
        return variant
    assert variant.typ == "match"
    return variant


def recordToVariants(record: str, variants_map: dict[Variant, Variant], pileup=None,
                     ignore_softclip: bool = False) -> list[Variant]:
    """Sorted, id-annotated segments of a record; a soft-clipped record yields nothing (:657-689)."""
    if pileup:
        raise NotImplementedError("pileup-based read error correction stays in the reference "
                                  "(the CLI path passes error_correction=False, main.py:149)")
    segments, soft_clip = recordToRawVariant(record)
    if not ignore_softclip and sum(soft_clip) > 0:
        return []
    return sorted(findVariantId(v, variants_map) for v in segments)


def getVariantsBoundary(read_variants: list[Variant], variants: list[Variant]) -> tuple[int, int]:
    """Index window [left, right) of the sorted variant table covered by the read (:692-713)."""
    first, last = read_variants[0], read_variants[-1]
    lo = Variant(ref=first.ref, pos=first.pos, typ="single", val="A")
    hi = Variant(ref=first.ref, pos=last.pos + last.length, typ="single", val="T")
    return bisect.bisect_left(variants, lo), bisect.bisect_left(variants, hi)


def getPNFromVariantList(read_variants: list[Variant], variants: list[Variant], exon_only: bool = False,
                         discard_novel_index: bool = True) -> tuple[list[Variant], list[Variant]]:
    """Positive variants = the non-match segments of the read; negative variants = every table
    variant inside the read's window that the read does not carry (:716-800)."""
    if not read_variants:
        return [], []
    left, right = getVariantsBoundary(read_variants, variants)
    read_end = read_variants[-1].pos + read_variants[-1].length
    assert left <= right
    if discard_novel_index:
        for kind in ("insertion", "deletion"):       # a novel indel is taken as a mapping error
            if any(v.typ == kind and v.id.startswith("nv") for v in read_variants):
                return [], []
    excluded = set()
    for v in read_variants:
        if v.val == "N":                              # base masked by error correction
            for base in "ATCG":
                alt = copy.deepcopy(v)
                alt.val = base
                excluded.add(alt)
    carried = [v for v in read_variants if v.typ != "match"]
    positives = [v for v in carried if v.in_exon] if exon_only else carried
    excluded.update(positives)
    negatives = []
    for v in variants[left:right]:
        if v in excluded or (exon_only and not v.in_exon):
            continue
        if v.typ == "deletion" and v.pos + v.val + 10 >= read_end:    # ambiguous near the read end
            continue
        negatives.append(v)
    return positives, negatives


def extractVariant(pair_reads: Iterable[tuple[str, str]], variants: list[Variant], pileup=None
                   ) -> ReadsAndVariantsData:
    """(left record, right record) pairs -> PairRead list + variant list incl. novel ones (:803-844)."""
    variants_map = {v: v for v in variants}
    reads = []
    for left_record, right_record in pair_reads:
        lv = recordToVariants(left_record, variants_map, pileup)
        rv = recordToVariants(right_record, variants_map, pileup)
        lp, ln = getPNFromVariantList(lv, variants)
        rp, rn = getPNFromVariantList(rv, variants)
        reads.append(PairRead(
            lpv=[v.id for v in lp if v.id is not None], lnv=[v.id for v in ln if v.id is not None],
            rpv=[v.id for v in rp if v.id is not None], rnv=[v.id for v in rn if v.id is not None],
            l_sam=left_record, r_sam=right_record, multiple=getNH(left_record),
            backbone=left_record.split("\t")[2]))
    logger.info(f"[Graph] Filterd pairs: {len(reads)}")
    return {"variants": list(variants_map.values()), "reads": reads}


def _write_sidecar(ext, output_prefix: str) -> None:
    """``{output_prefix}.gkpack.npz``: the genes of an extraction packed as the typing step wants them
    (variant correction on, single-mapped pairs: what main.alleleTyping asks for), so that typing a
    cohort does not parse the ``.json`` again."""
    from . import fastjson, packio
    packs = fastjson.packs_from_scan(ext.scan(), variant_correction=True, single_mapped_only=True)
    side = packio.sidecar_path(output_prefix)
    try:
        packio.save_packs(side, packs, packio.sidecar_meta(output_prefix))
    except ValueError as exc:
        # the sidecar is a cache: a gene beyond a capacity of the device path is reported per gene when the
        # sample is typed from its .json; the extraction itself goes on (and leaves no older sidecar behind)
        logger.warning(f"[Graph] No packed sidecar for {output_prefix}: {exc}")
        if os.path.exists(side):
            os.remove(side)


def extractVariantFromSam(index: str | list[Variant], sam_file: str, output_prefix: str | None,
                          error_correction: bool = False, num_editdist: int = 4, write_pack: bool = False):
    """``extractVariantFromBam`` (:904-940) over a name-sorted SAM file (what the reference's
    ``readBam`` pipes out of ``samtools sort -n | samtools view -h``, :205-225; running samtools stays
    in the reference): filter, call, annotate, and write ``{output_prefix}.json`` in the reference's
    layout.  The loop itself is the native batch routine (:mod:`kir_graph_b200.fastsam`); the
    returned ``SamExtract`` holds the same reads as arrays (``.scan()`` feeds the typing driver
    directly, ``output_prefix=None`` skips the JSON; ``write_pack`` adds the ``.gkpack.npz`` sidecar that
    ``main.cohortAlleleTyping`` loads instead of the JSON).  ``index``: HISAT2 index prefix, or the sorted
    variant table itself.  The ``.bam`` copies the reference also writes (:937-940) need samtools and
    are not produced."""
    if error_correction:
        raise NotImplementedError("pileup-based read error correction stays in the reference "
                                  "(the CLI path passes error_correction=False, main.py:149)")
    from . import fastsam
    variants = getVariants(index) if isinstance(index, str) else index
    ext = fastsam.extract_file(sam_file, variants, num_editdist, json_reads=output_prefix is not None)
    logger.info(f"[Graph] Filterd pairs: {ext.n_reads}")
    if output_prefix is not None:
        logger.debug(f"[Graph] Save allele per reads in {output_prefix}.json")
        ext.write_json(f"{output_prefix}.json")               # == writeReadsAndVariantsData(ext.reads_data(), ...)
        if write_pack:
            _write_sidecar(ext, output_prefix)
    return ext


def _samtools(args: list[str], capture: bool = False) -> str:
    """One local ``samtools`` call (the reference's runTool with the local engine,
    external_tools.py:161-186 -> utils.py:88-103); containers stay in the reference."""
    import subprocess
    try:
        proc = subprocess.run(["samtools", *args], capture_output=True, check=True, universal_newlines=True)
    except FileNotFoundError as exc:
        raise FileNotFoundError("samtools is not on PATH: pass a name-sorted SAM file to "
                                "extractVariantFromSam instead (mapping and BAM handling stay in the "
                                "reference)") from exc
    return proc.stdout if capture else ""


def readBam(bam_file: str) -> list[str]:
    """Name-sorted SAM lines of a bam file via a local samtools (:103-110)."""
    return _samtools(["sort", "-n", bam_file, "-O", "SAM"], capture=True).split("\n")


def readBamHeader(bam_file: str) -> str:
    """Header of a bam file via a local samtools (:113-118)."""
    return _samtools(["view", "-H", bam_file], capture=True)


def readPair(bam_file: str) -> Iterable[tuple[str, str]]:
    """(left record, right record) of every proper pair of a bam file (:228-276)."""
    return pairRecords(readBam(bam_file))


def saveSam(filename: str, header: str, reads: Iterable[PairRead]) -> None:
    """Header + both records of every read pair (:869-881)."""
    with open(filename, "w") as handle:
        handle.writelines(header)
        handle.writelines(r.l_sam + "\n" + r.r_sam + "\n" for r in reads)


def saveReadsToBam(reads_data: ReadsAndVariantsData, filename_prefix: str, bam_file: str,
                   filter_multi_mapped: bool = False) -> None:
    """``{filename_prefix}.bam`` (sorted, indexed) of the kept pairs, header taken from ``bam_file``
    (:884-901, samtobam utils.py:106-116)."""
    reads = reads_data["reads"]
    if filter_multi_mapped:
        reads = [r for r in reads if r.multiple == 1]
    saveSam(filename_prefix + ".sam", readBamHeader(bam_file), reads)
    _samtools(["sort", filename_prefix + ".sam", "-o", filename_prefix + ".bam"])
    _samtools(["index", filename_prefix + ".bam"])
    os.remove(filename_prefix + ".sam")


def extractVariantFromBam(index: str | list[Variant], bam_file: str, output_prefix: str,
                          error_correction: bool = True, num_editdist: int = 4, write_pack: bool = False):
    """The reference's entry point (:904-940), same arguments: reads ``samtools sort -n bam -O SAM``
    (readBam, :103-110), runs the native extraction loop, writes ``{output_prefix}.json`` and - as the
    reference does through samtools - ``{output_prefix}.bam`` and ``{output_prefix}.no_multi.bam``
    (saveReadsToBam, :884-901).  ``error_correction`` defaults to True as in the reference, which this
    implementation refuses (the CLI path passes False, main.py:149)."""
    if error_correction:
        raise NotImplementedError("pileup-based read error correction stays in the reference "
                                  "(the CLI path passes error_correction=False, main.py:149)")
    from . import fastsam
    variants = getVariants(index) if isinstance(index, str) else index
    text = _samtools(["sort", "-n", bam_file, "-O", "SAM"], capture=True)
    ext = fastsam.extract(text, variants, num_editdist, json_reads=True)
    logger.info(f"[Graph] Filterd pairs: {ext.n_reads}")
    logger.debug(f"[Graph] Save allele per reads in {output_prefix}.json")
    ext.write_json(f"{output_prefix}.json")
    if write_pack:
        _write_sidecar(ext, output_prefix)
    header = _samtools(["view", "-H", bam_file], capture=True)
    for prefix, keep in ((output_prefix, None), (output_prefix + ".no_multi", ext.multiple == 1)):
        with open(prefix + ".sam", "wb") as handle:                   # saveSam (:869-881)
            handle.write(header.encode("utf-8"))
            for r in range(ext.n_reads):
                if keep is None or keep[r]:
                    lo, ln, ro, rn = (int(x) for x in ext.span[r])
                    handle.write(ext.sam[lo:lo + ln] + b"\n" + ext.sam[ro:ro + rn] + b"\n")
        _samtools(["sort", prefix + ".sam", "-o", prefix + ".bam"])    # samtobam (utils.py:106-116)
        _samtools(["index", prefix + ".bam"])
        os.remove(prefix + ".sam")
    return ext


def pairRecords(lines: Iterable[str]) -> Iterable[tuple[str, str]]:
    """Pair up name-sorted SAM records (the body of the reference's readPair, :228-276; the
    ``samtools sort -n`` subprocess that feeds it stays in the reference)."""
    pending: dict[tuple[str, str, str, int], str] = {}
    for line in lines:
        if not line or line.startswith("@") or line.startswith("[bam_sort_core]"):
            continue
        name, flag_str, ref, pos, _, _, next_ref, next_pos = line.split("\t")[:8]
        if next_ref != "=":
            continue
        flag = int(flag_str)
        mate_key = (name, ref, next_pos, flag & 256)
        if mate_key in pending:
            mate = pending[mate_key]
            if ((int(mate.split("\t")[1]) | flag) & 192) != 192:
                logger.warning(f"[Graph] Read Pair strange case: {line} {mate}")
                continue
            del pending[mate_key]
            yield line, mate
        else:
            pending[(name, ref, pos, flag & 256)] = line


# --- index files (.snp / .link / .locus), reference: hisat2.py:121-225 -----------------------
def readLink(index: str) -> dict[str, list[str]]:
    out = {}
    with open(index + ".link") as handle:
        for line in handle:
            vid, alleles = line.strip().split("\t")
            out[vid] = alleles.split()
    return out


def readExons(index: str) -> dict[str, list[tuple[int, int]]]:
    out = {}
    with open(index + ".locus") as handle:
        for line in handle:
            gene, _, _, _, _, exons, _ = line.split("\t")
            out[gene] = [(int(e.split("-")[0]) - 1, int(e.split("-")[1]) - 1) for e in exons.split(" ")]
    return out


def readVariants(index: str) -> list[Variant]:
    out = []
    with open(index + ".snp") as handle:
        for line in handle:
            vid, typ, ref, pos, val = line.strip().split("\t")
            out.append(Variant(typ=typ, ref=ref, pos=int(pos), id=vid,
                               val=int(val) if typ == "deletion" else val))
    return out


def isInExon(exons: list[tuple[int, int]], variant: Variant) -> bool:
    for start, end in exons:
        if start <= variant.pos < end:
            return True
        if variant.typ == "deletion" and variant.pos < start and variant.pos + variant.val >= start:
            return True
    return False


def getVariants(index: str) -> list[Variant]:
    """Sorted variant table of a HISAT2 index with allele lists and exon flags (:183-203)."""
    variants = readVariants(index)
    links = readLink(index)
    exons = readExons(index)
    for v in variants:
        v.allele = links.get(v.id, [])
        v.in_exon = isInExon(exons[v.ref], v)
    return sorted(variants)
