"""
Fast path from name-sorted SAM text to per-read variant lists.

``hisat2.extractVariant`` follows the reference record by record (readPair -> filterRead ->
recordToVariants -> getPNFromVariantList, graphkir/hisat2.py:228-276, :541-578, :657-844): a
``Variant`` object per CIGAR segment and a ``PairRead`` per pair, 60-80 us per pair.  Here the host
routine ``gk_sam_extract`` of ``libgk_typing.so`` does the whole loop over the SAM text and returns
the four variant lists of every kept pair as CSR arrays plus the novel variants it met; objects are
built only on request (:meth:`SamExtract.reads_data`, used by the tests to compare with the object
path); the reference's ``.json`` is written natively as well (:meth:`SamExtract.write_json`, byte
for byte what ``writeReadsAndVariantsData`` writes).  :meth:`SamExtract.scan` presents the result as the
``fastjson.JsonScan`` the packing code takes, so SAM text -> ``GenePack`` needs no JSON at all.

Pileup-based read error correction is not covered (the CLI path runs with it off, main.py:149).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .fastjson import SCAN_LISTS, JsonScan
from .hisat2 import PairRead, ReadsAndVariantsData
from .msa2hisat import Variant

_TYP_CODE = {"insertion": 0, "single": 1, "deletion": 2}
_TYP_NAME = {code: name for name, code in _TYP_CODE.items()}
_ERRORS = {-3: lambda: NotImplementedError("Cannot typing with splicing"), -4: NotImplementedError,
           -5: AssertionError, -6: IndexError, -7: ValueError}


@dataclass
class SamExtract:
    """Kept read pairs of a SAM text as arrays (indices into ``table + novel``)."""

    sam: bytes
    table: list[Variant]                 # the index's sorted variant table, as passed in
    novel: list[Variant]                 # novel variants in creation order (ids nv<k>)
    refs: list[str]                      # backbone names: the table's, then new ones
    backbone: np.ndarray                 # int32 [R] index into ``refs``
    multiple: np.ndarray                 # int32 [R] NH tag of the left record
    span: np.ndarray                     # int64 [R, 4] offset, length of the left / right record
    offsets: dict[str, np.ndarray]       # list name -> int64 [R + 1]
    indices: dict[str, np.ndarray]       # list name -> int32
    n_strange: int                       # pairs skipped because the flags are not first + second mate
    reads_json: bytes | None = None      # the "reads" array of the .json, written natively (json_reads=True)

    @property
    def n_reads(self) -> int:
        return len(self.backbone)

    def variants(self) -> list[Variant]:
        """``list(variants_map.values())`` of extractVariant (:839-842): table variants (one per
        distinct key, the last duplicate), then the novel ones."""
        return list({v: v for v in self.table}.values()) + self.novel

    def ids(self) -> list[str]:
        return [str(v.id) for v in self.table] + [str(v.id) for v in self.novel]

    def reads_data(self) -> ReadsAndVariantsData:
        """The object form the reference writes to ``{prefix}.json``."""
        ids = self.ids()
        reads = []
        for r in range(self.n_reads):
            lists = {name: [ids[i] for i in self.indices[name][self.offsets[name][r]:self.offsets[name][r + 1]]]
                     for name in SCAN_LISTS}
            lo, ln, ro, rn = (int(x) for x in self.span[r])
            reads.append(PairRead(l_sam=self.sam[lo:lo + ln].decode("utf-8"), r_sam=self.sam[ro:ro + rn].decode("utf-8"),
                                  multiple=int(self.multiple[r]), backbone=self.refs[int(self.backbone[r])], **lists))
        return {"variants": self.variants(), "reads": reads}

    def write_json(self, filename: str) -> None:
        """``writeReadsAndVariantsData(self.reads_data(), filename)`` (hisat2.py:847-857), byte for byte,
        without building the read objects: the "reads" array comes from the native writer (needs
        ``extract(..., json_reads=True)``), only the variant table goes through ``json``."""
        if self.reads_json is None:
            raise ValueError("extract(..., json_reads=True) is needed for write_json")
        import json
        # asdict(v) field by field (the dataclass order of msa2hisat.Variant), without its deep copies
        rows = [{"pos": v.pos, "typ": v.typ, "ref": v.ref, "val": v.val, "id": v.id, "length": v.length,
                 "allele": list(v.allele), "freq": v.freq, "ignore": v.ignore, "in_exon": v.in_exon}
                for v in self.variants()]
        with open(filename, "wb") as handle:
            handle.write(b'{"variants": ')
            handle.write(json.dumps(rows).encode("ascii"))
            handle.write(b', "reads": [')
            handle.write(self.reads_json)
            handle.write(b"]}")

    def scan(self) -> JsonScan:
        """The same reads as the array form ``fastjson`` produces from a ``.variant.json``."""
        return JsonScan(self.backbone, self.multiple, self.offsets, self.indices, self.ids(), list(self.refs),
                        self.variants())


def _table_arrays(table: list[Variant]):
    refs: list[str] = []
    ref_of: dict[str, int] = {}
    n = len(table)
    v_ref = np.zeros(n, dtype=np.int32)
    v_pos = np.zeros(n, dtype=np.int32)
    v_typ = np.zeros(n, dtype=np.int32)
    v_int = np.zeros(n, dtype=np.int32)
    v_len = np.zeros(n, dtype=np.int32)
    vals: list[bytes] = []
    prev = None
    for i, v in enumerate(table):
        if v.typ not in _TYP_CODE:
            raise ValueError(f"variant table holds a {v.typ!r} record")
        ref = str(v.ref)
        if ref not in ref_of:
            ref_of[ref] = len(refs)
            refs.append(ref)
        if prev is not None and v < prev:
            raise ValueError("the variant table must be sorted (hisat2.getVariants returns it sorted; "
                             "getVariantsBoundary bisects it, graphkir/hisat2.py:692-713)")
        prev = v
        v_ref[i], v_pos[i], v_typ[i], v_len[i] = ref_of[ref], v.pos, _TYP_CODE[v.typ], v.length
        if v.typ == "deletion":
            v_int[i] = int(v.val)
            vals.append(b"")
        else:
            vals.append(str(v.val).encode("utf-8"))
    return refs, v_ref, v_pos, v_typ, v_int, v_len, vals


def _blob(items: list[bytes]) -> tuple[np.ndarray, bytes]:
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        np.cumsum([len(b) for b in items], out=off[1:])
    return off, b"".join(items)


def _strings(off: np.ndarray, blob: bytes) -> list[str]:
    return [blob[off[i]:off[i + 1]].decode("utf-8") for i in range(len(off) - 1)]


def extract(sam: bytes | str, variants: list[Variant], num_editdist: int = 4, json_reads: bool = False
            ) -> SamExtract:
    """Name-sorted SAM text (header lines allowed) + sorted index variants -> :class:`SamExtract`.

    Equivalent to ``extractVariant(filter(filterRead both, readPair(...)), variants)`` of the
    reference with ``pileup=None``; ``Variant.novel_id`` advances by the number of novel variants.
    A malformed record raises what the reference raises for it (NotImplementedError for splicing
    or an unknown CIGAR operation, AssertionError, IndexError, ValueError)."""
    if isinstance(sam, str):
        sam = sam.encode("utf-8")
    lib = _cabi.load()
    refs, v_ref, v_pos, v_typ, v_int, v_len, vals = _table_arrays(variants)
    val_off, val_blob = _blob(vals)
    ref_off, ref_blob = _blob([r.encode("utf-8") for r in refs])
    fn = lib.gk_sam_extract
    fn.restype = ctypes.c_void_p
    fn.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int32] + [ctypes.c_void_p] * 6 + [
        ctypes.c_char_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
        ctypes.POINTER(ctypes.c_int64)]
    lib.gk_sam_extract_free.argtypes = [ctypes.c_void_p]
    lib.gk_sam_extract_free.restype = None
    sizes = (ctypes.c_int64 * 12)()
    novel_base = Variant.novel_id
    handle = fn(sam, len(sam), len(variants), v_ref.ctypes.data, v_pos.ctypes.data, v_typ.ctypes.data,
                v_int.ctypes.data, v_len.ctypes.data, val_off.ctypes.data, val_blob, len(refs),
                ref_off.ctypes.data, ref_blob, novel_base, num_editdist, sizes)
    try:
        n, n_novel = int(sizes[0]), int(sizes[5])
        if sizes[9] != 0:
            # the novel variants met before the failing record keep their numbers, as in the reference
            Variant.novel_id = novel_base + n_novel
            err = _ERRORS.get(int(sizes[9]), ValueError)()
            err.args = (*err.args, f"SAM line {int(sizes[11])}")
            raise err
        backbone = np.zeros(n, dtype=np.int32)
        multiple = np.zeros(n, dtype=np.int32)
        span = np.zeros((n, 4), dtype=np.int64)
        offs = [np.zeros(n + 1, dtype=np.int64) for _ in range(4)]
        idxs = [np.zeros(max(int(sizes[1 + w]), 1), dtype=np.int32) for w in range(4)]
        nv = [np.zeros(max(n_novel, 1), dtype=np.int32) for _ in range(5)]
        nv_val_off = np.zeros(n_novel + 1, dtype=np.int64)
        nv_blob = ctypes.create_string_buffer(max(int(sizes[6]), 1))
        out_ref_off = np.zeros(int(sizes[7]) + 1, dtype=np.int64)
        out_ref_blob = ctypes.create_string_buffer(max(int(sizes[8]), 1))
        off_ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in offs])
        idx_ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in idxs])
        lib.gk_sam_extract_fill.argtypes = [ctypes.c_void_p] * 15
        rc = lib.gk_sam_extract_fill(handle, multiple.ctypes.data, backbone.ctypes.data, span.ctypes.data, off_ptrs,
                                     idx_ptrs, *[a.ctypes.data for a in nv], nv_val_off.ctypes.data, nv_blob,
                                     out_ref_off.ctypes.data, out_ref_blob)
        if rc != 0:
            raise ValueError(lib.gk_last_error().decode())
        reads_json = None
        if json_reads:
            id_off, id_blob = _blob([str(v.id).encode("utf-8") for v in variants])
            out, out_len = ctypes.c_char_p(), ctypes.c_int64()
            lib.gk_sam_extract_json.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p,
                                                ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p),
                                                ctypes.POINTER(ctypes.c_int64)]
            if lib.gk_sam_extract_json(handle, sam, id_off.ctypes.data, id_blob, novel_base, ctypes.byref(out),
                                       ctypes.byref(out_len)) != 0:
                raise UnicodeDecodeError("utf-8", b"", 0, 1, lib.gk_last_error().decode())
            reads_json = ctypes.string_at(out, out_len.value)
    finally:
        lib.gk_sam_extract_free(handle)
    all_refs = _strings(out_ref_off, out_ref_blob.raw)
    nv_vals = _strings(nv_val_off, nv_blob.raw)
    novel = []
    for j in range(n_novel):
        typ = _TYP_NAME[int(nv[2][j])]
        val = int(nv[3][j]) if typ == "deletion" else nv_vals[j]
        novel.append(Variant(pos=int(nv[1][j]), typ=typ, ref=all_refs[int(nv[0][j])], val=val,
                             id=f"nv{novel_base + j}", length=int(nv[4][j])))
    Variant.novel_id = novel_base + n_novel
    return SamExtract(sam, variants, novel, all_refs, backbone, multiple, span,
                      {name: offs[w] for w, name in enumerate(SCAN_LISTS)},
                      {name: idxs[w][: int(sizes[1 + w])] for w, name in enumerate(SCAN_LISTS)}, int(sizes[10]),
                      reads_json)


def extract_file(filename: str, variants: list[Variant], num_editdist: int = 4, json_reads: bool = False
                 ) -> SamExtract:
    """:func:`extract` over a name-sorted ``.sam`` file (``samtools sort -n | samtools view -h``
    output; running samtools stays in the reference, hisat2.py:205-225)."""
    with open(filename, "rb") as handle:
        return extract(handle.read(), variants, num_editdist, json_reads)
