"""
Fast path from ``{prefix}.variant.json`` to packed gene problems.

``kir_typing.TypingWithPosNegAllele`` follows the reference (``loadReadsAndVariantsData`` ->
``removeMultipleMapped`` -> ``groupReads`` -> one ``AlleleTyping`` per gene, kir_typing.py:92-97,
hisat2.py:859-866, :943-948): ``json.load`` of ~1 KB of SAM text per read pair, a dataclass per read,
then per-object packing - seconds per 200k-pair sample.  Here the C++ scanner of
``libgk_typing.so`` (``gk_json_scan``) extracts the four variant-id lists, ``backbone`` and
``multiple`` of every read straight from the JSON text as CSR arrays; only the small "variants"
array goes through ``json``.  The result is the same ``GenePack`` per gene that ``pack_gene`` builds
from objects (tests/test_fastjson.py compares them field by field).
"""
from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .msa2hisat import Variant
from .packing import LIST_NAMES, GenePack, ReadCSR, pack_gene_csr

SCAN_LISTS = ("lpv", "lnv", "rpv", "rnv")          # order of the scanner's four CSR lists


@dataclass
class JsonScan:
    """Reads of a ``.variant.json`` as arrays."""

    backbone: np.ndarray                 # int32 [R]  index into ``genes``
    multiple: np.ndarray                 # int32 [R]
    offsets: dict[str, np.ndarray]       # list name -> int64 [R + 1]
    indices: dict[str, np.ndarray]       # list name -> int32, index into ``ids``
    ids: list[str]                       # distinct variant ids
    genes: list[str]                     # distinct backbone names
    variants: list[Variant]              # the "variants" array (all genes)
    _id_index: dict | None = None

    @property
    def n_reads(self) -> int:
        return len(self.backbone)

    @property
    def id_index(self) -> dict[str, int]:
        """variant id string -> index into ``ids`` (built once)."""
        if self._id_index is None:
            self._id_index = {vid: i for i, vid in enumerate(self.ids)}
        return self._id_index


def _strings(off: np.ndarray, blob: bytes) -> list[str]:
    return [blob[off[i]:off[i + 1]].decode("utf-8") for i in range(len(off) - 1)]


def scan(filename: str) -> JsonScan:
    """Scan ``filename`` (``.json`` appended when missing, as loadReadsAndVariantsData does)."""
    if not filename.endswith(".json"):
        filename += ".json"
    with open(filename, "rb") as f:
        buf = f.read()
    return scan_bytes(buf)


def scan_bytes(buf: bytes) -> JsonScan:
    lib = _cabi.load()
    lib.gk_json_scan.restype = ctypes.c_void_p
    lib.gk_json_scan.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    lib.gk_json_free.argtypes = [ctypes.c_void_p]
    lib.gk_json_free.restype = None
    sizes = (ctypes.c_int64 * 11)()
    handle = lib.gk_json_scan(buf, len(buf), sizes)
    if not handle:
        raise ValueError(lib.gk_last_error().decode())
    try:
        n = int(sizes[0])
        backbone = np.zeros(n, dtype=np.int32)
        multiple = np.zeros(n, dtype=np.int32)
        offs = [np.zeros(n + 1, dtype=np.int64) for _ in range(4)]
        idxs = [np.zeros(max(int(sizes[1 + w]), 1), dtype=np.int32) for w in range(4)]
        id_off = np.zeros(int(sizes[5]) + 1, dtype=np.int64)
        id_blob = ctypes.create_string_buffer(max(int(sizes[6]), 1))
        gene_off = np.zeros(int(sizes[7]) + 1, dtype=np.int64)
        gene_blob = ctypes.create_string_buffer(max(int(sizes[8]), 1))
        off_ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in offs])
        idx_ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in idxs])
        lib.gk_json_fill.argtypes = [ctypes.c_void_p] * 9
        rc = lib.gk_json_fill(handle, backbone.ctypes.data, multiple.ctypes.data, off_ptrs, idx_ptrs,
                              id_off.ctypes.data, id_blob, gene_off.ctypes.data, gene_blob)
        if rc != 0:
            raise ValueError(lib.gk_last_error().decode())
    finally:
        lib.gk_json_free(handle)
    variants: list[Variant] = []
    if sizes[9] >= 0:
        variants = [Variant(**v) for v in json.loads(buf[int(sizes[9]):int(sizes[10])])]
    return JsonScan(
        backbone, multiple,
        {name: offs[w] for w, name in enumerate(SCAN_LISTS)},
        {name: idxs[w][: int(sizes[1 + w])] for w, name in enumerate(SCAN_LISTS)},
        _strings(id_off, id_blob.raw), _strings(gene_off, gene_blob.raw), variants)


def _id_lut(sc: JsonScan, vid_to_idx: dict[str, int]) -> np.ndarray:
    lut = np.full(len(sc.ids) + 1, -1, dtype=np.int64)
    for vid, j in vid_to_idx.items():
        i = sc.id_index.get(vid)
        if i is not None:
            lut[i] = j
    return lut


def gene_csr(sc: JsonScan, gene: str, vid_to_idx: dict[str, int], single_mapped_only: bool = True
             ) -> tuple[ReadCSR, np.ndarray]:
    """CSR lists (variant indices of ``gene``'s table) of the reads whose backbone is ``gene``, in
    file order, and their indices in the file.  ``single_mapped_only`` applies removeMultipleMapped
    (hisat2.py:943-948).  Raises KeyError for an id that is not in the gene's variant table, like
    ``self.variants[i]`` in the reference (typing_mulit_allele.py:364)."""
    if gene not in sc.genes:
        rows = np.zeros(0, dtype=np.int64)
    else:
        sel = sc.backbone == sc.genes.index(gene)
        if single_mapped_only:
            sel &= sc.multiple == 1
        rows = np.flatnonzero(sel)
    lut = _id_lut(sc, vid_to_idx)
    offsets, indices = {}, {}
    for name in LIST_NAMES:
        off = sc.offsets[name]
        lens = (off[1:] - off[:-1])[rows]
        new_off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=new_off[1:])
        # positions of the selected reads' ids in the flat list
        start = np.repeat(off[:-1][rows] - new_off[:-1], lens)
        flat = sc.indices[name][start + np.arange(int(new_off[-1]), dtype=np.int64)]
        mapped = lut[flat]
        if len(mapped) and mapped.min() < 0:
            raise KeyError(sc.ids[int(flat[int(np.argmin(mapped))])])
        offsets[name] = new_off
        indices[name] = mapped.astype(np.int32)
    return ReadCSR(len(rows), offsets, indices), rows


def group_variants(variants: list[Variant]) -> dict[str, list[Variant]]:
    """Variants per gene in file order (groupVariants, kir_typing.py:23-28)."""
    by_gene: dict[str, list[Variant]] = {}
    for v in variants:
        by_gene.setdefault(str(v.ref), []).append(v)
    return by_gene


def packs_from_scan(sc: JsonScan, genes: list[str] | None = None, variant_correction: bool = True,
                    no_empty: bool = True, single_mapped_only: bool = True) -> dict[str, GenePack]:
    """A gene whose reads exceed a capacity of the device path (``packing.CapacityError``) maps to the
    exception instead of a pack: the drivers call it ``fail`` with a warning, the other genes of the
    sample are typed."""
    from .packing import CapacityError
    packs: dict[str, GenePack] = {}
    for gene, variants in group_variants(sc.variants).items():
        if genes is not None and gene not in genes:
            continue
        ids = list({str(v.id): None for v in variants})
        csr, _ = gene_csr(sc, gene, {vid: i for i, vid in enumerate(ids)}, single_mapped_only)
        try:
            packs[gene] = pack_gene_csr(variants, csr, variant_correction=variant_correction, no_empty=no_empty,
                                        gene=gene)
        except CapacityError as exc:
            packs[gene] = exc
    return packs


def load_packs(filename: str, genes: list[str] | None = None, variant_correction: bool = True,
               no_empty: bool = True, single_mapped_only: bool = True) -> dict[str, GenePack]:
    """``.variant.json`` -> one ``GenePack`` per gene (insertion order = order of the variant table's
    genes, as groupVariants yields them, kir_typing.py:23-28)."""
    return packs_from_scan(scan(filename), genes, variant_correction, no_empty, single_mapped_only)
