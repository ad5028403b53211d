"""
Packed binary sidecar of a ``.variant.json``.

The JSON written by the extraction step carries the raw SAM lines (about 1 KB per read pair);
loading it and packing the reads dominates the wall clock once typing itself takes
milliseconds.  ``save_packs`` stores the packed per-gene arrays (``packing.GenePack``) next to
it as ``{prefix}.gkpack.npz``; ``load_packs`` restores them without touching the JSON.  The JSON
stays the public format (reference: graphkir/hisat2.py:847-866); the sidecar is a cache keyed by
the packing options.
"""
from __future__ import annotations

import json
import os
import zipfile

import numpy as np

from .hisat2 import loadReadsAndVariantsData, removeMultipleMapped
from .packing import GenePack, WirePack, pack_gene, wire_encode
from .synthetic import LIST_NAMES, ReadCSR

_ARRAYS = ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "kept_reads", "var_pos",
           "var_is_del", "obs_pos", "obs_neg")
_WIRE_ARRAYS = ("hdr", "stream", "neg_keep", "tile_stream", "tile_entry")


def _pack_arrays(p: GenePack) -> dict[str, np.ndarray]:
    """Every array of a pack that the sidecar stores, by name."""
    out = {name: np.asarray(getattr(p, name)) for name in _ARRAYS}
    for name in LIST_NAMES:
        out[f"csr_off_{name}"] = np.asarray(p.csr.offsets[name])
        out[f"csr_idx_{name}"] = np.asarray(p.csr.indices[name])
    wire = wire_encode(p)
    for name in _WIRE_ARRAYS:
        out[f"wire_{name}"] = np.asarray(getattr(wire, name))
    return out


def save_packs(path: str, packs: dict[str, GenePack], meta: dict | None = None, compress: bool = False) -> None:
    """Write ``{gene: GenePack}`` to one ``.npz``, the wire form of the reads (``packing.wire_encode``: what
    the cohort path copies to the device) included, so that a sample typed from its sidecar is read at file
    speed.  Layout (format 3): one member per array NAME holding the arrays of all genes back to back, and
    one JSON index with the names, shapes and scalars per gene - a few dozen zip members instead of one per
    gene and array (opening a member costs more than reading it), uncompressed by default (≈ 200 B per read
    pair against ≈ 1.2 KB of ``.json``).  Loading 50 k pairs: 20 ms against 230 ms for load + wire encoding
    from the compressed per-gene members of format 2 (build container's CPU)."""
    if any(not isinstance(p, GenePack) for p in packs.values()):
        raise ValueError("a gene exceeds a capacity of the device path (packing.CapacityError): no sidecar is "
                         "written, the .json stays the source")
    index = {"genes": list(packs), "meta": meta or {}, "per_gene": {}}
    columns: dict[str, list[np.ndarray]] = {}
    for g, p in packs.items():
        arrays = _pack_arrays(p)
        index["per_gene"][g] = {"gene": p.gene, "alleles": p.allele_names, "variants": p.variant_ids,
                                "var_val": p.var_val, "wire_n_entries": int(p.wire.n_entries),
                                "shapes": {name: list(a.shape) for name, a in arrays.items()}}
        for name, a in arrays.items():
            columns.setdefault(name, []).append(a.ravel())
    out = {"__index__": np.array(json.dumps(index))}
    for name, parts in columns.items():
        dtypes = {a.dtype for a in parts}
        if len(dtypes) != 1:
            raise ValueError(f"array {name} has different dtypes over the genes: {sorted(map(str, dtypes))}")
        out[name] = np.concatenate(parts) if parts else np.zeros(0)
    (np.savez_compressed if compress else np.savez)(path, **out)


def load_packs(path: str) -> tuple[dict[str, GenePack], dict]:
    data = np.load(path, allow_pickle=False)
    if "__index__" not in data.files:
        raise ValueError("sidecar of an older format")           # the caller falls back to the .json
    index = json.loads(str(data["__index__"]))
    columns = {name: data[name] for name in data.files if name != "__index__"}
    cursor = dict.fromkeys(columns, 0)
    packs = {}
    for g in index["genes"]:
        info = index["per_gene"][g]

        def take(name):
            shape = info["shapes"][name]
            size = int(np.prod(shape)) if shape else 1
            at = cursor[name]
            cursor[name] = at + size
            return columns[name][at:at + size].reshape(shape)

        arrays = {name: take(name) for name in info["shapes"]}
        csr = ReadCSR(len(arrays["k_obs"]),
                      {n: arrays[f"csr_off_{n}"] for n in LIST_NAMES},
                      {n: arrays[f"csr_idx_{n}"] for n in LIST_NAMES})
        p = GenePack(info["gene"], info["alleles"], info["variants"],
                     *(arrays[a] for a in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs",
                                           "kept_reads")), csr=csr)
        p.var_pos, p.var_is_del = arrays["var_pos"], arrays["var_is_del"]
        p.obs_pos, p.obs_neg = arrays["obs_pos"], arrays["obs_neg"]
        p.var_val = info["var_val"]
        p.wire = WirePack(arrays["wire_hdr"], arrays["wire_stream"], arrays["wire_neg_keep"],
                          int(info["wire_n_entries"]), arrays["wire_tile_stream"], arrays["wire_tile_entry"])
        packs[g] = p
    return packs, index["meta"]


SIDECAR_SUFFIX = ".gkpack.npz"
PACK_FORMAT = 3        # (3: wire form of the reads stored, uncompressed)  bumped whenever the packed arrays or their meaning change: older sidecars are rebuilt


def sidecar_path(filename_variant_json: str) -> str:
    """``{prefix}.json`` (or ``{prefix}``) -> ``{prefix}.gkpack.npz``."""
    prefix = filename_variant_json[:-5] if filename_variant_json.endswith(".json") else filename_variant_json
    return prefix + SIDECAR_SUFFIX


def sidecar_meta(filename_variant_json: str, variant_correction: bool = True, multiple: bool = False) -> dict:
    """What a sidecar records about how and from what it was packed.  ``json_size`` and
    ``json_mtime_ns`` tie it to the ``.json`` it sits next to: a ``.json`` that was re-extracted or
    edited - even to the same byte size - invalidates it (a copy that does not preserve the
    modification time merely rebuilds the sidecar); ``format`` is the packing-format version."""
    path = filename_variant_json if filename_variant_json.endswith(".json") else filename_variant_json + ".json"
    st = os.stat(path) if os.path.exists(path) else None
    return {"variant_correction": bool(variant_correction), "multiple": bool(multiple),
            "json_size": st.st_size if st else -1, "json_mtime_ns": st.st_mtime_ns if st else -1,
            "format": PACK_FORMAT}


def load_sample_packs(filename_variant_json: str, variant_correction: bool = True, multiple: bool = False
                      ) -> dict[str, GenePack]:
    """Packed genes of one sample: from ``{prefix}.gkpack.npz`` when it is there and was packed with
    the same options from a ``.json`` of the same size, else from the ``.json`` through the C++
    scanner (:mod:`kir_graph_b200.fastjson`).  Module-level and cheap to import: it is what the worker
    processes of ``main.cohortAlleleTyping`` run."""
    side = sidecar_path(filename_variant_json)
    if os.path.exists(side):
        try:
            packs, meta = load_packs(side)
            if meta == sidecar_meta(filename_variant_json, variant_correction, multiple):
                return packs
        except (OSError, ValueError, KeyError, EOFError, zipfile.BadZipFile):
            pass                                  # unreadable / truncated sidecar: fall through to the .json
    from . import fastjson
    return fastjson.load_packs(filename_variant_json, variant_correction=variant_correction,
                               single_mapped_only=not multiple)


def pack_variant_json(filename_variant_json: str, variant_correction: bool = True,
                      multiple: bool = False) -> dict[str, GenePack]:
    """``{prefix}.variant.json`` -> packed problems per gene (the work of
    ``TypingWithPosNegAllele.__init__`` + ``AlleleTyping.__init__`` up to the likelihood)."""
    from .kir_typing import groupReads, groupVariants
    data = loadReadsAndVariantsData(filename_variant_json)
    if not multiple:
        data = removeMultipleMapped(data)
    reads, variants = groupReads(data["reads"]), groupVariants(data["variants"])
    return {gene: pack_gene(reads.get(gene, []), vs, variant_correction=variant_correction, gene=gene)
            for gene, vs in variants.items()}
