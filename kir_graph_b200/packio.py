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

import numpy as np

from .hisat2 import loadReadsAndVariantsData, removeMultipleMapped
from .packing import GenePack, pack_gene
from .synthetic import LIST_NAMES, ReadCSR

_ARRAYS = ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "kept_reads", "var_pos",
           "var_is_del", "obs_pos", "obs_neg")


def save_packs(path: str, packs: dict[str, GenePack], meta: dict | None = None) -> None:
    """Write ``{gene: GenePack}`` to one compressed ``.npz``."""
    if any(not isinstance(p, GenePack) for p in packs.values()):
        raise ValueError("a gene exceeds a capacity of the device path (packing.CapacityError): no sidecar is "
                         "written, the .json stays the source")
    out = {"__genes__": np.array(json.dumps(list(packs)))}
    out["__meta__"] = np.array(json.dumps(meta or {}))
    for g, p in packs.items():
        out[f"{g}/names"] = np.array(json.dumps({"alleles": p.allele_names, "variants": p.variant_ids,
                                                  "var_val": p.var_val, "gene": p.gene}))
        for name in _ARRAYS:
            out[f"{g}/{name}"] = getattr(p, name)
        for name in LIST_NAMES:
            out[f"{g}/csr_off_{name}"] = p.csr.offsets[name]
            out[f"{g}/csr_idx_{name}"] = p.csr.indices[name]
    np.savez_compressed(path, **out)


def load_packs(path: str) -> tuple[dict[str, GenePack], dict]:
    data = np.load(path, allow_pickle=False)
    packs = {}
    for g in json.loads(str(data["__genes__"])):
        names = json.loads(str(data[f"{g}/names"]))
        csr = ReadCSR(len(data[f"{g}/k_obs"]),
                      {n: data[f"{g}/csr_off_{n}"] for n in LIST_NAMES},
                      {n: data[f"{g}/csr_idx_{n}"] for n in LIST_NAMES})
        p = GenePack(names["gene"], names["alleles"], names["variants"],
                     *(data[f"{g}/{a}"] for a in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg",
                                                  "k_obs", "kept_reads")), csr=csr)
        p.var_pos, p.var_is_del = data[f"{g}/var_pos"], data[f"{g}/var_is_del"]
        p.obs_pos, p.obs_neg = data[f"{g}/obs_pos"], data[f"{g}/obs_neg"]
        p.var_val = names["var_val"]
        packs[g] = p
    return packs, json.loads(str(data["__meta__"]))


SIDECAR_SUFFIX = ".gkpack.npz"
PACK_FORMAT = 2        # bumped whenever the packed arrays or their meaning change: older sidecars are rebuilt


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
        except (OSError, ValueError, KeyError):
            pass                                  # unreadable sidecar: fall through to the .json
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
