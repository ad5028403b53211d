"""
Host-side packing: (reads, variants) of one gene -> arrays the CUDA kernels eat.

This is the exact, integer part of ``AlleleTyping.__init__`` restated on flat
arrays (reference: graphkir/typing_mulit_allele.py:229-269):

* column order = sorted allele names collected from ``variant.allele``
  (:249-255, :283-285);
* ``errorCorrection`` (:302-338): per variant id count positive / negative
  observations over both mates; total < 3 drops the id from both polarities,
  a polarity whose share is < 0.2 is dropped;
* ``removeEmptyReads`` (:274-281);
* the likelihood of a read for an allele only depends on K_r (number of
  observations, duplicates across mates counted) and m[r, a] (how many of them
  disagree with the allele) (:294-300, :363-375), so the kernels get
    - ``mem_words[w, a]``: bit b set  <=>  allele a carries variant 32*w + b
      (word-major so that a warp reading 32 consecutive alleles of one word is
      one coalesced 128-byte access),
    - per read a short list of entries ``(word, pos_bits, neg_bits)``;
      ``m[r, a] = sum_e popc((pos_e & ~mem[word_e, a]) | (neg_e & mem[word_e, a]))``.
      An id observed c times in one polarity is spread over c entries, so
      multiplicities (a variant seen by both mates counts twice, :363-368) are
      exact.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import chain
from typing import Iterable, Sequence

import numpy as np

from .synthetic import LIST_NAMES, ReadCSR, SyntheticGene

MAX_OBS_PER_READ = 255  # m[r, a] is stored as one byte on the device


class CapacityError(ValueError):
    """A gene problem exceeds a capacity of the device path that the reference does not have (more than
    255 variant observations in one read pair, copy number above 8, ...).  Raised per gene: the drivers
    (``kir_typing.Typing.typing``, ``main.cohortAlleleTyping``) catch it, log the sample / gene and the
    limit, and call that gene ``fail`` instead of losing the whole sample or batch."""


@dataclass
class WirePack:
    """The reads of a problem in the wire format of csrc/gk_wire.cu (what ``gk_expand_reads`` turns back
    into observation entries on the device): per mate the window of the variant table, the positives
    inside it as a bitmap, excluded variants and outside positives."""

    hdr: np.ndarray          # uint16 [R]   entries the expansion emits | record length << 8 (0 = raw record)
    stream: np.ndarray       # uint16 [U]   records
    neg_keep: np.ndarray     # uint32 [W]   variants that occur as a negative in any read of the gene
    n_entries: int           # entries the expansion emits in total
    tile_stream: np.ndarray  # int64 [ceil(R / 128) + 1]  first record unit of each 128-read tile
    tile_entry: np.ndarray   # int64 [ceil(R / 128) + 1]  first entry of each tile

    @property
    def nbytes(self) -> int:
        return self.hdr.nbytes + self.stream.nbytes + self.neg_keep.nbytes


@dataclass
class GenePack:
    """Everything the device needs for one (sample, gene) typing problem."""

    gene: str
    allele_names: list[str]          # column id -> name (sorted)
    variant_ids: list[str]           # variant index -> id
    mem_words: np.ndarray            # uint32 [W, A]
    ent_off: np.ndarray              # int32 [R+1]
    ent_word: np.ndarray             # int32 [E]
    ent_pos: np.ndarray              # uint32 [E]
    ent_neg: np.ndarray              # uint32 [E]
    k_obs: np.ndarray                # int32 [R]   observations per read (K_r)
    kept_reads: np.ndarray           # int64 [R]   index of each packed read in the caller's list
    csr: ReadCSR | None = None       # post-correction lists (for homozygosity / EM / oracle)
    var_pos: np.ndarray | None = None      # int64 [V]  variant position
    var_val: list | None = None            # [V] str(variant.val)
    var_is_del: np.ndarray | None = None   # bool [V]  variant.typ == "deletion"
    obs_pos: np.ndarray | None = None      # int64 [V]  positive observations after correction
    obs_neg: np.ndarray | None = None      # int64 [V]  negative observations after correction
    # set on a read shard of a deep problem (shard_reads): the whole problem's read count / sum of K_r
    n_reads_total: int | None = None
    k_total: int | None = None
    wire: "WirePack | None" = None         # compact host->device form of the reads (wire_encode), cached

    @property
    def n_alleles(self) -> int:
        return len(self.allele_names)

    @property
    def n_variants(self) -> int:
        return len(self.variant_ids)

    @property
    def n_words(self) -> int:
        return self.mem_words.shape[0]

    @property
    def n_reads(self) -> int:
        return len(self.k_obs)

    @property
    def n_entries(self) -> int:
        return len(self.ent_word)


# ---------------------------------------------------------------------------
# object lists  <->  CSR
# ---------------------------------------------------------------------------
def csr_from_reads(reads: Sequence, vid_to_idx: dict[str, int]) -> ReadCSR:
    """Flatten ``read.lpv/rpv/lnv/rnv`` (lists of variant ids) into CSR arrays.

    Raises KeyError for an id that is not in the variant table, like the
    reference's ``self.variants[i]`` lookup (typing_mulit_allele.py:364).
    """
    n = len(reads)
    offsets, indices = {}, {}
    for name in LIST_NAMES:
        lens = np.fromiter((len(getattr(r, name)) for r in reads), dtype=np.int64, count=n)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.fromiter(
            (vid_to_idx[str(v)] for v in chain.from_iterable(getattr(r, name) for r in reads)),
            dtype=np.int32, count=int(off[-1]))
        offsets[name] = off
        indices[name] = flat
    return ReadCSR(n, offsets, indices)


def error_correction_masks(csr: ReadCSR, n_variants: int) -> tuple[np.ndarray, np.ndarray]:
    """Variant indices to drop from the positive / negative lists.

    reference: graphkir/typing_mulit_allele.py:304-325 (thresholds 3 and 0.2).
    """
    pos = (np.bincount(csr.indices["lpv"], minlength=n_variants)
           + np.bincount(csr.indices["rpv"], minlength=n_variants)).astype(np.int64)
    neg = (np.bincount(csr.indices["lnv"], minlength=n_variants)
           + np.bincount(csr.indices["rnv"], minlength=n_variants)).astype(np.int64)
    total = pos + neg
    seen = total > 0
    shallow = seen & (total < 3)
    safe_total = np.where(seen, total, 1)
    minor_pos = seen & ~shallow & ((pos / safe_total) < 0.2)
    minor_neg = seen & ~shallow & ((neg / safe_total) < 0.2)
    return shallow | minor_pos, shallow | minor_neg


def filter_csr(csr: ReadCSR, drop_pos: np.ndarray, drop_neg: np.ndarray) -> ReadCSR:
    """Remove dropped variant indices from each list, keeping list order."""
    offsets, indices = {}, {}
    for name in LIST_NAMES:
        drop = drop_pos if name in ("lpv", "rpv") else drop_neg
        idx = csr.indices[name]
        keep = ~drop[idx]
        off = csr.offsets[name]
        row = np.repeat(np.arange(csr.n_reads), off[1:] - off[:-1])
        lens = np.bincount(row[keep], minlength=csr.n_reads)
        new_off = np.zeros(csr.n_reads + 1, dtype=np.int64)
        np.cumsum(lens, out=new_off[1:])
        offsets[name] = new_off
        indices[name] = idx[keep]
    return ReadCSR(csr.n_reads, offsets, indices)


def nonempty_reads(csr: ReadCSR) -> np.ndarray:
    """Boolean mask of reads with at least one id in any list (:274-281)."""
    lens = np.zeros(csr.n_reads, dtype=np.int64)
    for name in LIST_NAMES:
        off = csr.offsets[name]
        lens += off[1:] - off[:-1]
    return lens > 0


# ---------------------------------------------------------------------------
# bit packing
# ---------------------------------------------------------------------------
def pack_membership(member: np.ndarray) -> np.ndarray:
    """bool [V, A] -> uint32 [ceil(V/32), A], bit b of word w = variant 32w+b."""
    n_var, n_allele = member.shape
    n_words = max(1, (n_var + 31) // 32)
    padded = np.zeros((n_words * 32, n_allele), dtype=np.uint32)
    padded[:n_var] = member
    shifts = (np.arange(32, dtype=np.uint32))[None, :, None]
    words = (padded.reshape(n_words, 32, n_allele) << shifts).sum(axis=1, dtype=np.uint64)
    return words.astype(np.uint32)


def pack_entries(csr: ReadCSR) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """CSR lists -> (ent_off, ent_word, ent_pos, ent_neg, k_obs) through the host routine of
    libgk_typing.so (``gk_pack_entries``); ``pack_entries_numpy`` is the array statement of the same
    thing (tests compare them)."""
    import ctypes
    from . import _cabi
    lib = _cabi.load()
    n = csr.n_reads
    offs = [np.ascontiguousarray(csr.offsets[name], dtype=np.int64) for name in LIST_NAMES]
    idxs = [np.ascontiguousarray(csr.indices[name], dtype=np.int32) for name in LIST_NAMES]
    total = int(sum(int(o[-1]) for o in offs))
    if total and (max(int(i.max(initial=0)) for i in idxs) >= 1 << 30 or n >= 1 << 30):
        raise CapacityError("problem too large for the 64-bit packing keys (2^30 reads or variants per gene)")
    pol = np.array([1 if name in ("lpv", "rpv") else 0 for name in LIST_NAMES], dtype=np.int32)
    ent_off = np.zeros(n + 1, dtype=np.int32)
    ent_word = np.zeros(max(total, 1), dtype=np.int32)
    ent_pos = np.zeros(max(total, 1), dtype=np.uint32)
    ent_neg = np.zeros(max(total, 1), dtype=np.uint32)
    k_obs = np.zeros(n, dtype=np.int32)
    lib.gk_pack_entries.restype = ctypes.c_int64
    lib.gk_pack_entries.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 8
    n_ent = lib.gk_pack_entries(
        n, (ctypes.c_void_p * 4)(*[o.ctypes.data for o in offs]), (ctypes.c_void_p * 4)(*[i.ctypes.data for i in idxs]),
        pol.ctypes.data, ent_off.ctypes.data, ent_word.ctypes.data, ent_pos.ctypes.data, ent_neg.ctypes.data,
        k_obs.ctypes.data)
    if n_ent < 0:
        raise ValueError(lib.gk_last_error().decode().split(": ", 1)[-1])
    return ent_off, ent_word[:n_ent].copy(), ent_pos[:n_ent].copy(), ent_neg[:n_ent].copy(), k_obs


def pack_entries_numpy(csr: ReadCSR) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Array statement of ``pack_entries`` (two sorts over all observations of the gene)."""
    n = csr.n_reads
    rows, pols, vids = [], [], []
    for name in LIST_NAMES:
        off = csr.offsets[name]
        rows.append(np.repeat(np.arange(n, dtype=np.int64), off[1:] - off[:-1]))
        idx = csr.indices[name].astype(np.int64)
        vids.append(idx)
        pols.append(np.full(len(idx), 1 if name in ("lpv", "rpv") else 0, dtype=np.int64))
    row = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    pol = np.concatenate(pols) if pols else np.zeros(0, np.int64)
    vid = np.concatenate(vids) if vids else np.zeros(0, np.int64)
    k_obs = np.bincount(row, minlength=n).astype(np.int32)
    if len(row) == 0:
        return (np.zeros(n + 1, np.int32), np.zeros(0, np.int32),
                np.zeros(0, np.uint32), np.zeros(0, np.uint32), k_obs)

    # occurrence rank of identical (read, polarity, variant) observations
    # (single 64-bit sort keys: row < 2^31, variant index < 2^30)
    if len(row) and (int(vid.max()) >= 1 << 30 or n >= 1 << 30):
        raise CapacityError("problem too large for the 64-bit packing keys (2^30 reads or variants per gene)")
    key = (row << 31) | (pol << 30) | vid
    order = np.argsort(key)
    key = key[order]
    row, pol, vid = row[order], pol[order], vid[order]
    new_run = np.ones(len(row), dtype=bool)
    new_run[1:] = key[1:] != key[:-1]
    run_start = np.maximum.accumulate(np.where(new_run, np.arange(len(row)), 0))
    rank = np.arange(len(row)) - run_start
    if len(rank) and int(rank.max()) >= 256:
        raise CapacityError("an observation is repeated more than 255 times in one read pair")

    word = vid >> 5
    bit = (np.uint32(1) << (vid & 31).astype(np.uint32)).astype(np.uint32)
    key = (row << 33) | (rank << 25) | word
    order = np.argsort(key)
    row, rank, word, bit, pol = row[order], rank[order], word[order], bit[order], pol[order]
    new_ent = np.ones(len(row), dtype=bool)
    new_ent[1:] = (row[1:] != row[:-1]) | (rank[1:] != rank[:-1]) | (word[1:] != word[:-1])
    starts = np.flatnonzero(new_ent)
    ent_pos = np.bitwise_or.reduceat(np.where(pol == 1, bit, np.uint32(0)), starts).astype(np.uint32)
    ent_neg = np.bitwise_or.reduceat(np.where(pol == 0, bit, np.uint32(0)), starts).astype(np.uint32)
    ent_word = word[starts].astype(np.int32)
    ent_row = row[starts]
    ent_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(ent_row, minlength=n), out=ent_off[1:])
    return ent_off.astype(np.int32), ent_word, ent_pos, ent_neg, k_obs


def _finish(gene: str, allele_names: list[str], variant_ids: list[str], member: np.ndarray,
            csr: ReadCSR, variant_correction: bool, no_empty: bool) -> tuple[GenePack, tuple]:
    drop = (np.zeros(len(variant_ids), bool), np.zeros(len(variant_ids), bool))
    if variant_correction:
        drop = error_correction_masks(csr, len(variant_ids))
        csr = filter_csr(csr, *drop)
    kept = np.arange(csr.n_reads, dtype=np.int64)
    if no_empty:
        mask = nonempty_reads(csr)
        kept = np.flatnonzero(mask)
        csr = csr.take(kept)
    ent_off, ent_word, ent_pos, ent_neg, k_obs = pack_entries(csr)
    if len(k_obs) and int(k_obs.max()) > MAX_OBS_PER_READ:
        raise CapacityError(
            f"gene {gene}: a read pair carries {int(k_obs.max())} variant observations; the device "
            f"mismatch matrix is one byte per cell (limit {MAX_OBS_PER_READ})")
    pack = GenePack(gene, allele_names, variant_ids, pack_membership(member),
                    ent_off, ent_word, ent_pos, ent_neg, k_obs, kept, csr)
    n_var = len(variant_ids)
    pack.obs_pos = (np.bincount(csr.indices["lpv"], minlength=n_var)
                    + np.bincount(csr.indices["rpv"], minlength=n_var)).astype(np.int64)
    pack.obs_neg = (np.bincount(csr.indices["lnv"], minlength=n_var)
                    + np.bincount(csr.indices["rnv"], minlength=n_var)).astype(np.int64)
    return pack, drop


def _variant_tables(variants: list):
    """(by_id, variant_ids, allele_names, member) of a gene's variant list (:253-255, :283-285)."""
    by_id: dict[str, object] = {str(v.id): v for v in variants}      # later duplicates win (:253)
    variant_ids = list(by_id.keys())
    vid_to_idx = {vid: i for i, vid in enumerate(variant_ids)}
    allele_names = sorted(set(chain.from_iterable(v.allele for v in variants)))   # (:254, :283-285)
    col = {name: i for i, name in enumerate(allele_names)}
    member = np.zeros((len(variant_ids), len(allele_names)), dtype=bool)
    for vid, i in vid_to_idx.items():
        for name in by_id[vid].allele:
            member[i, col[name]] = True
    return by_id, variant_ids, vid_to_idx, allele_names, member


def _annotate(pack: GenePack, by_id: dict, variant_ids: list[str]) -> None:
    pack.var_pos = np.array([by_id[v].pos for v in variant_ids], dtype=np.int64)
    pack.var_val = [str(by_id[v].val) for v in variant_ids]
    pack.var_is_del = np.array([by_id[v].typ == "deletion" for v in variant_ids], dtype=bool)


def pack_gene(reads: Sequence, variants: Iterable, variant_correction: bool = True,
              no_empty: bool = True, mutate_reads: bool = True, gene: str = "") -> GenePack:
    """Object-level entry: the work of ``AlleleTyping.__init__`` up to the likelihood.

    With ``mutate_reads`` the surviving ids are written back into the caller's
    read objects, reproducing the in-place side effect of the reference's
    ``errorCorrection`` (typing_mulit_allele.py:333-338).
    """
    variants = list(variants)
    by_id, variant_ids, vid_to_idx, allele_names, member = _variant_tables(variants)
    if not gene and variants:
        gene = str(variants[0].ref)

    csr = csr_from_reads(reads, vid_to_idx)
    pack, (drop_pos, drop_neg) = _finish(gene, allele_names, variant_ids, member, csr,
                                         variant_correction, no_empty)
    _annotate(pack, by_id, variant_ids)
    if variant_correction and mutate_reads:
        bad_pos = {variant_ids[i] for i in np.flatnonzero(drop_pos)}
        bad_neg = {variant_ids[i] for i in np.flatnonzero(drop_neg)}
        for read in reads:
            read.lpv = [v for v in read.lpv if v not in bad_pos]
            read.rpv = [v for v in read.rpv if v not in bad_pos]
            read.lnv = [v for v in read.lnv if v not in bad_neg]
            read.rnv = [v for v in read.rnv if v not in bad_neg]
    return pack


def pack_gene_csr(variants: Iterable, csr: ReadCSR, variant_correction: bool = True, no_empty: bool = True,
                  gene: str = "") -> GenePack:
    """Array-level entry: like ``pack_gene`` for reads that are already CSR lists of indices into the
    gene's variant table in ``dict.fromkeys(str(v.id) for v in variants)`` order (the fast
    ``.variant.json`` path, :mod:`kir_graph_b200.fastjson`)."""
    variants = list(variants)
    by_id, variant_ids, _, allele_names, member = _variant_tables(variants)
    if not gene and variants:
        gene = str(variants[0].ref)
    pack, _ = _finish(gene, allele_names, variant_ids, member, csr, variant_correction, no_empty)
    _annotate(pack, by_id, variant_ids)
    return pack


def pack_synthetic(gene: SyntheticGene, variant_correction: bool = True,
                   no_empty: bool = True) -> GenePack:
    """Array-level entry for :mod:`kir_graph_b200.synthetic` problems (no Python objects)."""
    has_variant = gene.member.any(axis=0)
    cols = [a for a in np.argsort(np.array(gene.allele_names, dtype=object), kind="stable")
            if has_variant[a]]
    names = [gene.allele_names[a] for a in cols]
    member = gene.member[:, cols]
    pack, _ = _finish(gene.gene, names, list(gene.variant_ids), member, gene.reads,
                      variant_correction, no_empty)
    n_var = gene.n_variants
    pack.var_pos = 25 * np.arange(n_var, dtype=np.int64)
    pack.var_val = ["ACGT"[v % 4] for v in range(n_var)]
    pack.var_is_del = np.zeros(n_var, dtype=bool)
    return pack


def wire_encode(pack: GenePack) -> WirePack:
    """Wire form of the pack's reads (cached on the pack) through ``gk_wire_encode``: 14 B per read
    pair on the cfg3 workload against 36 B of observation entries."""
    if pack.wire is not None:
        return pack.wire
    import ctypes
    from . import _cabi
    lib = _cabi.load()
    csr = pack.csr
    if csr is None:
        raise ValueError("the wire format is encoded from the per-read variant lists (pack.csr)")
    n = pack.n_reads
    offs = [np.ascontiguousarray(csr.offsets[name], dtype=np.int64) for name in LIST_NAMES]
    idxs = [np.ascontiguousarray(csr.indices[name], dtype=np.int32) for name in LIST_NAMES]
    n_words = pack.n_words
    keep_bits = np.zeros(n_words * 32, dtype=bool)
    # the negatives of the gene: from the lists themselves (a read shard carries the whole problem's
    # obs_neg, which is a superset and works as well; holes are relative to the mask that is shipped)
    seen = pack.obs_neg if pack.obs_neg is not None else (
        np.bincount(csr.indices["lnv"], minlength=pack.n_variants) + np.bincount(csr.indices["rnv"], minlength=pack.n_variants))
    keep_bits[: len(seen)] = np.asarray(seen) > 0
    neg_keep = (keep_bits.reshape(n_words, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)[None, :]) \
        .sum(axis=1, dtype=np.uint64).astype(np.uint32)
    ent_off = np.ascontiguousarray(pack.ent_off, dtype=np.int32)
    ent_word = np.ascontiguousarray(pack.ent_word, dtype=np.int32)
    ent_pos = np.ascontiguousarray(pack.ent_pos, dtype=np.uint32)
    ent_neg = np.ascontiguousarray(pack.ent_neg, dtype=np.uint32)
    fn = lib.gk_wire_encode
    fn.restype = ctypes.c_int64
    fn.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 9 + [ctypes.c_int64, ctypes.c_void_p]
    off_p = (ctypes.c_void_p * 4)(*[o.ctypes.data for o in offs])
    idx_p = (ctypes.c_void_p * 4)(*[i.ctypes.data for i in idxs])
    n_ent = ctypes.c_int64(0)
    common = (n, off_p, idx_p, neg_keep.ctypes.data, ent_off.ctypes.data, ent_word.ctypes.data,
              ent_pos.ctypes.data, ent_neg.ctypes.data)
    units = fn(*common, None, None, 0, ctypes.byref(n_ent))
    if units < 0:
        raise ValueError(lib.gk_last_error().decode().split(": ", 1)[-1])
    hdr = np.zeros(n, dtype=np.uint16)
    stream = np.zeros(max(int(units), 1), dtype=np.uint16)
    units2 = fn(*common, hdr.ctypes.data, stream.ctypes.data, int(units), ctypes.byref(n_ent))
    if units2 != units:
        raise ValueError(lib.gk_last_error().decode().split(": ", 1)[-1])
    stream = stream[: int(units)]
    # per 128-read tile: where its records and its entries start
    length = (hdr >> 8).astype(np.int64)
    count = (hdr & 255).astype(np.int64)
    length = np.where(length == 0, 5 * count, length)
    n_tiles = -(-n // 128)
    pad = n_tiles * 128 - n
    tile_stream = np.concatenate([[0], np.cumsum(np.pad(length, (0, pad)).reshape(n_tiles, 128).sum(axis=1))]) \
        if n_tiles else np.zeros(1, np.int64)
    tile_entry = np.concatenate([[0], np.cumsum(np.pad(count, (0, pad)).reshape(n_tiles, 128).sum(axis=1))]) \
        if n_tiles else np.zeros(1, np.int64)
    pack.wire = WirePack(hdr, stream, neg_keep, int(n_ent.value), tile_stream.astype(np.int64),
                         tile_entry.astype(np.int64))
    return pack.wire


def shard_reads(pack: GenePack, rank: int, world: int) -> GenePack:
    """Reads ``[R * rank // world, R * (rank + 1) // world)`` of a packed problem as a problem of its own:
    the read shard of one very deep sample on rank ``rank`` of ``world``.

    Everything that depends on all reads is decided before the split and carried along: the error
    correction (``errorCorrection`` counts observations over every read, typing_mulit_allele.py:302-338)
    has already filtered the lists, ``obs_pos`` / ``obs_neg`` (the homozygosity tallies) stay those of
    the whole problem, and ``n_reads_total`` / ``k_total`` hold the whole problem's read count and
    sum of K_r - the denominators of the read fractions and the constant of the log-likelihood.
    Partial scores, tie counts and column sums of the shards are integers and add up exactly."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    n = pack.n_reads
    lo, hi = n * rank // world, n * (rank + 1) // world
    e0, e1 = int(pack.ent_off[lo]), int(pack.ent_off[hi])
    k_eff = np.where(pack.k_obs == 0, 1, pack.k_obs).astype(np.int64)
    part = GenePack(pack.gene, pack.allele_names, pack.variant_ids, pack.mem_words,
                    (pack.ent_off[lo:hi + 1] - e0).astype(np.int32), pack.ent_word[e0:e1], pack.ent_pos[e0:e1],
                    pack.ent_neg[e0:e1], pack.k_obs[lo:hi], pack.kept_reads[lo:hi],
                    pack.csr.take(np.arange(lo, hi)) if pack.csr is not None else None,
                    pack.var_pos, pack.var_val, pack.var_is_del, pack.obs_pos, pack.obs_neg)
    part.n_reads_total = n if pack.n_reads_total is None else pack.n_reads_total
    part.k_total = int(k_eff.sum()) if pack.k_total is None else pack.k_total
    return part


def site_tallies(pack: GenePack) -> list[dict[str, int]]:
    """Per position: observation counts keyed by allele value (positives) or '*'+value
    (negatives), deletions skipped -- the tally of isHomozygous
    (reference: typing_mulit_allele.py:820-832) computed from the packed lists."""
    pos_cnt, neg_cnt = pack.obs_pos, pack.obs_neg
    sites: dict[int, dict[str, int]] = {}
    for v in np.flatnonzero(((pos_cnt + neg_cnt) > 0) & ~pack.var_is_del):
        site = sites.setdefault(int(pack.var_pos[v]), {})
        if pos_cnt[v]:
            key = pack.var_val[v]
            site[key] = site.get(key, 0) + int(pos_cnt[v])
        if neg_cnt[v]:
            key = "*" + pack.var_val[v]
            site[key] = site.get(key, 0) + int(neg_cnt[v])
    return list(sites.values())
