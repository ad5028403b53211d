"""
Batched typing engine: drives the CUDA kernels of libgk_typing.so.

``MatrixBatch``   uploads the packed problems of any number of (sample, gene)
                  pairs into pooled device buffers and runs the likelihood
                  kernel once (kernel (a)).
``SearchGroup``   advances any number of greedy top-N searches over those
                  matrices in lock step (kernels (b) and (c)): score ->
                  dedup/select -> rescore -> rank -> next P.  ``step()`` does one
                  copy-number step and reads its result back; ``run_pipeline()``
                  enqueues every step up front (grids sized from upper bounds,
                  work-item tables cached per batch) and reads back once.
``CudaBackend``   device memory, asynchronous staging of small host arrays,
                  pooled read-back buffers and the kernel launches; everything
                  it enqueues can be recorded into a CUDA graph.

Batching over genes and samples is what fills a 148-SM GPU: one WGS sample is
17 problems of a few thousand reads each.

The engine computes in exact integers (mismatch counts); ``typing_mulit_allele``
converts to the reference's float64 log10 units at the API boundary.
There is no CPU implementation of any kernel in this package.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _cabi
from ._cabi import (COUNT_ITEM_DTYPE, EXPAND_ITEM_DTYPE, GK_KB, GK_LIK_READS, GK_MAX_CN, GK_RT, LIK_ITEM_DTYPE,
                    MATRIX_DTYPE, P_ITEM_DTYPE, SCORE_ITEM_DTYPE, SEARCH_DTYPE, STEP_INFO_DTYPE)
from . import packing as _packing
from .packing import GenePack

import os as _os

SCORE_READ_CHUNK = 8192     # reads per scoring work item: 8192 * 255 < 2^24 keeps float32 sums exact
COUNT_READ_CHUNK = int(_os.environ.get("GK_COUNT_CHUNK", 16384))    # reads per rescoring work item
P_READ_CHUNK = int(_os.environ.get("GK_P_CHUNK", 2048))             # reads per P-writing work item (multiple of 128)
ALIVE_SLACK = 212           # alive sets beyond top_n the rescoring grids are sized for without a read-back
MAX_TOP_N = 2048

# reads travel host -> device in the wire format of csrc/gk_wire.cu (GK_WIRE=0: as 16-byte entries)
WIRE_DEFAULT = _os.environ.get("GK_WIRE", "1") != "0"
# packed 16-bit integer scoring path (False: FP32 sum of absolute differences); GK_PACKED=0/1 overrides
PACKED_DEFAULT = _os.environ.get("GK_PACKED", "1") != "0"
SCORE_SLOTS = int(_os.environ.get("GK_SCORE_SLOTS", 444))            # resident scoring CTAs: 148 SMs x 3
SCORE_ITEM_OVERHEAD = int(_os.environ.get("GK_SCORE_ITEM_OVERHEAD", 48))   # fixed cost of a work item, in reads
SCORE_CHUNK_CANDIDATES = (8192, 6144, 4096, 3072, 2048, 1536, 1024)
# Work-item tables from the C++ host routines of csrc/gk_plan.cu (row for row what the NumPy statements
# below produce; GK_PLAN=numpy keeps the NumPy path, which also serves the cases the routines do not cover)
PLAN_NATIVE = _os.environ.get("GK_PLAN", "native") != "numpy"


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def capacity_violation(n_alleles: int, cn: int, top_n: int) -> str | None:
    """Why a problem of ``n_alleles`` alleles typed at copy number ``cn`` with ``top_n`` kept sets does not
    fit the search kernels, or None.  (The per-read limit of 255 observations is checked when the reads
    are packed, ``packing.CapacityError``.)  Callers check this per gene up front so that one oversized
    gene is called ``fail`` with a warning instead of aborting a batch."""
    if cn > GK_MAX_CN:
        return f"copy number {cn} above {GK_MAX_CN} (GK_MAX_CN of the search kernels)"
    if not 1 <= top_n <= MAX_TOP_N:
        return f"top_n {top_n} outside 1..{MAX_TOP_N}"
    if n_alleles >= 65535:
        return f"{n_alleles} alleles exceed the 16-bit allele ids of the selection kernels"
    return None


# ---------------------------------------------------------------------------
# backend: device memory + kernel launches (torch is plumbing only)
# ---------------------------------------------------------------------------
class DeviceArena:
    """One device allocation handed out in aligned slices (a bump allocator).  A cohort stream that types a
    new batch every pass would otherwise pay a ``cudaMalloc`` for most of its ~40 buffers per sub-batch
    (sizes differ from batch to batch, so the caching allocator rarely has a fitting block: 0.2 s of a
    0.35 s pass): with ``CudaBackend.device_arena`` set, every buffer of a pass comes from the arena and
    ``reset()`` recycles all of them at once when the pass has been read back.  Requests that do not fit
    fall back to the regular allocator."""

    def __init__(self, backend: "CudaBackend", nbytes: int):
        self.buf = backend.torch.empty(int(nbytes), dtype=backend.torch.uint8, device=backend.device)
        self.nbytes = int(nbytes)
        self.pos = 0
        self.misses = 0

    def take(self, nbytes: int):
        pos = (self.pos + 511) & ~511
        if pos + nbytes > self.nbytes:
            self.misses += 1
            return None
        self.pos = pos + nbytes
        return self.buf[pos:pos + nbytes]

    def reset(self) -> None:
        self.pos = 0


class CudaBackend:
    """Device arrays are torch tensors; kernels are the C-ABI launchers."""

    def __init__(self, device: str | int | None = None):
        import torch
        if not torch.cuda.is_available():
            raise _cabi.GkError("kir_graph_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device(device if isinstance(device, str) else f"cuda:{device}")
        _cabi.load()
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.timing: dict[str, list] | None = None   # name -> [(start_event, end_event, work)]
        self._arenas: dict[int, list] = {}            # stream -> [pinned uint8 tensor, bump position]
        self.capturing = False                        # True while a CUDA graph is being recorded
        self.device_arena: DeviceArena | None = None  # when set, device buffers are slices of it
        self._stream_pool: list = []

    def streams(self, n: int, first: int = 0) -> list:
        """``n`` persistent CUDA streams (numbers ``first .. first + n - 1`` of the backend's pool): typers
        that are built per pass reuse them - and with them the page-locked staging arena of each stream -
        instead of creating streams (and pinning 16 MB) every time."""
        while len(self._stream_pool) < first + n:
            self._stream_pool.append(self.torch.cuda.Stream(device=self.device))
        return self._stream_pool[first:first + n]

    # Small host arrays (work-item tables, index lists) go through a page-locked staging arena so
    # that their copies are asynchronous: a copy from pageable memory synchronises the stream first,
    # which drains the GPU ~25 times per typing pass.  One arena per stream; it is rewound only after
    # that stream has been synchronised.
    ARENA_BYTES = 16 << 20
    ARENA_MAX_ITEM = 4 << 20

    def _stage(self, raw: np.ndarray):
        """Copy a flat uint8 host array into the arena of the current stream; pinned torch view."""
        stream = self.torch.cuda.current_stream(self.device)
        arena = self._arenas.get(stream.cuda_stream)
        if arena is None:
            arena = [self.torch.empty(self.ARENA_BYTES, dtype=self.torch.uint8).pin_memory(), 0]
            self._arenas[stream.cuda_stream] = arena
        pos = (arena[1] + 255) & ~255
        if pos + raw.size > self.ARENA_BYTES:
            stream.synchronize()                      # every copy staged so far has executed
            pos = 0
        view = arena[0][pos:pos + raw.size]
        view.numpy()[:] = raw
        arena[1] = pos + raw.size
        return view

    # --- memory ---------------------------------------------------------------
    _NP2T = {"uint8": "uint8", "int32": "int32", "uint32": "int32", "float32": "float32",
             "int64": "int64", "uint64": "int64", "float64": "float64", "uint16": "int16", "int16": "int16"}

    def _tdtype(self, dtype):
        return getattr(self.torch, self._NP2T[np.dtype(dtype).name])

    def zeros(self, n: int, dtype):
        if self.device_arena is not None:
            return self.empty(n, dtype).zero_()
        return self.torch.zeros(max(int(n), 1), dtype=self._tdtype(dtype), device=self.device)

    def empty(self, n: int, dtype):
        n, td = max(int(n), 1), self._tdtype(dtype)
        if self.device_arena is not None:
            raw = self.device_arena.take(n * np.dtype(self._NP2T[np.dtype(dtype).name]).itemsize)
            if raw is not None:
                return raw.view(td)
        return self.torch.empty(n, dtype=td, device=self.device)

    def upload(self, array: np.ndarray):
        if self.capturing:
            raise _cabi.GkError("host->device copy inside a CUDA-graph capture (launch plan not cached)")
        array = np.ascontiguousarray(array)
        if array.dtype.fields is not None or array.dtype.name not in self._NP2T:
            array = array.view(np.uint8)
        elif array.dtype.name == "uint32":
            array = array.view(np.int32)
        elif array.dtype.name == "uint64":
            array = array.view(np.int64)
        elif array.dtype.name == "uint16":
            array = array.view(np.int16)
        if array.size == 0:
            return self.zeros(1, array.dtype)
        self.h2d_bytes += array.nbytes
        flat = array.reshape(-1)
        if flat.nbytes <= self.ARENA_MAX_ITEM:
            staged = self._stage(flat.view(np.uint8)).view(self._tdtype(flat.dtype))
            return self.empty(flat.size, flat.dtype).copy_(staged, non_blocking=True)
        return self.empty(flat.size, flat.dtype).copy_(self.torch.from_numpy(flat), non_blocking=True)

    def copy_into(self, tensor, array: np.ndarray) -> None:
        """Asynchronous host->device copy into an existing device array of the same byte size."""
        flat = np.ascontiguousarray(array).reshape(-1)
        if flat.size == 0:
            return
        src = self.torch.from_numpy(flat.view(np.uint8))
        self.h2d_bytes += flat.nbytes
        tensor.view(self.torch.uint8)[: flat.nbytes].copy_(src, non_blocking=True)

    def download(self, tensor, dtype=None, count: int | None = None) -> np.ndarray:
        if count is not None:
            tensor = tensor[:count]
        out = tensor.cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out.view(dtype) if dtype is not None else out

    def download_async(self, tensors, sizes=None):
        """Start ONE device->host copy of several int32 device arrays (concatenated on the device
        into a page-locked buffer of the current stream); ``download_wait`` returns the arrays.
        With ``sizes`` the first argument is the already concatenated array."""
        torch = self.torch
        if sizes is not None:
            flat = tensors
        else:
            flat = torch.cat([t.reshape(-1) for t in tensors]) if len(tensors) != 1 else tensors[0].reshape(-1)
            sizes = [t.numel() for t in tensors]
        assert flat.dtype == torch.int32
        stream = torch.cuda.current_stream(self.device)
        # page-locked result buffers are pooled: one is busy from here until download_wait returns,
        # so several reads may be in flight (even on one stream) without sharing a buffer
        self._d2h_free = getattr(self, "_d2h_free", [])
        buf = None
        for i, cand in enumerate(self._d2h_free):
            if cand.numel() >= flat.numel():
                buf = self._d2h_free.pop(i)
                break
        if buf is None:
            buf = torch.empty(max(flat.numel(), 1 << 18), dtype=torch.int32).pin_memory()
        view = buf[: flat.numel()]
        view.copy_(flat, non_blocking=True)
        event = torch.cuda.Event()
        event.record(stream)
        return view, event, sizes, buf

    def download_wait(self, handle) -> list[np.ndarray]:
        view, event, sizes, buf = handle
        event.synchronize()
        host = view.numpy().copy()
        self._d2h_free.append(buf)
        self.d2h_bytes += host.nbytes
        out, pos = [], 0
        for n in sizes:
            out.append(host[pos:pos + n])
            pos += n
        return out

    def zero_(self, tensor) -> None:
        tensor.zero_()

    def sync(self) -> None:
        self.torch.cuda.synchronize(self.device)

    def gather_rows(self, tensor, row_len: int, rows: np.ndarray, cols: int | None = None) -> np.ndarray:
        """Download tensor.view(-1, row_len)[rows, :cols] (device-side gather, one copy)."""
        if len(rows) == 0:
            return np.zeros((0, cols or row_len), dtype=np.dtype(str(tensor.dtype).split(".")[-1]))
        idx = self.upload(np.asarray(rows, dtype=np.int64))
        self.h2d_bytes -= idx.numel() * 8              # index lists are not payload
        view = tensor.view(-1, row_len).index_select(0, idx)
        if cols is not None:
            view = view[:, :cols].contiguous()
        out = view.cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out

    def snapshot(self, tensor):
        """Device-side copy (asynchronous on the current stream)."""
        return tensor.clone()

    def gather_best_device(self, ids, score, info, rows: np.ndarray, top_n: int):
        """(ids row, score) of rank ``info.best_rank`` for each search in ``rows``; result stays on
        the device (no synchronisation)."""
        if hasattr(rows, "data_ptr"):
            idx = rows
        else:
            idx = self.upload(np.asarray(rows, dtype=np.int64))
            self.h2d_bytes -= idx.numel() * 8
        best = info.view(-1, STEP_INFO_DTYPE.itemsize // 4)[:, 6].index_select(0, idx).to(self.torch.int64)
        flat = idx * top_n + best
        return ids.view(-1, GK_MAX_CN).index_select(0, flat), score.index_select(0, flat)

    def pin(self, array: np.ndarray) -> np.ndarray:
        """Copy a host array into page-locked memory (for timed host->device copies)."""
        flat = np.ascontiguousarray(array).reshape(-1)
        raw = flat.view(np.uint8)
        t = self.torch.empty(max(raw.size, 1), dtype=self.torch.uint8).pin_memory()
        out = t.numpy()[: raw.size]
        out[:] = raw
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(t)
        return out.view(flat.dtype)

    # --- kernels ----------------------------------------------------------------
    def launch(self, name: str, *args, work: float = 0.0) -> None:
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        conv = [(_cabi.ptr(a.data_ptr()) if hasattr(a, "data_ptr") else a) for a in args]
        if self.timing is not None:
            start = self.torch.cuda.Event(enable_timing=True)
            end = self.torch.cuda.Event(enable_timing=True)
            start.record()
            _cabi.call(name, *conv, _cabi.ptr(stream))
            end.record()
            self.timing.setdefault(name, []).append((start, end, work))
        else:
            _cabi.call(name, *conv, _cabi.ptr(stream))
        self.launches += 1


_default_backends: dict = {}


def default_backend(device: int | None = None) -> CudaBackend:
    """The process-wide backend of a device (the current one by default).  Every class that is not handed
    a backend shares it - one page-locked staging arena and one set of read-back buffers per stream
    instead of a new 16 MB pinned allocation per AlleleTyping / MatrixBatch / EM call."""
    import torch
    if not torch.cuda.is_available():
        raise _cabi.GkError("kir_graph_b200 needs a CUDA device (no CPU fallback)")
    key = torch.cuda.current_device() if device is None else int(device)
    be = _default_backends.get(key)
    if be is None:
        be = _default_backends[key] = CudaBackend(key)
    return be


# ---------------------------------------------------------------------------
# matrices
# ---------------------------------------------------------------------------
class HostBatch:
    """Packed problems of a batch concatenated into host pools (the "packed host arrays"
    that the end-to-end path copies to the device).

    ``wire=True`` (default when every pack carries its per-read variant lists): the reads travel in
    the wire format of csrc/gk_wire.cu - per mate the window of the variant table, the positives as a
    bitmap, excluded variants - 14 B per read pair on the cfg3 workload, and ``gk_expand_reads``
    rebuilds the observation entries on the device.  ``wire=False``: the 16-byte entries themselves
    are built here and copied (what a caller that only has packed entries can do)."""

    def __init__(self, packs: list[GenePack], wire: bool | None = None):
        self.packs = packs
        self._m_max = None
        self._lik_items = None
        if wire is None:
            wire = WIRE_DEFAULT and all(p.csr is not None for p in packs)
        self.wire = bool(wire)
        n = len(packs)
        table = np.zeros(n, dtype=MATRIX_DTYPE)
        mem_off = entoff_off = ent_base = L_off = LT_off = col_off = 0
        wires = [_packing.wire_encode(p) for p in packs] if self.wire else None
        ent_bases = []
        for i, p in enumerate(packs):
            if p.n_reads and int(p.k_obs.astype(np.int64).sum()) >= 2 ** 32 - 1:
                raise ValueError("sum of observations per problem must stay below 2^32 (32-bit score atomics)")
            a_tile = 32
            n_ablk = max(1, -(-p.n_alleles // a_tile))
            r_pad = max(128, _round_up(p.n_reads, 128))
            n_total = getattr(p, "n_reads_total", None)        # set on a read shard (packing.shard_reads)
            table[i] = (mem_off, entoff_off, L_off, LT_off, col_off, p.n_reads, p.n_alleles, p.n_words,
                        r_pad, a_tile, n_ablk, p.n_reads if n_total is None else n_total,
                        min(int(p.k_obs.max(initial=0)), 255))
            ent_bases.append(ent_base)
            mem_off += p.n_words * n_ablk * a_tile
            entoff_off += p.n_reads + 1
            ent_base += wires[i].n_entries if self.wire else p.n_entries
            L_off += n_ablk * r_pad * a_tile
            LT_off += max(p.n_alleles, 1) * r_pad
            col_off += max(p.n_alleles, 1)
        if ent_base >= 2 ** 31:
            raise ValueError("entry pool exceeds 2^31 entries; split the batch")
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt))
        self.table = table
        self.n_entries = int(ent_base)
        self.n_offsets = int(entoff_off)
        # rows of the membership table are padded to whole 32-allele blocks: a lane of the likelihood
        # kernel loads the words of 1, 2 or 4 consecutive alleles with one aligned 32- / 64- / 128-bit load
        self.mem = np.zeros(int(mem_off), dtype=np.uint32)
        for i, p in enumerate(packs):
            stride = int(table["n_ablk"][i]) * int(table["a_tile"][i])
            view = self.mem[int(table["mem_off"][i]): int(table["mem_off"][i]) + p.n_words * stride]
            view.reshape(p.n_words, stride)[:, : p.n_alleles] = p.mem_words
        if self.wire:
            self.hdr = cat([w.hdr for w in wires], np.uint16)
            self.stream = cat([w.stream for w in wires], np.uint16)
            self.keep = cat([w.neg_keep for w in wires], np.uint32)
            if len(self.stream) >= 2 ** 32:
                raise ValueError("wire stream exceeds 2^32 units; split the batch")
            hdr_base = _excl_cumsum([p.n_reads for p in packs])
            keep_off = _excl_cumsum([len(w.neg_keep) for w in wires])
            stream_base = _excl_cumsum([len(w.stream) for w in wires])
            tiles = np.array([len(w.tile_stream) - 1 if p.n_reads else 1 for p, w in zip(packs, wires)], dtype=np.int64)
            items = np.zeros(int(tiles.sum()), dtype=EXPAND_ITEM_DTYPE)
            mat = np.repeat(np.arange(n, dtype=np.int64), tiles)
            local = np.arange(len(items), dtype=np.int64) - np.repeat(_excl_cumsum(tiles), tiles)
            items["matrix"] = mat
            items["r0"] = local * GK_LIK_READS
            items["hdr_base"] = hdr_base[mat]
            items["keep_off"] = keep_off[mat]
            items["stream_off"] = stream_base[mat] + cat([w.tile_stream[:max(len(w.tile_stream) - 1, 1)] for w in wires], np.int64)
            items["ent_off"] = np.asarray(ent_bases, dtype=np.int64)[mat] + cat(
                [w.tile_entry[:max(len(w.tile_entry) - 1, 1)] for w in wires], np.int64)
            self.xitems = items
        else:
            self.entoff = cat([p.ent_off.astype(np.int64) + b for p, b in zip(packs, ent_bases)], np.int32)
            # 16-byte entries {word * row stride of mem in bytes, pos, neg, 1 << 8 (r & 3)} (include/gk_typing.h)
            ent = np.zeros((self.n_entries, 4), dtype=np.uint32)
            for i, p in enumerate(packs):
                if not p.n_entries:
                    continue
                stride_bytes = int(table["n_ablk"][i]) * int(table["a_tile"][i]) * 4
                if int(p.ent_word.max(initial=0)) * stride_bytes >= 2 ** 32:
                    raise ValueError("membership table of one gene exceeds 4 GB")
                row = np.repeat(np.arange(p.n_reads, dtype=np.int64), np.diff(p.ent_off.astype(np.int64)))
                blk = ent[ent_bases[i]: ent_bases[i] + p.n_entries]
                blk[:, 0] = p.ent_word.astype(np.int64) * stride_bytes
                blk[:, 1] = p.ent_pos
                blk[:, 2] = p.ent_neg
                blk[:, 3] = np.uint32(1) << (8 * (row & 3)).astype(np.uint32)
            self.ent = ent.reshape(-1)
        self.k_max = int(max((int(p.k_obs.max()) for p in packs if p.n_reads), default=0))
        # sum of K_r over the reads of the whole problem (a read shard carries the global sum)
        self.k_total = np.array([int(np.where(p.k_obs == 0, 1, p.k_obs).astype(np.int64).sum())
                                 if getattr(p, "k_total", None) is None else int(p.k_total) for p in packs],
                                dtype=np.int64)
        self.L_size, self.LT_size, self.col_size = int(L_off), int(LT_off), int(col_off)

    @property
    def m_max(self) -> int:
        """Upper bound of every mismatch count of the batch: m[r, a] <= K_r <= 255."""
        if self._m_max is None:
            m = max([int(p.k_obs.max(initial=1)) for p in self.packs], default=1)
            self._m_max = min(max(m, 1), 255)
        return self._m_max

    @property
    def input_names(self) -> tuple[str, ...]:
        """Host pools that are copied to the device for every pass (besides the matrix table)."""
        return ("mem", "hdr", "stream", "keep", "xitems") if self.wire else ("mem", "entoff", "ent")

    def lik_items(self) -> np.ndarray:
        """Work items of the likelihood build (128-read tiles x groups of 4 allele blocks per problem): a
        function of the batch layout alone, built once per host batch - ``pin`` does it, so a typer built
        per pass from a prepared host batch finds it there."""
        if self._lik_items is None:
            t = self.table
            tiles = (t["r_pad"] // GK_LIK_READS).astype(np.int64)
            groups = -(-t["n_ablk"].astype(np.int64) // 4)          # 4 allele blocks per CTA
            per = tiles * groups
            total = int(per.sum())
            items = np.zeros(total, dtype=LIK_ITEM_DTYPE)
            mat = np.repeat(np.arange(len(t), dtype=np.int64), per)
            local = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(per) - per, per)
            items["matrix"] = mat
            items["a_blk"] = (local // tiles[mat]) * 4
            items["r0"] = (local % tiles[mat]) * GK_LIK_READS
            self._lik_items = items
            self.lik_matrix = mat                   # problem of every item (contiguous: per-pass flags are one gather)
        return self._lik_items

    def pin(self, backend) -> "HostBatch":
        for name in self.input_names:
            arr = getattr(self, name)
            raw = backend.pin(arr.view(np.uint8) if arr.dtype.fields is not None else arr)
            setattr(self, name, raw.view(arr.dtype) if arr.dtype.fields is not None else raw)
        self.lik_items()
        return self

    @property
    def nbytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in self.input_names) + self.table.nbytes


class MatrixBatch:
    """Likelihood data of a batch of gene problems, resident on one GPU."""

    def __init__(self, packs, backend=None, run: bool = True, packed: bool = PACKED_DEFAULT, reduce=None):
        """``packed``: score on 16-bit integer lanes (VIMNMX.U16x2 on the ALU pipe + IMAD on the FMA pipe,
        one issue slot per cell; ``L`` holds the pair (m, m), ``P`` is uint16) instead of FP32 (two FADDs
        per cell).  Both are exact for every supported input (counts <= 255).
        ``reduce(tensor)``: the packs are read shards of problems whose other reads live on other
        ranks (``packing.shard_reads``); the callable sums a device array over the ranks in place and
        is applied to the column sums after the likelihood build (integer sums: order independent)."""
        self.be = backend if backend is not None else default_backend()
        self.reduce = reduce
        host = packs if isinstance(packs, HostBatch) else HostBatch(list(packs))
        self.host = host
        self.half = bool(packed)          # name kept from the C ABI parameter (half_mode)
        self.packs = host.packs
        # m[r, a] <= K_r (observations of the read pair) <= 255; a 16-bit lane of the packed scoring
        # kernel therefore holds 65535 // m_max reads before it must be added to S
        self.flush_stages = max(1, min(65535 // host.m_max, SCORE_READ_CHUNK) // _cabi.GK_RT)
        self.table = host.table
        self.max_alleles = int(host.table["n_alleles"].max()) if len(host.table) else 0
        be = self.be
        self.d_table = be.upload(host.table)
        self.d_mem = be.upload(host.mem)
        if host.wire:                      # reads in the wire format: entries are rebuilt on the device
            self.d_hdr = be.upload(host.hdr)
            self.d_stream = be.upload(host.stream)
            self.d_keep = be.upload(host.keep)
            self.d_xitems = be.upload(host.xitems)
            self.d_entoff = be.empty(host.n_offsets, np.int32)
            self.d_ent = be.empty(4 * host.n_entries, np.uint32)
        else:
            self.d_entoff = be.upload(host.entoff)
            self.d_ent = be.upload(host.ent)
        self.d_L = be.empty(host.L_size, np.float32)
        self.d_LT = be.empty(host.LT_size, np.uint8)
        self.d_col = be.zeros(host.col_size, np.uint64)
        t = host.table
        self.n_cells = int((t["n_reads"].astype(np.int64) * t["n_alleles"]).sum())
        self.bytes_out = host.L_size * 4 + host.LT_size
        self._colsum_host: np.ndarray | None = None
        self.d_lik_items = None
        self._lik_key = None
        if run:
            self.run_likelihood()

    def reload(self) -> None:
        """Copy the host pools into the existing device buffers again (same batch layout): device
        addresses stay put, so a recorded CUDA graph and the cached launch plan remain valid."""
        self.be.copy_into(self.d_table, self.host.table)
        for name in self.host.input_names:
            self.be.copy_into(getattr(self, "d_" + name), getattr(self.host, name))
        self._colsum_host = None

    def lik_items(self) -> np.ndarray:
        """Work items of the likelihood build: a copy of the host batch's table (the caller sets flags)."""
        return self.host.lik_items().copy()

    def run_likelihood(self, colsum_only: np.ndarray | None = None) -> None:
        """Kernel (a) over every problem of the batch.  ``colsum_only`` (bool per problem): only the
        column sums are needed for those (they will be typed with one step), so ``L`` and ``LT`` are
        not written for them - they must not be read back or searched beyond the first step."""
        key = None if colsum_only is None or not np.any(colsum_only) else np.asarray(colsum_only, bool).tobytes()
        if self.d_lik_items is None or key != self._lik_key:
            items = self.host.lik_items()             # read only: copied when this pass sets flags
            if key is not None:
                items = items.copy()
                flag = np.where(np.asarray(colsum_only, bool), _cabi.GK_LIK_COLSUM_ONLY, 0).astype(items.dtype["flags"])
                items["flags"] = flag[self.host.lik_matrix]
            self.n_lik_items = len(items)
            self.d_lik_items = self.be.upload(items)
            self._lik_key = key
            t = self.table                    # bytes the kernel writes: L (4 B) + LT (1 B) per padded cell
            per = t["r_pad"].astype(np.int64) * (t["n_ablk"].astype(np.int64) * t["a_tile"] * 4 + t["n_alleles"])
            self.bytes_out = int(per.sum() if key is None else per[~np.asarray(colsum_only, bool)].sum())
        self.be.zero_(self.d_col)
        if self.host.wire:
            self.be.launch("gk_expand_reads", self.d_table, self.d_xitems, len(self.host.xitems), self.d_hdr,
                           self.d_stream, self.d_keep, self.d_entoff, self.d_ent, work=float(self.host.n_entries))
        self.be.launch("gk_likelihood", self.d_table, self.d_lik_items, self.n_lik_items, self.d_mem,
                       self.d_entoff, self.d_ent, self.d_L, self.d_LT, self.d_col, int(self.half),
                       work=float(self.n_cells))
        if self.reduce is not None:
            self.reduce(self.d_col)                  # read shards: column sums of the whole problem
        self._colsum_host = None

    # --- read-backs ------------------------------------------------------------
    def colsum_all(self) -> np.ndarray:
        if self._colsum_host is None:
            self._colsum_host = self.be.download(self.d_col, np.uint64).astype(np.int64)
        return self._colsum_host

    def colsum(self, i: int) -> np.ndarray:
        o = int(self.table["col_off"][i])
        return self.colsum_all()[o:o + int(self.table["n_alleles"][i])]

    def mismatch_counts(self, i: int) -> np.ndarray:
        """m[r, a] as uint8 [R, A], read back from the allele-major device copy."""
        t = self.table[i]
        o, a, rp, r = int(t["LT_off"]), int(t["n_alleles"]), int(t["r_pad"]), int(t["n_reads"])
        flat = self.be.download(self.d_LT[o:o + a * rp], np.uint8)
        return np.ascontiguousarray(flat.reshape(a, rp)[:, :r].T)

    def group_pattern(self, i: int, allele_ids) -> np.ndarray:
        """uint32 [R]: bit t set where allele ``allele_ids[t]`` attains the smallest mismatch count of
        the read among the given alleles (= the row maximum of ``probs``; novel_discover.py:62-64)."""
        ids = np.asarray(allele_ids, dtype=np.int32)
        t = self.table[i]
        r, a = int(t["n_reads"]), int(t["n_alleles"])
        if not 1 <= len(ids) <= 32:
            raise ValueError("between 1 and 32 alleles can be compared")
        if len(ids) and (ids.min() < 0 or ids.max() >= a):
            raise ValueError("allele id out of range")
        if r == 0:
            return np.zeros(0, dtype=np.uint32)
        be = self.be
        d_pat = be.empty(r, np.uint32)
        be.launch("gk_group_reads", self.d_table, int(i), r, be.upload(ids), len(ids), self.d_LT, d_pat)
        return be.download(d_pat, np.uint32)

    def blocked_counts(self, i: int) -> np.ndarray:
        """m[r, a] decoded from the blocked float32 copy (what the scoring kernel reads)."""
        t = self.table[i]
        o, a, rp, r = int(t["L_off"]), int(t["n_alleles"]), int(t["r_pad"]), int(t["n_reads"])
        tile, nb = int(t["a_tile"]), int(t["n_ablk"])
        flat = self.be.download(self.d_L[o:o + nb * rp * tile], np.float32)
        if self.half:
            pair = flat.view(np.uint16).reshape(-1, 2)
            assert np.array_equal(pair[:, 0], pair[:, 1])
            flat = pair[:, 0].astype(np.float32)
        # row-blocked [r_blk][a_blk][GK_RT][a_tile] -> [r, a]
        return flat.reshape(rp // GK_RT, nb, GK_RT, tile).transpose(0, 2, 1, 3).reshape(rp, nb * tile)[:r, :a]


# ---------------------------------------------------------------------------
# searches
# ---------------------------------------------------------------------------
@dataclass
class StepOutput:
    """Integer result of one copy-number step of one search."""

    n: int
    ids: np.ndarray          # int32 [K, n]
    score: np.ndarray        # int64 [K]   sum over reads of the min mismatch count
    cnt: np.ndarray          # int64 [K, n, n]  cnt[k, t, q-1]
    flat: np.ndarray         # int32 [K]   flat candidate index k_prev * n_cand + j
    n_unique: int
    n_alive: int
    cut: int
    tie_flags: int


class StepBatch:
    """Results of one step for every collected search, as bulk arrays."""

    def __init__(self, n: int, searches: np.ndarray, info: np.ndarray, ids, score, cnt, flat):
        self.n = n
        self.searches = searches                     # search index of each row below
        self.info = info                             # GkStepInfo [n_search]
        self.ids, self.score, self.cnt, self.flat = ids, score, cnt, flat
        self._row = {int(s): i for i, s in enumerate(searches)}

    def __contains__(self, s: int) -> bool:
        return int(s) in self._row

    def __getitem__(self, s: int) -> StepOutput:
        i = self._row[int(s)]
        n, inf = self.n, self.info[int(s)]
        k = int(inf["n_kept"])
        return StepOutput(
            n=n, ids=self.ids[i, :k, :n].copy(), score=self.score[i, :k].astype(np.int64),
            cnt=self.cnt[i, :k].astype(np.int64).reshape(k, n, n), flat=self.flat[i, :k].copy(),
            n_unique=int(inf["n_unique"]), n_alive=int(inf["n_alive"]), cut=int(inf["cut"]),
            tie_flags=int(inf["tie_flags"]))

    def keys(self):
        return self._row.keys()

    def __iter__(self):
        return iter(self._row)


class BestBatch:
    """Per collected search only the set chosen by selectBest (computed on the device)."""

    def __init__(self, n: int, searches: np.ndarray, info: np.ndarray, ids: np.ndarray, score: np.ndarray):
        self.n = n
        self.searches = searches
        self.info = info
        self.ids = ids          # int32 [len(searches), n]
        self.score = score      # int64 [len(searches)]


# scoring tile modes (gk_score.cu): 128 / 64 wide with vector loads, 16 / 32 / 48 wide remainders
MODE_F8, MODE_F4, MODE_S1, MODE_S2, MODE_S3 = 0, 1, 2, 3, 4
SHAPE_WARP_SPLIT = 1 << 16      # GkScoreItem.shape flag: warp-split tile of the packed scoring kernel


def _tile_counts(n16: np.ndarray) -> np.ndarray:
    """Tiles needed to cover n16 groups of 16: 128-wide, then 64, then a 16/32/48 remainder."""
    rem = n16 % 8
    return n16 // 8 + (rem >= 4) + ((rem % 4) > 0)


def _tile_decode(n16: np.ndarray, it: np.ndarray, blocks_per_128: int):
    """(first block, mode) of tile ``it`` in the cover of ``n16`` groups; a block is
    128 / blocks_per_128 wide (64 for kept sets, 32 for alleles)."""
    full = n16 // 8
    rem = n16 % 8
    has64 = rem >= 4
    tail = rem % 4
    is_full = it < full
    is_64 = ~is_full & has64 & (it == full)
    blk = np.where(is_full, it * blocks_per_128,
                   np.where(is_64, full * blocks_per_128,
                            full * blocks_per_128 + np.where(has64, blocks_per_128 // 2, 0)))
    mode = np.where(is_full, MODE_F8, np.where(is_64, MODE_F4, MODE_S1 + tail - 1))
    return blk, mode


class SearchGroup:
    """Greedy searches (one per entry of ``matrix_ids``) advancing in lock step."""

    def __init__(self, batch: MatrixBatch, matrix_ids, top_n: int, col_shard: tuple[int, int] | None = None,
                 reduce_scores=None, read_shard: bool = False):
        """Two ways to spread one deep problem over several ranks (both exact: every partial result is
        an integer sum, so the sharded run is bit-identical to the unsharded one):

        ``col_shard=(rank, world)`` scores only this rank's share of the candidate-column tiles;
        ``reduce_scores(tensor)`` must then sum the score pool over the ranks in place (one small
        collective per copy-number step); ``L``, ``LT``, ``P`` and everything after scoring are
        replicated.

        ``read_shard=True``: the matrices of ``batch`` hold only this rank's reads
        (``packing.shard_reads``, ``MatrixBatch(reduce=)``), so every read-streaming kernel -
        likelihood, scoring, tie counting, the ``P`` writer - works on ``R / world`` reads and nothing
        is replicated; ``reduce_scores`` sums the partial score pool and the partial tie counts
        (two small collectives per step); dedup, cut and ranking run on the complete sums on every
        rank and give identical kept sets everywhere."""
        self.col_shard = col_shard
        self.read_shard = bool(read_shard)
        self.reduce_scores = reduce_scores
        if (self.read_shard or (col_shard is not None and col_shard[1] > 1)) and reduce_scores is None:
            raise ValueError("sharding needs a reduce_scores callable")
        if self.read_shard and col_shard is not None:
            raise ValueError("shard either the candidate columns or the reads, not both")
        self.collective_bytes = 0            # bytes handed to reduce_scores since construction
        if not 1 <= top_n <= MAX_TOP_N:
            raise _packing.CapacityError(f"top_n must be in 1..{MAX_TOP_N}")
        self.batch = batch
        self.be = batch.be
        self.top_n = int(top_n)
        self.matrix_ids = np.asarray(matrix_ids, dtype=np.int32)
        ns = len(self.matrix_ids)
        self.n_search = ns
        mt = batch.table[self.matrix_ids]
        self.mt = mt
        self.R = mt["n_reads"].astype(np.int64)
        self.r_pad = mt["r_pad"].astype(np.int64)
        self.A = mt["n_alleles"].astype(np.int64)
        self.n_ablk = mt["n_ablk"].astype(np.int64)
        self.a_tile = mt["a_tile"].astype(np.int64)
        self.n_kblk_max = -(-self.top_n // GK_KB)

        tab = np.zeros(ns, dtype=SEARCH_DTYPE)
        cand_cap = np.maximum(self.A, 1)
        p_size = self.n_kblk_max * self.r_pad * GK_KB
        s_stride = self.n_ablk * self.a_tile
        s_size = self.n_kblk_max * GK_KB * s_stride
        flag_size = _round_up_arr(self.top_n * cand_cap, 16)       # uint32 per flat candidate
        alive_cap = np.maximum(self.top_n, (self.top_n * cand_cap) // 5)
        tab["P_off"] = _excl_cumsum(p_size)
        tab["S_off"] = _excl_cumsum(s_size)
        tab["cand_off"] = _excl_cumsum(cand_cap)
        tab["flag_off"] = _excl_cumsum(flag_size)
        tab["alive_off"] = _excl_cumsum(alive_cap)
        tab["matrix"] = self.matrix_ids
        tab["s_stride"] = s_stride
        tab["alive_cap"] = alive_cap
        tab["n_kblk"] = self.n_kblk_max
        self.tab = tab
        self.cand_cap = cand_cap
        be = self.be
        self.d_P = None
        self._p_size = int(p_size.sum())
        self.d_S = be.zeros(int(s_size.sum()), np.uint32)
        self.d_flag = be.empty(int(flag_size.sum()), np.uint32)
        self.d_alive = be.empty(int(alive_cap.sum()), np.int32)
        self.d_keys = be.empty(3 * int(alive_cap.sum()), np.uint64)
        out_rows = ns * self.top_n
        self.d_ids = [be.zeros(out_rows * GK_MAX_CN, np.int32), be.zeros(out_rows * GK_MAX_CN, np.int32)]
        self.d_score = [be.zeros(out_rows, np.uint32), be.zeros(out_rows, np.uint32)]
        self.d_cnt_out = be.zeros(out_rows * GK_MAX_CN * GK_MAX_CN, np.uint32)
        self.d_flat = be.zeros(out_rows, np.int32)
        self.d_info = be.zeros(ns * (STEP_INFO_DTYPE.itemsize // 4), np.int32)
        self.d_kept = be.zeros(ns, np.int32)
        # default candidates: every allele of the gene
        total = int(cand_cap.sum())
        self._default_pool = (np.arange(total, dtype=np.int64)
                              - np.repeat(tab["cand_off"], cand_cap)).astype(np.int32)
        self._pool = self._default_pool.copy()
        self.d_cand = be.upload(self._pool)
        self._pool_dirty = False
        self._plan_key = None                 # (step vector, matrix table) of the cached launch plan
        self._plan: dict = {}
        self._plan_on = False                 # True inside run_pipeline only
        self.reset()

    def _reduce(self, tensor) -> None:
        self.collective_bytes += int(tensor.numel() if hasattr(tensor, "numel") else tensor.size) * 4
        self.reduce_scores(tensor)

    def reset(self) -> None:
        """Start all searches over (buffers are reused)."""
        self.n = 0
        self.cur = 0
        self.kept = np.zeros(self.n_search, dtype=np.int32)
        self.score_cells = 0
        self.n_cand = self.A.copy()                       # current candidate count per search
        self.restricted: dict[int, np.ndarray] = {}       # search -> explicit candidate ids
        if self._pool_dirty:
            self._pool = self._default_pool.copy()
            self.d_cand = self.be.upload(self._pool)
            self._pool_dirty = False

    # --- launch plan of the pipelined run ------------------------------------------------
    def _planned(self, key, build):
        """Work-item tables of ``run_pipeline`` depend only on the step vector (their grids are sized
        from upper bounds), so a batch that is typed again reuses the device copies."""
        if not self._plan_on:
            return build()
        hit = self._plan.get(key)
        if hit is None:
            hit = self._plan[key] = build()
        return hit

    # --- candidates ------------------------------------------------------------------
    def _set_candidates(self, cands, active: np.ndarray) -> None:
        changed = False
        if cands is not None:
            for s, c in enumerate(cands):
                if not active[s]:
                    continue
                o = int(self.tab["cand_off"][s])
                a = int(self.A[s])
                if c is None:
                    if s in self.restricted:
                        del self.restricted[s]
                        self._pool[o:o + a] = np.arange(a, dtype=np.int32)
                        self.n_cand[s] = a
                        changed = True
                    continue
                c = np.asarray(c, dtype=np.int32)
                if len(c) > self.cand_cap[s]:
                    raise ValueError("more candidates than alleles in the gene")
                if len(c) and (c.min() < 0 or c.max() >= a):
                    raise ValueError("candidate allele id out of range")
                self._pool[o:o + len(c)] = c
                self.n_cand[s] = len(c)
                self.restricted[s] = c
                changed = True
        if changed:
            self.d_cand = self.be.upload(self._pool)
            self._pool_dirty = True
        self.tab["n_cand"] = np.where(active, self.n_cand, 0)

    # --- work items ---------------------------------------------------------------------
    @staticmethod
    def _product_items(counts_per_search: list[np.ndarray]):
        """Decode a flat index over the per-search product of several counts."""
        total_per = np.ones_like(counts_per_search[0])
        for c in counts_per_search:
            total_per = total_per * c
        total = int(total_per.sum())
        search = np.repeat(np.arange(len(total_per), dtype=np.int64), total_per)
        local = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(total_per) - total_per, total_per)
        parts = []
        for c in reversed(counts_per_search):
            cs = c[search]
            parts.append(local % np.maximum(cs, 1))
            local = local // np.maximum(cs, 1)
        return search, list(reversed(parts))

    def _restricted_a_tiles(self, s: int) -> list[tuple[int, int]]:
        """(a_blk, mode) CTA tiles covering the 32-column blocks that hold candidates of ``s``."""
        blocks = np.unique(self.restricted[s] // 32)
        tiles, i = [], 0
        while i < len(blocks):
            run = 1
            while i + run < len(blocks) and blocks[i + run] == blocks[i] + run and run < 4:
                run += 1
            if run == 4:
                tiles.append((int(blocks[i]), MODE_F8)); i += 4
            elif run >= 2:
                tiles.append((int(blocks[i]), MODE_F4)); i += 2
            else:
                tiles.append((int(blocks[i]), MODE_S2)); i += 1
        return tiles

    # --- packed scoring path: exact cover of ragged kept-set / allele counts ---------------------------
    # Row groups (of 8 kept sets) a warp-split tile can cover -> (G', log2 WK), and how a remainder of g groups
    # is cut: pieces that give a lane at least 6 rows where possible (a lane's shared-memory loads per minimum
    # are (2 G' + 2 TA') / (3 G' TA'): a 16-row piece needs twice the loads per cell of a 32-row one), measured
    # with tools/probe_score_shapes.py
    # _W_CUT[g] = the pieces (G', log2 WK) that cover a remainder of g row groups, in order; a piece spans
    # (8 G') << log2 WK rows and the last one may reach beyond the kept sets (P and S are allocated in whole
    # 64-row blocks and rows >= the kept count are never read back).
    # Measured per remainder size with tools/tune_row_cuts.py (profiles/r02_row_cuts.txt): one piece of four
    # (or two) warps per row block beats an exact cover by smaller pieces even when it computes up to 37 %
    # padding rows - a warp of a narrow piece has a quarter of the reads of every stage and pays the
    # per-stage overhead four times as often.
    _W_CUT = {1: ((1, 0),), 2: ((2, 0),), 3: ((3, 0),), 4: ((4, 0),), 5: ((3, 1),), 6: ((3, 1),), 7: ((2, 2),),
              8: ((4, 1),), 9: ((3, 1), (3, 0)), 10: ((3, 2),), 11: ((3, 2),), 12: ((3, 2),), 13: ((4, 2),),
              14: ((4, 2),), 15: ((4, 2),), 16: ((4, 2),)}

    _cut_cache = None

    @classmethod
    def _cut_table(cls) -> np.ndarray:
        """``_W_CUT`` as the int32 [17][3][2] array of csrc/gk_plan.cu (rebuilt when the table is replaced)."""
        key = tuple(sorted(cls._W_CUT.items()))
        if cls._cut_cache is None or cls._cut_cache[0] != key:
            table = np.zeros((17, 3, 2), dtype=np.int32)
            for g, pieces in cls._W_CUT.items():
                for i, (gp, wk) in enumerate(pieces):
                    table[g, i] = (gp, wk)
            cls._cut_cache = (key, table)
        return cls._cut_cache[1]

    @classmethod
    def _row_pieces(cls, k: int, kind: str) -> list[tuple[int, int, int]]:
        """Row tiles covering ``k`` kept sets as (first row, shape code without the column part, rows).
        kind "F": under a full-width (128-allele) column tile - 128-row tiles, then the whole 32-row groups
        of the remainder as one full-width tile; kind "H": what is left of that remainder (< 32 rows), as a
        warp-split tile (used over the two 64-allele halves); kind "W": under a warp-split column tile -
        128-row tiles, then the remainder cut into the spans a warp-split tile offers (``_W_CUT``).  Every
        piece starts at a multiple of 8 rows; kinds "F" and "H" pad nothing beyond 8 ceil(k / 8), kind "W"
        may pad the remainder up to the next multiple of 32 rows where that is faster (never beyond the
        64-row blocks P and S are allocated in)."""
        out = []
        n_full, rem = divmod(k, 128)
        if kind == "F":
            out += [(128 * i, (4 + 4) | (MODE_F8 << 8), 128) for i in range(n_full)]
            if rem // 32:
                out.append((128 * n_full, (4 + rem // 32) | (MODE_F8 << 8), 32 * (rem // 32)))
            return out
        if kind == "H":
            if rem % 32:
                gp = -(-(rem % 32) // 8)
                out.append((128 * n_full + 32 * (rem // 32), gp | SHAPE_WARP_SPLIT, 8 * gp))
            return out
        out += [(128 * i, 4 | (2 << 4) | SHAPE_WARP_SPLIT, 128) for i in range(n_full)]
        at = 128 * n_full
        for gp, wk in cls._W_CUT.get(-(-rem // 8), ()):
            out.append((at, gp | (wk << 4) | SHAPE_WARP_SPLIT, (8 * gp) << wk))
            at += (8 * gp) << wk
        return out

    def _packed_tiles(self, live: np.ndarray, kept: np.ndarray):
        """(search index into ``live``, k_blk, a_blk, shape, rows, columns) of every (row piece x column
        tile) of the packed scoring path, vectorised over the searches."""
        A = self.A[live]
        a8 = -(-A // 8)
        n128, rem8 = a8 // 16, a8 % 16
        # column tiles: (search, first 32-column block, kind 0 = F / 1 = W / 2 = H, columns / 8)
        cs, cb, ck, ct = [], [], [], []
        j_all = np.arange(len(live), dtype=np.int64)
        plain = np.array([int(s) not in self.restricted for s in live], dtype=bool)
        jf = np.repeat(j_all[plain], n128[plain])                      # full 128-column blocks
        bf = 4 * (np.arange(len(jf), dtype=np.int64) - np.repeat(_excl_cumsum(n128[plain]), n128[plain]))
        cs += [jf, jf, jf]; cb += [bf, bf, bf + 2]
        ck += [np.zeros(len(jf), np.int64), np.full(len(jf), 2), np.full(len(jf), 2)]
        ct += [np.full(len(jf), 16), np.full(len(jf), 8), np.full(len(jf), 8)]
        # one or two warp-split column tiles after the full blocks; two tiles start a 32-column block apart
        # and are cut as evenly as that allows (9 groups of 8 columns = 4 + 5, not 8 + 1)
        first = np.where(rem8 <= 8, rem8, np.where(rem8 <= 12, 4, 8))
        for extra, lo, width in ((rem8 > 0, np.zeros_like(rem8), first), (rem8 > 8, first, rem8 - first)):
            sel = plain & extra
            cs.append(j_all[sel]); cb.append(4 * n128[sel] + lo[sel] // 4)
            ck.append(np.ones(int(sel.sum()), np.int64)); ct.append(width[sel])
        for j in j_all[~plain]:                                        # restricted candidates: explicit block runs
            for a_blk, mode in self._restricted_a_tiles(int(live[j])):
                kinds = [(0, 16, 0), (2, 8, 0), (2, 8, 2)] if mode == MODE_F8 else \
                    [(1, 8 if mode == MODE_F4 else 4, 0)]
                for kind, t8, shift in kinds:
                    cs.append(np.array([j])); cb.append(np.array([a_blk + shift]))
                    ck.append(np.array([kind])); ct.append(np.array([t8]))
        cs, cb, ck, ct = (np.concatenate(x).astype(np.int64) for x in (cs, cb, ck, ct))
        # row pieces per (kept count, kind), looked up from a small table
        keys = kept[live][cs] * 3 + ck
        uniq, inv = np.unique(keys, return_inverse=True)
        table, base, count = [], [], []
        for key in uniq:
            pieces = self._row_pieces(int(key) // 3, "FWH"[int(key) % 3])
            base.append(len(table)); count.append(len(pieces)); table += pieces
        table = np.array(table, dtype=np.int64).reshape(-1, 3)
        base, count = np.array(base, dtype=np.int64), np.array(count, dtype=np.int64)
        per = count[inv]
        col = np.repeat(np.arange(len(cs), dtype=np.int64), per)
        piece = table[np.repeat(base[inv], per) + (np.arange(len(col), dtype=np.int64) - np.repeat(_excl_cumsum(per), per))]
        start, code, rows = piece[:, 0], piece[:, 1], piece[:, 2]
        split = (code & SHAPE_WARP_SPLIT) != 0
        shape = np.where(split, code | (ct[col] << 8) | (((start % GK_KB) // 8) << 20), code)
        return cs[col], start // GK_KB, cb[col], shape, rows, np.where(ck[col] == 0, 128, 8 * ct[col])

    def _score_items(self, active_idx: np.ndarray) -> np.ndarray:
        kept = self.kept.astype(np.int64)
        live = active_idx[(kept[active_idx] > 0) & (self.n_cand[active_idx] > 0)]
        self._step_cells = int((kept[live] * self.n_cand[live] * self.R[live]).sum())
        if not len(live):
            return np.zeros(0, dtype=SCORE_ITEM_DTYPE)
        half = self.batch.half
        r16 = _round_up_arr(self.R[live], _cabi.GK_RT)
        custom = {}
        native = (PLAN_NATIVE and half and not self.restricted
                  and (self.col_shard is None or self.col_shard[1] <= 1))
        if native:
            lib = _cabi.load()
            A_l = np.ascontiguousarray(self.A[live], dtype=np.int64)
            kept_l = np.ascontiguousarray(kept[live], dtype=np.int64)
            cut = self._cut_table()
            tiles = np.zeros(len(live), dtype=np.int64)
            lib.gk_plan_score_tiles(len(live), A_l.ctypes.data, kept_l.ctypes.data, cut.ctypes.data, tiles.ctypes.data)
        elif half:
            # exact cover of the ragged kept-set and allele counts (granularity 8 x 8)
            t_search, t_kblk, t_ablk, t_shape, t_rows, t_cols = self._packed_tiles(live, kept)
            tiles = np.bincount(t_search, minlength=len(live)).astype(np.int64)
        else:
            k16 = -(-kept[live] // 16)
            a16 = -(-self.A[live] // 16)
            n_kt = _tile_counts(k16)
            n_at = _tile_counts(a16)
            for j, s in enumerate(live):
                if int(s) in self.restricted:
                    custom[j] = self._restricted_a_tiles(int(s))
                    n_at[j] = len(custom[j])
            tiles = n_kt * n_at
        # Reads per work item: a launch lasts (work per CTA slot) + (one item) at best, so large chunks
        # leave a tail when a slot holds only a few of them (one deep problem spread over 444 slots,
        # or a read shard of it), and small chunks pay the pipeline fill and the final atomics of an
        # item (about SCORE_ITEM_OVERHEAD reads' worth) more often.  Pick the candidate with the
        # smallest estimate of both.
        best = None
        for cand_chunk in SCORE_CHUNK_CANDIDATES:
            if cand_chunk > SCORE_READ_CHUNK:
                continue
            n_ch = np.maximum(1, -(-r16 // cand_chunk))
            per_item = r16 / n_ch                                   # reads per item of each search
            est = float((tiles * n_ch * (per_item + SCORE_ITEM_OVERHEAD)).sum()) / SCORE_SLOTS + float(per_item.max())
            if best is None or est < best[0] * 0.995:               # prefer the larger chunk on a tie
                best = (est, cand_chunk)
        chunk = best[1] if best is not None else SCORE_READ_CHUNK
        if _os.environ.get("GK_SCORE_CHUNK"):                       # sweeps (tools/sweep_params.sh)
            chunk = int(_os.environ["GK_SCORE_CHUNK"])
        n_ch = np.maximum(1, -(-r16 // chunk))
        if native:
            r16_l = np.ascontiguousarray(r16, dtype=np.int64)
            ids = np.ascontiguousarray(live, dtype=np.int32)
            items = np.zeros(int((tiles * n_ch).sum()), dtype=SCORE_ITEM_DTYPE)
            got = lib.gk_plan_score_items(len(live), ids.ctypes.data, A_l.ctypes.data, kept_l.ctypes.data,
                                          r16_l.ctypes.data, int(chunk), cut.ctypes.data, items.ctypes.data, len(items))
            if got != len(items):
                raise RuntimeError(f"gk_plan_score_items wrote {got} rows, {len(items)} expected")
            return items
        mode_span = np.array([128, 64, 16, 32, 48, 32, 64, 96, 128], dtype=np.int64)
        if half:
            per = n_ch[t_search]
            tile = np.repeat(np.arange(len(t_search), dtype=np.int64), per)
            ich = np.arange(len(tile), dtype=np.int64) - np.repeat(_excl_cumsum(per), per)
            search, k_blk, a_blk, shape = t_search[tile], t_kblk[tile], t_ablk[tile], t_shape[tile]
            kspan, aspan = t_rows[tile], t_cols[tile]
            # column tile of an item (for column sharding): the 128-allele block, or the ragged tile after them
            n128s = ((-(-self.A[live] // 8)) // 16)[search]
            iat = np.where(a_blk < 4 * n128s, a_blk // 4, (1 << 20) | a_blk)
            items = np.zeros(len(search), dtype=SCORE_ITEM_DTYPE)
            items["search"] = live[search]
        else:
            search, (ikt, iat, ich) = self._product_items([n_kt, n_at, n_ch])
            items = np.zeros(len(search), dtype=SCORE_ITEM_DTYPE)
            items["search"] = live[search]
            k_blk, k_mode = _tile_decode(k16[search], ikt, 2)
            a_blk, a_mode = _tile_decode(a16[search], iat, 4)
            for j, tiles_j in custom.items():
                sel = np.flatnonzero(search == j)
                t = np.array(tiles_j, dtype=np.int64)
                a_blk[sel] = t[iat[sel], 0]
                a_mode[sel] = t[iat[sel], 1]
            kspan, aspan = mode_span[k_mode], mode_span[a_mode]
            shape = k_mode | (a_mode << 8)
        items["k_blk"] = k_blk
        items["a_blk"] = a_blk
        items["r0"] = ich * chunk
        items["r1"] = np.minimum((ich + 1) * chunk, r16[search])
        items["shape"] = shape
        if self.col_shard is not None and self.col_shard[1] > 1:
            rank, world = self.col_shard
            # column tiles go to the least-loaded rank, widest first (deterministic on every rank)
            owner = np.zeros(len(search), dtype=np.int64)
            for j in range(len(live)):
                sel = np.flatnonzero(search == j)
                if not len(sel):
                    continue
                tiles_j = np.unique(iat[sel])
                width = {int(t): int(aspan[sel][iat[sel] == t].max()) for t in tiles_j}
                load = [0] * world
                assign = {}
                for t in sorted(width, key=lambda t: (-width[t], t)):
                    r = min(range(world), key=lambda q: (load[q], q))
                    assign[t] = r
                    load[r] += width[t]
                owner[sel] = np.array([assign[int(t)] for t in iat[sel]])
            mine = owner == rank
            k0 = k_blk * GK_KB + (((shape >> 20) & 7) * 8 if half else 0)
            a0 = a_blk * 32
            rows_u = np.minimum(kspan, kept[live][search] - k0)
            cols_u = np.minimum(aspan, self.A[live][search] - a0)
            reads = np.minimum(items["r1"], self.R[live][search]) - items["r0"]
            useful = np.maximum(rows_u, 0) * np.maximum(cols_u, 0) * np.maximum(reads, 0)
            self._step_cells = int(useful[mine].sum())
            items, kspan, aspan = items[mine], kspan[mine], aspan[mine]
        order = np.argsort(-((items["r1"] - items["r0"]).astype(np.int64) * kspan * aspan),
                           kind="stable")
        return items[order]

    @staticmethod
    def _grid_items(idx, count, scale: int, extent, chunk: int, dtype) -> np.ndarray:
        """{search, index * scale, r0, r1} rows through gk_plan_grid_items (see the NumPy statements in
        ``_p_items`` / ``_count_items``)."""
        lib = _cabi.load()
        ids = np.ascontiguousarray(idx, dtype=np.int32)
        count = np.ascontiguousarray(count, dtype=np.int64)
        extent = np.ascontiguousarray(extent, dtype=np.int64)
        n_rows = int((count * np.maximum(1, -(-extent // chunk))).sum())
        items = np.zeros(n_rows, dtype=dtype)
        got = lib.gk_plan_grid_items(len(ids), ids.ctypes.data, count.ctypes.data, int(scale), extent.ctypes.data,
                                     int(chunk), items.ctypes.data, n_rows)
        if got != n_rows:
            raise RuntimeError(f"gk_plan_grid_items wrote {got} rows, {n_rows} expected")
        return items

    def _p_items(self, idx: np.ndarray) -> np.ndarray:
        kept = self.kept.astype(np.int64)
        idx = idx[kept[idx] > 0]
        if not len(idx):
            return np.zeros(0, dtype=P_ITEM_DTYPE)
        n_k64 = -(-kept[idx] // GK_KB)
        if PLAN_NATIVE:
            return self._grid_items(idx, n_k64, 1, self.r_pad[idx], P_READ_CHUNK, P_ITEM_DTYPE)
        n_rt = -(-self.r_pad[idx] // P_READ_CHUNK)
        search, (ik, ir) = self._product_items([n_k64, n_rt])
        items = np.zeros(len(search), dtype=P_ITEM_DTYPE)
        items["search"] = idx[search]
        items["k_blk"] = ik
        items["r0"] = ir * P_READ_CHUNK
        items["r1"] = np.minimum((ir + 1) * P_READ_CHUNK, self.r_pad[idx][search])
        return items

    def _count_items(self, active_idx: np.ndarray, n_alive: np.ndarray) -> np.ndarray:
        idx = active_idx[n_alive[active_idx] > 0]
        if not len(idx):
            return np.zeros(0, dtype=COUNT_ITEM_DTYPE)
        n_f = -(-n_alive[idx] // 8)
        r16 = _round_up_arr(self.R[idx], 16)
        if PLAN_NATIVE:
            return self._grid_items(idx, n_f, 8, r16, COUNT_READ_CHUNK, COUNT_ITEM_DTYPE)
        n_ch = np.maximum(1, -(-r16 // COUNT_READ_CHUNK))
        search, (jf, ich) = self._product_items([n_f, n_ch])
        items = np.zeros(len(search), dtype=COUNT_ITEM_DTYPE)
        items["search"] = idx[search]
        items["f0"] = jf * 8
        items["r0"] = ich * COUNT_READ_CHUNK
        items["r1"] = np.minimum((ich + 1) * COUNT_READ_CHUNK, r16[search])
        return items

    def _write_p(self, idx: np.ndarray, n_set: int) -> None:
        def plan():
            items = self._p_items(idx)
            return (self.be.upload(items) if len(items) else None, len(items),
                    float((items["r1"] - items["r0"]).sum()) * GK_KB)
        d_items, n_items, work = self._planned(("write_p", n_set), plan)
        if not n_items:
            return
        if self.d_P is None:
            self.d_P = self.be.empty(self._p_size, np.uint16 if self.batch.half else np.float32)
        self.be.launch("gk_write_p", self.batch.d_table, self.d_tab, d_items, n_items, self.top_n, n_set,
                       self.d_kept, self.d_ids[self.cur], self.batch.d_LT, self.d_P, int(self.batch.half),
                       work=work)

    def _collect(self, rows: np.ndarray, n: int, info: np.ndarray) -> StepBatch:
        be, tn = self.be, self.top_n
        ids = be.gather_rows(self.d_ids[self.cur], tn * GK_MAX_CN, rows).reshape(len(rows), tn, GK_MAX_CN)
        score = be.gather_rows(self.d_score[self.cur], tn, rows).view(np.uint32)
        cnt = be.gather_rows(self.d_cnt_out[: self.n_search * tn * n * n], tn * n * n, rows) \
            .view(np.uint32).reshape(len(rows), tn, n * n)
        flat = be.gather_rows(self.d_flat, tn, rows)
        return StepBatch(n, rows, info, ids, score, cnt, flat)

    # --- one copy-number step ---------------------------------------------------------
    def _rescore_and_rank(self, active_idx: np.ndarray, f_cap: np.ndarray, n: int, new: int) -> None:
        """gk_rescore_count over the first f_cap[s] alive sets of every search, then gk_rank."""
        be, bt = self.be, self.batch

        def plan():
            self.tab["cnt_off"] = _excl_cumsum(f_cap * n * n)
            tab = self.tab.copy()
            tab["alive_cap"] = f_cap             # the kernels clamp the alive count to this capacity
            c_items = self._count_items(active_idx, f_cap)
            return (be.upload(tab), be.upload(c_items), len(c_items),
                    be.empty(int((f_cap * n * n).sum()), np.uint32), float((f_cap * self.R).sum()) * n)
        self.d_tab, d_citems, n_citems, d_cnt, work = self._planned(("count", n), plan)
        be.zero_(d_cnt)
        be.launch("gk_rescore_count", bt.d_table, self.d_tab, d_citems, n_citems, self.top_n, n,
                  self.d_info, self.d_ids[self.cur], self.d_cand, self.d_alive, bt.d_LT, d_cnt,
                  work=work)
        if self.read_shard:
            self._reduce(d_cnt)                      # tie counts over the reads of the other ranks
        be.launch("gk_rank", bt.d_table, self.d_tab, self.n_search, self.top_n, n, self.d_ids[self.cur],
                  self.d_cand, self.d_alive, self.d_S, d_cnt, bt.d_col, self.d_score[self.cur], self.d_keys,
                  self.d_ids[new], self.d_score[new], self.d_cnt_out, self.d_flat, self.d_info, self.d_kept,
                  int(bt.half))

    def _collect_best(self, rows: np.ndarray, n: int, info: np.ndarray) -> BestBatch:
        flat = rows.astype(np.int64) * self.top_n + info["best_rank"][rows].astype(np.int64)
        ids = self.be.gather_rows(self.d_ids[self.cur], GK_MAX_CN, flat)[:, :n]
        score = self.be.gather_rows(self.d_score[self.cur], 1, flat).view(np.uint32).reshape(-1).astype(np.int64)
        return BestBatch(n, rows, info, ids, score)

    def step(self, cands=None, active=None, need_next=None, collect=None, best_only: bool = False):
        """Advance the active searches by one allele.

        cands      per-search candidate allele ids (None = every allele of the gene)
        active     bool per search (default all)
        need_next  bool per search: will be stepped again (P is only written for those)
        collect    bool per search: download the full step output (default: the active ones)
        best_only  download only the set picked by selectBest (info.best_rank) per collected search
        """
        ns = self.n_search
        active = np.ones(ns, bool) if active is None else np.asarray(active, bool)
        need_next = active.copy() if need_next is None else (np.asarray(need_next, bool) & active)
        collect = active if collect is None else (np.asarray(collect, bool) & active)
        be, bt = self.be, self.batch
        n = self.n + 1
        if n > GK_MAX_CN:
            raise _packing.CapacityError(f"copy number above {GK_MAX_CN} is not supported by the search kernels")
        active_idx = np.flatnonzero(active)
        self._set_candidates(cands, active)
        self.d_tab = be.upload(self.tab)
        new = 1 - self.cur

        if n == 1:
            be.launch("gk_first_step", bt.d_table, self.d_tab, ns, self.top_n, bt.d_col, self.d_cand,
                      self.d_ids[new], self.d_score[new], self.d_cnt_out, self.d_flat, self.d_info, self.d_kept)
        else:
            items = self._score_items(active_idx)
            be.zero_(self.d_S)
            d_items = be.upload(items)
            be.launch("gk_score", bt.d_table, self.d_tab, d_items, len(items), bt.d_L, self.d_P, self.d_S,
                      int(bt.half), int(bt.flush_stages), None, work=float(self._step_cells))
            self.score_cells += self._step_cells
            if self.reduce_scores is not None:
                self._reduce(self.d_S)                   # sum of the per-rank column slices / read shards
            be.launch("gk_select", bt.d_table, self.d_tab, ns, self.top_n, n - 1, max(bt.max_alleles, 1),
                      int(self.cand_cap.max()), self.d_kept, self.d_ids[self.cur], self.d_cand, self.d_S,
                      bt.d_col, self.d_score[self.cur], self.d_flag, self.d_alive, self.d_info, int(bt.half))
            # Rescoring grids are sized for top_n + slack alive sets per search, so no read-back
            # is needed between selection and ranking; a search with more alive sets (a large
            # exact tie at the cut) makes the step fall back to exactly sized grids below.
            f_cap = np.where(active, np.minimum(self.tab["alive_cap"], self.top_n + ALIVE_SLACK), 0).astype(np.int64)
            self._rescore_and_rank(active_idx, f_cap, n, new)
            info = be.download(self.d_info, None).view(STEP_INFO_DTYPE).copy()   # read-back: counts, calls
            n_alive = np.minimum(info["n_alive"], self.tab["alive_cap"]).astype(np.int64)
            n_alive[~active] = 0
            if np.any(n_alive > f_cap):
                self._rescore_and_rank(active_idx, n_alive, n, new)
                info = None
        self.cur = new
        if n == 1 or info is None:
            info = be.download(self.d_info, None).view(STEP_INFO_DTYPE).copy()
        self.kept[active_idx] = info["n_kept"][active_idx]
        self.n = n
        nxt = np.flatnonzero(need_next)
        if len(nxt):
            self._write_p(nxt, n)
        if best_only:
            return self._collect_best(np.flatnonzero(collect), n, info)
        return self._collect(np.flatnonzero(collect), n, info)

    def run_pipeline(self, steps: np.ndarray):
        """All copy-number steps of all searches with no host round trip in between.

        ``steps[s]`` is the number of steps of search ``s`` (all alleles are candidates).  Grids
        are sized from upper bounds of the kept-set counts (``min(top_n, K * A)``; tiles beyond
        the real count exit at once) and from ``top_n + ALIVE_SLACK`` alive sets, so the host
        enqueues every launch up front and reads back once at the end.  Returns
        ``(ids [S, max_step], score [S], info per finishing step [S])`` of the set picked by
        selectBest for every search, or ``None`` when a search had more alive sets than the
        pre-sized grids (exact tie at a cut larger than the slack): the caller then repeats the
        run step by step."""
        return self.run_pipeline_finish(self.run_pipeline_start(steps))

    def run_pipeline_start(self, steps: np.ndarray):
        """Enqueue every launch of ``run_pipeline`` on the current stream without synchronising;
        ``run_pipeline_finish`` reads the results back.  Lets a caller enqueue several groups (on
        several streams) before it blocks on the first."""
        core = self.run_pipeline_enqueue(steps)
        return core[:-1] + (self.be.download_async(core[-1]),)

    def run_pipeline_enqueue(self, steps: np.ndarray):
        """The launches of ``run_pipeline_start`` without the read-back: the last element of the
        returned tuple is the list of device arrays to read instead of the download handle (the
        part a CUDA graph can capture)."""
        bt = self.batch
        steps = np.asarray(steps, dtype=np.int64)
        self.reset()
        max_step = int(steps.max(initial=0))
        if max_step > GK_MAX_CN:
            raise _packing.CapacityError(f"copy number above {GK_MAX_CN} is not supported by the search kernels")
        key = (steps.tobytes(), id(bt.host))
        if key != self._plan_key:
            self._plan_key, self._plan, self._plan_host = key, {}, bt.host     # (the reference pins the id)
        self._plan_on = True
        try:
            return self._enqueue_pipeline(steps, max_step)
        finally:
            self._plan_on = False

    def _enqueue_pipeline(self, steps: np.ndarray, max_step: int):
        be, bt, ns = self.be, self.batch, self.n_search
        k_ub = np.zeros(ns, dtype=np.int64)
        snaps, finals, f_caps, actives = [], [], [], []
        self._pipe_cells = 0.0
        for n in range(1, max_step + 1):
            active = steps >= n
            active_idx = np.flatnonzero(active)
            self.tab["n_cand"] = np.where(active, self.n_cand, 0)
            self.d_tab = self._planned(("tab", n), lambda: be.upload(self.tab))
            new = 1 - self.cur
            if n == 1:
                be.launch("gk_first_step", bt.d_table, self.d_tab, ns, self.top_n, bt.d_col, self.d_cand,
                          self.d_ids[new], self.d_score[new], self.d_cnt_out, self.d_flat, self.d_info,
                          self.d_kept)
                k_ub = np.minimum(self.top_n, self.n_cand)
                f_cap = np.zeros(ns, dtype=np.int64)
            else:
                self.kept = np.where(active, k_ub, 0).astype(np.int32)        # upper bounds size the grids

                def score_plan():
                    items = self._score_items(active_idx)
                    return be.upload(items), len(items), float(self._step_cells)
                d_items, n_items, step_cells = self._planned(("score", n), score_plan)
                self._pipe_cells += step_cells
                be.zero_(self.d_S)
                be.launch("gk_score", bt.d_table, self.d_tab, d_items, n_items, bt.d_L, self.d_P, self.d_S,
                          int(bt.half), int(bt.flush_stages), self.d_kept, work=step_cells)
                if self.reduce_scores is not None:
                    self._reduce(self.d_S)
                be.launch("gk_select", bt.d_table, self.d_tab, ns, self.top_n, n - 1, max(bt.max_alleles, 1),
                          int(self.cand_cap.max()), self.d_kept, self.d_ids[self.cur], self.d_cand, self.d_S,
                          bt.d_col, self.d_score[self.cur], self.d_flag, self.d_alive, self.d_info, int(bt.half))
                f_cap = np.where(active, np.minimum(self.tab["alive_cap"], self.top_n + ALIVE_SLACK), 0).astype(np.int64)
                self._rescore_and_rank(active_idx, f_cap, n, new)
                k_ub = np.minimum(self.top_n, k_ub * self.n_cand)
            self.cur = new
            self.n = n
            snaps.append(be.snapshot(self.d_info))
            f_caps.append(f_cap)
            actives.append(active)
            done = np.flatnonzero(steps == n)
            if len(done):
                d_done = self._planned(("done", n), lambda: be.upload(done.astype(np.int64)))
                finals.append((n, done) + tuple(be.gather_best_device(self.d_ids[self.cur], self.d_score[self.cur],
                                                                      self.d_info, d_done, self.top_n)))
            nxt = np.flatnonzero(steps > n)
            if len(nxt):
                self.kept = np.where(steps > n, k_ub, 0).astype(np.int32)
                self._write_p(nxt, n)
        # ---- the single read-back: one device->host copy of everything the host needs --------
        tensors = list(snaps)
        for _, _, d_ids, d_score in finals:
            tensors += [d_ids, d_score]
        finals = [(n, done) for n, done, _, _ in finals]
        return steps, max_step, len(snaps), finals, f_caps, actives, tensors

    def run_pipeline_finish(self, pending):
        """Wait for the read-back of a pipelined run (see ``run_pipeline``)."""
        be, ns = self.be, self.n_search
        steps, max_step, n_snaps, finals, f_caps, actives, handle = pending
        arrays = be.download_wait(handle)
        infos = [a.view(STEP_INFO_DTYPE).copy() for a in arrays[:n_snaps]]
        arrays = arrays[n_snaps:]
        cells = 0
        for i, info in enumerate(infos):
            n = i + 1
            if n >= 2:
                alive = np.where(actives[i], np.minimum(info["n_alive"], self.tab["alive_cap"]), 0)
                if np.any(alive > f_caps[i]):
                    return None
                prev_kept = infos[i - 1]["n_kept"].astype(np.int64)
                cells += int((np.where(actives[i], prev_kept, 0) * self.n_cand * self.R).sum())
        if self.col_shard is not None and self.col_shard[1] > 1:
            # sharded columns: this rank's share, from the work items (sized from the kept-set upper
            # bounds, which are exact whenever a step keeps top_n sets)
            cells = int(self._pipe_cells)
        self.score_cells = cells
        out_ids = np.full((ns, max(max_step, 1)), -1, dtype=np.int64)
        out_score = np.zeros(ns, dtype=np.int64)
        out_info = np.zeros(ns, dtype=STEP_INFO_DTYPE)
        for j, (n, done) in enumerate(finals):
            ids = arrays[2 * j].view(np.int32).reshape(len(done), GK_MAX_CN)[:, :n]
            out_ids[done, :n] = ids
            out_score[done] = arrays[2 * j + 1].view(np.uint32).astype(np.int64)
            out_info[done] = infos[n - 1][done]
            # tie flags are sticky over the steps of a search: a tie at a cut of an earlier step
            # decides which sets the later steps grow
            for earlier in infos[: n - 1]:
                out_info["tie_flags"][done] |= earlier["tie_flags"][done]
        self.kept = infos[-1]["n_kept"].astype(np.int32) if infos else self.kept
        return out_ids, out_score, out_info

    def restore(self, s: int, ids: np.ndarray, p_colsum: np.ndarray) -> None:
        """Re-seed search ``s`` with kept sets ``ids`` [K, n] and their sum_r P[r, k]
        (``p_colsum``, the score of each set); P itself is rebuilt by gk_write_p."""
        ids = np.asarray(ids, dtype=np.int32)
        k, n = ids.shape
        sc = self.be.download(self.d_score[self.cur], np.uint32).reshape(self.n_search, self.top_n).copy()
        sc[s] = 0
        sc[s, :k] = np.asarray(p_colsum, dtype=np.uint32)
        self.d_score[self.cur] = self.be.upload(sc)
        if k > self.top_n or n > GK_MAX_CN:
            raise ValueError("kept sets do not fit this search")
        host = self.be.download(self.d_ids[self.cur], np.int32).reshape(self.n_search, self.top_n, GK_MAX_CN).copy()
        host[s] = 0
        host[s, :k, :n] = ids
        self.d_ids[self.cur] = self.be.upload(host)
        self.kept[s] = k
        self.d_kept = self.be.upload(self.kept)
        self.n = n
        self.tab["n_cand"] = self.n_cand
        self.d_tab = self.be.upload(self.tab)
        self._write_p(np.array([s]), n)

    def materialize_p(self, s: int, ids: np.ndarray) -> np.ndarray:
        """min over members of m[r, id] for arbitrary id sets of search ``s`` -> int [R, K].

        Runs gk_write_p into a scratch buffer (used for the lazy ``allele_prob``)."""
        ids = np.asarray(ids, dtype=np.int32)
        k, n = ids.shape
        be, bt = self.be, self.batch
        if k > self.top_n:
            raise ValueError("more sets than top_n")
        n_kblk = max(1, -(-k // GK_KB))
        r_pad, r = int(self.r_pad[s]), int(self.R[s])
        tab = self.tab[s:s + 1].copy()
        tab["P_off"] = 0
        tab["n_kblk"] = n_kblk
        buf = np.zeros((self.top_n, GK_MAX_CN), dtype=np.int32)
        buf[:k, :n] = ids
        d_ids = be.upload(buf)
        d_tab = be.upload(tab)
        d_kept = be.upload(np.array([k], dtype=np.int32))
        kb, rr = np.meshgrid(np.arange(n_kblk), np.arange(0, r_pad, P_READ_CHUNK), indexing="ij")
        items = np.zeros(kb.size, dtype=P_ITEM_DTYPE)
        items["k_blk"] = kb.reshape(-1)
        items["r0"] = rr.reshape(-1)
        items["r1"] = np.minimum(rr.reshape(-1) + P_READ_CHUNK, r_pad)
        d_items = be.upload(items)
        p_dtype = np.uint16 if bt.half else np.float32
        d_P = be.empty(n_kblk * r_pad * GK_KB, p_dtype)
        be.launch("gk_write_p", bt.d_table, d_tab, d_items, len(items), self.top_n, n, d_kept, d_ids,
                  bt.d_LT, d_P, int(bt.half))
        p = be.download(d_P, p_dtype).reshape(r_pad // GK_RT, n_kblk, GK_RT, GK_KB)
        return p.transpose(0, 2, 1, 3).reshape(r_pad, n_kblk * GK_KB)[:r, :k].astype(np.int64)


def _excl_cumsum(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    out = np.zeros(len(x), dtype=np.int64)
    if len(x) > 1:
        np.cumsum(x[:-1], out=out[1:])
    return out


def _round_up_arr(x: np.ndarray, m: int) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    return (x + m - 1) // m * m
