"""
Batched typing engine: drives the CUDA kernels of libgk_typing.so.

``MatrixBatch``   uploads the packed problems of any number of (sample, gene)
                  pairs into pooled device buffers and runs the likelihood
                  kernel once (kernel (a)).
``SearchGroup``   advances any number of greedy top-N searches over those
                  matrices in lock step, one copy-number step per call
                  (kernels (b) and (c)): score -> dedup/select -> rescore ->
                  rank -> next P.  Work-item lists are built on the host with
                  NumPy; two small device->host reads per step size the grids.

Batching over genes and samples is what fills a 148-SM GPU: one WGS sample is
17 problems of a few thousand reads each.

The engine computes in exact integers (mismatch counts); ``typing_mulit_allele``
converts to the reference's float64 log10 units at the API boundary.
There is no CPU implementation of any kernel in this package.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _cabi
from ._cabi import (COUNT_ITEM_DTYPE, GK_KB, GK_LIK_READS, GK_MAX_CN, LIK_ITEM_DTYPE, MATRIX_DTYPE,
                    P_ITEM_DTYPE, SCORE_ITEM_DTYPE, SEARCH_DTYPE, STEP_INFO_DTYPE)
from .packing import GenePack

SCORE_READ_CHUNK = 8192     # reads per scoring work item: 8192 * 255 < 2^24 keeps float32 sums exact
COUNT_READ_CHUNK = 16384    # reads per rescoring work item
MAX_TOP_N = 2048


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pick_a_tile(n_alleles: int) -> int:
    for tile in (16, 32, 64):
        if n_alleles <= tile:
            return tile
    return 128


# ---------------------------------------------------------------------------
# backend: device memory + kernel launches (torch is plumbing only)
# ---------------------------------------------------------------------------
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
        self.timing: dict[str, list] | None = None   # name -> [(start_event, end_event, work)]

    # --- memory ---------------------------------------------------------------
    _NP2T = {"uint8": "uint8", "int32": "int32", "uint32": "int32", "float32": "float32",
             "int64": "int64", "uint64": "int64"}

    def _tdtype(self, dtype):
        return getattr(self.torch, self._NP2T[np.dtype(dtype).name])

    def zeros(self, n: int, dtype):
        return self.torch.zeros(max(int(n), 1), dtype=self._tdtype(dtype), device=self.device)

    def empty(self, n: int, dtype):
        return self.torch.empty(max(int(n), 1), dtype=self._tdtype(dtype), device=self.device)

    def upload(self, array: np.ndarray):
        array = np.ascontiguousarray(array)
        if array.dtype.fields is not None or array.dtype.name not in self._NP2T:
            array = array.view(np.uint8)
        elif array.dtype.name == "uint32":
            array = array.view(np.int32)
        elif array.dtype.name == "uint64":
            array = array.view(np.int64)
        if array.size == 0:
            return self.zeros(1, array.dtype)
        return self.torch.from_numpy(array.reshape(-1)).to(self.device, non_blocking=False)

    def download(self, tensor, dtype=None, count: int | None = None) -> np.ndarray:
        if count is not None:
            tensor = tensor[:count]
        out = tensor.cpu().numpy()
        return out.view(dtype) if dtype is not None else out

    def zero_(self, tensor) -> None:
        tensor.zero_()

    def sync(self) -> None:
        self.torch.cuda.synchronize(self.device)

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


# ---------------------------------------------------------------------------
# matrices
# ---------------------------------------------------------------------------
class MatrixBatch:
    """Likelihood data of a batch of gene problems, resident on one GPU."""

    def __init__(self, packs: list[GenePack], backend=None, run: bool = True):
        self.be = backend if backend is not None else CudaBackend()
        self.packs = packs
        n = len(packs)
        table = np.zeros(n, dtype=MATRIX_DTYPE)
        mem_off = entoff_off = ent_base = L_off = LT_off = col_off = 0
        entoffs = []
        for i, p in enumerate(packs):
            if p.n_reads and int(p.k_obs.astype(np.int64).sum()) >= 2 ** 32 - 1:
                raise ValueError("sum of observations per problem must stay below 2^32 (32-bit score atomics)")
            a_tile = pick_a_tile(p.n_alleles)
            n_ablk = max(1, -(-p.n_alleles // a_tile))
            r_pad = max(128, _round_up(p.n_reads, 128))
            table[i] = (mem_off, entoff_off, L_off, LT_off, col_off, p.n_reads, p.n_alleles, p.n_words,
                        r_pad, a_tile, n_ablk)
            entoffs.append(p.ent_off.astype(np.int64) + ent_base)
            mem_off += p.mem_words.size
            entoff_off += p.n_reads + 1
            ent_base += p.n_entries
            L_off += n_ablk * r_pad * a_tile
            LT_off += max(p.n_alleles, 1) * r_pad
            col_off += max(p.n_alleles, 1)
        if ent_base >= 2 ** 31:
            raise ValueError("entry pool exceeds 2^31 entries; split the batch")
        self.table = table
        self.max_alleles = int(table["n_alleles"].max()) if n else 0
        be = self.be
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt))
        self.d_table = be.upload(table)
        self.d_mem = be.upload(cat([p.mem_words.reshape(-1) for p in packs], np.uint32))
        self.d_entoff = be.upload(cat(entoffs, np.int32))
        self.d_ent_word = be.upload(cat([p.ent_word for p in packs], np.int32))
        self.d_ent_pos = be.upload(cat([p.ent_pos for p in packs], np.uint32))
        self.d_ent_neg = be.upload(cat([p.ent_neg for p in packs], np.uint32))
        self.d_L = be.empty(L_off, np.float32)
        self.d_LT = be.empty(LT_off, np.uint8)
        self.d_col = be.zeros(col_off, np.uint64)
        self.n_cells = int((table["n_reads"].astype(np.int64) * table["n_alleles"]).sum())
        self.bytes_out = int(L_off) * 4 + int(LT_off)
        self._colsum_host: np.ndarray | None = None
        if run:
            self.run_likelihood()

    def lik_items(self) -> np.ndarray:
        t = self.table
        tiles = t["r_pad"] // GK_LIK_READS
        per = tiles * t["n_ablk"]
        total = int(per.sum())
        items = np.zeros(total, dtype=LIK_ITEM_DTYPE)
        mat = np.repeat(np.arange(len(t), dtype=np.int32), per)
        start = np.repeat(np.cumsum(per) - per, per)
        local = np.arange(total) - start
        items["matrix"] = mat
        items["a_blk"] = local // tiles[mat]
        items["r0"] = (local % tiles[mat]) * GK_LIK_READS
        return items

    def run_likelihood(self) -> None:
        items = self.lik_items()
        self.be.zero_(self.d_col)
        self.d_lik_items = self.be.upload(items)
        self.be.launch("gk_likelihood", self.d_table, self.d_lik_items, len(items), self.d_mem, self.d_entoff,
                       self.d_ent_word, self.d_ent_pos, self.d_ent_neg, self.d_L, self.d_LT, self.d_col,
                       work=float(self.n_cells))
        self._colsum_host = None

    # --- read-backs ------------------------------------------------------------
    def colsum(self, i: int) -> np.ndarray:
        if self._colsum_host is None:
            self._colsum_host = self.be.download(self.d_col, np.uint64).astype(np.int64)
        o = int(self.table["col_off"][i])
        return self._colsum_host[o:o + int(self.table["n_alleles"][i])]

    def mismatch_counts(self, i: int) -> np.ndarray:
        """m[r, a] as uint8 [R, A], read back from the allele-major device copy."""
        t = self.table[i]
        o, a, rp, r = int(t["LT_off"]), int(t["n_alleles"]), int(t["r_pad"]), int(t["n_reads"])
        flat = self.be.download(self.d_LT[o:o + a * rp], np.uint8)
        return np.ascontiguousarray(flat.reshape(a, rp)[:, :r].T)

    def blocked_counts(self, i: int) -> np.ndarray:
        """m[r, a] decoded from the blocked float32 copy (what the scoring kernel reads)."""
        t = self.table[i]
        o, a, rp, r = int(t["L_off"]), int(t["n_alleles"]), int(t["r_pad"]), int(t["n_reads"])
        tile, nb = int(t["a_tile"]), int(t["n_ablk"])
        flat = self.be.download(self.d_L[o:o + nb * rp * tile], np.float32)
        return flat.reshape(nb, rp, tile).transpose(1, 0, 2).reshape(rp, nb * tile)[:r, :a]


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


class SearchGroup:
    """Greedy searches (one per entry of ``matrix_ids``) advancing in lock step."""

    def __init__(self, batch: MatrixBatch, matrix_ids, top_n: int, max_cn: int = GK_MAX_CN):
        if not 1 <= top_n <= MAX_TOP_N:
            raise ValueError(f"top_n must be in 1..{MAX_TOP_N}")
        self.batch = batch
        self.be = batch.be
        self.top_n = int(top_n)
        self.matrix_ids = np.asarray(matrix_ids, dtype=np.int32)
        ns = len(self.matrix_ids)
        self.n_search = ns
        mt = batch.table[self.matrix_ids]
        self.mt = mt
        self.n_kblk_max = -(-self.top_n // GK_KB)
        self.n = 0
        self.kept = np.zeros(ns, dtype=np.int32)
        self.score_cells = 0

        tab = np.zeros(ns, dtype=SEARCH_DTYPE)
        cand_cap = np.maximum(mt["n_alleles"], 1).astype(np.int64)
        p_size = self.n_kblk_max * mt["r_pad"].astype(np.int64) * GK_KB
        s_stride = (mt["n_ablk"] * mt["a_tile"]).astype(np.int64)
        s_size = self.n_kblk_max * GK_KB * s_stride
        flag_size = _round_up_arr(self.top_n * cand_cap, 16)
        alive_cap = np.maximum(self.top_n, (self.top_n * cand_cap) // 5)
        tab["P_off"] = _excl_cumsum(p_size)
        tab["S_off"] = _excl_cumsum(s_size)
        tab["cand_off"] = _excl_cumsum(cand_cap)
        tab["flag_off"] = _excl_cumsum(flag_size)
        tab["alive_off"] = _excl_cumsum(alive_cap)
        tab["matrix"] = self.matrix_ids
        tab["s_stride"] = s_stride
        tab["alive_cap"] = alive_cap
        self.tab = tab
        self.cand_cap = cand_cap
        be = self.be
        self.d_P = None
        self._p_size = int(p_size.sum())
        self.d_S = be.zeros(int(s_size.sum()), np.uint32)
        self.d_cand = be.zeros(int(cand_cap.sum()), np.int32)
        self.d_flag = be.empty(int(flag_size.sum()), np.uint8)
        self.d_alive = be.empty(int(alive_cap.sum()), np.int32)
        self.d_keys = be.empty(3 * int(alive_cap.sum()), np.uint64)
        out_rows = ns * self.top_n
        self.d_ids = [be.zeros(out_rows * GK_MAX_CN, np.int32), be.zeros(out_rows * GK_MAX_CN, np.int32)]
        self.d_score = be.zeros(out_rows, np.uint32)
        self.d_cnt_out = be.zeros(out_rows * GK_MAX_CN * GK_MAX_CN, np.uint32)
        self.d_flat = be.zeros(out_rows, np.int32)
        self.d_info = be.zeros(ns * (STEP_INFO_DTYPE.itemsize // 4), np.int32)
        self.d_kept = be.zeros(ns, np.int32)
        self.cur = 0  # index of the ids buffer holding the current kept sets
        self.cands: list[np.ndarray] = [np.zeros(0, np.int32)] * ns

    # --- helpers -----------------------------------------------------------------
    def _set_candidates(self, cands, active: np.ndarray) -> None:
        pool = np.zeros(int(self.cand_cap.sum()), dtype=np.int32)
        for s in range(self.n_search):
            if not active[s]:
                self.tab["n_cand"][s] = 0
                continue
            c = cands[s]
            a = int(self.mt["n_alleles"][s])
            c = np.arange(a, dtype=np.int32) if c is None else np.asarray(c, dtype=np.int32)
            if len(c) > self.cand_cap[s]:
                raise ValueError("more candidates than alleles in the gene")
            if len(c) and (c.min() < 0 or c.max() >= a):
                raise ValueError("candidate allele id out of range")
            o = int(self.tab["cand_off"][s])
            pool[o:o + len(c)] = c
            self.tab["n_cand"][s] = len(c)
            self.cands[s] = c
        self.d_cand = self.be.upload(pool)

    def _read_chunks(self, s_idx: np.ndarray, chunk: int, align: int):
        """(search, r0, r1) triples covering [0, round_up(R, align)) per search."""
        out_s, out_r0, out_r1 = [], [], []
        r_end = _round_up_arr(self.mt["n_reads"][s_idx].astype(np.int64), align)
        n_chunk = np.maximum(1, -(-r_end // chunk))
        for s, re, nc in zip(s_idx, r_end, n_chunk):
            r0 = np.arange(nc, dtype=np.int64) * chunk
            out_s.append(np.full(nc, s, dtype=np.int64))
            out_r0.append(r0)
            out_r1.append(np.minimum(r0 + chunk, re))
        if not out_s:
            z = np.zeros(0, np.int64)
            return z, z, z
        return np.concatenate(out_s), np.concatenate(out_r0), np.concatenate(out_r1)

    def _score_items(self, active_idx: np.ndarray) -> np.ndarray:
        rows = []
        cells = 0
        for s in active_idx:
            k = int(self.kept[s])
            c = self.cands[s]
            if k == 0 or len(c) == 0:
                continue
            a_tile = int(self.mt["a_tile"][s])
            ablks = np.unique(c // a_tile)
            n_kblk = -(-k // GK_KB)
            cs, r0, r1 = self._read_chunks(np.array([s]), SCORE_READ_CHUNK, _cabi.GK_RT)
            kb, ab, ch = np.meshgrid(np.arange(n_kblk), ablks, np.arange(len(r0)), indexing="ij")
            item = np.zeros(kb.size, dtype=SCORE_ITEM_DTYPE)
            item["search"] = s
            item["k_blk"] = kb.reshape(-1)
            item["a_blk"] = ab.reshape(-1)
            item["r0"] = r0[ch.reshape(-1)]
            item["r1"] = r1[ch.reshape(-1)]
            rows.append(item)
            cells += k * len(c) * int(self.mt["n_reads"][s])
        self._step_cells = cells
        if not rows:
            return np.zeros(0, dtype=SCORE_ITEM_DTYPE)
        items = np.concatenate(rows)
        # longest items first
        order = np.argsort(-(items["r1"] - items["r0"]).astype(np.int64) * self.mt["a_tile"][items["search"]],
                           kind="stable")
        return items[order]

    def _p_items(self, idx: np.ndarray) -> np.ndarray:
        rows = []
        for s in idx:
            k = int(self.kept[s])
            if k == 0:
                continue
            n_kblk = -(-k // GK_KB)
            r0 = np.arange(0, int(self.mt["r_pad"][s]), 128)
            kb, rr = np.meshgrid(np.arange(n_kblk), r0, indexing="ij")
            item = np.zeros(kb.size, dtype=P_ITEM_DTYPE)
            item["search"] = s
            item["k_blk"] = kb.reshape(-1)
            item["r0"] = rr.reshape(-1)
            rows.append(item)
        return np.concatenate(rows) if rows else np.zeros(0, dtype=P_ITEM_DTYPE)

    def _write_p(self, idx: np.ndarray, n_set: int) -> None:
        items = self._p_items(idx)
        if not len(items):
            return
        if self.d_P is None:
            self.d_P = self.be.empty(self._p_size, np.float32)
        d_items = self.be.upload(items)
        self.be.launch("gk_write_p", self.batch.d_table, self.d_tab, d_items, len(items), self.top_n, n_set,
                       self.d_kept, self.d_ids[self.cur], self.batch.d_LT, self.d_P,
                       work=float(len(items)) * 128 * 128)

    def _collect(self, active_idx: np.ndarray, n: int) -> dict[int, StepOutput]:
        be = self.be
        ns, tn = self.n_search, self.top_n
        info = be.download(self.d_info, None).view(STEP_INFO_DTYPE)
        ids = be.download(self.d_ids[self.cur], np.int32).reshape(ns, tn, GK_MAX_CN)
        score = be.download(self.d_score, np.uint32).reshape(ns, tn)
        cnt = be.download(self.d_cnt_out, np.uint32).reshape(ns, tn, GK_MAX_CN * GK_MAX_CN)
        flat = be.download(self.d_flat, np.int32).reshape(ns, tn)
        out = {}
        for s in active_idx:
            k = int(info["n_kept"][s])
            self.kept[s] = k
            out[int(s)] = StepOutput(
                n=n, ids=ids[s, :k, :n].copy(), score=score[s, :k].astype(np.int64),
                cnt=cnt[s, :k, :n * n].astype(np.int64).reshape(k, n, n), flat=flat[s, :k].copy(),
                n_unique=int(info["n_unique"][s]), n_alive=int(info["n_alive"][s]),
                cut=int(info["cut"][s]), tie_flags=int(info["tie_flags"][s]))
        return out

    # --- one copy-number step ---------------------------------------------------------
    def step(self, cands=None, active=None, need_next=None) -> dict[int, StepOutput]:
        """Advance the active searches by one allele.

        cands      per-search candidate allele ids (None = every allele of the gene)
        active     bool per search (default all)
        need_next  bool per search: will be stepped again (P is only written for those)
        """
        ns = self.n_search
        active = np.ones(ns, bool) if active is None else np.asarray(active, bool)
        need_next = active.copy() if need_next is None else (np.asarray(need_next, bool) & active)
        cands = [None] * ns if cands is None else cands
        be, bt = self.be, self.batch
        n = self.n + 1
        if n > GK_MAX_CN:
            raise ValueError(f"copy number above {GK_MAX_CN} is not supported by the search kernels")
        active_idx = np.flatnonzero(active)
        self._set_candidates(cands, active)
        self.d_tab = be.upload(self.tab)
        new = 1 - self.cur

        if n == 1:
            be.launch("gk_first_step", bt.d_table, self.d_tab, ns, self.top_n, bt.d_col, self.d_cand,
                      self.d_ids[new], self.d_score, self.d_cnt_out, self.d_flat, self.d_info, self.d_kept)
        else:
            items = self._score_items(active_idx)
            be.zero_(self.d_S)
            d_items = be.upload(items)
            be.launch("gk_score", bt.d_table, self.d_tab, d_items, len(items), bt.d_L, self.d_P, self.d_S,
                      work=float(self._step_cells))
            self.score_cells += self._step_cells
            be.launch("gk_select", bt.d_table, self.d_tab, ns, self.top_n, n - 1, max(bt.max_alleles, 1),
                      self.d_kept, self.d_ids[self.cur], self.d_cand, self.d_S, self.d_flag, self.d_alive,
                      self.d_info)
            info = be.download(self.d_info, None).view(STEP_INFO_DTYPE)     # sync 1: alive counts
            n_alive = np.minimum(info["n_alive"], self.tab["alive_cap"]).astype(np.int64)
            n_alive[~active] = 0
            self.tab["cnt_off"] = _excl_cumsum(n_alive * n * n)
            self.d_tab = be.upload(self.tab)
            d_cnt = be.zeros(int((n_alive * n * n).sum()), np.uint32)
            rows = []
            for s in active_idx:
                f = int(n_alive[s])
                if f == 0:
                    continue
                _, r0, r1 = self._read_chunks(np.array([s]), COUNT_READ_CHUNK, 16)
                f0 = np.arange(0, f, 8)
                ff, ch = np.meshgrid(f0, np.arange(len(r0)), indexing="ij")
                item = np.zeros(ff.size, dtype=COUNT_ITEM_DTYPE)
                item["search"] = s
                item["f0"] = ff.reshape(-1)
                item["r0"] = r0[ch.reshape(-1)]
                item["r1"] = r1[ch.reshape(-1)]
                rows.append(item)
            c_items = np.concatenate(rows) if rows else np.zeros(0, dtype=COUNT_ITEM_DTYPE)
            d_citems = be.upload(c_items)
            be.launch("gk_rescore_count", bt.d_table, self.d_tab, d_citems, len(c_items), self.top_n, n,
                      self.d_info, self.d_ids[self.cur], self.d_cand, self.d_alive, bt.d_LT, d_cnt,
                      work=float((n_alive * self.mt["n_reads"]).sum()) * n)
            be.launch("gk_rank", bt.d_table, self.d_tab, ns, self.top_n, n, self.d_ids[self.cur], self.d_cand,
                      self.d_alive, self.d_S, d_cnt, bt.d_col, self.d_keys, self.d_ids[new], self.d_score,
                      self.d_cnt_out, self.d_flat, self.d_info, self.d_kept)
        self.cur = new
        out = self._collect(active_idx, n)                                    # sync 2: step results
        self.n = n
        nxt = np.flatnonzero(need_next)
        if len(nxt):
            self._write_p(nxt, n)
        return out

    def restore(self, s: int, ids: np.ndarray) -> None:
        """Re-seed search ``s`` with kept sets ``ids`` [K, n] (device state is a pure function
        of the kept ids: P is rebuilt by gk_write_p)."""
        ids = np.asarray(ids, dtype=np.int32)
        k, n = ids.shape
        if k > self.top_n or n > GK_MAX_CN:
            raise ValueError("kept sets do not fit this search")
        host = self.be.download(self.d_ids[self.cur], np.int32).reshape(self.n_search, self.top_n, GK_MAX_CN).copy()
        host[s] = 0
        host[s, :k, :n] = ids
        self.d_ids[self.cur] = self.be.upload(host)
        self.kept[s] = k
        self.d_kept = self.be.upload(self.kept)
        self.n = n
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
        r_pad, r = int(self.mt["r_pad"][s]), int(self.mt["n_reads"][s])
        tab = self.tab[s:s + 1].copy()
        tab["P_off"] = 0
        buf = np.zeros((self.top_n, GK_MAX_CN), dtype=np.int32)
        buf[:k, :n] = ids
        d_ids = be.upload(buf)
        d_tab = be.upload(tab)
        d_kept = be.upload(np.array([k], dtype=np.int32))
        kb, rr = np.meshgrid(np.arange(n_kblk), np.arange(0, r_pad, 128), indexing="ij")
        items = np.zeros(kb.size, dtype=P_ITEM_DTYPE)
        items["k_blk"] = kb.reshape(-1)
        items["r0"] = rr.reshape(-1)
        d_items = be.upload(items)
        d_P = be.empty(n_kblk * r_pad * GK_KB, np.float32)
        be.launch("gk_write_p", bt.d_table, d_tab, d_items, len(items), self.top_n, n, d_kept, d_ids,
                  bt.d_LT, d_P)
        p = be.download(d_P, np.float32).reshape(n_kblk, r_pad, GK_KB)
        return p.transpose(1, 0, 2).reshape(r_pad, n_kblk * GK_KB)[:r, :k].astype(np.int64)


def _excl_cumsum(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    out = np.zeros(len(x), dtype=np.int64)
    if len(x) > 1:
        np.cumsum(x[:-1], out=out[1:])
    return out


def _round_up_arr(x: np.ndarray, m: int) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    return (x + m - 1) // m * m
