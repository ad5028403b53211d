"""
TEST DOUBLE -- a NumPy statement of what each CUDA launcher of libgk_typing.so must
compute on the pooled buffers, item by item.  It lives in tests/ only:

* `-m "not gpu"` tests run the engine's host orchestration (work-item lists, pool
  offsets, step sequencing) against it and compare with the oracle;
* `-m gpu` kernel tests run the same launch on the GPU and on this double and compare
  the output pools byte for byte.

``kir_graph_b200`` never imports it; the product path has no CPU implementation.
"""
from __future__ import annotations

import numpy as np

from kir_graph_b200._cabi import (COUNT_ITEM_DTYPE, EXPAND_ITEM_DTYPE, GK_KB, GK_LIK_READS, GK_MAX_CN, GK_RT, LIK_ITEM_DTYPE,
                                  MATRIX_DTYPE, P_ITEM_DTYPE, SCORE_ITEM_DTYPE, SEARCH_DTYPE, STEP_INFO_DTYPE,
                                  EM_PROBLEM_DTYPE)

LCM = [1, 1, 2, 6, 12, 60, 60, 420, 840]


def popcount32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    out = np.zeros(x.shape, dtype=np.int64)
    for b in range(32):
        out += (x >> np.uint32(b)) & np.uint32(1)
    return out


class FakeBackend:
    def __init__(self):
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.timing = None
        self.log: list[str] = []

    # --- memory ---------------------------------------------------------------
    def zeros(self, n, dtype):
        return np.zeros(max(int(n), 1), dtype=dtype)

    def empty(self, n, dtype):
        # poison so that reads of never-written memory show up in tests
        arr = np.zeros(max(int(n), 1), dtype=dtype)
        arr.view(np.uint8)[:] = 0xA5
        return arr

    def upload(self, array):
        array = np.ascontiguousarray(array)
        if array.size == 0:
            return np.zeros(1, dtype=array.dtype)
        return array.reshape(-1).copy()

    def copy_into(self, tensor, array):
        flat = np.ascontiguousarray(array).reshape(-1)
        if flat.size:
            tensor.view(np.uint8)[: flat.nbytes] = flat.view(np.uint8)

    def download(self, tensor, dtype=None, count=None):
        if count is not None:
            tensor = tensor[:count]
        out = np.array(tensor, copy=True)
        return out.view(dtype) if dtype is not None else out

    def download_async(self, tensors):
        return [np.array(t, copy=True).reshape(-1).view(np.int32) for t in tensors]

    def download_wait(self, handle):
        return handle

    def zero_(self, tensor):
        tensor.view(np.uint8)[:] = 0

    def sync(self):
        pass

    def gather_rows(self, tensor, row_len, rows, cols=None):
        view = tensor.reshape(-1, row_len)[np.asarray(rows, dtype=np.int64)]
        return np.array(view[:, :cols] if cols is not None else view, copy=True)

    def pin(self, array):
        return np.ascontiguousarray(array).reshape(-1)

    def snapshot(self, tensor):
        return tensor.copy()

    def gather_best_device(self, ids, score, info, rows, top_n):
        """(ids row, score) of rank info.best_rank for each search in ``rows``; stays on the 'device'."""
        best = info.view(STEP_INFO_DTYPE)["best_rank"][rows].astype(np.int64)
        flat = np.asarray(rows, dtype=np.int64) * top_n + best
        return ids.reshape(-1, GK_MAX_CN)[flat].copy(), score[flat].copy()

    def launch(self, name, *args, work=0.0):
        self.launches += 1
        self.log.append(name)
        getattr(self, name)(*args)

    # --- kernel (a) --------------------------------------------------------------
    def gk_expand_reads(self, table, items, n_items, hdr, stream, keep, entoff, ent):
        """Wire format -> entry offsets and 16-byte entries (csrc/gk_wire.cu), read by read."""
        table = table.view(MATRIX_DTYPE)
        items = items.view(EXPAND_ITEM_DTYPE)[:n_items]
        hdr, stream, keep = hdr.view(np.uint16), stream.view(np.uint16), keep.view(np.uint32)
        ent = ent.view(np.uint32)
        ent = ent[: len(ent) // 4 * 4].reshape(-1, 4)

        def mate(rec):
            lo, x = int(rec[0]), int(rec[1])
            n, n_out, n_hole = x & 255, (x >> 8) & 15, x >> 12
            nb = (n + 15) // 16
            pos = [lo + i for i in range(n) if (int(rec[2 + i // 16]) >> (i % 16)) & 1]
            holes = [lo + ((int(rec[2 + nb + h // 2]) >> (8 * (h % 2))) & 255) for h in range(n_hole)]
            outs = [int(v) for v in rec[2 + nb + (n_hole + 1) // 2: 2 + nb + (n_hole + 1) // 2 + n_out]]
            words = {}
            for v in range(lo, lo + n):
                w, bit = v >> 5, 1 << (v & 31)
                p, ng = words.setdefault(w, [0, 0])
                if v in pos:
                    words[w][0] |= bit
                elif v not in holes and (int(keep_g[w]) >> (v & 31)) & 1:
                    words[w][1] |= bit
            return words, outs, 2 + nb + (n_hole + 1) // 2 + n_out

        for it in items:
            M = table[it["matrix"]]
            R, stride_bytes = int(M["n_reads"]), int(M["n_ablk"]) * int(M["a_tile"]) * 4
            keep_g = keep[int(it["keep_off"]):]
            at, rec_at = int(it["ent_off"]), int(it["stream_off"])
            eo = entoff[int(M["entoff_off"]):]
            if R == 0 and int(it["r0"]) == 0:
                eo[0] = at
            for r in range(int(it["r0"]), min(int(it["r0"]) + GK_LIK_READS, R)):
                h = int(hdr[int(it["hdr_base"]) + r])
                n_ent, length = h & 255, h >> 8
                eo[r] = at
                out = []
                if length == 0:                                     # raw record: the entries themselves
                    length = 5 * n_ent
                    rec = stream[rec_at: rec_at + length].astype(np.int64)
                    for e in range(n_ent):
                        out.append((int(rec[5 * e]), int(rec[5 * e + 1] | (rec[5 * e + 2] << 16)),
                                    int(rec[5 * e + 3] | (rec[5 * e + 4] << 16))))
                else:
                    rec = stream[rec_at: rec_at + length]
                    lw, louts, used = mate(rec)
                    rw, routs, used2 = mate(rec[used:])
                    assert used + used2 == length, "record length"
                    for w in sorted(set(lw) | set(rw)):
                        lp, ln = lw.get(w, (0, 0))
                        rp, rn = rw.get(w, (0, 0))
                        ov = (lp | ln) & (rp | rn)
                        pp, nn = lp | (rp & ~ov), ln | (rn & ~ov)
                        if pp | nn:
                            out.append((w, pp, nn))
                        if ov:
                            out.append((w, rp & ov, rn & ov))
                    out += [(v >> 5, 1 << (v & 31), 0) for v in louts + routs]
                assert len(out) == n_ent, f"read {r}: {len(out)} entries emitted, header says {n_ent}"
                for w, pp, nn in out:
                    ent[at] = (w * stride_bytes, pp, nn, 1 << (8 * (r & 3)))
                    at += 1
                rec_at += length
                if r == R - 1:
                    eo[R] = at

    def gk_likelihood(self, table, items, n_items, mem, entoff, ent, L, LT, col, half_mode):
        table = table.view(MATRIX_DTYPE)
        items = items.view(LIK_ITEM_DTYPE)[:n_items]
        ent = ent.view(np.uint32)
        ent = ent[: len(ent) // 4 * 4].reshape(-1, 4)
        ent_pos, ent_neg = ent[:, 1], ent[:, 2]
        for it in items:
            M = table[it["matrix"]]
            A, a_tile, rp, R = int(M["n_alleles"]), int(M["a_tile"]), int(M["r_pad"]), int(M["n_reads"])
            r0 = int(it["r0"])
            eo = entoff[M["entoff_off"]: M["entoff_off"] + R + 1]
            stride = int(M["n_ablk"]) * a_tile                 # rows of mem are padded to whole allele blocks
            memv = mem[M["mem_off"]: M["mem_off"] + int(M["n_words"]) * stride].reshape(int(M["n_words"]), stride)
            colsum_only = bool(int(it["flags"]) & 1)           # GK_LIK_COLSUM_ONLY: neither L nor LT is written
            for blk in range(int(it["a_blk"]), min(int(it["a_blk"]) + 4, int(M["n_ablk"]))):
                a0 = blk * a_tile
                a_hi = min(a0 + a_tile, A)
                tile = np.zeros((GK_LIK_READS, a_tile), dtype=np.int64)
                for rl in range(GK_LIK_READS):
                    r = r0 + rl
                    if r >= R:
                        continue
                    for e in range(eo[r], eo[r + 1]):
                        assert int(ent[e, 3]) == 1 << (8 * (r & 3)), "entry tag"
                        mw = memv[int(ent[e, 0]) // (stride * 4), a0:a_hi]
                        x = (np.uint32(ent_pos[e]) & ~mw) | (np.uint32(ent_neg[e]) & mw)
                        tile[rl, : a_hi - a0] += popcount32(x)
                for a in range(a0, a_hi):
                    col[int(M["col_off"]) + a] += np.uint64(tile[:, a - a0].sum())
                if colsum_only:
                    continue
                for rb in range(GK_LIK_READS // GK_RT):      # row-blocked: [r_blk][a_blk][GK_RT][a_tile]
                    base = int(M["L_off"]) + ((r0 // GK_RT + rb) * int(M["n_ablk"]) + blk) * GK_RT * a_tile
                    part = tile[rb * GK_RT:(rb + 1) * GK_RT].reshape(-1)
                    if half_mode:      # the 16-bit pair (m, m) in every 4-byte slot
                        L.view(np.uint16)[2 * base: 2 * (base + GK_RT * a_tile)] = \
                            np.repeat(part.astype(np.uint16), 2)
                    else:
                        L[base: base + GK_RT * a_tile] = part.astype(np.float32)
                for a in range(a0, a_hi):
                    o = int(M["LT_off"]) + a * rp + r0
                    LT[o: o + GK_LIK_READS] = tile[:, a - a0].astype(np.uint8)

    # --- helpers -------------------------------------------------------------------
    @staticmethod
    def _L_view(M, L, half_mode=0):
        nb, rp, tile = int(M["n_ablk"]), int(M["r_pad"]), int(M["a_tile"])
        o = int(M["L_off"])
        flat = L[o: o + nb * rp * tile]
        if half_mode:
            flat = flat.view(np.uint16)[0::2].astype(np.float32)
        return flat.reshape(rp // GK_RT, nb, GK_RT, tile).transpose(1, 0, 2, 3).reshape(nb, rp, tile)

    @staticmethod
    def _P_view(X, M, P):
        """P of one search as a writable view [r_blk][k_blk][GK_RT][GK_KB]."""
        nkb, rp = int(X["n_kblk"]), int(M["r_pad"])
        o = int(X["P_off"])
        return P[o: o + nkb * rp * GK_KB].reshape(rp // GK_RT, nkb, GK_RT, GK_KB)

    @staticmethod
    def _LT_view(M, LT):
        A, rp = int(M["n_alleles"]), int(M["r_pad"])
        o = int(M["LT_off"])
        return LT[o: o + A * rp].reshape(A, rp)

    # --- CN = 1 -----------------------------------------------------------------------
    def gk_first_step(self, table, stab, n_search, top_n, col, cand_pool, ids_out, score_out, cnt_out,
                      flat_out, info, kept):
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        info = info.view(STEP_INFO_DTYPE)
        ids_out = ids_out.reshape(n_search, top_n, GK_MAX_CN)
        score_out = score_out.reshape(n_search, top_n)
        cnt_out = cnt_out[: n_search * top_n].reshape(n_search, top_n, 1)
        flat_out = flat_out.reshape(n_search, top_n)
        for s in range(n_search):
            X = stab[s]
            M = table[X["matrix"]]
            C = int(X["n_cand"])
            cand = cand_pool[X["cand_off"]: X["cand_off"] + C]
            sc = col[int(M["col_off"]) + cand].astype(np.uint64)
            order = np.lexsort((np.arange(C), sc))
            k = min(C, top_n)
            for rank in range(k):
                j = order[rank]
                ids_out[s, rank, 0] = cand[j]
                score_out[s, rank] = sc[j]
                cnt_out[s, rank, 0] = M["n_reads_total"]
                flat_out[s, rank] = j
            flags = 0
            if C > top_n and sc[order[top_n - 1]] == sc[order[top_n]]:
                flags |= 2
            if C > 1 and sc[order[0]] == sc[order[1]]:
                flags |= 4
            bar = sc[order[top_n - 1]] if C >= top_n else 0xFFFFFFFF
            info[s] = (k, C, k, top_n, np.uint32(bar & 0xFFFFFFFF), flags, 0, 0)
            kept[s] = k

    # --- kernel (b) ---------------------------------------------------------------------
    MODE_SPAN = {0: 128, 1: 64, 2: 16, 3: 32, 4: 48, 5: 32, 6: 64, 7: 96, 8: 128}   # F8 F4 S1 S2 S3 | half G1..G4

    def gk_score(self, table, stab, items, n_items, L, P, S, half_mode, flush_stages, kept_count):
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        items = items.view(SCORE_ITEM_DTYPE)[:n_items]
        for it in items:
            if kept_count is not None and int(it["k_blk"]) * GK_KB >= int(kept_count[it["search"]]):
                continue
            X = stab[it["search"]]
            M = table[X["matrix"]]
            rp, tile = int(M["r_pad"]), int(M["a_tile"])
            r0, r1 = int(it["r0"]), int(it["r1"])
            shape = int(it["shape"])
            assert flush_stages >= 1
            assert P.dtype == (np.uint16 if half_mode else np.float32)
            if shape & (1 << 16):      # warp-split tile: WK warps x G' groups of 8 rows, 8 TA' columns
                assert half_mode
                gp, wk, ta8 = shape & 0xF, 1 << ((shape >> 4) & 0xF), (shape >> 8) & 0xFF
                assert 1 <= gp <= 4 and wk in (1, 2, 4) and 1 <= ta8 <= 8
                kspan, aspan = 8 * gp * wk, 8 * ta8
                koff = ((shape >> 20) & 7) * 8
            else:
                koff = 0
                assert (shape & 0xFF >= 5) == bool(half_mode)
                kspan = self.MODE_SPAN[shape & 0xFF]
                aspan = self.MODE_SPAN[(shape >> 8) & 0xFF]
            assert tile == 32 and r0 % GK_RT == 0 and r1 % GK_RT == 0 and r1 <= rp and r1 > r0
            kw, aw = -(-(koff + kspan) // GK_KB), -(-aspan // tile)
            assert koff + kspan <= 2 * GK_KB
            assert int(it["a_blk"]) + aw <= int(M["n_ablk"])
            stride = int(X["s_stride"])
            assert int(it["k_blk"]) + kw <= int(X["n_kblk"])
            Pv = self._P_view(X, M, P)[r0 // GK_RT: r1 // GK_RT]
            Pt = np.concatenate([Pv[:, kb].reshape(r1 - r0, GK_KB)
                                 for kb in range(int(it["k_blk"]), int(it["k_blk"]) + kw)],
                                axis=1)[:, koff:koff + kspan].astype(np.float32)
            Lt = np.concatenate([self._L_view(M, L, half_mode)[ab, r0:r1, :]
                                 for ab in range(int(it["a_blk"]), int(it["a_blk"]) + aw)], axis=1)[:, :aspan]
            if half_mode:      # a 16-bit lane sums `flush_stages` stages of min(L, P) <= max L
                assert flush_stages * GK_RT * int(Lt.max(initial=0)) <= 65535
            if half_mode:      # packed integer path accumulates the min-sum itself
                part = np.minimum(Lt[:, None, :], Pt[:, :, None]).sum(axis=0)
            else:              # FP32 path accumulates the sum of absolute differences
                part = np.abs(Lt[:, None, :] - Pt[:, :, None]).sum(axis=0)      # [kspan, aspan]
            assert part.max(initial=0) < 2 ** 24
            for kl in range(kspan):
                o = int(X["S_off"]) + (int(it["k_blk"]) * GK_KB + koff + kl) * stride + int(it["a_blk"]) * tile
                S[o: o + aspan] += part[kl].astype(np.uint32)

    # --- kernel (c), part 1 ----------------------------------------------------------------
    @staticmethod
    def _min_sum(S, col, sprev, X, M, k, a, direct):
        d = int(S[int(X["S_off"]) + k * int(X["s_stride"]) + a])
        return d if direct else (int(sprev[k]) + int(col[int(M["col_off"]) + a]) - d) // 2

    def gk_select(self, table, stab, n_search, top_n, n_prev, max_alleles, max_cand, kept, ids_prev, cand_pool,
                  S, col, score_prev, flag, alive, info, direct):
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        info = info.view(STEP_INFO_DTYPE)
        ids_prev = ids_prev.reshape(n_search, top_n, GK_MAX_CN)
        for s in range(n_search):
            X = stab[s]
            M = table[X["matrix"]]
            sprev = score_prev[s * top_n: (s + 1) * top_n]
            K, C = int(kept[s]), int(X["n_cand"])
            N = K * C
            cand = cand_pool[X["cand_off"]: X["cand_off"] + C]
            stride = int(X["s_stride"])
            seen = {}
            uniq = np.zeros(N, dtype=bool)
            score = np.zeros(N, dtype=np.uint32)
            for i in range(N):
                k, j = divmod(i, C)
                key = tuple(sorted(list(ids_prev[s, k, :n_prev]) + [cand[j]]))
                if key not in seen:
                    seen[key] = i
                    uniq[i] = True
                score[i] = self._min_sum(S, col, sprev, X, M, k, cand[j], direct)
            assert C <= max_cand and flag.dtype == np.uint32
            flag[X["flag_off"]: X["flag_off"] + N] = np.where(uniq, score, np.uint32(0xFFFFFFFF))
            n_unique = int(uniq.sum())
            cut = max(top_n, n_unique // 5)
            us = np.flatnonzero(uniq)
            if n_unique > top_n:
                bar = np.sort(score[us])[top_n - 1]
                c_less = int((score[us] < bar).sum())
                c_eq = int((score[us] == bar).sum())
            else:
                bar, c_less, c_eq = np.uint32(0xFFFFFFFF), n_unique, 0
            take_eq = max(0, min(cut - c_less, c_eq))
            out, taken = [], 0
            for i in us:
                if score[i] < bar:
                    out.append(i)
                elif score[i] == bar and taken < take_eq:
                    out.append(i)
                    taken += 1
            out = out[: int(X["alive_cap"])]
            alive[X["alive_off"]: X["alive_off"] + len(out)] = out
            info[s] = (0, n_unique, len(out), cut, bar, 1 if c_eq > take_eq else 0, 0, 0)

    # --- rescoring ------------------------------------------------------------------------
    def _alive_sets(self, X, s, n, top_n, ids_prev, cand_pool, alive, n_alive):
        C = int(X["n_cand"])
        cand = cand_pool[X["cand_off"]: X["cand_off"] + C]
        flat = alive[X["alive_off"]: X["alive_off"] + n_alive]
        k, j = flat // max(C, 1), flat % max(C, 1)
        ids = np.concatenate([ids_prev[s, k, : n - 1], cand[j][:, None]], axis=1) if n_alive else np.zeros((0, n), int)
        return flat, ids

    def gk_rescore_count(self, table, stab, items, n_items, top_n, n_set, info, ids_prev, cand_pool, alive,
                         LT, cnt):
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        info = info.view(STEP_INFO_DTYPE)
        items = items.view(COUNT_ITEM_DTYPE)[:n_items]
        ids_prev = ids_prev.reshape(len(stab), top_n, GK_MAX_CN)
        n = n_set
        for it in items:
            s = int(it["search"])
            X = stab[s]
            M = table[X["matrix"]]
            n_alive = min(int(info[s]["n_alive"]), int(X["alive_cap"]))
            _, ids = self._alive_sets(X, s, n, top_n, ids_prev, cand_pool, alive, n_alive)
            r0, r1 = int(it["r0"]), min(int(it["r1"]), int(M["n_reads"]))
            assert int(it["r0"]) % 16 == 0 and int(it["r1"]) % 16 == 0
            if r1 <= r0:
                continue
            m = self._LT_view(M, LT)
            for f in range(int(it["f0"]), min(int(it["f0"]) + 8, n_alive)):
                g = m[ids[f], r0:r1].astype(np.int64)              # [n, r]
                mn = g.min(axis=0)
                eq = g == mn[None, :]
                q = eq.sum(axis=0)
                for t in range(n):
                    for qq in range(1, n + 1):
                        cnt[int(X["cnt_off"]) + (f * n + t) * n + qq - 1] += np.uint32((eq[t] & (q == qq)).sum())

    # --- kernel (c), part 2 -----------------------------------------------------------------
    def gk_rank(self, table, stab, n_search, top_n, n_set, ids_prev, cand_pool, alive, S, cnt, col, score_prev,
                keys, ids_out, score_out, cnt_out, flat_out, info, kept_out, direct):
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        info = info.view(STEP_INFO_DTYPE)
        ids_prev = ids_prev.reshape(n_search, top_n, GK_MAX_CN)
        ids_out = ids_out.reshape(n_search, top_n, GK_MAX_CN)
        score_out = score_out.reshape(n_search, top_n)
        cnt_out = cnt_out[: n_search * top_n * n_set * n_set].reshape(n_search, top_n, n_set * n_set)
        flat_out = flat_out.reshape(n_search, top_n)
        n = n_set
        for s in range(n_search):
            X = stab[s]
            M = table[X["matrix"]]
            F = min(int(info[s]["n_alive"]), int(X["alive_cap"]))
            flat, ids = self._alive_sets(X, s, n, top_n, ids_prev, cand_pool, alive, F)
            C = int(X["n_cand"])
            stride = int(X["s_stride"])
            colv = col[int(M["col_off"]): int(M["col_off"]) + int(M["n_alleles"])].astype(np.int64)
            cn = cnt[int(X["cnt_off"]): int(X["cnt_off"]) + F * n * n].astype(np.int64).reshape(F, n, n)
            w = np.array([LCM[n] // q for q in range(1, n + 1)], dtype=np.int64)
            num = (cn * w[None, None, :]).sum(axis=2)
            even = int(M["n_reads_total"]) * LCM[n] // n
            uneven = np.abs(num - even).sum(axis=1)
            sprev = score_prev[s * top_n: (s + 1) * top_n]
            sc = np.array([self._min_sum(S, col, sprev, X, M, i // C, ids[f, -1], direct) for f, i in enumerate(flat)],
                          dtype=np.int64)
            cs = colv[ids].sum(axis=1) if F else np.zeros(0, np.int64)
            order = np.lexsort((np.arange(F), uneven, cs, sc))
            k = min(F, top_n)
            for rank in range(k):
                f = order[rank]
                ids_out[s, rank, :n] = ids[f]
                score_out[s, rank] = sc[f]
                cnt_out[s, rank, :] = cn[f].reshape(-1)
                flat_out[s, rank] = flat[f]
            flags = int(info[s]["tie_flags"])
            if F > top_n and sc[order[top_n - 1]] == sc[order[top_n]]:
                flags |= 2
            if F > 1 and sc[order[0]] == sc[order[1]]:
                flags |= 4
            best = 0
            for rank in range(k):
                if np.all(2 * n * num[order[rank]] >= int(M["n_reads_total"]) * LCM[n]):
                    best = rank
                    break
            # bit3: selectBest's verdict on a rank it looks at hinges on how tied reads are counted
            passing = [rank for rank in range(k) if np.all(2 * n * num[order[rank]] >= int(M["n_reads_total"]) * LCM[n])]
            last = passing[0] + 1 if passing else k
            reads = int(M["n_reads_total"])
            for rank in range(last):
                c = cn[order[rank]]
                if np.any((2 * n * c[:, 0] < reads) & (2 * n * c.sum(axis=1) >= reads)):
                    flags |= 8
            info[s]["n_kept"] = k
            info[s]["best_rank"] = best
            info[s]["tie_flags"] = flags
            kept_out[s] = k

    def gk_write_p(self, table, stab, items, n_items, top_n, n_set, kept, ids, LT, P, half_mode):
        assert P.dtype == (np.uint16 if half_mode else np.float32)
        table = table.view(MATRIX_DTYPE)
        stab = stab.view(SEARCH_DTYPE)
        items = items.view(P_ITEM_DTYPE)[:n_items]
        ids = ids.reshape(-1, top_n, GK_MAX_CN)
        for it in items:
            s = int(it["search"])
            X = stab[s]
            M = table[X["matrix"]]
            rp = int(M["r_pad"])
            K = int(kept[s])
            m = self._LT_view(M, LT)
            assert int(it["r0"]) % 128 == 0 and int(it["r1"]) % 128 == 0 and int(it["r1"]) <= rp
            for r0 in range(int(it["r0"]), int(it["r1"]), 128):
                tile = np.zeros((128, GK_KB), dtype=np.float32)
                for kl in range(GK_KB):
                    k = int(it["k_blk"]) * GK_KB + kl
                    if k < K:
                        tile[:, kl] = m[ids[s, k, :n_set], r0:r0 + 128].min(axis=0)
                self._P_view(X, M, P)[r0 // GK_RT: (r0 + 128) // GK_RT, int(it["k_blk"])] = \
                    tile.reshape(128 // GK_RT, GK_RT, GK_KB)

    # --- read grouping (section 8f rank 3) ---------------------------------------------------------
    def gk_cn_fit(self, x, density, bases, n_base, bin_num, max_cn, start_base, base_dev, y0_dev, dev_decay,
                  dev_decay_neg, space, likelihood, prob_out):
        """likelihood[b] = sum_i log(max_n pdf_n(x_i; bases[b]) * space + 1e-9) * density_i, element by element."""
        with np.errstate(all="ignore"):
            for b in range(n_base):
                rows = []
                for n in range(max_cn):
                    if start_base == 1:
                        loc, scale = (0.0, base_dev * y0_dev) if n == 0 else (bases[b] * n, base_dev * (dev_decay * (n - 1) + 1))
                    else:
                        loc = bases[b] * n
                        scale = base_dev * (dev_decay_neg * (start_base - n) + 1) if n < start_base \
                            else base_dev * (dev_decay * (n - start_base) + 1)
                    if scale > 0:
                        y = (x[:bin_num] - loc) / scale
                        rows.append(np.exp(-(y * y) / 2.0) / 2.5066282746310002 / scale * space)
                    else:
                        rows.append(np.full(bin_num, np.nan))
                rows = np.array(rows)
                if prob_out is not None:
                    prob_out[b * max_cn * bin_num: (b + 1) * max_cn * bin_num] = rows.reshape(-1)
                terms = np.log(rows.max(axis=0) + 1e-9) * density[:bin_num]
                total = 0.0
                for t in terms:                     # plain left-to-right sum (the kernel's order differs: tests compare with a tolerance)
                    total += t
                likelihood[b] = total

    def gk_group_reads(self, table, matrix, n_reads, ids, n_ids, LT, pattern):
        M = table.view(MATRIX_DTYPE)[matrix]
        m = self._LT_view(M, LT)[np.asarray(ids[:n_ids], dtype=np.int64), :n_reads].astype(np.int64)
        is_min = m == m.min(axis=0, keepdims=True)
        pattern.view(np.uint32)[:n_reads] = (is_min.astype(np.uint64) << np.arange(n_ids, dtype=np.uint64)[:, None]) \
            .sum(axis=0).astype(np.uint32)

    # --- EM path -------------------------------------------------------------------------------
    def gk_em_compat(self, membT, n_aw, n_alleles, off_lp, idx_lp, off_ln, idx_ln, off_rp, idx_rp, off_rn,
                     idx_rn, n_reads, compat):
        membT = membT.reshape(-1, n_aw)
        out = compat.reshape(-1, n_aw)
        full = np.full(n_aw, 0xFFFFFFFF, dtype=np.uint32)
        if n_alleles & 31:
            full[-1] = np.uint32((1 << (n_alleles & 31)) - 1)
        for r in range(n_reads):
            mates = []
            for (op, ip, on, in_) in ((off_lp, idx_lp, off_ln, idx_ln), (off_rp, idx_rp, off_rn, idx_rn)):
                acc = np.zeros(n_aw, dtype=np.uint32)
                if op[r + 1] > op[r]:
                    acc = full.copy()
                    for e in range(op[r], op[r + 1]):
                        acc &= membT[ip[e]]
                    for e in range(on[r], on[r + 1]):
                        acc &= ~membT[in_[e]]
                mates.append(acc)
            both = mates[0] & mates[1]
            out[r] = both if both.any() else (mates[0] | mates[1])

    def gk_em_squarem(self, problems, n_problems, row_pool, wgt_pool, len_pool, out_pool, iters_out, iter_max,
                      diff_threshold):
        problems = problems.view(EM_PROBLEM_DTYPE)
        for i in range(n_problems):
            E = problems[i]
            U, A, W = int(E["n_rows"]), int(E["n_alleles"]), int(E["n_awords"])
            rows = row_pool[E["row_off"]: E["row_off"] + U * W].reshape(U, W)
            wgt = wgt_pool[E["wgt_off"]: E["wgt_off"] + U].astype(np.float64)
            length = len_pool[E["len_off"]: E["len_off"] + A]
            bits = np.zeros((U, A))
            for a in range(A):
                bits[:, a] = (rows[:, a >> 5] >> np.uint32(a & 31)) & 1

            def step(p):
                b = (bits * p).sum(axis=1)
                binv = np.divide(wgt, b, out=np.zeros(U), where=b != 0)
                acc = (bits * binv[:, None]).sum(axis=0) * p / length
                return acc / acc.sum()

            p = step(np.ones(A))
            it = 0
            while it < iter_max:
                p1 = step(p)
                p2 = step(p1)
                r = p1 - p
                v = p2 - p1 - r
                rr, vv = (r * r).sum(), (v * v).sum()
                if vv > 0.0:
                    g = -np.sqrt(rr / vv)
                    p1 = step(np.maximum(p - r * g * 2 + v * (g * g), 0))
                if np.abs(p - p1).sum() <= diff_threshold:
                    break
                p = p1
                it += 1
            out_pool[E["out_off"]: E["out_off"] + A] = p
            iters_out[i] = it
