"""
Batched typing of many (sample, gene) problems on one GPU, and sharding of a
cohort over the GPUs of one box.

The reference types samples and genes one after the other in Python
(graphkir/main.py:171-220 -> kir_typing.py:42-62, :103-132).  Here every problem
of a batch advances through the copy-number steps together, so each step is a
handful of grouped kernel launches over all problems; samples are independent,
so a cohort shards across ranks with no communication on the data path (one
process per GPU; rank r takes samples r, r + world, ...).

``BatchTyper``   one resident batch: ``start()`` enqueues a whole pass (from the third
                 pass of the same batch on: one CUDA-graph replay), ``finish()`` reads
                 back once and forms the calls.
``CohortTyper``  sub-batches on their own streams, all started before the first is
                 finished, so that host->device copies and the host-side call phase
                 of one overlap the kernels of the others.
``PassPipeline`` consecutive passes with two (or more) in flight over replicas of a
                 typer: double buffering of whole passes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import engine
from .packing import GenePack, site_tallies
from .typing_mulit_allele import C_HIT, C_MISS, _lcm_upto, _no_hetero_site, isHetrozygous


@dataclass(slots=True)
class GeneCall:
    """Typing outcome of one gene problem."""

    gene: str
    cn: int
    alleles: list[str]            # called allele names (len == cn); "fail" entries when no reads
    n_reads: int
    homozygous: bool
    best_rank: int = 0
    value: float = float("nan")   # log10 likelihood of the called set
    tie_flags: int = 0
    score: int = 0                # integer mismatch score of the called set
    ids: list[int] = field(default_factory=list)


def decide_homozygous(pack: GenePack, cn: int) -> bool:
    """force_homo rule of kir_typing.py:106 + isHomozygous (typing_mulit_allele.py:807-857)."""
    if isHetrozygous(pack.gene):
        return False
    if cn <= 1:
        return False
    return _no_hetero_site(site_tallies(pack), cn)


def _site_summary(p: GenePack):
    """Per position of one problem, everything of isHomozygous (typing_mulit_allele.py:807-857) that does
    not depend on the copy number: ``(second, candidate, broken)`` with ``second`` the share of the
    runner-up allele value at the site and ``candidate`` whether the site is looked at (more than one
    value, a positive one among them, depth >= 20, two values passing count > 3 and share > 0.1).  A
    function of the packed observation counts alone, so it is computed once per pack and cached."""
    cached = getattr(p, "_site_summary", None)
    if cached is not None:
        return cached
    keep = ~p.var_is_del
    _, site_id = np.unique(p.var_pos, return_inverse=True)
    _, val_code = np.unique(np.array(p.var_val, dtype=object).astype(str), return_inverse=True)
    site, key, pos, neg = site_id[keep], val_code[keep], p.obs_pos[keep], p.obs_neg[keep]
    n_site = int(site_id.max()) + 1 if len(site_id) else 0
    # one entry per (site, value, polarity) with a non-zero count; equal keys at a site add up
    ent_site = np.concatenate([site, site])
    ent_key = np.concatenate([key * 2, key * 2 + 1])
    ent_cnt = np.concatenate([pos, neg])
    nz = ent_cnt > 0
    ent_site, ent_key, ent_cnt = ent_site[nz], ent_key[nz], ent_cnt[nz]
    combo = ent_site * (int(ent_key.max(initial=0)) + 2) + ent_key
    uniq, inv = np.unique(combo, return_inverse=True)
    e_cnt = np.bincount(inv, weights=ent_cnt, minlength=len(uniq)).astype(np.int64)
    first = np.zeros(len(uniq), dtype=np.int64)
    first[inv] = np.arange(len(inv))
    e_site = ent_site[first]
    e_negative = (ent_key[first] % 2) == 1
    n_keys = np.bincount(e_site, minlength=n_site)
    any_pos = np.bincount(e_site, weights=(~e_negative).astype(np.float64), minlength=n_site) > 0
    big = e_cnt > 3                                                  # drop low coverage (:844)
    depth = np.bincount(e_site[big], weights=e_cnt[big], minlength=n_site).astype(np.int64)
    share = e_cnt / np.maximum(depth[e_site], 1)
    strong = big & (share > 0.1)                                     # (:850)
    n_strong = np.bincount(e_site[strong], minlength=n_site)
    order = np.lexsort((-e_cnt, e_site))
    s_sorted, strong_sorted, share_sorted = e_site[order], strong[order], share[order]
    s_strong, sh_strong = s_sorted[strong_sorted], share_sorted[strong_sorted]
    start = np.searchsorted(s_strong, np.arange(n_site), side="left")
    has2 = n_strong >= 2
    second = np.zeros(n_site)                                        # runner-up share per site
    second[has2] = sh_strong[start[has2] + 1]
    considered = (n_keys > 1) & any_pos & (depth >= 20)
    candidate = considered & has2
    # the fourth element is all a batch needs: the runner-up shares of the sites that are looked at (about a
    # tenth of the sites of a 30x sample)
    out = (second, candidate, bool(np.any(considered & (n_strong == 0))), second[candidate])
    p._site_summary = out
    return out


def undecidable_homozygosity(pack: GenePack, cn: int) -> bool:
    """isHomozygous would index an empty list for this problem (a site that is looked at without any value
    passing the share filter, typing_mulit_allele.py:850-853): copy number above 1, not heterozygous by name."""
    return bool(cn > 1 and not isHetrozygous(pack.gene) and _site_summary(pack)[2])


class HomozygosityIndex:
    """isHomozygous (typing_mulit_allele.py:807-857) for a whole batch at once.

    The per-site summaries are a property of each pack (``_site_summary``, cached on it: the per-variant
    observation counts are part of the packed input); a batch concatenates them and ``decide`` is pure
    array arithmetic over all sites of all problems."""

    def __init__(self, packs: list[GenePack]):
        parts = [_site_summary(p) for p in packs]
        # only the sites isHomozygous looks at take part (``second`` of the candidate sites of every problem)
        shares = [x[3] for x in parts]
        self.second = np.concatenate(shares) if shares else np.zeros(0, np.float64)
        self.broken = np.array([x[2] for x in parts], dtype=bool)
        self.genes = [p.gene for p in packs]
        self.site_owner = np.repeat(np.arange(len(packs), dtype=np.int64), [len(x) for x in shares])
        self.n_pack = len(packs)
        self.forced_hetero = np.array([isHetrozygous(p.gene) for p in packs], dtype=bool)

    def decide(self, cns: np.ndarray) -> np.ndarray:
        # a site that is looked at without any value passing the share filter: the reference indexes an empty
        # list there (:853) - for the problems it runs isHomozygous on, i.e. copy number > 1 and not one of the
        # always-heterozygous genes
        failing = self.broken & (cns > 1) & ~self.forced_hetero
        if failing.any():
            raise IndexError("list index out of range (isHomozygous: " +
                             ", ".join(self.genes[i] for i in np.flatnonzero(failing)[:5]) + ")")
        threshold = 1 / (np.maximum(cns, 1) * 2)
        hetero_site = self.second > threshold[self.site_owner]
        hits = np.bincount(self.site_owner[hetero_site], minlength=self.n_pack)
        return (hits == 0) & (cns > 1) & ~self.forced_hetero


def _select_best(frac_num: np.ndarray, n_reads: int, n: int) -> int:
    """First rank whose every fraction >= 1/(2n), else 0 (TypingResult.selectBest)."""
    fraction = frac_num / float(n_reads * _lcm_upto(n))
    ok = np.all(fraction >= (1 / n) / 2, axis=1)
    hits = np.flatnonzero(ok)
    return int(hits[0]) if len(hits) else 0


class BatchTyper:
    """Types a fixed batch of gene problems; device buffers are reused across ``run`` calls."""

    def __init__(self, packs: list[GenePack], cns: list[int], top_n: int = 300, backend=None,
                 host_batch: engine.HostBatch | None = None, col_shard: tuple[int, int] | None = None,
                 reduce_scores=None, read_shard: bool = False):
        """``col_shard`` / ``read_shard`` + ``reduce_scores``: one very deep sample spread over several
        ranks, by candidate columns or by reads (see :class:`engine.SearchGroup`; with ``read_shard``
        the packs are this rank's ``packing.shard_reads`` parts)."""
        self.be = backend if backend is not None else engine.default_backend()
        self.col_shard, self.reduce_scores, self.read_shard = col_shard, reduce_scores, bool(read_shard)
        self.packs = packs
        self.cns = np.asarray(cns, dtype=np.int64)
        self.top_n = top_n
        self.host = host_batch if host_batch is not None else engine.HostBatch(packs)
        if len(self.host.packs) != len(packs) or any(a is not b for a, b in zip(self.host.packs, packs)):
            raise ValueError("host_batch was packed from other problems than the ones to type")
        # Functions of the packed input alone live on the host batch (like its pools): a typer that is built
        # per pass from prepared host batches (a cohort stream) does not redo them.
        # (``host_batch`` holds the same problems: checked above.)
        static = getattr(self.host, "_typer_static", None)
        if static is None or len(static[0]) != len(packs):
            n_reads = np.array([p.n_reads if p.n_reads_total is None else p.n_reads_total for p in packs],
                               dtype=np.int64)       # reads of the whole problem (a read shard may hold none)
            typable = (n_reads > 0) & np.array([p.n_alleles > 0 for p in packs], dtype=bool)
            static = (packs, HomozygosityIndex(packs), n_reads, typable,
                      [(p.gene, p.allele_names, r) for p, r in zip(packs, n_reads.tolist())])
            self.host._typer_static = static
        _, self.homo_index, self.n_reads, typable, per_problem = static
        self.live = np.flatnonzero((self.cns > 0) & typable)
        # per problem: what the call phase needs without attribute look-ups in its loop
        self._static = [(gene, names, r, c) for (gene, names, r), c in zip(per_problem, self.cns.tolist())]
        self.batch: engine.MatrixBatch | None = None
        self.group: engine.SearchGroup | None = None
        self.score_cells = 0
        self.homo = np.zeros(len(packs), dtype=bool)
        self._homo_cache: np.ndarray | None = None
        self.pipelined = True        # enqueue all steps up front (one read-back); False = read back per step
        self._pending = None
        # A batch that is typed repeatedly replays its launches as one CUDA graph from the third
        # pass on (the work-item tables are cached by then, so nothing is uploaded inside the graph).
        self.use_graph = hasattr(self.be, "torch") and os.environ.get("GK_GRAPH", "1") != "0"
        self._graph = None           # (CUDAGraph, pending core, group state, launches)
        self._graph_runs = 0

    def upload(self) -> None:
        """Host pools -> device (the end-to-end path times this; the resident path does it once)."""
        if self.batch is not None and self.batch.host is self.host:
            self.batch.reload()                # same layout: device buffers (and a recorded graph) are kept
            return
        self.batch = engine.MatrixBatch(self.host, backend=self.be, run=False,
                                        reduce=self.reduce_scores if self.read_shard else None)
        self._graph, self._graph_runs = None, 0          # a captured graph names the old buffers
        if self.group is not None:
            self.group.batch = self.batch      # same tables and offsets: search buffers are reused

    def upload_and_run(self) -> list[GeneCall]:
        self.upload()
        return self.run()

    def upload_and_start(self):
        self.upload()
        return self.start()

    def run(self) -> list[GeneCall]:
        """Likelihood build + greedy search + calls for every problem of the batch."""
        self.start()
        return self.finish()

    def _colsum_only(self) -> np.ndarray:
        """Problems whose likelihood matrix is never read: not typed at all, or typed with one step
        (CN 1, homozygous shortcut) - the first step needs the column sums only."""
        if self._homo_cache is None:           # the copy numbers are fixed for this batch
            self._homo_cache = self.homo_index.decide(self.cns) & (self.n_reads > 0)
        only = np.ones(len(self.packs), dtype=bool)
        only[self.live] = np.where(self._homo_cache[self.live], 1, self.cns[self.live]) <= 1
        return only

    def _start_eager(self) -> None:
        batch = self.batch
        batch.run_likelihood(self._colsum_only())
        if self.group is None:
            self.group = engine.SearchGroup(batch, self.live, self.top_n, col_shard=self.col_shard,
                                            reduce_scores=self.reduce_scores, read_shard=self.read_shard)
        else:
            self.group.reset()
        if self._homo_cache is None:           # the copy numbers are fixed for this batch
            self._homo_cache = self.homo_index.decide(self.cns) & (self.n_reads > 0)
        self.homo = self._homo_cache
        steps = np.where(self.homo[self.live], 1, self.cns[self.live])
        self._pending = None
        if self.pipelined:
            self._pending = self.group.run_pipeline_start(steps)

    def start(self):
        """Enqueue the likelihood build and (pipelined mode) every search launch on the current
        stream without waiting for the device; ``finish`` reads back and forms the calls.  Returns
        the token of this pass (``finish(token)``): a second pass may be started before the first
        is finished - the launches of both are ordered on the stream and every pass has its own
        page-locked read-back buffer."""
        self._start()
        return self._pending

    def _start(self) -> None:
        if self.batch is None:
            self.upload()
        graphable = (self.use_graph and self.pipelined and self.col_shard is None and not self.read_shard
                     and self.group is not None and self.be.timing is None)
        if not graphable:
            self._graph_runs += 1
            self._start_eager()
            return
        if self._graph is None and self._graph_runs >= 2:
            self._capture()
        if self._graph is None:
            self._graph_runs += 1
            self._start_eager()
            return
        graph, core, flat, sizes, state, launches = self._graph
        graph.replay()
        self.be.launches += launches
        self.batch._colsum_host = None
        group = self.group
        group.reset()
        group.cur, group.n = state
        self._pending = core + (self.be.download_async(flat, sizes),)

    def _capture(self) -> None:
        """Record one pass (likelihood build + every search launch) into a CUDA graph."""
        torch = self.be.torch
        steps = np.where(self.homo[self.live], 1, self.cns[self.live])
        launches0 = self.be.launches
        try:
            torch.cuda.current_stream(self.be.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            self.be.capturing = True                  # an upload inside the graph would replay stale bytes
            with torch.cuda.graph(graph):
                self.batch.run_likelihood(self._colsum_only())
                core = self.group.run_pipeline_enqueue(steps)
                flat = torch.cat([t.reshape(-1) for t in core[-1]])
            self._graph = (graph, core[:-1], flat, [t.numel() for t in core[-1]],
                           (self.group.cur, self.group.n), self.be.launches - launches0)
        except Exception as exc:                      # capture is an optimisation: fall back to eager launches
            self.use_graph = False
            self._graph = None
            self.graph_error = repr(exc)
            torch.cuda.synchronize(self.be.device)
        finally:
            self.be.capturing = False
        self.be.launches = launches0

    def finish(self, pending=None) -> list[GeneCall]:
        """Read back and form the calls of the pass ``pending`` (default: the last one started)."""
        if pending is None:
            pending = self._pending
        group = self.group
        cn_live = self.cns[self.live]
        homo_live = self.homo[self.live]
        steps = np.where(homo_live, 1, cn_live)

        # per search: (best rank, ids of the called set, score, tie flags)
        n_live = len(self.live)
        best = np.zeros(n_live, dtype=np.int64)
        score = np.zeros(n_live, dtype=np.int64)
        flags = np.zeros(n_live, dtype=np.int64)
        called = np.full((n_live, max(int(cn_live.max(initial=1)), 1)), -1, dtype=np.int64)
        piped = group.run_pipeline_finish(pending) if pending is not None else None
        if pending is self._pending:
            self._pending = None
        if piped is not None:
            ids, score, info = piped
            kept = info["n_kept"].astype(np.int64)
            best = np.where(homo_live, 0, info["best_rank"]).astype(np.int64)
            flags = info["tie_flags"].astype(np.int64)
            ids = ids.copy()
            ids[kept == 0] = -1
            called[:, : ids.shape[1]] = ids
            called[homo_live, 1:] = called[homo_live, :1]                    # homozygous shortcut (:423-454)
        else:
            group.reset()
            sticky = np.zeros(n_live, dtype=np.int64)          # tie flags accumulate over the steps of a search
            for step in range(1, int(steps.max(initial=0)) + 1):
                out = group.step(active=steps >= step, need_next=steps > step, collect=steps == step,
                                 best_only=True)
                sticky |= np.where(steps >= step, out.info["tie_flags"].astype(np.int64), 0)
                rows = out.searches
                if not len(rows):
                    continue
                n = out.n
                kept = out.info["n_kept"][rows].astype(np.int64)
                best[rows] = np.where(homo_live[rows], 0, out.info["best_rank"][rows])
                score[rows] = out.score
                flags[rows] = sticky[rows]
                ids = out.ids.astype(np.int64)
                ids[kept == 0] = -1
                if n == 1:                                                   # homozygous shortcut (:423-454)
                    called[rows] = ids[:, :1]
                else:
                    called[rows, :n] = ids
        self.score_cells = group.score_cells

        where = np.full(len(self.packs), -1, dtype=np.int64)
        where[self.live] = np.arange(n_live)
        mult = np.where(self.homo, self.cns, 1)
        sc_all = np.zeros(len(self.packs), dtype=np.int64)
        sc_all[self.live] = score
        value = (self.host.k_total * C_HIT + sc_all * (C_MISS - C_HIT)) * mult
        sc_all = sc_all * mult
        calls: list[GeneCall] = []
        append = calls.append
        called_l, best_l, flags_l = called.tolist(), best.tolist(), flags.tolist()
        value_l, sc_l, homo_l, where_l = value.tolist(), sc_all.tolist(), self.homo.tolist(), where.tolist()
        for i, (gene, names, n_reads, cn) in enumerate(self._static):
            if cn == 0:
                continue
            s = where_l[i]
            if s < 0 or called_l[s][0] < 0:
                append(GeneCall(gene, cn, ["fail"] * cn, n_reads, False))
                continue
            ids = called_l[s][:cn]
            append(GeneCall(gene, cn, [names[a] for a in ids], n_reads, homo_l[i],
                            best_l[s], value_l[i], flags_l[s], sc_l[i], ids))
        return calls


class CohortTyper:
    """Splits a batch into ``n_parts`` sub-batches, each on its own CUDA stream.  One host thread
    enqueues every part (uploads, likelihood build, all search steps) before it blocks on the
    first read-back, so the host->device copies and the host-side call phase of one part overlap
    the kernels of the others.  Results come back in input order."""

    def __init__(self, packs: list[GenePack], cns: list[int], top_n: int = 300, backend=None,
                 n_parts: int = 2, group_size: int = 1, col_shard: tuple[int, int] | None = None,
                 reduce_scores=None, own_stream: bool = False, host_batches: list | None = None,
                 read_shard: bool = False, streams: list | None = None):
        """``group_size`` consecutive problems (e.g. the 17 genes of a sample) stay in one part.
        ``own_stream``: a stream of its own even for a single part (replicas of a ``PassPipeline``
        overlap on the device only if they do not share the current stream).  ``host_batches``: the
        packed host pools of another typer over the same problems and parts (shared, read only)."""
        self.be = backend if backend is not None else engine.default_backend()
        sharded = read_shard or (col_shard is not None and col_shard[1] > 1)
        if sharded:
            n_parts = 1                          # one collective stream: keep the parts serial
        n = len(packs)
        n_groups = max(1, -(-n // group_size))
        n_parts = max(1, min(n_parts, n_groups))
        bounds = [(g * n_groups // n_parts) * group_size for g in range(n_parts)] + [n]
        self.slices = [slice(bounds[i], min(bounds[i + 1], n)) for i in range(n_parts)]
        if host_batches is not None and len(host_batches) != n_parts:
            raise ValueError("host_batches must hold one HostBatch per part")
        self.parts = [BatchTyper(packs[sl], list(cns)[sl], top_n=top_n, backend=self.be, col_shard=col_shard,
                                 reduce_scores=reduce_scores, read_shard=read_shard,
                                 host_batch=host_batches[i] if host_batches is not None else None)
                      for i, sl in enumerate(self.slices)]
        self.streams = None
        if (n_parts > 1 or own_stream) and not sharded and hasattr(self.be, "torch"):
            # ``streams``: persistent streams of the caller (``CudaBackend.streams``) for typers built per pass
            self.streams = list(streams[: len(self.parts)]) if streams is not None else \
                [self.be.torch.cuda.Stream(device=self.be.device) for _ in self.parts]

    @property
    def score_cells(self) -> int:
        return sum(p.score_cells for p in self.parts)

    @property
    def host_nbytes(self) -> int:
        return sum(p.host.nbytes for p in self.parts)

    def pin(self) -> None:
        for p in self.parts:
            p.host.pin(self.be)

    def _each(self, fn_name: str, args: list | None = None) -> list:
        """``fn_name`` of every part, each under its own stream (no synchronisation)."""
        args = [()] * len(self.parts) if args is None else [(a,) for a in args]
        if self.streams is None:
            return [getattr(part, fn_name)(*a) for part, a in zip(self.parts, args)]
        torch = self.be.torch
        cur = torch.cuda.current_stream(self.be.device)
        out = []
        for part, st, a in zip(self.parts, self.streams, args):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                out.append(getattr(part, fn_name)(*a))
        return out

    def _join(self) -> None:
        if self.streams is not None:
            cur = self.be.torch.cuda.current_stream(self.be.device)
            for st in self.streams:
                cur.wait_stream(st)

    def upload(self) -> None:
        self._each("upload")
        self._join()

    def run_serial(self) -> list[GeneCall]:
        """All parts one after the other on the current stream (used to time kernels in isolation)."""
        calls: list[GeneCall] = []
        for part in self.parts:
            calls.extend(part.run())
        return calls

    def run(self) -> list[GeneCall]:
        self._each("start")                  # every part enqueued ...
        calls: list[GeneCall] = []
        for part in self._each("finish"):    # ... before the first read-back blocks the host
            calls.extend(part)
        self._join()
        return calls

    def upload_and_run(self) -> list[GeneCall]:
        """End-to-end pass: host pools -> device, typing, calls (copies of one part overlap the
        kernels of the others)."""
        self._each("upload_and_start")
        calls: list[GeneCall] = []
        for part in self._each("finish"):
            calls.extend(part)
        self._join()
        return calls

    def start_pass(self, upload: bool = False) -> list:
        """Enqueue one pass of every part (with ``upload``: the host->device copies first) and
        return its token without blocking; ``finish_pass(token)`` yields the calls."""
        return self._each("upload_and_start" if upload else "start")

    def finish_pass(self, token: list) -> list[GeneCall]:
        calls: list[GeneCall] = []
        for part in self._each("finish", token):
            calls.extend(part)
        self._join()
        return calls

    def replica(self, packs: list[GenePack], cns: list[int], top_n: int, group_size: int = 1) -> "CohortTyper":
        """A second typer over the same problems: same parts and host pools (shared, page-locked
        once), device buffers and streams of its own."""
        return CohortTyper(packs, cns, top_n=top_n, backend=self.be, n_parts=len(self.parts), group_size=group_size,
                           own_stream=True, host_batches=[p.host for p in self.parts])


class PassPipeline:
    """Consecutive passes of a cohort with ``len(typers)`` of them in flight: pass ``i`` runs on
    ``typers[i % depth]``, and the host blocks on a pass only when its typer is needed again (or at
    ``drain``).  With one typer used twice in a row the host-side call phase of a pass overlaps the
    kernels of the next; with replicas (device buffers and streams of their own) the next pass's
    host->device copies and kernels also overlap the latency-bound tail of the current one
    (selection and ranking are short serial chains) - double buffering of whole passes.  Calls come
    back in submission order."""

    def __init__(self, typers: list[CohortTyper], upload: bool = False, depth: int | None = None):
        self.typers = typers
        self.upload = upload
        self.depth = max(1, depth if depth is not None else len(typers))
        self._in_flight: list[tuple[CohortTyper, list]] = []
        self._next = 0

    def submit(self) -> list[GeneCall] | None:
        """Enqueue the next pass; returns the calls of the oldest pass if it had to be finished
        to make room, else ``None``."""
        done = None
        if len(self._in_flight) >= self.depth:
            typer, token = self._in_flight.pop(0)
            done = typer.finish_pass(token)
        typer = self.typers[self._next % len(self.typers)]
        self._next += 1
        self._in_flight.append((typer, typer.start_pass(self.upload)))
        return done

    def drain(self) -> list[list[GeneCall]]:
        out = []
        while self._in_flight:
            typer, token = self._in_flight.pop(0)
            out.append(typer.finish_pass(token))
        return out


def shard(items: list, rank: int, world: int) -> list:
    """Round-robin assignment of cohort samples to ranks (no data-path communication)."""
    return items[rank::world]
