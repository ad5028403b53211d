#!/usr/bin/env python
"""
Benchmark of the allele-typing hot path (BASELINE.json metric: read x candidate GCells/s and
samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cohort|wgs30x|deep]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the unmodified reference (oracle/_ref) on every host core

One step = one pass of the hot path over the whole batch: likelihood build (kernel a), every
copy-number step of the greedy search (kernels b, c) and the allele calls.

  value   whole-job scoring GCells/s with the packed inputs already resident in HBM (a batch that is
          typed repeatedly replays its launches as one CUDA graph from the third pass on)
  e2e     the same pass through the public host API (``CohortTyper.start_pass(upload=True)`` /
          ``finish_pass``): packed host arrays in pinned memory -> device, results back to the host,
          every step; sub-batches on their own streams so that the copies of one overlap the kernels
          of the others
  Consecutive steps are pipelined (``cohort.PassPipeline``, ``--pipeline-depth``): step i+1 is enqueued
  on a replica of the typer - device buffers and streams of its own - before the host reads back
  step i, as a prefetching input pipeline would; all K steps, with their copies, kernels, read-backs
  and host-side calls, complete inside the timed region, and every step's calls are compared with the
  warm-up pass afterwards (``parity.every_timed_pass_equals_warmup_calls``).
  roofline        dominant kernel (gk_score), CUDA events around every launch of extra passes run
                  right after the timed region (eager launches, one stream); bound "issue_nontensor": one
                  issue slot per cell (2 VIMNMX.U16x2 on the ALU pipe + 2 IMAD on the FMA pipe per 4 cells),
                  with the fraction of the measured peak of that mix and of round 1's ALU-only denominator
                  beside it (score_roofline below)
  cpu_baseline    the reference itself (kind "reference": oracle/_ref, the byte-compiled graphkir typing path,
                  likelihood build included, one process per sample on every host core) over a bounded
                  sample; kind "port" (the oracle's NumPy restatement) only where oracle/_ref is absent

Workloads: cohort = cfg5 (96 x cfg3, sample-sharded over the ranks, strong
scaling), wgs30x = cfg3 (one sample, 17 genes), deep = cfg4 (2M reads x 1000 alleles, CN 6).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SM_COUNT = 148
FP32_LANES_PER_SM = 128


# ---------------------------------------------------------------------------
# workload construction (host, before CUDA is touched so that fork is safe)
# ---------------------------------------------------------------------------
def _make_sample(args):
    seed, scale = args
    from kir_graph_b200 import packing, synthetic
    genes = synthetic.make_wgs30x_sample(seed=seed, scale=scale)
    packs = [packing.pack_synthetic(g) for g in genes]
    return packs, [g.cn for g in genes], [sorted(g.allele_names[t] for t in g.truth) for g in genes]


def build_cohort(seeds, scale, workers):
    jobs = [(s, scale) for s in seeds]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            out = pool.map(_make_sample, jobs)
    else:
        out = [_make_sample(j) for j in jobs]
    packs, cns, truth = [], [], []
    for p, c, t in out:
        packs += p
        cns += c
        truth += t
    return packs, cns, truth


def build_deep(n_reads, n_allele, cn):
    from kir_graph_b200 import packing, synthetic
    gene = synthetic.make_deep_sample(n_reads=n_reads, n_allele=n_allele, n_var=8 * n_allele, cn=cn)
    pack = packing.pack_synthetic(gene)
    return [pack], [cn], [sorted(gene.allele_names[t] for t in gene.truth)]


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, interval_ms: int = 200):
        self.index = index
        self.interval_ms = interval_ms          # the profiling recipe's -lms 200
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.interval_ms),
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref: graphkir byte-compiled by oracle/make_ref.py), or - only
# when that is absent - the oracle's float64 restatement
# ---------------------------------------------------------------------------
def _cpu_type_sample(args):
    """Type one synthetic cfg3 sample (read-scaled) on one core; returns (scoring cells, seconds of the
    whole typing path, seconds of the search alone, genes, genes whose call equals the generator's truth).

    kind "reference": per gene ``AlleleTyping(reads, variants, top_n=...).typing(cn)`` of the imported,
    unmodified reference - constructor (errorCorrection, reads2AlleleProb, log10:
    typing_mulit_allele.py:229-269) and greedy search (:383-598) - on the generator's objects; this
    process never loads libgk_typing.so.  kind "port": the oracle's F64Search (search only)."""
    seed, scale, top_n, kind = args
    from kir_graph_b200 import synthetic
    genes = synthetic.make_wgs30x_sample(seed=seed, scale=scale)
    cells, t_all, t_search, ok = 0, 0.0, 0.0, 0
    if kind == "reference":
        from oracle import ref_loader
        tma = ref_loader.load()[0]
        for g in genes:
            reads, variants = ref_loader.to_ref_objects(*g.to_objects())
            t0 = time.perf_counter()
            typ = tma.AlleleTyping(reads, variants, top_n=top_n)
            t1 = time.perf_counter()
            res = typ.typing(g.cn)
            t2 = time.perf_counter()
            t_all += t2 - t0
            t_search += t2 - t1
            n_reads, n_alleles = (typ.probs.shape if typ.probs.ndim == 2 else (0, 0))
            if len(typ.result) == g.cn:                      # not the homozygous shortcut: cn - 1 scored steps
                cells += sum(len(r.value) for r in typ.result[:-1]) * n_reads * n_alleles
            ok += sorted(res.selectBest()) == sorted(g.allele_names[t] for t in g.truth)
        return cells, t_all, t_search, len(genes), ok
    from kir_graph_b200 import packing
    from oracle import typing_oracle as orc
    prepared = []
    for g in genes:
        pack = packing.pack_synthetic(g)
        member = g.member[:, [g.allele_names.index(n) for n in pack.allele_names]]
        m = np.zeros((pack.n_reads, pack.n_alleles), dtype=np.int64)
        for name in ("lpv", "rpv", "lnv", "rnv"):
            off, idx = pack.csr.offsets[name], pack.csr.indices[name]
            row = np.repeat(np.arange(pack.n_reads), np.diff(off))
            np.add.at(m, row, (~member[idx] if name in ("lpv", "rpv") else member[idx]).astype(np.int64))
        prepared.append((orc.log_probs_from_counts(m, pack.k_obs.astype(np.int64)), g.cn))
    t0 = time.perf_counter()
    for lp, cn in prepared:
        search = orc.F64Search(lp, top_n=top_n, read_chunk=1024)
        for step in range(cn):
            if step:
                cells += len(search.result[-1].value) * lp.shape[1] * lp.shape[0]
            search.add_candidate()
    t_all = t_search = time.perf_counter() - t0
    return cells, t_all, t_search, len(genes), -1


def cpu_reference_kind() -> str:
    from oracle import ref_loader
    return "reference" if ref_loader.available() else "port"


class CpuArm:
    """``cores`` worker processes (spawned once, reused by every step), each typing one read-scaled
    cfg3 sample per step - the reference's own cohort recipe is one process per sample
    (research/test_speed.graphkir.par.sh:14)."""

    def __init__(self, cores, scale, top_n):
        import multiprocessing as mp
        self.cores, self.scale, self.top_n = cores, scale, top_n
        self.kind = cpu_reference_kind()
        self.pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None    # CUDA may be initialised: no fork
        self.next_seed = 100

    def step(self):
        """(cells, wall seconds of the slowest worker's typing, search-only seconds, samples, genes, genes ok)"""
        jobs = [(self.next_seed + i, self.scale, self.top_n, self.kind) for i in range(self.cores)]
        self.next_seed += self.cores
        out = self.pool.map(_cpu_type_sample, jobs) if self.pool is not None else [_cpu_type_sample(jobs[0])]
        return (sum(o[0] for o in out), max(o[1] for o in out), max(o[2] for o in out), len(jobs),
                sum(o[3] for o in out), sum(o[4] for o in out))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

    def describe(self) -> str:
        reads = int(200000 * self.scale)
        if self.kind == "reference":
            return (f"{self.cores} processes x 1 synthetic cfg3 sample each per step at {self.scale:g} of the reads "
                    f"({reads} read pairs, 17 genes, top_n={self.top_n}); the unmodified reference (oracle/_ref: graphkir "
                    "byte-compiled from /root/reference) AlleleTyping(reads, variants).typing(cn) per gene, likelihood "
                    "build included; time of the slowest process")
        return (f"{self.cores} processes x 1 synthetic cfg3 sample each per step at {self.scale:g} of the reads ({reads} "
                f"read pairs, 17 genes, top_n={self.top_n}), oracle F64Search (reference NumPy expressions, 1024-read "
                "chunks); search only (oracle/_ref not built)")


def run_reference(args, rank):
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1), 64))
    arm = CpuArm(cores, args.cpu_scale, args.top_n)
    for _ in range(args.warmup):
        arm.step()
    cells = wall = search = 0.0
    samples = genes = ok = 0
    steps = max(1, args.steps)
    for _ in range(steps):
        c, w, sw, n, g, k = arm.step()
        cells += c
        wall += w
        search += sw
        samples += n
        genes += g
        ok += k
    arm.close()
    value = cells / wall / 1e9
    line = {
        "impl": "reference", "metric": "allele-typing read x candidate GCells/s", "value": value,
        "unit": "GCells/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg5 cohort: {args.samples} synthetic 30x WGS samples x 200000 read pairs x 900 alleles / 17 "
                               f"genes, CN<=4, top_n={args.top_n}, sample-sharded",
                   "top_n": args.top_n,
                   "bounded_sample": f"each step types {cores} samples (one per host core) at {args.cpu_scale:g} of the reads "
                                     "instead of 96 full samples: the full step is ~10 core-minutes per sample in the "
                                     "reference; GCells/s and samples/s are rates, and the reference's cost is linear in "
                                     "the reads (likelihood) and in reads x candidates (search)"},
        "samples_per_s": samples * args.cpu_scale / wall,
        "search_only_value": cells / search / 1e9 if search else None,
        "parity": {"genes_matching_generator_truth": ok, "genes": genes},
        "cpu_baseline": {"value": value, "unit": "GCells/s", "cores": cores, "kind": arm.kind, "sample": arm.describe()},
        "e2e": {"value": value, "unit": "GCells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def load_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def score_roofline(cells, ms, packed, sm_max_mhz):
    """Roofline entry of the scoring kernel.  Its bound is non-tensor instruction issue, not HBM or
    the tensor cores.  The packed path spends 4 instructions on 4 cells: 2 VIMNMX.U16x2 on the ALU pipe
    (64 lanes/clk/SM) and 2 IMAD (acc = min * 1 + acc) on the FMA pipe, i.e. ONE issue slot per cell
    (128 lanes/clk/SM) and half an ALU-pipe instruction per cell: both limits are 128 cells/clk/SM.
    (Until round 2 the adds were IADD3 on the ALU pipe as well: 0.75 ALU instructions per cell,
    85.3 cells/clk/SM; ``alu_only_formulation`` keeps that denominator for comparison with round 1.)
    The FP32 path issues 2 FADD per cell on the FMA pipe.  The FP32-equivalent fraction (2 ops per
    cell against the FP32 non-tensor peak) is what BASELINE.json's target is stated in."""
    cps = cells / (ms * 1e-3) if ms else 0.0
    fp32_peak = SM_COUNT * FP32_LANES_PER_SM * sm_max_mhz * 1e6 / 1e12          # T lane-ops/s
    src = (f"148 SM x {{lanes}} lanes/clk x {sm_max_mhz:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); not in "
           "MEASURED_PEAKS.json, which only holds HBM and bf16 tensor peaks")
    fp32 = {"achieved": 2.0 * cps / 1e12, "peak": fp32_peak, "frac": 2.0 * cps / 1e12 / fp32_peak}
    measured = {}
    try:                                         # the same instruction mix timed by tools/micro/mixpipe.cu on this pool
        measured = json.load(open(os.path.join(ROOT, "profiles", "nontensor_peaks.json")))
    except (OSError, ValueError):
        pass
    if packed:
        peak = fp32_peak                                                         # issue: 128 lanes/clk/SM
        ach = 1.0 * cps / 1e12
        out = {"kernel": "gk_score_packed_kernel", "bound": "issue_nontensor", "achieved": ach, "peak": peak,
               "unit": "T instruction lanes/s (2 VIMNMX.U16x2 on the ALU pipe + 2 IMAD on the FMA pipe per 4 cells)",
               "frac": ach / peak, "cells_per_s": cps, "fp32_nontensor_equiv": fp32,
               "alu_only_formulation": {"achieved": 0.75 * cps / 1e12, "peak": fp32_peak / 2.0,
                                        "frac": 0.75 * cps / 1e12 / (fp32_peak / 2.0),
                                        "note": "round-1 denominator: 2 VIMNMX.U16x2 + 1 IADD3 per 4 cells, all on the "
                                                "64 lanes/clk/SM ALU pipe; above 1 means faster than that formulation can run"},
               "peak_source": src.format(lanes=128)}
        if measured.get("u16x2_min_imad_tcells_per_s"):
            m_peak = float(measured["u16x2_min_imad_tcells_per_s"])
            out["measured_peak"] = {"achieved": cps / 1e12, "peak": m_peak, "unit": "TCells/s", "frac": cps / 1e12 / m_peak,
                                    "source": "tools/micro/mixpipe.cu (same VIMNMX.U16x2 + IMAD mix, operands from shared "
                                              "memory, accumulators in registers) on a B200 of this pool: "
                                              "profiles/r02_mixpipe.txt"}
        return out
    return {"kernel": "gk_score_kernel", "bound": "fp32_nontensor", "achieved": fp32["achieved"], "peak": fp32_peak,
            "unit": "T FP32 ops/s (2 FADD per cell: d = p - l, acc += |d|)", "frac": fp32["frac"],
            "cells_per_s": cps, "fp32_nontensor_equiv": fp32, "peak_source": src.format(lanes=128)}



# ---------------------------------------------------------------------------
# cfg4: one very deep sample over all ranks
# ---------------------------------------------------------------------------
def deep_leg(args, be, rank, world, timed, packed, sm_max, shard, steps=3, parity=True):
    """One deep sample (cfg4) typed by all ranks together; returns the ``deep`` block (rank 0) or None.

    shard = "reads": every rank holds R / world reads (``packing.shard_reads``); likelihood, scoring, tie
                     counting and the P writer all work on the local reads, the partial score matrix and
                     the partial tie counts are summed with two small all-reduces per copy-number step
                     (plus one of the column sums), nothing is replicated.
    shard = "cols":  candidate-column tiles of the scoring kernel are dealt to the ranks, L / LT / P and
                     every other kernel are replicated, one all-reduce of the score matrix per step.
    ``calls_equal_unsharded``: every copy-number step of the sharded search (kept allele ids, scores,
    tie counts, tie flags, N_uniq) against the same code run unsharded on rank 0."""
    import torch
    import torch.distributed as dist
    from kir_graph_b200 import cohort, engine, packing
    t0 = time.perf_counter()
    d_packs, d_cns, d_truth = build_deep(args.deep_reads, args.deep_alleles, args.deep_cn)
    t_build = time.perf_counter() - t0
    reduce_scores = None
    if world > 1:
        def reduce_scores(t):                       # integer sums over NVLink: exact, order independent
            dist.all_reduce(t)
    kw = {}
    my_packs = d_packs
    if world > 1 and shard == "reads":
        my_packs = [packing.shard_reads(p, rank, world) for p in d_packs]
        kw = dict(read_shard=True, reduce_scores=reduce_scores)
    elif world > 1:
        kw = dict(col_shard=(rank, world), reduce_scores=reduce_scores)
    deep = cohort.CohortTyper(my_packs, d_cns, top_n=args.top_n, backend=be, n_parts=1, **kw)
    deep.pin()
    deep.upload()
    for _ in range(3):                              # unsharded: the third pass records the CUDA graph
        d_calls = deep.run()
    d_ms = timed(deep.run, steps) / steps
    group = deep.parts[0].group
    coll0 = group.collective_bytes
    deep.run()
    coll_bytes = group.collective_bytes - coll0
    h0 = be.h2d_bytes
    e_ms = timed(deep.upload_and_run, steps) / steps
    h2d = (be.h2d_bytes - h0) / steps
    # per-kernel times of one more pass (CUDA events around every launch; collectives excluded)
    torch.cuda.synchronize()
    be.timing = {}
    deep.run_serial()
    torch.cuda.synchronize()
    timing, be.timing = be.timing, None
    k_ms = {k: sum(a.elapsed_time(b) for a, b, _ in v) for k, v in timing.items()}
    evs = timing.get("gk_score", [])
    d_ms_s = sum(a.elapsed_time(b) for a, b, _ in evs)
    d_work = sum(w for _, _, w in evs)
    agg = torch.tensor([deep.score_cells, h2d, d_work], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
    mx = torch.tensor([d_ms_s] + [k_ms.get(k, 0.0) for k in sorted(k_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    cells_all, h2d_all, work_all = [float(x) for x in agg.tolist()]

    equal = None
    if parity and world > 1:
        # every step of the sharded search against the unsharded one (rank 0 holds both)
        part = deep.parts[0]
        sg = engine.SearchGroup(part.batch, [0], args.top_n, reduce_scores=reduce_scores, **{
            k: v for k, v in kw.items() if k in ("read_shard", "col_shard")})
        cn = int(d_cns[0])
        mine = [sg.step(need_next=[i + 1 < cn])[0] for i in range(cn)]
        del sg
        if rank == 0:
            whole = engine.MatrixBatch(d_packs, backend=be)
            wg = engine.SearchGroup(whole, [0], args.top_n)
            ref = [wg.step(need_next=[i + 1 < cn])[0] for i in range(cn)]
            equal = all(np.array_equal(a.ids, b.ids) and np.array_equal(a.score, b.score)
                        and np.array_equal(a.cnt, b.cnt) and a.tie_flags == b.tie_flags
                        and a.n_unique == b.n_unique and a.cut == b.cut for a, b in zip(mine, ref))
            del wg, whole
        if world > 1:
            dist.barrier()
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del deep
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    names = sorted(k_ms)
    return {
        "workload": f"cfg4 deep: {args.deep_reads} read pairs x {args.deep_alleles} alleles, CN {args.deep_cn}, "
                    f"top_n={args.top_n}",
        "n_gpus": world, "sharding": "none" if world == 1 else shard,
        "value": cells_all / (d_ms * 1e-3) / 1e9, "unit": "GCells/s", "ms_per_step": d_ms, "steps": steps,
        "e2e": {"value": cells_all / (e_ms * 1e-3) / 1e9, "unit": "GCells/s", "ms_per_step": e_ms,
                "h2d_bytes_per_step": h2d_all},
        "collective": None if world == 1 else {
            "kind": "ncclAllReduce (sum of uint32 / uint64 partial integer sums)",
            "bytes_per_pass_per_rank": coll_bytes, "per_cn_step": "score matrix S" + (
                " + tie counts of the alive sets" if shard == "reads" else ""),
            "calls_per_pass": (2 * (int(d_cns[0]) - 1) + 1) if shard == "reads" else int(d_cns[0]) - 1},
        "calls_match_truth": sorted(d_calls[0].alleles) == d_truth[0],
        "calls_equal_unsharded": equal,
        "kernel_ms_per_step_max_over_ranks": dict(zip(names, [float(x) for x in mx.tolist()[1:]])),
        "roofline": dict(score_roofline(work_all / world, float(mx[0].item()), packed, sm_max), launches=len(evs),
                         note="cells of one rank (mean) over the slowest rank's scoring time"),
        "device_mem_gib_rank0": mem, "build_s": t_build,
    }


# ---------------------------------------------------------------------------
# cold / streaming cohort throughput: every pass types a DIFFERENT cohort
# ---------------------------------------------------------------------------
COLD_SETS = 3


def build_cold_sets(args, rank, world, workers, n_sets=COLD_SETS):
    """The cohorts of the e2e_cold measurement (other seeds than the warm benchmark's), packed on the host
    before CUDA is initialised (the worker pool forks)."""
    sets = []
    for k in range(n_sets):
        seeds = list(range(5000 + 1000 * k, 5000 + 1000 * k + args.samples))[rank::world]
        sets.append(build_cohort(seeds, args.scale, workers))
    return sets


def cold_leg(args, be, rank, world, timed, group_size, parts, packed_sets):
    """``e2e_cold``: each timed pass types a cohort the GPU has not seen - other samples, other read counts
    per gene.  ``n_sets`` cohorts (other seeds than the warm benchmark's) are packed and their host pools
    page-locked beforehand, together with what depends on the packed input alone (per-site summaries of
    ``isHomozygous``, the tile layout of the likelihood build; host preparation is reported in ``host_prep``);
    inside the timed region every pass constructs its ``CohortTyper`` from those prepared host batches -
    homozygosity decisions for its copy numbers, work-item tables, device buffers, every launch issued
    eagerly: no launch plan, CUDA graph or decision survives from another pass - copies the inputs, types,
    and returns the calls on the host.  Two passes are in flight: the host builds
    and enqueues pass i + 1 while the GPU runs pass i (what a streaming cohort does)."""
    import torch
    import torch.distributed as dist
    from kir_graph_b200 import cohort, engine
    sets = []
    n_sets = len(packed_sets)
    for packs, cns, truth in packed_sets:
        probe = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=parts, group_size=group_size)
        probe.pin()
        hosts = [p.host for p in probe.parts]
        del probe
        sets.append((packs, cns, truth, hosts))
    # the packed cohorts and their page-locked pools live for the whole stream: out of the garbage collector's
    # view, like the packed cohort of the warm legs (a pass creates a few thousand short-lived objects - typers,
    # work-item tables, calls - and every full collection otherwise re-traverses the packs of all the cohorts)
    import gc
    gc.collect()
    gc.freeze()
    state = {"i": 0, "pending": None}
    results = []
    # device buffers of a pass are slices of one of two arenas (two passes in flight), recycled as a whole
    # when the pass two steps earlier has been read back; streams (and their staging arenas) are persistent
    free_b, _ = torch.cuda.mem_get_info()
    arena_bytes = int(min(0.4 * free_b, 48 * 2 ** 30))
    arenas = [engine.DeviceArena(be, arena_bytes) for _ in range(2)]
    stream_sets = [be.streams(parts, first=k * parts) for k in range(2)]

    def step():
        packs, cns, _, hosts = sets[state["i"] % n_sets]
        slot = state["i"] % 2
        arenas[slot].reset()
        be.device_arena = arenas[slot]
        try:
            typer = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=parts,
                                       group_size=group_size, own_stream=True, host_batches=hosts,
                                       streams=stream_sets[slot])
            token = typer.start_pass(upload=True)
        finally:
            be.device_arena = None
        if state["pending"] is not None:
            prev, prev_token, k = state["pending"]
            results.append((k, prev.finish_pass(prev_token), prev.score_cells))
        state["pending"] = (typer, token, state["i"] % n_sets)
        state["i"] += 1

    def finalize():
        if state["pending"] is not None:
            prev, prev_token, k = state["pending"]
            results.append((k, prev.finish_pass(prev_token), prev.score_cells))
            state["pending"] = None

    for _ in range(2):                      # CUDA context, allocator pools and kernel modules are warm; no typer is
        step()
    finalize()
    results.clear()
    state["i"] = 0
    h0 = be.h2d_bytes
    l0 = be.launches
    ms = timed(step, args.steps, finalize) / args.steps
    h2d = (be.h2d_bytes - h0) / args.steps
    launches = (be.launches - l0) / args.steps
    ok = genes = 0
    cells = 0.0
    for k, calls, c in results:
        truth = sets[k][2]
        ok += sum(sorted(x.alleles) == t for x, t in zip(calls, truth))
        genes += len(truth)
        cells += c
    agg = torch.tensor([cells / max(len(results), 1), h2d, ok, genes, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
    cells_all, h2d_all, ok_all, genes_all, launches_all = [float(x) for x in agg.tolist()]
    del sets
    gc.unfreeze()
    if rank != 0:
        return None
    return {"value": cells_all / (ms * 1e-3) / 1e9, "unit": "GCells/s", "ms_per_step": ms,
            "samples_per_s": args.samples / (ms * 1e-3), "h2d_bytes_per_step": h2d_all,
            "gpu_launches_per_step": launches_all, "cohorts_rotated": n_sets, "passes_in_flight": 2,
            "sub_batches": parts, "device_arena_gib": arena_bytes / 2 ** 30,
            "device_arena_misses": sum(a.misses for a in arenas),
            "parity": {"genes_matching_generator_truth": int(ok_all), "genes": int(genes_all)},
            "timed_region": "per pass: CohortTyper constructed from the prepared host batch of a cohort not seen "
                            "before (page-locked packed pools and what depends on the packed input alone: per-site "
                            "summaries of isHomozygous, the tile layout of the likelihood build) - homozygosity "
                            "decisions for the pass's copy numbers, work-item tables of every search step, device "
                            "buffers - host->device copies, every kernel launched eagerly, read-back, calls on the "
                            "host; no launch plan, graph or device buffer content reused; the packed cohorts are "
                            "frozen out of the garbage collector's view (gc.freeze) as in the warm legs"}


def guarded(world, fn, *a, **kw):
    """A secondary block of the line (cold, deep, host_prep, api_e2e, cn_model, cpu_baseline) must not cost the
    headline numbers measured before it: on one GPU its failure is recorded in the block ({"error": ...}, traceback
    on stderr) and the line is still printed.  With several ranks a block may hold collectives, where a rank that
    went on alone would leave the others waiting: there the failure stays fatal."""
    if world > 1:
        return fn(*a, **kw)
    try:
        return fn(*a, **kw)
    except Exception as exc:                                          # noqa: BLE001 - recorded, not swallowed
        import traceback
        traceback.print_exc()
        return {"error": f"{type(exc).__name__}: {exc}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cohort", choices=["cohort", "wgs30x", "deep"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the reads per sample (tests)")
    ap.add_argument("--top-n", type=int, default=300)
    ap.add_argument("--deep-reads", type=int, default=2_000_000)
    ap.add_argument("--deep-alleles", type=int, default=1000)
    ap.add_argument("--deep-cn", type=int, default=6)
    ap.add_argument("--cpu-scale", type=float, default=0.1,
                    help="fraction of the reads of a cfg3 sample each CPU-arm process types per step")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cold", action="store_true", help="skip the e2e_cold measurement (a different cohort every pass)")
    ap.add_argument("--no-host", action="store_true", help="skip the host_prep / api_e2e blocks (1-GPU run only)")
    ap.add_argument("--no-deep", action="store_true",
                    help="skip the nested cfg4 deep-sample measurement of the default 1-GPU cohort run")
    ap.add_argument("--deep-shard", default="reads", choices=["reads", "cols", "both"],
                    help="how the deep sample (cfg4) is spread over the ranks: by reads (nothing replicated, two "
                         "small all-reduces per step), by candidate columns (L / LT / P replicated, one all-reduce "
                         "per step), or both one after the other (the second as deep_cols)")
    ap.add_argument("--parts", type=int, default=1, help="sub-batches (streams) of the resident pass")
    ap.add_argument("--pipeline-depth", type=int, default=0,
                    help="consecutive passes (steps) in flight: pass i+1 is enqueued on a replica of the typer "
                         "(device buffers and streams of its own) before the host reads back pass i; 1 = one "
                         "pass at a time; 0 = 3, or 4 when this rank holds fewer than 16 samples (the "
                         "selection / ranking chains of a small batch leave more of the GPU idle)")
    ap.add_argument("--e2e-parts", type=int, default=6,
                    help="sub-batches (streams) of the end-to-end pass: the host->device copies of one overlap "
                         "the kernels of the others")
    ap.add_argument("--clock-interval-ms", type=int, default=200,
                    help="nvidia-smi sampling period during the timed region (profiling recipe: 200)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    # ---- host: build this rank's share of the workload -------------------------
    t_build = time.perf_counter()
    workers = max(1, (os.cpu_count() or 1) // max(world, 1))
    if args.workload == "cohort":
        seeds = list(range(100, 100 + args.samples))[rank::world]
        packs, cns, truth = build_cohort(seeds, args.scale, workers)
        n_samples_total, n_samples_local = args.samples, len(seeds)
        desc = (f"cfg5 cohort: {args.samples} synthetic 30x WGS samples x {int(200000 * args.scale)} read pairs x 900 "
                f"alleles / 17 genes, CN<=4, top_n={args.top_n}, sample-sharded")
    elif args.workload == "wgs30x":
        packs, cns, truth = build_cohort([3], args.scale, 1)
        n_samples_total = n_samples_local = 1
        desc = f"cfg3: one synthetic 30x WGS sample, {int(200000 * args.scale)} read pairs x 900 alleles / 17 genes"
    t_build = time.perf_counter() - t_build
    cold_sets = build_cold_sets(args, rank, world, workers) if args.workload == "cohort" and not args.no_cold else None

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from kir_graph_b200 import cohort, engine
    be = engine.CudaBackend(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finalize=None):
        """CUDA events on the launching stream around `steps` calls (and `finalize`, which ends the
        passes still in flight: its streams are joined into the launching one); max over ranks."""
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        if finalize is not None:
            finalize()
        end.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    if args.workload == "deep":
        # cfg4 alone: the deep block is the line
        peaks = load_peaks()
        sm_max = float(peaks.get("sm_max_mhz") or 1965.0)
        packed = bool(engine.PACKED_DEFAULT)
        sampler = ClockSampler(local_rank, args.clock_interval_ms)
        sampler.start()
        launches0 = be.launches
        modes = ["reads", "cols"] if args.deep_shard == "both" else [args.deep_shard]
        blocks = {}
        for mode in (modes if world > 1 else modes[:1]):
            blocks[mode] = deep_leg(args, be, rank, world, timed, packed, sm_max, mode, steps=args.steps)
        clocks = sampler.stop()
        if rank == 0:
            first = blocks[modes[0]]
            line = {"metric": "allele-typing read x candidate GCells/s", "value": first["value"], "unit": "GCells/s",
                    "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": first["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "u16" if packed else "f32", "data": "synthetic",
                    "config": {"workload": first["workload"], "top_n": args.top_n, "sharding": first["sharding"],
                               "l2": "inputs larger than L2 (no flush needed)"},
                    "e2e": first["e2e"], "gpu_launches": be.launches - launches0, "clocks": clocks,
                    "roofline": first["roofline"], "deep": first}
            for mode in modes[1:]:
                if mode in blocks:
                    line["deep_" + mode] = blocks[mode]
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    group_size = 17
    depth = args.pipeline_depth if args.pipeline_depth > 0 else (4 if n_samples_local < 16 else 3)
    col_shard = reduce_scores = None
    typer = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=args.parts,
                               group_size=group_size, col_shard=col_shard,
                               reduce_scores=reduce_scores, own_stream=depth > 1)
    typer.pin()
    # The packed cohort is a large, long-lived heap (thousands of arrays); keep the cyclic garbage
    # collector from re-traversing it every time the per-run result objects trigger a collection
    # (measured: 2.5 ms of a 28 ms pass).
    import gc
    gc.collect()
    gc.freeze()

    # ---- resident: inputs already in HBM ---------------------------------------------
    sampler = ClockSampler(local_rank, args.clock_interval_ms)
    sampler.start()
    typer.upload()
    calls = None
    for _ in range(args.warmup):
        calls = typer.run()
    cells_per_step = typer.score_cells
    lik_cells = sum(p.batch.n_cells for p in typer.parts)
    lik_bytes = sum(p.batch.bytes_out for p in typer.parts)
    n_parts = len(typer.parts)
    key = lambda cs: [(c.gene, c.alleles, c.score) for c in cs]
    pipe_ok = True

    def pipelined(typers, upload):
        """(step, finalize) of a PassPipeline over `typers`; every pass's calls are checked against
        the warm-up pass outside the timed region."""
        pipe = cohort.PassPipeline(typers, upload=upload)
        results = []

        def step():
            done = pipe.submit()
            if done is not None:
                results.append(done)

        def finalize():
            results.extend(pipe.drain())
        return step, finalize, results

    replicas = [typer]
    for _ in range(depth - 1):
        twin = typer.replica(packs, cns, args.top_n, group_size)
        twin.upload()
        for _ in range(3):                       # the third pass records the CUDA graph
            twin.run()
        replicas.append(twin)
    gc.collect()
    gc.freeze()
    launches0 = be.launches
    if depth > 1:
        step, finalize, results = pipelined(replicas, upload=False)
        for _ in range(depth):                   # fill and drain once untimed
            step()
        finalize()
        results.clear()
        launches0 = be.launches
        ms_total = timed(step, args.steps, finalize)
        pipe_ok = len(results) == args.steps and all(key(r) == key(calls) for r in results)
        del results[:]
    else:
        ms_total = timed(typer.run, args.steps)
    launches = be.launches - launches0
    clocks = sampler.stop()
    mem_resident = torch.cuda.max_memory_allocated() / 2 ** 30
    del replicas[1:]
    # kernel timing pass: same work, sub-batches one after the other on one stream so that the
    # CUDA events around each launch do not overlap other kernels
    torch.cuda.synchronize()
    be.timing = {}
    roof_steps = max(1, min(3, args.steps))
    for _ in range(roof_steps):
        typer.run_serial()
    torch.cuda.synchronize()
    timing = be.timing
    be.timing = None

    # ---- end to end: pinned host arrays -> device -> calls on the host ------------------
    e2e_typer = typer
    if args.e2e_parts != len(typer.parts) and col_shard is None:
        # at least 12 samples per sub-batch: smaller ones cost more in fixed per-part work than the
        # copy/compute overlap returns
        e2e_parts = max(1, min(args.e2e_parts, n_samples_local // 12))
        cand = cohort.CohortTyper(packs, cns, top_n=args.top_n, backend=be, n_parts=e2e_parts,
                                  group_size=group_size, own_stream=depth > 1)
        if len(cand.parts) != len(typer.parts):
            e2e_typer = cand
            e2e_typer.pin()
            gc.collect()
            gc.freeze()
    twin = None
    if e2e_typer is not typer:                   # the resident typers' device buffers are not needed any more
        replicas.clear()
        typer = None
        gc.unfreeze()
        gc.collect()
        gc.freeze()
    torch.cuda.empty_cache()

    e2e_replicas = [e2e_typer]
    for _ in range(depth - 1):
        e2e_replicas.append(e2e_typer.replica(packs, cns, args.top_n, group_size))
    for t in e2e_replicas:
        for _ in range(max(3, args.warmup)):     # the third pass of a batch records its CUDA graph
            t.upload_and_run()
    gc.collect()
    gc.freeze()
    e2e_ok = True
    if depth > 1:
        step, finalize, results = pipelined(e2e_replicas, upload=True)
        for _ in range(depth):
            step()
        finalize()
        results.clear()
        h0, d0 = be.h2d_bytes, be.d2h_bytes
        ms_e2e = timed(step, args.steps, finalize)
        e2e_ok = len(results) == args.steps and all(key(r) == key(calls) for r in results)
        del results[:]
    else:
        h0, d0 = be.h2d_bytes, be.d2h_bytes
        ms_e2e = timed(e2e_typer.upload_and_run, args.steps)
    h2d = (be.h2d_bytes - h0) / args.steps
    d2h = (be.d2h_bytes - d0) / args.steps
    mem_e2e = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- cold: a different cohort every pass ------------------------------------------------
    cold = None
    if args.workload == "cohort" and not args.no_cold:
        e2e_replicas.clear()
        e2e_typer_parts = len(e2e_typer.parts)
        e2e_typer = typer = twin = None
        gc.unfreeze()
        gc.collect()
        torch.cuda.empty_cache()
        cold = guarded(world, cold_leg, args, be, rank, world, timed, group_size, 1, cold_sets)   # one sub-batch: least host work per pass
        be.device_arena = None
        cold_sets = None
        gc.collect()
        torch.cuda.empty_cache()
    else:
        e2e_typer_parts = len(e2e_typer.parts)

    # ---- aggregate over ranks --------------------------------------------------------------
    agg = torch.tensor([cells_per_step, lik_cells, launches, h2d, d2h, n_samples_local], dtype=torch.float64,
                       device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    cells_all, lik_all, launches_all, h2d_all, d2h_all, samples_all = [float(x) for x in agg.tolist()]

    def kernel_stats(name):
        evs = timing.get(name, [])
        ms = sum(s.elapsed_time(e) for s, e, _ in evs)
        work = sum(w for _, _, w in evs)
        return ms, work, len(evs)

    peaks = load_peaks()
    sm_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    packed = bool(engine.PACKED_DEFAULT)
    line = None
    if rank == 0:
        ms_s, work_s, n_s = kernel_stats("gk_score")
        ms_l, work_l, n_l = kernel_stats("gk_likelihood")
        traffic = None
        try:                                     # DRAM bytes per launch of the same kernel from an ncu capture
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            entry = prof.get(args.workload, {}).get("gk_score_packed_kernel" if packed else "gk_score_kernel")
            if entry and world == 1:
                traffic = entry["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs") or 6650.0)
        lik_gbs = (lik_bytes * n_l / n_parts) / (ms_l * 1e-3) / 1e9 if ms_l else 0.0
        step_ms = ms_total / args.steps
        value = cells_all / (step_ms * 1e-3) / 1e9
        e2e_ms = ms_e2e / args.steps
        truth_ok = sum(sorted(c.alleles) == t for c, t in zip(calls, truth)) if calls else 0
        line = {
            "metric": "allele-typing read x candidate GCells/s", "value": value, "unit": "GCells/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong" if args.workload == "cohort" else "replicas",
            "vs_baseline": None, "dtype": "u16" if packed else "f32", "data": "synthetic",
            "config": {"workload": desc, "top_n": args.top_n, "l2": "inputs larger than L2 (no flush needed)",
                       "timed_region": "likelihood build + all CN steps + calls, packed inputs resident in HBM",
                       "concurrent_sub_batches": n_parts, "e2e_sub_batches": e2e_typer_parts,
                       "passes_in_flight": depth,
                       "pipelining": ("step i+1 is enqueued on a second set of device buffers and streams before the "
                                      "host reads back step i (double-buffered passes; every step's copies, kernels, "
                                      "read-back and calls lie inside the timed region)") if depth > 1 else "none"},
            "samples_per_s": samples_all / (step_ms * 1e-3),
            "e2e": {"value": cells_all / (e2e_ms * 1e-3) / 1e9, "unit": "GCells/s",
                    "samples_per_s": samples_all / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": dict(
                score_roofline(work_s, ms_s, packed, sm_max), traffic=traffic, launches=n_s,
                kernel_ms_per_step=ms_s / roof_steps,
                share_of_step=(ms_s / roof_steps) / (ms_total / args.steps) if ms_total else None,
                measured=f"CUDA events around every launch of {roof_steps} extra step(s) run right after the timed "
                         "region with the sub-batches serialised on one stream"),
            "roofline_likelihood": {
                "kernel": "gk_likelihood_kernel", "bound": "hbm", "achieved": lik_gbs, "peak": hbm_peak,
                "unit": "GB/s", "frac": lik_gbs / hbm_peak if hbm_peak else None, "traffic": None,
                "cells_per_s": work_l / (ms_l * 1e-3) if ms_l else 0.0, "bytes_per_cell": 5,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            },
            "kernel_ms_per_step": {k: kernel_stats(k)[0] / roof_steps for k in timing},
            "parity": {"genes_matching_generator_truth": truth_ok, "genes": len(truth),
                       "every_timed_pass_equals_warmup_calls": bool(pipe_ok and e2e_ok)},
            "device_mem_gib": {"after_resident": mem_resident, "after_e2e": mem_e2e},
            "build_s": t_build,
        }
        if cold is not None:
            line["e2e_cold"] = cold
    if args.workload == "cohort" and not args.no_deep:
        # second shape of the same path: one very deep sample (cfg4) typed by all ranks together
        typer = e2e_typer = twin = None
        del e2e_replicas[:], replicas[:]
        gc.unfreeze()
        gc.collect()
        torch.cuda.empty_cache()
        modes = ["reads", "cols"] if args.deep_shard == "both" else [args.deep_shard]
        for i, mode in enumerate(modes if world > 1 else modes[:1]):
            block = guarded(world, deep_leg, args, be, rank, world, timed, packed, sm_max, mode)
            be.timing = None
            if rank == 0:
                line["deep" if i == 0 else "deep_" + mode] = block
    if rank == 0:
        if world == 1 and not args.no_host:
            from tools import host_numbers
            line["host_prep"] = guarded(world, host_numbers.host_prep)
            line["api_e2e"] = guarded(world, host_numbers.api_e2e, args.cpu_scale, backend=be)
            line["cn_model"] = guarded(world, host_numbers.cn_model, args.samples, backend=be)
        if world == 1 and not args.no_cpu_baseline:
            def cpu_baseline():
                cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1), 64))
                arm = CpuArm(cores, args.cpu_scale, args.top_n)
                try:
                    c_cells, c_wall, c_search, c_n, _, _ = arm.step()
                finally:
                    arm.close()
                return {"value": c_cells / c_wall / 1e9, "unit": "GCells/s", "cores": cores, "kind": arm.kind,
                        "samples_per_s": c_n * args.cpu_scale / c_wall,
                        "search_only_value": c_cells / c_search / 1e9 if c_search else None, "sample": arm.describe()}
            line["cpu_baseline"] = guarded(world, cpu_baseline)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
