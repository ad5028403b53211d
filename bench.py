#!/usr/bin/env python
"""
Benchmark of the allele-typing hot path (BASELINE.json metric: read x candidate GCells/s and
samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cohort|wgs30x|deep]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle's float64 NumPy restatement

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
                  right after the timed region (eager launches, one stream)
  cpu_baseline    the oracle (kind "port": reference NumPy expressions, read-chunked) on host cores

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
# CPU arm: the oracle's float64 restatement of the reference path
# ---------------------------------------------------------------------------
def _cpu_type_sample(args):
    """Type one synthetic cfg3 sample (read-scaled) with the reference's NumPy expressions."""
    seed, scale, top_n = args
    from kir_graph_b200 import packing, synthetic
    from oracle import typing_oracle as orc
    genes = synthetic.make_wgs30x_sample(seed=seed, scale=scale)
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
    cells = 0
    for lp, cn in prepared:
        search = orc.F64Search(lp, top_n=top_n, read_chunk=1024)
        for step in range(cn):
            if step:
                cells += len(search.result[-1].value) * lp.shape[1] * lp.shape[0]
            search.add_candidate()
    return cells, time.perf_counter() - t0


def cpu_reference_step(cores, scale, top_n, first_seed=100):
    """``cores`` worker processes each type one read-scaled sample; returns (cells, wall seconds, samples)."""
    jobs = [(first_seed + i, scale, top_n) for i in range(cores)]
    t0 = time.perf_counter()
    if cores > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(cores) as pool:      # CUDA may already be initialised: no fork
            out = pool.map(_cpu_type_sample, jobs)
    else:
        out = [_cpu_type_sample(jobs[0])]
    wall = max(t for _, t in out)      # typing time only (generation excluded), slowest worker
    _ = time.perf_counter() - t0
    return sum(c for c, _ in out), wall, len(jobs)


def run_reference(args, rank):
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1), 64))
    scale = args.cpu_scale
    for _ in range(args.warmup):
        pass                                   # nothing to warm: fresh processes every step
    cells = wall = 0.0
    samples = 0
    for _ in range(max(1, args.steps)):
        c, w, n = cpu_reference_step(cores, scale, args.top_n)
        cells += c
        wall += w
        samples += n
    value = cells / wall / 1e9
    sample_desc = (f"{cores} processes x 1 synthetic cfg3 sample each at {scale:g} of the reads "
                   f"({int(200000 * scale)} read pairs, 17 genes, top_n={args.top_n}), oracle F64Search "
                   f"(reference NumPy expressions, 1024-read chunks); typing time only")
    line = {
        "impl": "reference", "metric": "allele-typing read x candidate GCells/s", "value": value,
        "unit": "GCells/s", "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cohort (cfg5: 96 x cfg3 synthetic 30x WGS samples), bounded sample", "top_n": args.top_n},
        "samples_per_s": samples * scale / wall,
        "cpu_baseline": {"value": value, "unit": "GCells/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "GCells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def score_roofline(cells, ms, packed, sm_max_mhz):
    """Roofline entry of the scoring kernel.  Its bound is non-tensor instruction issue, not HBM or
    the tensor cores: the packed path issues 3 ALU-pipe instructions (2 VIMNMX.U16x2 + 1 IADD3)
    per 4 cells on the half-rate (64 lanes/clk/SM) ALU pipe, the FP32 path 2 FADD per cell on the
    128 lanes/clk/SM FMA pipe.  The FP32-equivalent fraction (2 ops per cell against the FP32
    non-tensor peak) is what BASELINE.json's target is stated in."""
    cps = cells / (ms * 1e-3) if ms else 0.0
    fp32_peak = SM_COUNT * FP32_LANES_PER_SM * sm_max_mhz * 1e6 / 1e12          # T lane-ops/s
    src = (f"148 SM x {{lanes}} lanes/clk x {sm_max_mhz:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); not in "
           "MEASURED_PEAKS.json, which only holds HBM and bf16 tensor peaks")
    fp32 = {"achieved": 2.0 * cps / 1e12, "peak": fp32_peak, "frac": 2.0 * cps / 1e12 / fp32_peak}
    if packed:
        peak = fp32_peak / 2.0                                                   # ALU pipe: 64 lanes/clk/SM
        ach = 0.75 * cps / 1e12
        return {"kernel": "gk_score_packed_kernel", "bound": "alu_nontensor", "achieved": ach, "peak": peak,
                "unit": "T ALU lane-ops/s (2 VIMNMX.U16x2 + 1 IADD3 per 4 cells)", "frac": ach / peak,
                "cells_per_s": cps, "fp32_nontensor_equiv": fp32, "peak_source": src.format(lanes=64)}
    return {"kernel": "gk_score_kernel", "bound": "fp32_nontensor", "achieved": fp32["achieved"], "peak": fp32_peak,
            "unit": "T FP32 ops/s (2 FADD per cell: d = p - l, acc += |d|)", "frac": fp32["frac"],
            "cells_per_s": cps, "fp32_nontensor_equiv": fp32, "peak_source": src.format(lanes=128)}


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
    ap.add_argument("--cpu-scale", type=float, default=0.5)
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-deep", action="store_true",
                    help="skip the nested cfg4 deep-sample measurement of the default 1-GPU cohort run")
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
    else:
        packs, cns, truth = build_deep(args.deep_reads, args.deep_alleles, args.deep_cn)
        n_samples_total = n_samples_local = 1
        desc = (f"cfg4 deep: {args.deep_reads} read pairs x {args.deep_alleles} alleles, CN {args.deep_cn}, "
                f"top_n={args.top_n}")
    t_build = time.perf_counter() - t_build

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from kir_graph_b200 import cohort, engine
    be = engine.CudaBackend(local_rank)
    col_shard = reduce_scores = None
    if args.workload == "deep" and world > 1:
        col_shard = (rank, world)

        def reduce_scores(d_S):                      # one small collective per copy-number step
            dist.all_reduce(d_S)
    group_size = 17 if args.workload != "deep" else 1
    depth = args.pipeline_depth if args.pipeline_depth > 0 else (4 if n_samples_local < 16 else 3)
    if args.workload == "deep" and args.pipeline_depth <= 0:
        depth = 2                                # 150 ms of saturated kernels per pass: only the host phase to hide
    if col_shard is not None:
        depth = 1                                # one collective stream: passes stay serial
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

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        sm_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        ms_s, work_s, n_s = kernel_stats("gk_score")
        ms_l, work_l, n_l = kernel_stats("gk_likelihood")
        packed = bool(engine.PACKED_DEFAULT)
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
                       "concurrent_sub_batches": n_parts, "e2e_sub_batches": len(e2e_typer.parts),
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
        if world == 1 and args.workload == "cohort" and not args.no_deep:
            # second shape of the same path: one very deep sample (cfg4), scoring kernel dominated
            del typer, e2e_typer, e2e_replicas, replicas
            torch.cuda.empty_cache()
            d_packs, d_cns, d_truth = build_deep(args.deep_reads, args.deep_alleles, args.deep_cn)
            deep = cohort.CohortTyper(d_packs, d_cns, top_n=args.top_n, backend=be, n_parts=1)
            deep.upload()
            for _ in range(3):                       # the third pass records the CUDA graph
                d_calls = deep.run()
            d_steps = 3
            d_ms = timed(deep.run, d_steps) / d_steps
            be.timing = {}
            deep.run_serial()
            torch.cuda.synchronize()
            evs = be.timing.get("gk_score", [])
            be.timing = None
            d_ms_s = sum(a.elapsed_time(b) for a, b, _ in evs)
            d_work = sum(w for _, _, w in evs)
            line["deep"] = {
                "workload": f"cfg4 deep: {args.deep_reads} read pairs x {args.deep_alleles} alleles, CN {args.deep_cn}, "
                            f"top_n={args.top_n}",
                "value": deep.score_cells / (d_ms * 1e-3) / 1e9, "unit": "GCells/s", "ms_per_step": d_ms,
                "steps": d_steps, "calls_match_truth": sorted(d_calls[0].alleles) == d_truth[0],
                "roofline": dict(score_roofline(d_work, d_ms_s, packed, sm_max), launches=len(evs)),
            }
        if world == 1 and not args.no_cpu_baseline:
            cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1), 64))
            c_cells, c_wall, c_n = cpu_reference_step(cores, args.cpu_scale, args.top_n)
            line["cpu_baseline"] = {
                "value": c_cells / c_wall / 1e9, "unit": "GCells/s", "cores": cores, "kind": "port",
                "samples_per_s": c_n * args.cpu_scale / c_wall,
                "sample": f"{cores} processes x 1 synthetic cfg3 sample each at {args.cpu_scale:g} of the reads, "
                          "oracle F64Search (reference NumPy expressions, 1024-read chunks), typing time only"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
