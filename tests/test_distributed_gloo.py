"""Multi-rank paths on CPU (gloo, world_size 2): candidate-column sharding with one score
reduction per step, and cohort sample sharding.  Kernels are the NumPy test double; this
covers the host-side sharding logic only (the GPU path runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kir_graph_b200 import cohort, engine, packing, synthetic
from oracle import typing_oracle as orc
from tests.fake_backend import FakeBackend


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- one deep problem, candidate columns sharded -------------------------------
        gene = synthetic.make_gene([21, 0], "KIRSH*BACKBONE", 300, 2400, 3, 400, homo_prob=0.0)
        pack = packing.pack_synthetic(gene)
        be = FakeBackend()
        batch = engine.MatrixBatch([pack], backend=be)

        def reduce_scores(d_S):
            t = torch.from_numpy(d_S.view(np.int32))
            dist.all_reduce(t)

        group = engine.SearchGroup(batch, [0], 40, col_shard=(rank, world), reduce_scores=reduce_scores)
        steps = [group.step(need_next=[i < 2])[0] for i in range(3)]
        cells = torch.tensor([group.score_cells], dtype=torch.int64)
        dist.all_reduce(cells)
        # --- the same problem with its READS sharded: nothing replicated, two reductions per step ----
        part = packing.shard_reads(pack, rank, world)
        rbatch = engine.MatrixBatch([part], backend=be, reduce=reduce_scores)
        rgroup = engine.SearchGroup(rbatch, [0], 40, reduce_scores=reduce_scores, read_shard=True)
        rsteps = [rgroup.step(need_next=[i < 2])[0] for i in range(3)]
        rcells = torch.tensor([rgroup.score_cells], dtype=torch.int64)
        dist.all_reduce(rcells)
        read_sharded = ([(s.ids.tolist(), s.score.tolist(), s.cnt.tolist(), s.tie_flags, s.n_unique) for s in rsteps],
                        int(rcells.item()), rbatch.colsum(0).tolist(), part.n_reads)
        # ... and through the batched typer (pipelined run, one read-back)
        rcalls = cohort.BatchTyper([part], [3], top_n=40, backend=FakeBackend(), reduce_scores=reduce_scores,
                                   read_shard=True).run()
        read_sharded += ([(c.alleles, c.score, c.value, c.tie_flags, c.n_reads) for c in rcalls],)
        # --- cohort: samples dealt round-robin, results gathered on rank 0 -----------------
        samples = list(range(4))
        mine = cohort.shard(samples, rank, world)
        calls = {}
        for sidx in mine:
            genes = synthetic.make_wgs30x_sample(seed=200 + sidx, total_reads=1500)[:5]
            packs = [packing.pack_synthetic(g) for g in genes]
            res = cohort.BatchTyper(packs, [g.cn for g in genes], top_n=20, backend=FakeBackend()).run()
            calls[sidx] = [c.alleles for c in res]
        gathered = [None] * world
        dist.all_gather_object(gathered, calls)
        if rank == 0:
            merged = {}
            for part in gathered:
                merged.update(part)
            out.put(("ok", [(s.ids.tolist(), s.score.tolist()) for s in steps], int(cells.item()), merged, read_sharded))
        else:
            out.put(("ok", [(s.ids.tolist(), s.score.tolist()) for s in steps], -1, None, read_sharded))
    except Exception as exc:  # pragma: no cover
        out.put(("error", repr(exc), 0, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_column_sharding_and_cohort_sharding_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in results), results
    # every rank ends with identical kept sets, equal to the unsharded oracle
    gene = synthetic.make_gene([21, 0], "KIRSH*BACKBONE", 300, 2400, 3, 400, homo_prob=0.0)
    pack = packing.pack_synthetic(gene)
    m = engine.MatrixBatch([pack], backend=FakeBackend()).mismatch_counts(0)
    search = orc.IntSearch(m.astype(np.int64), pack.k_obs, top_n=40)
    ref = [search.add_candidate() for _ in range(3)]
    for res in results:
        for (ids, score), want in zip(res[1], ref):
            assert ids == want.allele_id.tolist() and score == want.score.tolist()
    # read sharding: both ranks hold the complete, identical step results of the unsharded search
    assert sorted(r[4][3] for r in results) == [pack.n_reads // 2, pack.n_reads - pack.n_reads // 2]
    for res in results:
        steps, cells, colsum, _, calls = res[4]
        assert colsum == m.astype(np.int64).sum(axis=0).tolist()
        for (ids, score, cnt, flags, n_unique), want in zip(steps, ref):
            assert ids == want.allele_id.tolist() and score == want.score.tolist()
            n = len(ids[0])
            w = np.array([orc.lcm_upto(n) // q for q in range(1, n + 1)])
            assert ((np.array(cnt).reshape(len(ids), n, n) * w[None, None, :]).sum(axis=2) == want.frac_num).all()
            assert n_unique == want.n_unique
        assert cells == sum(len(ref[i].score) * pack.n_alleles * pack.n_reads for i in range(2))
        whole = cohort.BatchTyper([pack], [3], top_n=40, backend=FakeBackend()).run()
        assert calls == [(c.alleles, c.score, c.value, c.tie_flags, c.n_reads) for c in whole]
    total_cells = max(r[2] for r in results)
    assert total_cells == sum(len(ref[i].score) * pack.n_alleles * pack.n_reads for i in range(2))
    merged = [r[3] for r in results if r[3] is not None][0]
    assert sorted(merged) == [0, 1, 2, 3]
    for sidx, got in merged.items():
        genes = synthetic.make_wgs30x_sample(seed=200 + sidx, total_reads=1500)[:5]
        packs = [packing.pack_synthetic(g) for g in genes]
        want = cohort.BatchTyper(packs, [g.cn for g in genes], top_n=20, backend=FakeBackend()).run()
        assert got == [c.alleles for c in want]


def _cohort_worker(rank, world, port, folder, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        from kir_graph_b200 import main
        os.chdir(folder)
        names = json.load(open("names.json"))
        files = main.cohortAlleleTyping(names, [n + ".depth.cn.tsv" for n in names], "full", rank=dist.get_rank(),
                                        world=dist.get_world_size(), _backend=FakeBackend())
        dist.barrier()                       # every rank has written its samples' files
        if rank == 0:
            main.mergeAllele(files, "cohort.allele.tsv")
        out.put(("ok", files))
    except Exception as exc:  # pragma: no cover
        out.put(("error", repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_cohort_entry_world2_merges_the_references_file(tmp_path):
    """main.cohortAlleleTyping on two ranks (samples 0, 2 / sample 1), rank 0 merges after a barrier:
    cohort.allele.tsv equals the file the reference's alleleTyping + mergeAllele wrote
    (tests/golden/main_tsv.json.gz)."""
    import json
    import pandas as pd
    from kir_graph_b200.hisat2 import PairRead, writeReadsAndVariantsData
    from kir_graph_b200.msa2hisat import Variant
    from tests.helpers import load_golden
    data = load_golden("main_tsv")
    names = []
    for inp in data["inputs"]:
        base = str(tmp_path / inp["name"])
        writeReadsAndVariantsData({"variants": [Variant(**v) for v in inp["variants"]],
                                   "reads": [PairRead(**r) for r in inp["reads"]]}, base + ".json")
        pd.DataFrame({"gene": list(inp["cn"]), "cn": list(inp["cn"].values())}).to_csv(
            base + ".depth.cn.tsv", sep="\t", index=False)
        names.append(inp["name"])
    json.dump(names, open(tmp_path / "names.json", "w"))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cohort_worker, args=(r, 2, port, str(tmp_path), out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in results), results
    want = data["methods"]["full"]
    assert all(r[1] == want["files"] for r in results)
    assert open(tmp_path / "cohort.allele.tsv").read() == want["merged"]
    assert [open(tmp_path / f).read() for f in want["files"]] == want["tsv"]
