"""One deep problem spread over two GPUs with NCCL (``python -m pytest -m gpu`` on a box with >= 2 GPUs,
e.g. ``gpurun --gpus 2``): read sharding and candidate-column sharding of ``engine.SearchGroup`` through the
CUDA kernels against the oracle's exact integer search, step by step (ids, scores, fraction numerators,
N_uniq, tie flags), and the batched typer's calls against the unsharded run.  The host-side sharding logic
alone is covered on CPU by tests/test_distributed_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SPEC = dict(a=300, v=2400, cn=4, r=20000, top_n=100)


def _gene():
    from kir_graph_b200 import packing, synthetic
    gene = synthetic.make_gene([31, 7], "KIRNCCL*BACKBONE", SPEC["a"], SPEC["v"], SPEC["cn"], SPEC["r"], homo_prob=0.0)
    return packing.pack_synthetic(gene)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from kir_graph_b200 import cohort, engine, packing
        pack = _gene()
        be = engine.CudaBackend(rank)

        def reduce_scores(t):
            dist.all_reduce(t)

        cn, top_n = SPEC["cn"], SPEC["top_n"]
        dump = lambda steps: [(s.ids.tolist(), s.score.tolist(), s.cnt.tolist(), s.tie_flags, s.n_unique, s.cut)
                              for s in steps]
        # reads sharded: nothing replicated
        part = packing.shard_reads(pack, rank, world)
        rb = engine.MatrixBatch([part], backend=be, reduce=reduce_scores)
        rg = engine.SearchGroup(rb, [0], top_n, reduce_scores=reduce_scores, read_shard=True)
        by_reads = dump([rg.step(need_next=[i + 1 < cn])[0] for i in range(cn)])
        colsum = rb.colsum(0).tolist()
        # candidate columns sharded: L / LT / P replicated
        cb = engine.MatrixBatch([pack], backend=be)
        cg = engine.SearchGroup(cb, [0], top_n, col_shard=(rank, world), reduce_scores=reduce_scores)
        by_cols = dump([cg.step(need_next=[i + 1 < cn])[0] for i in range(cn)])
        # batched typer, pipelined (every launch and collective enqueued up front, one read-back)
        typer = cohort.BatchTyper([part], [cn], top_n=top_n, backend=be, reduce_scores=reduce_scores, read_shard=True)
        calls = [(c.alleles, c.score, c.value, c.tie_flags, c.best_rank, c.n_reads) for c in typer.run()]
        whole = None
        if rank == 0:
            wt = cohort.BatchTyper([pack], [cn], top_n=top_n, backend=be)
            whole = [(c.alleles, c.score, c.value, c.tie_flags, c.best_rank, c.n_reads) for c in wt.run()]
        dist.barrier()
        out.put(("ok", rank, by_reads, by_cols, colsum, calls, whole, be.launches))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put(("error", rank, traceback.format_exc() + repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_deep_problem_sharded_over_two_gpus_equals_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from kir_graph_b200 import engine
    from oracle import typing_oracle as orc
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in results), results
    pack = _gene()
    m = engine.MatrixBatch([pack], backend=engine.CudaBackend(0)).mismatch_counts(0).astype(np.int64)
    search = orc.IntSearch(m, pack.k_obs, top_n=SPEC["top_n"])
    ref = [search.add_candidate() for _ in range(SPEC["cn"])]
    for res in results:
        _, rank, by_reads, by_cols, colsum, calls, whole, launches = res
        assert launches > 0
        assert colsum == m.sum(axis=0).tolist()
        for steps in (by_reads, by_cols):
            for (ids, score, cnt, flags, n_unique, cut), want in zip(steps, ref):
                n = len(ids[0])
                assert ids == want.allele_id.tolist() and score == want.score.tolist()
                w = np.array([orc.lcm_upto(n) // q for q in range(1, n + 1)])
                assert ((np.array(cnt).reshape(len(ids), n, n) * w[None, None, :]).sum(axis=2) == want.frac_num).all()
                assert n_unique == want.n_unique
        assert by_reads == by_cols
    assert results[0][5] == results[1][5]                       # both ranks form the same calls ...
    whole = [r[6] for r in results if r[6] is not None][0]
    assert results[0][5] == whole                                # ... equal to the unsharded run
