"""CPU tests of the multi-GPU host logic: block-column partition, sub-problem patterns, assembling X.
world_size-2 processes over gloo; the per-rank solver is the ORACLE here (no GPU in this container),
the product's ShardedBsrsv uses exactly the same ShardSpec/scatter logic on top of the C-ABI."""
import os
import socket

import numpy as np
import pytest

import orclib as O
from tfqmrgpu_b200 import problems as P, sharded as S


def test_partition_is_contiguous_balanced_and_complete():
    rng = np.random.default_rng(0)
    for ncols, world in [(1, 1), (1, 2), (2, 2), (7, 2), (16, 4), (16, 8), (5, 8), (33, 8)]:
        colindx = rng.integers(0, ncols, size=500)
        colindx[:ncols] = np.arange(ncols)
        ranges = S.partition_columns(colindx, ncols, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == ncols
        for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 == b0 and a0 <= a1
        nonempty = sum(1 for a, b in ranges if b > a)
        assert nonempty == min(world, ncols)


def test_shard_patterns_cover_the_problem_once():
    prob = P.random_system(14, 4, 4, ncols=9, seed=5, unsorted=True)
    seen = np.zeros(prob.X.nnzb, int)
    for r in range(3):
        sp = S.ShardSpec(prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, r, 3)
        seen[sp.selX] += 1
        assert sp.rpX[-1] == sp.selX.size == sp.ciX.size and sp.rpB[-1] == sp.selB.size
        # every kept B block still finds its X block in the same row
        rowsX = np.repeat(np.arange(prob.mb), np.diff(sp.rpX)); rowsB = np.repeat(np.arange(prob.mb), np.diff(sp.rpB))
        for rb, cb in zip(rowsB, sp.ciB):
            assert np.any((rowsX == rb) & (sp.ciX == cb))
    assert np.all(seen == 1)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    prob = P.random_system(12, 4, 8, ncols=6, seed=9, unsorted=True)
    vA = P.interleave(prob.A.val, np.float64); vB = P.interleave(prob.B.val, np.float64)
    A_int = O.import_blocks(vA, prob.A.nnzb, 4, 4, var="A")
    B_int = O.import_blocks(vB, prob.B.nnzb, 4, 8, var="B")
    n3 = prob.X.nnzb*2*4*8
    v3_global = np.random.default_rng(1).random(n3).astype(np.float32).reshape(prob.X.nnzb, -1)
    sp = S.ShardSpec(prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, rank, world)
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
    assert pl.status == 0
    o = O.solve(pl, 4, 8, A_int, B_int[sp.selB], v3_global[sp.selX], 1e-9, 100)
    # gather: pad to the largest shard, all_gather, scatter into the caller's order (same as gather_x)
    counts = [S.ShardSpec(prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, r, world).selX.size for r in range(world)]
    nmax = max(counts)
    pad = torch.zeros((nmax, 2, 4, 8), dtype=torch.float64); pad[:sp.selX.size] = torch.from_numpy(o["X"])
    allx = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allx, pad)
    its = torch.tensor([o["iterations"]]); dist.all_reduce(its, op=dist.ReduceOp.MAX)
    sels = [S.ShardSpec(prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind, r, world).selX for r in range(world)]
    Xg = S.scatter_shards([allx[r][:counts[r]].numpy() for r in range(world)], sels, prob.X.nnzb)
    if rank == 0:
        q.put((Xg, int(its.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_rank():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    Xg, its = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-rank solve of the whole problem with the same global v3
    prob = P.random_system(12, 4, 8, ncols=6, seed=9, unsorted=True)
    vA = P.interleave(prob.A.val, np.float64); vB = P.interleave(prob.B.val, np.float64)
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    v3 = np.random.default_rng(1).random(prob.X.nnzb*2*4*8).astype(np.float32)
    o = O.solve(pl, 4, 8, O.import_blocks(vA, prob.A.nnzb, 4, 4, var="A"), O.import_blocks(vB, prob.B.nnzb, 4, 8, var="B"),
                v3, 1e-9, 100)
    assert o["status"] == 0 and abs(its - o["iterations"]) <= 1
    # columns are independent: shards reproduce the 1-rank columns up to the (possibly earlier) stop
    assert np.abs(Xg - o["X"]).max() <= 10*1e-9*np.abs(o["X"]).max()
