"""Multi-GPU host logic on CPU: read sharding across ranks (world_size 2 over gloo). The oracle
stands in for the per-rank engine; what is tested is that shards are contiguous, balanced, cover
every read once and that rank-local results land in the right output slices."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import golden, read_fasta_reads
from sbwt_b200.sharding import output_bounds, shard_batch, shard_bounds
from sbwt_b200.testing import synth


def test_shard_bounds_properties():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 400, size=1000)
    off = np.concatenate([[7], 7 + np.cumsum(lens)]).astype(np.int64)
    for world in (1, 2, 3, 8, 64):
        b = shard_bounds(off, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        bases = [off[r1] - off[r0] for r0, r1 in b]
        assert max(bases) - min(bases) <= 2 * 400 or world > 8
        ob = output_bounds(off, 31, b)
        assert ob[0][0] == 0 and all(ob[i][1] == ob[i + 1][0] for i in range(world - 1))
        assert ob[-1][1] == int(np.maximum(lens - 30, 0).sum())
    assert shard_bounds(np.zeros(1, np.int64), 4) == [(0, 0)] * 4


def _worker(rank, world, port, index_path, reads, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, off = synth.ragged_to_batch(reads)
    idx = oracle.OracleIndex(index_path)
    sa, so, (r0, r1) = shard_batch(a, off, rank, world)
    mine = idx.query_batch(sa, so, streaming=True)
    ob = output_bounds(off, idx.k, shard_bounds(off, world))
    out = torch.full((ob[-1][1],), -9, dtype=torch.int64)
    out[ob[rank][0]:ob[rank][1]] = torch.from_numpy(mine)
    # results are disjoint slices: a max-reduce assembles them (test-only; the product writes to disjoint host ranges)
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    t = torch.tensor([float(r1 - r0)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["out"] = out.numpy().copy()
        ret["reads"] = int(t.item())
    dist.destroy_process_group()


def test_two_ranks_gloo():
    reads = read_fasta_reads(golden("small_k31", "reads.fna"))
    index_path = golden("small_k31", "index.sbwt")
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, index_path, reads, ret), nprocs=2, join=True)
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(index_path).query_batch(a, off, streaming=True)
    np.testing.assert_array_equal(ret["out"], want)
    assert ret["reads"] == len(reads)
