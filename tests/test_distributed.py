"""The N>1 path on CPU: two gloo ranks shard one work-list, 'score' their shards with a stand-in
that only depends on the records of each job, and rank 0 reassembles the vectors by job index."""
import os
import socket
import sys
import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genomealignmenttools_b200 import sharding, synth  # noqa: E402
from genomealignmenttools_b200.records import ali_bases, job_block_counts  # noqa: E402


def stand_in(jobs, total, blocks):
    """Deterministic per-job digest of exactly the records the job references."""
    counts = job_block_counts(jobs, total)
    job_of = np.repeat(np.arange(len(jobs)), counts)
    within = np.arange(total) - np.repeat(jobs["blockPtr"].astype(np.int64), counts)
    b = blocks[jobs["firstBlock"].astype(np.int64)[job_of] + within]
    h = b["tStart"].astype(np.int64) * 3 + b["qStart"].astype(np.int64) * 5 + b["size"].astype(np.int64) * 7 + within
    g = np.zeros(len(jobs), dtype=np.int64)
    np.add.at(g, job_of, h)
    return g, g // 2 + jobs["tSeq"].astype(np.int64)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs, total, blocks = synth.make_chains([3_000_000, 1_000_000], [2_500_000, 900_000], 60_000, seed=77)
    part, ali = sharding.assign_jobs(jobs, total, blocks, world)
    idx, shard, shard_total = sharding.take_shard(jobs, total, part, rank)
    cj, cb = sharding.compact_blocks(shard, shard_total, blocks)
    assert np.array_equal(ali_bases(cj, shard_total, cb), ali[idx])      # the compacted shard is the same work
    g, l = stand_in(cj, shard_total, cb)
    G, L = sharding.gather_scores(len(jobs), idx, g, l, dist)
    if rank == 0:
        wg, wl = stand_in(jobs, total, blocks)
        np.save(out, np.array([np.array_equal(G, wg), np.array_equal(L, wl), len(idx), len(jobs)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ok_g, ok_l, n0, n = np.load(out)
    assert ok_g and ok_l
    assert 0 < n0 < n


def _exchange_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = ([3_000_000, 1_000_000], [2_500_000, 900_000])
    jobs, total, blocks = synth.make_chains(*sizes, 20_000 + 7_000 * rank, seed=100 + rank)      # parts of different length
    parts = sharding.exchange_parts(dist, world, "cpu", jobs, blocks)
    ok = len(parts) == world
    for r, (pj, pb) in enumerate(parts):
        wj, wt, wb = synth.make_chains(*sizes, 20_000 + 7_000 * r, seed=100 + r)
        ok = ok and np.array_equal(pj, wj) and np.array_equal(pb, wb)
    np.save(out % rank, np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_of_workload_parts(tmp_path):
    """bench.py at N > 1: each rank generates one part of the chain set and the ranks swap them."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "ok%d.npy")
    mp.spawn(_exchange_worker, args=(2, port, out), nprocs=2, join=True)
    assert np.load(out % 0)[0] and np.load(out % 1)[0]


def test_assign_single_gpu_is_identity():
    jobs, total, blocks = synth.make_chains([300_000], [250_000], 5_000, seed=5)
    part, ali = sharding.assign_jobs(jobs, total, blocks, 1)
    assert part.max() == 0
    idx, shard, st = sharding.take_shard(jobs, total, part, 0)
    assert np.array_equal(shard, jobs) and st == total
    g, l = sharding.gather_scores(len(jobs), idx, ali, ali)
    assert np.array_equal(g, ali)
