"""Multi-GPU decomposition of a work-list: jobs are independent, so a batch is cut
into per-GPU shards by greedy longest-processing-time on aligned bases, every GPU holds a full
genome copy, and the scores come back by job index.  No collective on the data path; the only
communication is the gather of the result vectors (torch.distributed, gloo or nccl)."""
import ctypes
import os
import numpy as np
from .records import JOB_DTYPE, ali_bases, job_block_counts

_HOST = None


def _host():
    global _HOST
    if _HOST is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgathost.so")
        if not os.path.exists(path):
            raise ImportError("%s is missing: run __graft_entry__.build()" % path)
        lib = ctypes.CDLL(path)
        lib.gathost_shard_jobs.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_void_p]
        _HOST = lib
    return _HOST


def assign_jobs(jobs, total_job_blocks, blocks, parts):
    """part[j] = GPU of job j (greedy LPT on aligned bases, gathost::shardJobs)."""
    jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
    ali = np.ascontiguousarray(ali_bases(jobs, total_job_blocks, blocks), dtype=np.int64)
    part = np.zeros(len(jobs), dtype=np.uint32)
    if parts > 1 and len(jobs):
        rc = _host().gathost_shard_jobs(jobs.ctypes.data, len(jobs), int(total_job_blocks), ali.ctypes.data, int(parts),
                                        part.ctypes.data)
        if rc != 0:
            raise RuntimeError("gathost_shard_jobs failed")
    return part, ali


def take_shard(jobs, total_job_blocks, part, rank):
    """The jobs of `rank` with their CSR row pointers rebuilt (block records stay shared)."""
    idx = np.nonzero(part == rank)[0]
    counts = job_block_counts(jobs, total_job_blocks)[idx]
    shard = jobs[idx].copy()
    ptr = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    shard["blockPtr"] = ptr[:-1]
    return idx, shard, int(ptr[-1])


def compact_blocks(shard_jobs, shard_total, blocks):
    """Copy only the records a shard references (one contiguous range per job) and re-point the jobs."""
    counts = job_block_counts(shard_jobs, shard_total)
    first = shard_jobs["firstBlock"].astype(np.int64)
    ptr = shard_jobs["blockPtr"].astype(np.int64)
    within = np.arange(shard_total, dtype=np.int64) - np.repeat(ptr, counts)
    src = np.repeat(first, counts) + within
    out_jobs = shard_jobs.copy()
    out_jobs["firstBlock"] = ptr
    return out_jobs, blocks[src]


def gather_scores(n_jobs, idx, global_scores, local_scores, dist=None, dst=0):
    """Scatter every rank's (idx, global, local) back into job order on rank `dst`."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        g = np.zeros(n_jobs, dtype=np.int64); l = np.zeros(n_jobs, dtype=np.int64)
        g[idx] = global_scores; l[idx] = local_scores
        return g, l
    payload = (np.asarray(idx), np.asarray(global_scores), np.asarray(local_scores))
    gathered = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None, None
    g = np.zeros(n_jobs, dtype=np.int64); l = np.zeros(n_jobs, dtype=np.int64)
    for i, gg, ll in gathered:
        g[i] = gg; l[i] = ll
    return g, l


def exchange_parts(dist, world, device, jobs, blocks):
    """Set-up helper of bench.py at N > 1: every rank generated one part (jobs, blocks) of a synthetic chain set; all ranks end
    up with every part, in rank order (an all_gather of raw bytes, padded to the longest part).  `device`: where the
    collective's tensors live ("cuda" for NCCL, "cpu" for gloo)."""
    import torch
    cols = []
    for arr in (jobs, blocks):
        dtype = arr.dtype
        raw = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(device)
        n = torch.tensor([raw.numel()], dtype=torch.int64, device=device)
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        buf = torch.zeros(max(sizes), dtype=torch.uint8, device=device)
        buf[:raw.numel()] = raw
        got = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(got, buf)
        cols.append([g[:k].cpu().numpy().view(dtype).copy() for g, k in zip(got, sizes)])
    return list(zip(*cols))
