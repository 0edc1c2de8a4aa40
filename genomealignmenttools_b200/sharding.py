"""Multi-GPU decomposition of a work-list: jobs are independent, so a batch is cut
into per-GPU shards by greedy longest-processing-time on aligned bases, every GPU holds a full
genome copy, and the scores come back by job index.  No collective on the data path; the only
communication is the gather of the result vectors (torch.distributed, gloo or nccl)."""
import ctypes
import os
import numpy as np
from .records import JOB_DTYPE, BLOCK_JOINED, NO_CLIP_START, NO_CLIP_END, ali_bases, job_block_counts

# gat_tuple (include/gat.h): a part of a chain as a map on the local-score state
TUPLE_DTYPE = np.dtype([("d", "<i8"), ("c", "<i8"), ("e", "<i8"), ("f", "<i8")])
TUPLE_MIN_BLOCKS = 1024       # GAT_TUPLE_MIN_BLOCKS

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


def split_giant_jobs(jobs, total_job_blocks, blocks, parts, ali=None, share=None):
    """a whole-chain job whose aligned bases exceed total / (4 parts) is cut at block boundaries into
    pieces of about that many bases, so that greedy balancing can spread it over the GPUs.  A piece never starts with a
    JOINED record and never holds fewer than TUPLE_MIN_BLOCKS blocks (its tuple comes from the fix-up kernel).
    Returns (piece_jobs, origin, first_piece): piece_jobs is a job list in which every giant job is replaced by its
    pieces (same records, same order; blockPtr rebuilt), origin[k] = index of the job piece k belongs to, first_piece[j]
    = index of the first piece of job j (+ sentinel).  Jobs that clip or share records are never cut."""
    jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
    counts = job_block_counts(jobs, total_job_blocks)
    if ali is None:
        ali = ali_bases(jobs, total_job_blocks, blocks)
    limit = int(share) if share else max(1, int(ali.sum()) // (4 * max(1, parts)))
    whole = (jobs["clipStart"] == NO_CLIP_START) & (jobs["clipEnd"] == NO_CLIP_END)
    giant = np.nonzero(whole & (ali > limit) & (counts >= 2 * TUPLE_MIN_BLOCKS))[0] if parts > 1 else []
    pieces_of = {}
    for j in giant:
        fb, n = int(jobs["firstBlock"][j]), int(counts[j])
        size = (blocks["size"][fb:fb + n] & np.uint32(0x7FFFFFFF)).astype(np.int64)
        joined = (blocks["size"][fb:fb + n] & np.uint32(BLOCK_JOINED)) != 0
        csum = np.cumsum(size)
        k = int(np.ceil(csum[-1] / limit))
        cuts = [0]
        for i in range(1, k):
            c = int(np.searchsorted(csum, csum[-1] * i // k))      # first block of the next piece
            while c < n and joined[c]:
                c += 1
            if c - cuts[-1] >= TUPLE_MIN_BLOCKS and n - c >= TUPLE_MIN_BLOCKS:
                cuts.append(c)
        if len(cuts) > 1:
            pieces_of[int(j)] = cuts
    n_pieces = len(jobs) + sum(len(c) - 1 for c in pieces_of.values())
    out = np.zeros(n_pieces, dtype=JOB_DTYPE)
    origin = np.zeros(n_pieces, dtype=np.int64)
    first_piece = np.zeros(len(jobs) + 1, dtype=np.int64)
    # number of pieces per job -> positions
    npj = np.ones(len(jobs), dtype=np.int64)
    for j, cuts in pieces_of.items():
        npj[j] = len(cuts)
    np.cumsum(npj, out=first_piece[1:])
    origin[:] = np.repeat(np.arange(len(jobs)), npj)
    out[:] = jobs[origin]
    piece_counts = counts[origin].copy()
    for j, cuts in pieces_of.items():
        p0 = first_piece[j]
        ends = cuts[1:] + [int(counts[j])]
        for i, (a, b) in enumerate(zip(cuts, ends)):
            out["firstBlock"][p0 + i] = jobs["firstBlock"][j] + a
            piece_counts[p0 + i] = b - a
    ptr = np.zeros(n_pieces + 1, dtype=np.int64)
    np.cumsum(piece_counts, out=ptr[1:])
    out["blockPtr"] = ptr[:-1]
    return out, origin, first_piece


def join_pieces(jobs, total_job_blocks, blocks, piece_jobs, origin, first_piece, piece_global, piece_local, piece_tuples, gap_cost):
    """Scores of the original jobs from the scores of their pieces: a job that was not cut takes its piece's scores, a cut
    one joins the tuples of its pieces in order (gat_tuple_join) with the gap cost between neighbouring pieces, which is
    gapCalcCost of the two blocks either side of the cut (gap_cost(dq, dt), e.g. Scoring.gap.cost).
    piece_tuples: {piece index: TUPLE_DTYPE record} for the pieces of cut jobs."""
    from . import _native
    lib = _native.load()
    n = len(first_piece) - 1
    g = np.zeros(n, dtype=np.int64); l = np.zeros(n, dtype=np.int64)
    single = np.diff(first_piece) == 1
    g[single] = piece_global[first_piece[:-1][single]]
    l[single] = piece_local[first_piece[:-1][single]]
    counts = job_block_counts(piece_jobs, total_job_blocks)
    for j in np.nonzero(~single)[0]:
        p0, p1 = int(first_piece[j]), int(first_piece[j + 1])
        acc = np.array(piece_tuples[p0], dtype=TUPLE_DTYPE).reshape(1).copy()
        for p in range(p0 + 1, p1):
            prev_last = blocks[int(piece_jobs["firstBlock"][p - 1]) + int(counts[p - 1]) - 1]
            first = blocks[int(piece_jobs["firstBlock"][p])]
            psz = int(prev_last["size"] & np.uint32(0x7FFFFFFF))
            dq = int(first["qStart"]) - (int(prev_last["qStart"]) + psz)
            dt = int(first["tStart"]) - (int(prev_last["tStart"]) + psz)
            nxt = np.array(piece_tuples[p], dtype=TUPLE_DTYPE).reshape(1).copy()
            lib.gat_tuple_join(acc.ctypes.data, int(gap_cost(dq, dt)), nxt.ctypes.data)
        gg = ctypes.c_int64(); ll = ctypes.c_int64()
        lib.gat_tuple_scores(acc.ctypes.data, ctypes.byref(gg), ctypes.byref(ll))
        g[j], l[j] = gg.value, ll.value
    return g, l
