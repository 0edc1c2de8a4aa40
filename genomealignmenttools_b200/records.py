"""Work-list record layouts, byte-identical to include/gat.h (gat_block, gat_job, gat_nrun)."""
import numpy as np

BLOCK_DTYPE = np.dtype([("tStart", "<i4"), ("qStart", "<i4"), ("size", "<u4")])
JOB_DTYPE = np.dtype([("tSeq", "<u4"), ("qSeq", "<u4"), ("firstBlock", "<u4"), ("blockPtr", "<u4"),
                      ("clipStart", "<i4"), ("clipEnd", "<i4")])
NRUN_DTYPE = np.dtype([("seq", "<u4"), ("start", "<u4"), ("len", "<u4")])
assert BLOCK_DTYPE.itemsize == 12 and JOB_DTYPE.itemsize == 24 and NRUN_DTYPE.itemsize == 12

QSEQ_MINUS = 0x80000000
BLOCK_JOINED = 0x80000000
NO_CLIP_START = -(2 ** 31)
NO_CLIP_END = 2 ** 31 - 1


def jobs_from_counts(t_seq, q_seq, minus, n_blocks, first_block=None, clip_start=None, clip_end=None):
    """Build a JOB_DTYPE array from per-job columns; blockPtr is the exclusive prefix sum."""
    n_blocks = np.asarray(n_blocks, dtype=np.int64)
    jobs = np.zeros(len(n_blocks), dtype=JOB_DTYPE)
    ptr = np.zeros(len(n_blocks) + 1, dtype=np.int64)
    np.cumsum(n_blocks, out=ptr[1:])
    jobs["tSeq"] = t_seq
    jobs["qSeq"] = np.asarray(q_seq, dtype=np.uint32) | (np.asarray(minus, dtype=np.uint32) << np.uint32(31))
    jobs["blockPtr"] = ptr[:-1]
    jobs["firstBlock"] = ptr[:-1] if first_block is None else first_block
    jobs["clipStart"] = NO_CLIP_START if clip_start is None else clip_start
    jobs["clipEnd"] = NO_CLIP_END if clip_end is None else clip_end
    return jobs, int(ptr[-1])


def job_block_counts(jobs, total_job_blocks):
    ptr = np.append(jobs["blockPtr"].astype(np.int64), np.int64(total_job_blocks))
    return np.diff(ptr)


def ali_bases(jobs, total_job_blocks, blocks):
    """aliBases of chainCalcScoreLocal (src/scoreChain/scoreChain.c:180-182): the sum of clipped
    block sizes per job.  Plain host arithmetic on the work-list, as SURVEY.md 8b prescribes."""
    counts = job_block_counts(jobs, total_job_blocks)
    out = np.zeros(len(jobs), dtype=np.int64)
    if len(jobs) == 0 or counts.sum() == 0:
        return out
    job_of = np.repeat(np.arange(len(jobs)), counts)
    within = np.arange(counts.sum()) - np.repeat(jobs["blockPtr"].astype(np.int64), counts)
    idx = jobs["firstBlock"].astype(np.int64)[job_of] + within
    ts = blocks["tStart"].astype(np.int64)[idx]
    te = ts + (blocks["size"][idx] & np.uint32(0x7FFFFFFF)).astype(np.int64)
    ts = np.maximum(ts, jobs["clipStart"].astype(np.int64)[job_of])
    te = np.minimum(te, jobs["clipEnd"].astype(np.int64)[job_of])
    np.add.at(out, job_of, te - ts)
    return out


SPLIT_BASES = 4096   # include/gat.h GAT_SPLIT_BASES


def split_long_blocks(jobs, total_job_blocks, blocks, max_bases=SPLIT_BASES):
    """Cut blocks longer than max_bases into JOINED records (what gathost::buildRecords does for the
    tools).  Only for work-lists whose jobs tile the block array (firstBlock == blockPtr)."""
    jobs = np.asarray(jobs); blocks = np.asarray(blocks)
    assert np.array_equal(jobs["firstBlock"], jobs["blockPtr"]), "jobs must tile the block array"
    size = (blocks["size"] & np.uint32(0x7FFFFFFF)).astype(np.int64)
    pieces = np.maximum(1, (size + max_bases - 1) // max_bases)
    if pieces.max() == 1:
        return jobs, total_job_blocks, blocks
    first = np.zeros(len(blocks) + 1, dtype=np.int64)
    np.cumsum(pieces, out=first[1:])
    src = np.repeat(np.arange(len(blocks)), pieces)
    k = np.arange(first[-1]) - first[src]
    out = np.zeros(first[-1], dtype=BLOCK_DTYPE)
    out["tStart"] = blocks["tStart"][src] + k * max_bases
    out["qStart"] = blocks["qStart"][src] + k * max_bases
    out["size"] = np.minimum(max_bases, size[src] - k * max_bases).astype(np.uint32) | np.where(k > 0, np.uint32(BLOCK_JOINED), np.uint32(0))
    new_jobs = jobs.copy()
    new_jobs["firstBlock"] = first[jobs["firstBlock"].astype(np.int64)]
    new_jobs["blockPtr"] = new_jobs["firstBlock"]
    return new_jobs, int(first[-1]), out

# gat_xpair (include/gat.h): one pair of overlapping adjacent blocks for gat_crossover
XPAIR_DTYPE = np.dtype([("tSeq", "<u4"), ("qSeq", "<u4"), ("leftTEnd", "<i4"), ("leftQEnd", "<i4"),
                        ("rightTStart", "<i4"), ("rightQStart", "<i4"), ("overlap", "<i4")])
