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


# ---- compact work-list of gat_score_compact (include/gat.h): gat_cblock, gat_cjob, gat_cabs
CBLOCK_DTYPE = np.dtype([("size", "<u2"), ("dt", "<u2"), ("dq", "<u2")])
CJOB_DTYPE = np.dtype([("blockPtr", "<u4"), ("tSeq", "<u2"), ("qSeq", "<u2")])
CABS_DTYPE = np.dtype([("tStart", "<i4"), ("qStart", "<i4")])
assert CBLOCK_DTYPE.itemsize == 6 and CJOB_DTYPE.itemsize == 8 and CABS_DTYPE.itemsize == 8
CBLOCK_ABS, CBLOCK_JOINED, CBLOCK_MAX_SIZE, CJOB_MINUS, CGROUP = 0x8000, 0x4000, 0x3FFF, 0x8000, 1024


def pack_compact(jobs, total_job_blocks, blocks):
    """Whole-chain work-list (firstBlock == blockPtr, no clip; records of at most CBLOCK_MAX_SIZE bases: run
    split_long_blocks first) -> (cjobs, cblocks, abs, anchors) of gat_score_compact.  A block is stored as its size and
    the gap in front of it -- the numbers of a .chain file's "size dt dq" lines -- unless it opens a chain or the gap
    does not fit 16 bits; then its start goes to the absolute table."""
    jobs = np.asarray(jobs); blocks = np.asarray(blocks)
    n = int(total_job_blocks)
    if n != len(blocks) or not np.array_equal(jobs["firstBlock"], jobs["blockPtr"]):
        raise ValueError("pack_compact wants jobs that tile the block array")
    if np.any(jobs["clipStart"] != NO_CLIP_START) or np.any(jobs["clipEnd"] != NO_CLIP_END):
        raise ValueError("pack_compact wants unclipped jobs")
    if np.any(jobs["tSeq"] > 0xFFFF) or np.any((jobs["qSeq"] & np.uint32(0x7FFFFFFF)) > 0x7FFF):
        raise ValueError("pack_compact: sequence index beyond 16 bits")
    size = (blocks["size"] & np.uint32(0x7FFFFFFF)).astype(np.int64)
    if np.any(size > CBLOCK_MAX_SIZE):
        raise ValueError("pack_compact: record longer than %d bases" % CBLOCK_MAX_SIZE)
    joined = (blocks["size"] >> np.uint32(31)).astype(bool)
    ts = blocks["tStart"].astype(np.int64); qs = blocks["qStart"].astype(np.int64)
    dt = np.zeros(n, dtype=np.int64); dq = np.zeros(n, dtype=np.int64)
    dt[1:] = ts[1:] - (ts[:-1] + size[:-1]); dq[1:] = qs[1:] - (qs[:-1] + size[:-1])
    is_abs = (dt < 0) | (dt > 0xFFFF) | (dq < 0) | (dq > 0xFFFF)
    counts = job_block_counts(jobs, n)
    is_abs[jobs["blockPtr"][counts > 0].astype(np.int64)] = True      # first block of every chain
    abs_ix = np.cumsum(is_abs) - 1
    cb = np.zeros(n, dtype=CBLOCK_DTYPE)
    cb["size"] = size | np.where(joined, CBLOCK_JOINED, 0) | np.where(is_abs, CBLOCK_ABS, 0)
    cb["dt"] = np.where(is_abs, abs_ix & 0xFFFF, dt)
    cb["dq"] = np.where(is_abs, abs_ix >> 16, dq)
    if is_abs.sum() > 0 and abs_ix[-1] >> 32:
        raise ValueError("pack_compact: more than 2^32 absolute records")
    ab = np.zeros(int(is_abs.sum()), dtype=CABS_DTYPE)
    ab["tStart"] = ts[is_abs]; ab["qStart"] = qs[is_abs]
    anchors = np.zeros((n + CGROUP - 1) // CGROUP, dtype=CABS_DTYPE)
    anchors["tStart"] = ts[::CGROUP]; anchors["qStart"] = qs[::CGROUP]
    cj = np.zeros(len(jobs), dtype=CJOB_DTYPE)
    cj["blockPtr"] = jobs["blockPtr"]; cj["tSeq"] = jobs["tSeq"]
    cj["qSeq"] = (jobs["qSeq"] & np.uint32(0x7FFF)) | np.where(jobs["qSeq"] >> np.uint32(31), CJOB_MINUS, 0).astype(np.uint32)
    return cj, cb, ab, anchors


# ---- 4-byte work-list of gat_score_packed (include/gat.h): gat_pblock words
PBLOCK_MAX_SIZE, PBLOCK_JOINED, PBLOCK_ABS, PBLOCK_MAX_GAP = 0xFFF, 0x1000, 0x2000, 511


def pack_packed(jobs, total_job_blocks, blocks):
    """Whole-chain work-list -> (cjobs, pblocks, abs, anchors, abs_base) of gat_score_packed: one 32-bit word per block (size 12
    bits, JOINED, ABS, dt and dq 9 bits each).  Blocks of more than 4095 bases are cut into JOINED pieces here; a block that opens a
    chain or whose gap is negative or above 511 on either side takes the next entry of the absolute table (no index stored)."""
    jobs, n, blocks = split_long_blocks(jobs, total_job_blocks, blocks, PBLOCK_MAX_SIZE)
    jobs = np.asarray(jobs); blocks = np.asarray(blocks)
    if n != len(blocks) or not np.array_equal(jobs["firstBlock"], jobs["blockPtr"]):
        raise ValueError("pack_packed wants jobs that tile the block array")
    if np.any(jobs["clipStart"] != NO_CLIP_START) or np.any(jobs["clipEnd"] != NO_CLIP_END):
        raise ValueError("pack_packed wants unclipped jobs")
    if np.any(jobs["tSeq"] > 0xFFFF) or np.any((jobs["qSeq"] & np.uint32(0x7FFFFFFF)) > 0x7FFF):
        raise ValueError("pack_packed: sequence index beyond 16 bits")
    size = (blocks["size"] & np.uint32(0x7FFFFFFF)).astype(np.int64)
    joined = (blocks["size"] >> np.uint32(31)).astype(bool)
    ts = blocks["tStart"].astype(np.int64); qs = blocks["qStart"].astype(np.int64)
    dt = np.zeros(n, dtype=np.int64); dq = np.zeros(n, dtype=np.int64)
    dt[1:] = ts[1:] - (ts[:-1] + size[:-1]); dq[1:] = qs[1:] - (qs[:-1] + size[:-1])
    is_abs = (dt < 0) | (dt > PBLOCK_MAX_GAP) | (dq < 0) | (dq > PBLOCK_MAX_GAP)
    counts = job_block_counts(jobs, n)
    is_abs[jobs["blockPtr"][counts > 0].astype(np.int64)] = True      # first block of every chain
    pb = (size | np.where(joined, PBLOCK_JOINED, 0) | np.where(is_abs, PBLOCK_ABS, 0)
          | (np.where(is_abs, 0, dt) << 14) | (np.where(is_abs, 0, dq) << 23)).astype(np.uint32)
    ab = np.zeros(int(is_abs.sum()), dtype=CABS_DTYPE)
    ab["tStart"] = ts[is_abs]; ab["qStart"] = qs[is_abs]
    n_groups = (n + CGROUP - 1) // CGROUP
    anchors = np.zeros(n_groups, dtype=CABS_DTYPE)
    anchors["tStart"] = ts[::CGROUP]; anchors["qStart"] = qs[::CGROUP]
    before = np.concatenate([[0], np.cumsum(is_abs)])                 # absolute records in front of record i
    abs_base = before[np.arange(n_groups) * CGROUP].astype(np.uint32)
    cj = np.zeros(len(jobs), dtype=CJOB_DTYPE)
    cj["blockPtr"] = jobs["blockPtr"]; cj["tSeq"] = jobs["tSeq"]
    cj["qSeq"] = (jobs["qSeq"] & np.uint32(0x7FFF)) | np.where(jobs["qSeq"] >> np.uint32(31), CJOB_MINUS, 0).astype(np.uint32)
    return cj, pb, ab, anchors, abs_base


def unpack_packed(cjobs, pblocks, ab, anchors, abs_base):
    """Host restatement of the device expansion of a packed list (tests) -> (jobs, total, blocks)."""
    n = len(pblocks)
    w = pblocks.astype(np.int64)
    size = w & PBLOCK_MAX_SIZE
    is_abs = (w & PBLOCK_ABS) != 0
    dt = (w >> 14) & 0x1FF; dq = w >> 23
    ts = np.zeros(n, dtype=np.int64); qs = np.zeros(n, dtype=np.int64)
    k = 0
    for i in range(n):
        if i % CGROUP == 0:
            k = int(abs_base[i // CGROUP])
            ts[i], qs[i] = anchors["tStart"][i // CGROUP], anchors["qStart"][i // CGROUP]
        elif is_abs[i]:
            ts[i], qs[i] = ab["tStart"][k], ab["qStart"][k]
        else:
            ts[i] = ts[i - 1] + size[i - 1] + dt[i]; qs[i] = qs[i - 1] + size[i - 1] + dq[i]
        if is_abs[i]:
            k += 1
    blocks = np.zeros(n, dtype=BLOCK_DTYPE)
    blocks["tStart"] = ts; blocks["qStart"] = qs
    blocks["size"] = size.astype(np.uint32) | np.where(w & PBLOCK_JOINED, BLOCK_JOINED, 0).astype(np.uint32)
    jobs = np.zeros(len(cjobs), dtype=JOB_DTYPE)
    jobs["tSeq"] = cjobs["tSeq"]
    jobs["qSeq"] = (cjobs["qSeq"] & 0x7FFF).astype(np.uint32) | np.where(cjobs["qSeq"] & CJOB_MINUS, QSEQ_MINUS, 0).astype(np.uint32)
    jobs["firstBlock"] = cjobs["blockPtr"]; jobs["blockPtr"] = cjobs["blockPtr"]
    jobs["clipStart"] = NO_CLIP_START; jobs["clipEnd"] = NO_CLIP_END
    return jobs, n, blocks


def unpack_compact(cjobs, cblocks, ab, anchors):
    """Host restatement of the device expansion (tests): compact work-list -> (jobs, total, blocks)."""
    n = len(cblocks)
    size = (cblocks["size"] & CBLOCK_MAX_SIZE).astype(np.int64)
    is_abs = (cblocks["size"] & CBLOCK_ABS) != 0
    ix = cblocks["dt"].astype(np.int64) | (cblocks["dq"].astype(np.int64) << 16)
    ts = np.zeros(n, dtype=np.int64); qs = np.zeros(n, dtype=np.int64)
    for i in range(n):
        if i % CGROUP == 0:
            ts[i], qs[i] = anchors["tStart"][i // CGROUP], anchors["qStart"][i // CGROUP]
        elif is_abs[i]:
            ts[i], qs[i] = ab["tStart"][ix[i]], ab["qStart"][ix[i]]
        else:
            ts[i] = ts[i - 1] + size[i - 1] + int(cblocks["dt"][i]); qs[i] = qs[i - 1] + size[i - 1] + int(cblocks["dq"][i])
    blocks = np.zeros(n, dtype=BLOCK_DTYPE)
    blocks["tStart"] = ts; blocks["qStart"] = qs
    blocks["size"] = size.astype(np.uint32) | np.where(cblocks["size"] & CBLOCK_JOINED, BLOCK_JOINED, 0).astype(np.uint32)
    jobs = np.zeros(len(cjobs), dtype=JOB_DTYPE)
    jobs["tSeq"] = cjobs["tSeq"]
    jobs["qSeq"] = (cjobs["qSeq"] & 0x7FFF).astype(np.uint32) | np.where(cjobs["qSeq"] & CJOB_MINUS, QSEQ_MINUS, 0).astype(np.uint32)
    jobs["firstBlock"] = cjobs["blockPtr"]; jobs["blockPtr"] = cjobs["blockPtr"]
    jobs["clipStart"] = NO_CLIP_START; jobs["clipEnd"] = NO_CLIP_END
    return jobs, n, blocks
