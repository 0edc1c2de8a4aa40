"""Synthetic workloads at real chrom.sizes: random .2bit payloads, heavy-tailed
chain sets with both strands, N runs, and a query made homologous along the chain blocks.

Used by tests (small) and bench.py (genome-wide).  Everything is seeded and reproducible."""
import ctypes
import numpy as np
from . import _native
from .records import BLOCK_DTYPE, JOB_DTYPE, NRUN_DTYPE, NO_CLIP_START, NO_CLIP_END
from .twobit import PackedGenome

# example/hg38.chrom.sizes, example/mm10.chrom.sizes: the primary assemblies (the reference ships
# the full lists incl. unplaced scaffolds; bench.py reads those from tests/golden/*.chrom.sizes)


def read_chrom_sizes(path):
    names, sizes = [], []
    with open(path) as f:
        for line in f:
            w = line.split()
            if len(w) >= 2:
                names.append(w[0]); sizes.append(int(w[1]))
    return names, np.asarray(sizes, dtype=np.int64)


def random_genome(names, sizes, seed, telomere_n=0):
    """Uniform random bases for every sequence + an N run of `telomere_n` bases at both ends."""
    synth = _native.load_synth()
    sizes = np.asarray(sizes, dtype=np.int64)
    nbytes = (sizes + 3) // 4
    offs = np.zeros(len(sizes), dtype=np.uint64)
    offs[1:] = np.cumsum(nbytes)[:-1]
    packed = np.empty(int(nbytes.sum()), dtype=np.uint8)
    synth.gat_synth_fill(packed.ctypes.data, packed.size, int(seed))
    runs = []
    if telomere_n:
        for i, s in enumerate(sizes):
            k = int(min(telomere_n, s // 4))
            if k:
                runs.append((i, 0, k)); runs.append((i, int(s) - k, k))
    n_runs = np.array(runs, dtype=NRUN_DTYPE) if runs else None
    return PackedGenome(names, sizes, packed, offs, n_runs)


def _chain_batch(rng, t_sizes, q_sizes, n_blocks_target, mean_log_len, sigma_log_len, max_len, zipf_s,
                 max_chain_blocks, minus_fraction, gap_mu, gap_sigma, max_gap, double_gap_fraction):
    # blocks per chain: Zipf, capped
    counts = []
    have = 0
    while have < n_blocks_target:
        c = np.minimum(rng.zipf(zipf_s, size=max(1024, n_blocks_target // 8)), max_chain_blocks)
        counts.append(c); have += int(c.sum())
    counts = np.concatenate(counts).astype(np.int64)
    cut = int(np.searchsorted(np.cumsum(counts), n_blocks_target)) + 1
    counts = counts[:cut]
    n_chains = len(counts)
    nb = int(counts.sum())
    size = np.minimum(np.ceil(rng.lognormal(mean_log_len, sigma_log_len, nb)), max_len).astype(np.int64)
    gap = np.clip(np.ceil(rng.lognormal(gap_mu, gap_sigma, nb)), 1, max_gap).astype(np.int64)
    kind = rng.random(nb)
    lo = (1.0 - double_gap_fraction) / 2
    dt = np.where(kind < lo, 0, gap)                       # q-only gap: dt = 0
    dq = np.where((kind >= lo) & (kind < 2 * lo), 0, gap)  # t-only gap: dq = 0
    both = kind >= 2 * lo
    dq = np.where(both, np.clip(np.ceil(rng.lognormal(gap_mu, gap_sigma, nb)), 1, max_gap).astype(np.int64), dq)
    chain_of = np.repeat(np.arange(n_chains), counts)
    first = np.zeros(n_chains, dtype=np.int64); first[1:] = np.cumsum(counts)[:-1]
    # offsets of every block inside its chain (the gap after a chain's last block is never used)
    t_step = size + dt; q_step = size + dq
    t_cum = np.cumsum(t_step) - t_step; q_cum = np.cumsum(q_step) - q_step
    t_off = t_cum - t_cum[first][chain_of]; q_off = q_cum - q_cum[first][chain_of]
    # sequences: weighted by size; a chain keeps the prefix of its blocks that fits
    t_seq = rng.choice(len(t_sizes), size=n_chains, p=t_sizes / t_sizes.sum())
    q_seq = rng.choice(len(q_sizes), size=n_chains, p=q_sizes / q_sizes.sum())
    limit = (np.minimum(t_sizes[t_seq], q_sizes[q_seq]) * 0.9).astype(np.int64)
    keep = (np.maximum(t_off, q_off) + size) <= limit[chain_of]
    counts2 = np.bincount(chain_of[keep], minlength=n_chains)
    alive = counts2 > 0
    size, t_off, q_off, chain_of = size[keep], t_off[keep], q_off[keep], chain_of[keep]
    chain_of = (np.cumsum(alive) - 1)[chain_of]
    counts2 = counts2[alive]; t_seq = t_seq[alive]; q_seq = q_seq[alive]
    n_chains = len(counts2)
    first = np.zeros(n_chains, dtype=np.int64); first[1:] = np.cumsum(counts2)[:-1]
    last = first + counts2 - 1
    t_span = t_off[last] + size[last]; q_span = q_off[last] + size[last]
    t0 = (rng.random(n_chains) * (t_sizes[t_seq] - t_span + 1)).astype(np.int64)
    q0 = (rng.random(n_chains) * (q_sizes[q_seq] - q_span + 1)).astype(np.int64)
    minus = rng.random(n_chains) < minus_fraction
    return counts2, t_seq, q_seq, minus, t0[chain_of] + t_off, q0[chain_of] + q_off, size


def make_chains(t_sizes, q_sizes, n_blocks_target, seed, mean_log_len=3.6, sigma_log_len=1.1,
                max_len=30000, zipf_s=2.0, max_chain_blocks=1000000, minus_fraction=0.5,
                gap_mu=2.0, gap_sigma=2.0, max_gap=500000, double_gap_fraction=0.10):
    """Chain set with the statistics of SURVEY.md 8d (config 1/5).  Returns (jobs, total, blocks)."""
    rng = np.random.default_rng(seed)
    t_sizes = np.asarray(t_sizes, dtype=np.int64); q_sizes = np.asarray(q_sizes, dtype=np.int64)
    parts, have = [], 0
    while have < n_blocks_target:
        part = _chain_batch(rng, t_sizes, q_sizes, n_blocks_target - have, mean_log_len, sigma_log_len, max_len,
                            zipf_s, max_chain_blocks, minus_fraction, gap_mu, gap_sigma, max_gap, double_gap_fraction)
        parts.append(part); have += int(part[0].sum())
    counts, t_seq, q_seq, minus, ts, qs, size = (np.concatenate([p[i] for p in parts]) for i in range(7))
    n_chains = len(counts)
    first = np.zeros(n_chains, dtype=np.int64); first[1:] = np.cumsum(counts)[:-1]
    blocks = np.zeros(len(size), dtype=BLOCK_DTYPE)
    blocks["tStart"] = ts; blocks["qStart"] = qs; blocks["size"] = size
    jobs = np.zeros(n_chains, dtype=JOB_DTYPE)
    jobs["tSeq"] = t_seq
    jobs["qSeq"] = q_seq.astype(np.uint32) | (minus.astype(np.uint32) << np.uint32(31))
    jobs["firstBlock"] = first; jobs["blockPtr"] = first
    jobs["clipStart"] = NO_CLIP_START; jobs["clipEnd"] = NO_CLIP_END
    return jobs, len(blocks), blocks


def plant_homology(t_genome, q_genome, jobs, total, blocks, subst=0.30, seed=1):
    """Make the query a mutated copy of the target along every block (in place)."""
    synth = _native.load_synth()
    jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE); blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    synth.gat_synth_plant(t_genome.packed.ctypes.data, t_genome.byte_offsets.ctypes.data,
                          q_genome.packed.ctypes.data, q_genome.byte_offsets.ctypes.data,
                          q_genome.sizes.ctypes.data, jobs.ctypes.data, len(jobs), int(total),
                          blocks.ctypes.data, int(round(subst * 65536)), int(seed))


def sprinkle_n_runs(genome, side, jobs, blocks, fraction, seed, max_len=50):
    """Put a 1..max_len bp N run inside `fraction` of the blocks (target or query side)."""
    rng = np.random.default_rng(seed)
    nb = len(blocks)
    pick = np.nonzero(rng.random(nb) < fraction)[0]
    if len(pick) == 0:
        return
    # which job owns each picked block (whole-chain jobs tile the block array)
    job_of = np.searchsorted(jobs["firstBlock"].astype(np.int64), pick, side="right") - 1
    size = blocks["size"][pick].astype(np.int64)
    off = (rng.random(len(pick)) * size).astype(np.int64)
    ln = np.minimum(rng.integers(1, max_len + 1, len(pick)), size - off)
    if side == "t":
        seq = jobs["tSeq"][job_of].astype(np.int64)
        start = blocks["tStart"][pick].astype(np.int64) + off
    else:
        seq = (jobs["qSeq"][job_of] & np.uint32(0x7FFFFFFF)).astype(np.int64)
        minus = (jobs["qSeq"][job_of] >> np.uint32(31)).astype(bool)
        p = blocks["qStart"][pick].astype(np.int64) + off
        qsz = genome.sizes.astype(np.int64)[seq]
        start = np.where(minus, qsz - p - ln, p)
    new = np.zeros(len(pick), dtype=NRUN_DTYPE)
    new["seq"] = seq; new["start"] = start; new["len"] = ln
    runs = np.concatenate([genome.n_runs, new])
    # .2bit wants runs sorted and non-overlapping per sequence: merge
    order = np.lexsort((runs["start"], runs["seq"]))
    runs = runs[order]
    s = runs["start"].astype(np.int64); e = s + runs["len"].astype(np.int64); q = runs["seq"].astype(np.int64)
    merged = []
    cs, ce, cq = s[0], e[0], q[0]
    for i in range(1, len(runs)):
        if q[i] == cq and s[i] <= ce:
            ce = max(ce, e[i])
        else:
            merged.append((cq, cs, ce - cs)); cs, ce, cq = s[i], e[i], q[i]
    merged.append((cq, cs, ce - cs))
    genome.n_runs = np.array(merged, dtype=NRUN_DTYPE)


class Workload:
    def __init__(self, t_genome, q_genome, jobs, total, blocks):
        self.t, self.q, self.jobs, self.total, self.blocks = t_genome, q_genome, jobs, total, blocks

    @property
    def aligned_bp(self):
        return int((self.blocks["size"] & np.uint32(0x7FFFFFFF)).astype(np.int64).sum())

    def algorithmic_bytes(self):
        """0.5 B per aligned base pair + 12 B per block + 40 B per job."""
        # blocks cut into JOINED records for load balance still count once
        return 0.5 * self.aligned_bp + 12.0 * getattr(self, "alg_blocks", self.total) + 40.0 * len(self.jobs)


def make_workload(t_names, t_sizes, q_names, q_sizes, n_blocks, seed, telomere_n=10000, n_fraction=0.001,
                  subst=0.30, **chain_kw):
    t = random_genome(t_names, t_sizes, seed * 7919 + 1, telomere_n)
    q = random_genome(q_names, q_sizes, seed * 7919 + 2, telomere_n)
    jobs, total, blocks = make_chains(t_sizes, q_sizes, n_blocks, seed * 7919 + 3, **chain_kw)
    plant_homology(t, q, jobs, total, blocks, subst, seed * 7919 + 4)
    if n_fraction:
        sprinkle_n_runs(t, "t", jobs, blocks, n_fraction / 2, seed * 7919 + 5)
        sprinkle_n_runs(q, "q", jobs, blocks, n_fraction / 2, seed * 7919 + 6)
    return Workload(t, q, jobs, total, blocks)
