#!/usr/bin/env python
"""bench.py -- aligned bp scored/s of the chain-rescoring hot path on N B200s of one node.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): a genome-wide
synthetic hg38 x mm10 chain set -- all 455 / 66 sequences of example/{hg38,mm10}.chrom.sizes,
~10 M blocks per GPU, heavy-tailed chain sizes, both strands, N runs, homologous query -- scored
with the default matrix and -linearGap=medium (global + local score per chain, as scoreChain does).

A step = one pass of the hot path over the rank's whole work-list.  `value` times the kernels with
the work-list resident in HBM; `e2e` times the public call gat_score_packed() with pinned HOST buffers
(H2D of the work-list + kernels + D2H of the scores) every step.  At N > 1 the SAME ~10 M-block set is cut
N ways (strong scaling, the default): chains above 1/(4N) of the aligned bases are cut at block boundaries
 and the pieces are balanced over the GPUs by greedy aligned-base load; every GPU holds a full
genome copy and there is no collective on the data path; time is the max over ranks.  `--scaling weak` scores
N x 10 M blocks instead (round 1's measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--blocks B] [--scaling strong|weak] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
METRIC = "aligned bp scored/s"
UNIT = "Gbp/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- workload
def build_workload(blocks_per_gpu, world=1, plant=True, mean_log_len=3.6, max_len=30000, rank=0, exchange=None):
    """The whole job for `world` GPUs: world x blocks_per_gpu job-blocks (weak scaling), generated in
    `world` seeded parts so that N=1 is exactly part 0.  Every rank ends up with the identical set: either
    it generates all parts itself, or (exchange given) only part `rank` and the ranks swap parts -- the chain
    generator is the slow step of the set-up (about 10 s and 3 GB per part)."""
    from genomealignmenttools_b200 import synth
    tn, ts = synth.read_chrom_sizes(os.path.join(GOLDEN, "example", "hg38.chrom.sizes"))
    qn, qs = synth.read_chrom_sizes(os.path.join(GOLDEN, "example", "mm10.chrom.sizes"))
    t0 = time.time()
    t = synth.random_genome(tn, ts, 0x5EED0001, telomere_n=10000)
    q = synth.random_genome(qn, qs, 0x5EED0002, telomere_n=10000)

    def chains_of(part):
        jobs, n, blocks = synth.make_chains(ts, qs, blocks_per_gpu, seed=0x5EED0050 + part, mean_log_len=mean_log_len, max_len=max_len)
        return jobs, blocks

    if exchange is not None and world > 1:
        parts = exchange(*chains_of(rank))
    else:
        parts = [chains_of(part) for part in range(world)]
    job_parts, block_parts, total = [], [], 0
    for part, (jobs, blocks) in enumerate(parts):
        n = len(blocks)
        if plant:
            synth.plant_homology(t, q, jobs, n, blocks, 0.30, 0x5EED0060 + part)
        synth.sprinkle_n_runs(t, "t", jobs, blocks, 0.0005, 0x5EED0070 + part)
        synth.sprinkle_n_runs(q, "q", jobs, blocks, 0.0005, 0x5EED0080 + part)
        jobs["firstBlock"] += total
        jobs["blockPtr"] += total
        total += n
        job_parts.append(jobs); block_parts.append(blocks)
    jobs = np.concatenate(job_parts) if world > 1 else job_parts[0]
    blocks = np.concatenate(block_parts) if world > 1 else block_parts[0]
    w = synth.Workload(t, q, jobs, total, blocks)
    w.t_names, w.q_names = tn, qn
    log("workload: %d chains, %d blocks, %.1f Mbp aligned for %d GPU(s), built in %.1f s"
        % (len(jobs), total, w.aligned_bp / 1e6, world, time.time() - t0))
    return w


def shard_workload(w, rank, world):
    """This rank's share: chains above 1/(4 world) of the aligned bases are cut into pieces, the pieces
    are balanced over the GPUs by greedy aligned-base load, records are compacted per shard.  Returns the shard, the
    indices of its jobs among the pieces, and the split (piece jobs, origin, first piece per job)."""
    from genomealignmenttools_b200 import sharding, synth
    if world == 1:
        return w, np.arange(len(w.jobs)), None
    pieces, origin, first_piece = sharding.split_giant_jobs(w.jobs, w.total, w.blocks, world)
    part, _ = sharding.assign_jobs(pieces, w.total, w.blocks, world)
    idx, shard, shard_total = sharding.take_shard(pieces, w.total, part, rank)
    sj, sb = sharding.compact_blocks(shard, shard_total, w.blocks)
    mine = synth.Workload(w.t, w.q, sj, shard_total, sb)
    mine.t_names, mine.q_names = w.t_names, w.q_names
    cut = np.diff(first_piece) > 1
    log("[rank %d] shard: %d jobs (%d pieces of the %d cut chains), %d blocks, %.1f Mbp aligned"
        % (rank, len(sj), int(cut[origin[idx]].sum()), int(cut.sum()), shard_total, mine.aligned_bp / 1e6))
    return mine, idx, (pieces, origin, first_piece)


def h2d_ceiling(torch, device, barrier, nbytes=256 << 20, reps=5):
    """This rank's pinned host -> device copy rate while every rank copies at once (GB/s): what bounds `e2e`."""
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def source_hash():
    """Hash of the kernel sources: a DRAM-traffic figure measured under ncu (profiles/traffic.json, written by
    tools/measure_traffic.sh) is only quoted for the code it was measured on."""
    import hashlib, re
    h = hashlib.sha256()
    for f in ("gat_tiles.cuh", "gat_kernels.cuh", "gat_capi.cu"):
        text = open(os.path.join(ROOT, "genomealignmenttools_b200", "csrc", f), "r", errors="replace").read()
        # the code, not its commentary: // comments, /* */ comments and blank lines do not count (no string literal of these
        # files holds "//" or "/*")
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        lines = [re.sub(r"//.*$", "", ln).rstrip() for ln in text.splitlines()]
        h.update("\n".join(ln for ln in lines if ln.strip()).encode())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            inside = t_begin - 0.05 <= ts <= t_end + 0.15
            try:
                if inside:
                    sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def sample_for_cpu(w, cores, max_aligned_bp):
    """A bounded sample of the same workload for the host cores: whole chains on `cores` mid-size
    target chromosomes x 6 query chromosomes (bounds the reference's whole-chromosome unpacking),
    one shard per target chromosome = one reference process per core."""
    t_pick = [w.t_names.index(n) for n in ("chr8", "chr9", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15",
                                           "chr16", "chr17", "chr18", "chr19", "chr20", "chr21", "chr22", "chrX")
              if n in w.t_names][:max(1, cores)]
    q_pick = [w.q_names.index(n) for n in ("chr1", "chr2", "chr3", "chr4", "chr5", "chr6") if n in w.q_names]
    qs = w.jobs["qSeq"] & np.uint32(0x7FFFFFFF)
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    csum = np.concatenate([[0], np.cumsum(w.blocks["size"].astype(np.int64))])
    job_bp = csum[w.jobs["firstBlock"].astype(np.int64) + counts] - csum[w.jobs["firstBlock"].astype(np.int64)]
    shards = []
    per_shard = max_aligned_bp // max(1, len(t_pick))
    for ti in t_pick:
        sel = np.nonzero((w.jobs["tSeq"] == ti) & np.isin(qs, q_pick))[0]
        keep = sel[np.cumsum(job_bp[sel]) <= per_shard]
        if len(keep) == 0 and len(sel):
            keep = sel[:1]
        shards.append(keep)
    return t_pick, q_pick, shards


def write_shard_files(w, t_pick, q_pick, shards, d):
    """Sample .2bit files holding only the touched sequences + one .chain per shard."""
    from genomealignmenttools_b200.twobit import PackedGenome
    from genomealignmenttools_b200 import chainio

    def subset(g, names, pick):
        chunks, offs, cur = [], [], 0
        for i in pick:
            nb = (int(g.sizes[i]) + 3) // 4
            chunks.append(g.packed[int(g.byte_offsets[i]):int(g.byte_offsets[i]) + nb]); offs.append(cur); cur += nb
        runs = g.n_runs[np.isin(g.n_runs["seq"], pick)].copy()
        remap = {old: new for new, old in enumerate(pick)}
        runs["seq"] = [remap[int(s)] for s in runs["seq"]]
        return PackedGenome([names[i] for i in pick], g.sizes[pick], np.concatenate(chunks), offs, runs)

    t_path, q_path = os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit")
    subset(w.t, w.t_names, t_pick).write_2bit(t_path)
    subset(w.q, w.q_names, q_pick).write_2bit(q_path)
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    paths = []
    for k, sel in enumerate(shards):
        heads = []
        for j in sel:
            job = w.jobs[j]
            fb, nb = int(job["firstBlock"]), int(counts[j])
            first, last = w.blocks[fb], w.blocks[fb + nb - 1]
            ti, qi = int(job["tSeq"]), int(job["qSeq"] & 0x7FFFFFFF)
            heads.append((0, w.t_names[ti], int(w.t.sizes[ti]), int(first["tStart"]), int(last["tStart"]) + int(last["size"]),
                          w.q_names[qi], int(w.q.sizes[qi]), "-" if job["qSeq"] >> 31 else "+",
                          int(first["qStart"]), int(last["qStart"]) + int(last["size"]), int(j) + 1))
        p = os.path.join(d, "shard%d.chain" % k)
        chainio.write_chains(p, heads, w.blocks, w.jobs["firstBlock"][sel], counts[sel])
        paths.append(p)
    return t_path, q_path, paths


def run_reference_pass(t_path, q_path, chain_paths, reps, want_scores=False):
    """One ref_driver process per shard, all at once (the reference is single-threaded).  Returns
    (aligned bp, wall seconds of the slowest process' best scoring pass, per-process dicts)."""
    procs = []
    for p in chain_paths:
        cmd = [REF_DRIVER, p, t_path, q_path, "medium", "-", str(reps)]
        if want_scores:
            cmd.append(p + ".scores")
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for pr in procs:
        so, se = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("ref_driver failed: " + se[-500:])
        outs.append(json.loads(so.strip().splitlines()[-1]))
    bp = sum(o["aligned_bp"] for o in outs)
    return bp, max(o["score_s_best"] for o in outs), outs


def cpu_reference(w, steps, warmup, max_aligned_bp, gpu_scores=None):
    cores = os.cpu_count() or 1
    if not os.path.exists(REF_DRIVER):
        return None
    t_pick, q_pick, shards = sample_for_cpu(w, cores, max_aligned_bp)
    with tempfile.TemporaryDirectory() as d:
        t_path, q_path, chain_paths = write_shard_files(w, t_pick, q_pick, shards, d)
        t0 = time.time()
        bp, sec, outs = run_reference_pass(t_path, q_path, chain_paths, max(1, steps + warmup), want_scores=True)
        wall = time.time() - t0
        checked = mismatches = 0
        if gpu_scores is not None:
            g, l = gpu_scores
            for p in chain_paths:
                rows = np.loadtxt(p + ".scores", dtype=np.int64, ndmin=2)
                if rows.size == 0:
                    continue
                j = rows[:, 0] - 1
                mismatches += int((g[j] != rows[:, 1]).sum() + (l[j] != rows[:, 2]).sum())
                checked += len(j)
    n_chains = sum(o["chains"] for o in outs)
    return {"value": bp / sec / 1e9, "unit": UNIT, "cores": len(chain_paths), "kind": "reference",
            "sample": "%d whole chains (%d blocks, %.1f Mbp aligned) of the same workload on %d target x %d query "
                      "chromosomes; unmodified getChainScore (chainCalcScore + chainCalcScoreLocal) of "
                      "src/scoreChain/scoreChain.c in memory, one process per target chromosome, best of %d passes; "
                      ".2bit unpack + chain parsing excluded (%.1f s wall incl. them)"
                      % (n_chains, sum(o["blocks"] for o in outs), bp / 1e6, len(t_pick), len(q_pick),
                         max(1, steps + warmup), wall),
            "host_cores": cores, "parity_checked_jobs": checked, "parity_mismatches": mismatches,
            "ms_per_step": sec * 1e3, "aligned_bp": bp}


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank on the CPUs of its GPU's NUMA node, so that the pinned work-list buffers of the e2e leg are
    first-touched next to the GPU (8 ranks copying 50 GB/s each across sockets do not scale).  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(gpu_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=10_000_000, help="job-blocks per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration: 5 = the genome-wide set the metric is quoted on (default); 1-4 = scoreChain chr1, "
                         "chainNet -rescore fills, chainCleaner sub-chains, distant-species short blocks (tools/bench_configs.py)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the same --blocks set cut N ways (default), weak = N x --blocks")
    ap.add_argument("--cpu-sample-mbp", type=float, default=160.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mean-log-len", type=float, default=3.6, help="block length ~ lognormal(mu, 1.1); 3.6 = the benchmark (mean 67 bp)")
    ap.add_argument("--max-len", type=int, default=30000)
    ap.add_argument("--fold", type=int, default=0, help="experiment: fold block coordinates into the first FOLD bases of their sequences (cache-resident genome)")
    ap.add_argument("--split", type=int, default=0, help="cut blocks longer than this into JOINED records (0 = as generated)")
    args = ap.parse_args()
    if args.config != 5:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        return bench_configs.run(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    config = {"workload": "genome-wide synthetic hg38 x mm10 chain set (BASELINE.json configs[4]): all sequences of "
                          "example/{hg38,mm10}.chrom.sizes, %d job-blocks per GPU, default matrix, linearGap medium, "
                          "global+local score per chain" % args.blocks,
              "blocks_per_gpu": args.blocks if (world == 1 or args.scaling == "weak") else args.blocks // world,
              "total_blocks": args.blocks * (world if args.scaling == "weak" else 1),
              "parallelism": ("x%d GPUs, %s scaling: one set of %d job-blocks, chains above 1/(4N) of the aligned bases cut at block "
                              "boundaries, pieces balanced by greedy aligned-base load, full genome copy per GPU, no collective"
                              % (world, args.scaling if world > 1 else "single GPU", args.blocks * (world if args.scaling == "weak" else 1))),
              "l2": "inputs (work-list + touched genome sectors, ~0.6 GB) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        if not os.path.exists(REF_DRIVER):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built"}))
            return
        w = build_workload(args.blocks, 1, plant=True)
        res = cpu_reference(w, args.steps, args.warmup, int(args.cpu_sample_mbp * 1e6))
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from genomealignmenttools_b200 import ChainScorer, Scoring
    from genomealignmenttools_b200.engine import PinnedArray
    from genomealignmenttools_b200.records import JOB_DTYPE, BLOCK_DTYPE

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def exchange(jobs, blocks):
        """Set-up only: every rank generated one part of the chain set; swap them (NCCL all_gather of raw bytes)."""
        from genomealignmenttools_b200.sharding import exchange_parts
        parts = exchange_parts(dist, world, "cuda", jobs, blocks)
        torch.cuda.empty_cache()
        return parts

    weak = world > 1 and args.scaling == "weak"
    whole = build_workload(args.blocks, world if weak else 1, mean_log_len=args.mean_log_len, max_len=args.max_len, rank=rank,
                           exchange=exchange if weak else None)
    w, shard_idx, split = shard_workload(whole, rank, world)
    part_local = np.zeros(0, dtype=np.uint32)     # jobs of my shard that are pieces of a cut chain: their tuples are wanted
    if split is not None:
        cut = np.diff(split[2]) > 1
        part_local = np.nonzero(cut[split[1][shard_idx]])[0].astype(np.uint32)
    if args.split:
        from genomealignmenttools_b200.records import split_long_blocks
        bp = w.aligned_bp
        w.alg_blocks = w.total
        w.jobs, w.total, w.blocks = split_long_blocks(w.jobs, w.total, w.blocks, args.split)
        assert w.aligned_bp == bp
    if args.fold:
        small = w.blocks["size"] < 300
        w.blocks["tStart"][small] %= args.fold
        w.blocks["qStart"][small] %= args.fold
    # our kernels launch on this torch stream, so torch's CUDA events bracket them
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    sc = ChainScorer(local_rank, stream=stream.cuda_stream)
    t0 = time.time()
    sc.load_genome("t", w.t)
    sc.load_genome("q", w.q)
    sc.set_scoring(Scoring(None, "medium"))
    upload_s = time.time() - t0
    log("[rank %d] genomes resident in HBM after %.1f s (NUMA node %s)" % (rank, upload_s, numa))

    # ---- device-resident timing (value)
    wl = sc.upload(w.jobs, w.total, w.blocks)
    for _ in range(warmup):
        wl.run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        wl.run()
    e1.record(stream)
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = sc.stats()["kernel_launches"] * args.steps
    long_streamed = int(sc.stats()["long_streamed"])      # which instantiation the library picked for this list
    part_tuples = sc.request_tuples(part_local) if len(part_local) else None      # (one more, untimed pass when pieces are present)
    if part_tuples is not None:
        wl.run()
    g_res, l_res = wl.results()
    part_tuples = part_tuples.copy() if part_tuples is not None else None

    # ---- kernel-only pass for the roofline (events around the scoring kernel, same stream)
    sc.set_profiling(True)
    kms = []
    for _ in range(args.steps):
        wl.run()
        sc.synchronize()
        kms.append(sc.stats()["score_kernel_ms"])
    sc.set_profiling(False)
    kernel_ms = float(np.mean(kms))
    # ---- end-to-end through the public calls with pinned host buffers (e2e): every step copies the work-list in,
    # runs the kernels and copies the scores out.  Headline: gat_score_compact(), the work-list as a .chain file
    # stores it (size + gaps, 6 bytes per block; expanded on the device).  gat_score() with 12-byte absolute
    # records is timed beside it.
    from genomealignmenttools_b200.records import (pack_compact, split_long_blocks, CJOB_DTYPE, CBLOCK_DTYPE, CABS_DTYPE)
    e2e_steps = max(3, min(args.steps, 10))

    def time_calls(call):
        for _ in range(2):
            call()
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.time()
        ea.record(stream)
        for _ in range(e2e_steps):
            call()
        eb.record(stream)
        barrier()
        wall_ms = (time.time() - tw0) * 1e3 / e2e_steps
        return max(ea.elapsed_time(eb) / e2e_steps, wall_ms)       # the calls block: wall time is the honest one

    def parts_of(call):         # where a step goes (one extra, profiled call; not part of a timed region)
        sc.set_profiling(True)
        call()
        st = sc.stats()
        sc.set_profiling(False)
        return float(st["h2d_ms"]), float(st["all_kernels_ms"])

    pg, pl = PinnedArray(len(w.jobs), np.int64), PinnedArray(len(w.jobs), np.int64)
    pj, pb = PinnedArray(len(w.jobs), JOB_DTYPE), PinnedArray(len(w.blocks), BLOCK_DTYPE)
    pj.array[:] = w.jobs
    pb.array[:] = w.blocks
    plain_call = lambda: sc.score(pj.array, w.total, pb.array, pg.array, pl.array)
    plain_ms = time_calls(plain_call)
    assert np.array_equal(pg.array, g_res) and np.array_equal(pl.array, l_res), "e2e and resident results differ"
    plain_bytes = int(w.jobs.nbytes + w.blocks.nbytes)
    pj.free(); pb.free()

    cj, cb, ab, an = pack_compact(*split_long_blocks(w.jobs, w.total, w.blocks, 4096))
    pins = [PinnedArray(len(a), d) for a, d in ((cj, CJOB_DTYPE), (cb, CBLOCK_DTYPE), (ab, CABS_DTYPE), (an, CABS_DTYPE))]
    for pin, a in zip(pins, (cj, cb, ab, an)):
        pin.array[:] = a
    pg.array[:] = 0; pl.array[:] = 0
    e2e_tuples = [None]

    def compact_call():
        if len(part_local):
            e2e_tuples[0] = sc.request_tuples(part_local)
        sc.score_compact(pins[0].array, pins[1].array, pins[2].array, pins[3].array, pg.array, pl.array)

    compact_ms = time_calls(compact_call)
    assert np.array_equal(pg.array, g_res) and np.array_equal(pl.array, l_res), "compact e2e and resident results differ"
    assert part_tuples is None or np.array_equal(e2e_tuples[0], part_tuples), "compact e2e and resident tuples differ"
    compact6_bytes = int(cj.nbytes + cb.nbytes + ab.nbytes + an.nbytes)
    for pin in pins:
        pin.free()

    # headline: gat_score_packed(), one 32-bit word per block (size, gap in front on both sequences), absolute records for
    # chain starts and gaps that do not fit
    from genomealignmenttools_b200.records import pack_packed
    packed = pack_packed(w.jobs, w.total, w.blocks)
    pins = [PinnedArray(len(a), d) for a, d in zip(packed, (CJOB_DTYPE, np.uint32, CABS_DTYPE, CABS_DTYPE, np.uint32))]
    for pin, a in zip(pins, packed):
        pin.array[:] = a
    pg.array[:] = 0; pl.array[:] = 0

    def packed_call():
        if len(part_local):
            e2e_tuples[0] = sc.request_tuples(part_local)
        sc.score_packed(pins[0].array, pins[1].array, pins[2].array, pins[3].array, pins[4].array, pg.array, pl.array)

    e2e_ms = time_calls(packed_call)
    assert np.array_equal(pg.array, g_res) and np.array_equal(pl.array, l_res), "packed e2e and resident results differ"
    assert part_tuples is None or np.array_equal(e2e_tuples[0], part_tuples), "packed e2e and resident tuples differ"
    compact_bytes = int(sum(a.nbytes for a in packed))
    h2d_ms, kern_ms = parts_of(packed_call)
    # (a profiled call runs unsliced on one stream; the timed calls overlap the copy of a slice with the kernels of the one before)
    e2e_parts = {"h2d_and_expand_ms_unsliced": round(h2d_ms, 4), "kernels_ms_unsliced": round(kern_ms, 4)}
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    ceiling = h2d_ceiling(torch, torch.device("cuda", local_rank), barrier)

    # ---- N > 1: rank 0 joins the pieces of the cut chains and checks every score against one GPU scoring the whole set
    scaling_check = None
    if world > 1 and split is not None:
        from genomealignmenttools_b200 import sharding
        payload = (shard_idx, g_res, l_res, shard_idx[part_local], part_tuples)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
        if rank == 0:
            pieces, origin, first_piece = split
            pg_all = np.zeros(len(pieces), dtype=np.int64); pl_all = np.zeros(len(pieces), dtype=np.int64)
            tuples = {}
            for idx, gg, ll, pidx, ptup in gathered:
                pg_all[idx] = gg; pl_all[idx] = ll
                for k, pi in enumerate(pidx):
                    tuples[int(pi)] = ptup[k]
            scoring = Scoring(None, "medium")
            jg, jl = sharding.join_pieces(whole.jobs, whole.total, whole.blocks, pieces, origin, first_piece, pg_all, pl_all, tuples,
                                          scoring.gap.cost)
            one_g, one_l = sc.score(whole.jobs, whole.total, whole.blocks)
            scaling_check = {"jobs": int(len(jg)), "cut_chains": int((np.diff(first_piece) > 1).sum()),
                             "pieces": int(len(pieces) - len(jg) + (np.diff(first_piece) > 1).sum()),
                             "mismatches_vs_one_gpu": int((jg != one_g).sum() + (jl != one_l).sum())}
            log("scaling check: %s" % scaling_check)

    per_step_ms = ms / args.steps
    stats = torch.tensor([per_step_ms, e2e_ms, kernel_ms, float(w.aligned_bp), float(w.algorithmic_bytes()), plain_ms, ceiling, compact_ms],
                         dtype=torch.float64, device="cuda")
    alg_bytes = float(w.algorithmic_bytes())
    ceiling_sum = ceiling
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        per_step_ms, e2e_ms, kernel_ms, plain_ms, compact_ms = mx[0].item(), mx[1].item(), mx[2].item(), mx[5].item(), mx[7].item()
        total_bp = sm[3].item()
        alg_bytes = sm[4].item() / world          # per GPU (mean) against the slowest GPU's kernel time
        ceiling_sum = sm[6].item()
    else:
        total_bp = float(w.aligned_bp)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes of one launch of the scoring kernel, measured under ncu by tools/measure_traffic.sh on the default N=1
        # workload; quoted only while the kernel sources are the ones it was measured on (else null)
        traffic, traffic_note = None, "not measured for this code / workload (tools/measure_traffic.sh)"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if (tj.get("source_hash") == source_hash() and world == 1 and tj.get("blocks") == args.blocks
                    and args.mean_log_len == 3.6 and not args.split and not args.fold):
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("how")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_bp / (per_step_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
            "e2e": {"value": total_bp / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": compact_bytes, "d2h_bytes_per_step": int(16 * len(w.jobs) + 32 * len(part_local)),
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "parts_rank0": e2e_parts,
                    # pinned host -> device rate of this box with all ranks copying at once, and how much of the e2e step the
                    # bare copy of the step's bytes at that rate would take
                    "h2d_ceiling_gbs_all_ranks": round(ceiling_sum, 1), "h2d_ceiling_gbs_rank0": round(ceiling, 1),
                    "h2d_time_fraction_of_step": round((compact_bytes / (ceiling * 1e9)) / (e2e_ms * 1e-3), 3),
                    "call": "gat_score_packed: work-list as a .chain file stores it (size and the gaps in front, one 32-bit word per block; 8-byte chains; absolute table for chain starts and gaps above 511), expanded on the device",
                    "compact_records": {"call": "gat_score_compact: 6-byte size/gap blocks", "value": total_bp / (compact_ms * 1e-3) / 1e9,
                                        "ms_per_step": compact_ms, "h2d_bytes_per_step": compact6_bytes},
                    "plain_records": {"call": "gat_score: 12-byte absolute blocks, 24-byte jobs", "value": total_bp / (plain_ms * 1e-3) / 1e9,
                                      "ms_per_step": plain_ms, "h2d_bytes_per_step": plain_bytes}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                         "kernel": "scoreTilesKernel<SYM=1,PLAIN=1,LONG=%d>" % long_streamed, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes": "SURVEY 8d: 0.5 B per aligned bp + 12 B per block + 40 B per job"},
            "clocks": clocks,
            "aligned_bp_per_gpu": w.aligned_bp, "chains_per_gpu": int(len(w.jobs)), "genome_upload_s": round(upload_s, 3),
        }
        if scaling_check is not None:
            line["scaling_check"] = scaling_check
            if scaling_check["mismatches_vs_one_gpu"]:
                raise SystemExit("bench: %d scores of the sharded run differ from one GPU" % scaling_check["mismatches_vs_one_gpu"])
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.time()
            res = cpu_reference(w, 2, 1, int(args.cpu_sample_mbp * 1e6), gpu_scores=(g_res, l_res))
            if res is not None:
                line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores",
                                                            "parity_checked_jobs", "parity_mismatches")}
                log("cpu baseline leg took %.1f s" % (time.time() - t0))
                if res["parity_mismatches"]:
                    raise SystemExit("bench: %d GPU scores differ from the reference" % res["parity_mismatches"])
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref/ref_driver missing"}
        print(json.dumps(line))
    wl.free()
    for p in [pg, pl] + pins:
        p.free()
    sc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
