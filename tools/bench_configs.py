"""bench.py --config {1,2,3,4}: measured lines for the other configurations of BASELINE.json (config 5 is bench.py's
own default).  Run by the builder on a GPU box; the JSON lines are kept under profiles/.

  1  scoreChain, hg38 chr1 x mm10 chromosomes, default matrix, -linearGap=medium: whole chains
  2  chainNet -rescore on the same chains: every partial fill of the target net as a clipped job (jobs share records)
  3  chainCleaner on the same chains, -linearGap=loose: the batches of suspect sub-chains (4 per suspect)
  4  distant species: hg38 chr2 x danRer10 chr22, HoxD55.q, -linearGap=loose, 10^6 blocks of mean 38 bp

Every line: the work-list exactly as the tool sends it (configs 2 and 3: dumped by the drop-in tool itself through
GAT_DUMP_WORKLIST), resident in HBM, K timed steps of the kernels (CUDA events) -> `value` and `roofline`; `e2e` through
gat_score() with pinned host buffers; the CPU leg runs the UNMODIFIED reference tool (oracle/_ref) on the same files on
one host core, as shipped; parity = the tool outputs are byte-identical (configs 1-3) / every score equals the reference's
(config 4)."""
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
OURDIR = os.path.join(ROOT, "bin")
EX = os.path.join(ROOT, "tests", "golden", "example")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def timed(cmd, **kw):
    t0 = time.time()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    dt = time.time() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-2000:])
        raise SystemExit("%s failed with %d" % (cmd[0], r.returncode))
    return dt, r.stderr.decode()


def differing_lines(a, b):
    la, lb = open(a).read().split("\n"), open(b).read().split("\n")
    return sum(1 for x, y in zip(la, lb) if x != y) + abs(len(la) - len(lb))


def gpu_line(w_t, w_q, scoring, jobs, total, blocks, steps, warmup):
    """Kernel-only timing of a resident work-list + e2e through gat_score(); returns a dict of measurements."""
    import torch
    from genomealignmenttools_b200 import ChainScorer
    from genomealignmenttools_b200.engine import PinnedArray
    from genomealignmenttools_b200.records import JOB_DTYPE, BLOCK_DTYPE, ali_bases
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(stream)
    sc = ChainScorer(0, stream=stream.cuda_stream)
    sc.load_genome("t", w_t); sc.load_genome("q", w_q); sc.set_scoring(scoring)
    wl = sc.upload(jobs, total, blocks)
    for _ in range(max(3, warmup)):
        wl.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        wl.run()
    e1.record(stream)
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / steps
    launches = sc.stats()["kernel_launches"] * steps
    g, l = wl.results()
    sc.set_profiling(True)
    kms = []
    for _ in range(steps):
        wl.run(); sc.synchronize()
        kms.append(sc.stats()["score_kernel_ms"])
    sc.set_profiling(False)
    pj, pb = PinnedArray(len(jobs), JOB_DTYPE), PinnedArray(len(blocks), BLOCK_DTYPE)
    pg, pl = PinnedArray(len(jobs), np.int64), PinnedArray(len(jobs), np.int64)
    pj.array[:] = jobs; pb.array[:] = blocks
    for _ in range(2):
        sc.score(pj.array, total, pb.array, pg.array, pl.array)
    t0 = time.time()
    reps = max(3, min(steps, 10))
    for _ in range(reps):
        sc.score(pj.array, total, pb.array, pg.array, pl.array)
    e2e_ms = (time.time() - t0) * 1e3 / reps
    assert np.array_equal(pg.array, g) and np.array_equal(pl.array, l)
    bp = int(ali_bases(jobs, total, blocks).sum())
    alg = 0.5 * bp + 12.0 * total + 40.0 * len(jobs)
    wl.free(); sc.close()
    for p in (pj, pb, pg, pl):
        p.free()
    return {"global": g, "local": l, "step_ms": step_ms, "kernel_ms": float(np.mean(kms)), "e2e_ms": e2e_ms, "aligned_bp": bp,
            "alg_bytes": alg, "launches": int(launches), "h2d": int(jobs.nbytes + blocks.nbytes), "d2h": int(16 * len(jobs))}


def emit(args, cfg, workload, m, cpu, parity, extra=None):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = m["alg_bytes"] / (m["kernel_ms"] * 1e-3) / 1e9
    line = {"metric": "aligned bp scored/s", "value": m["aligned_bp"] / (m["step_ms"] * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": m["step_ms"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": cfg,
                       "l2": "work-list + touched genome lines: see aligned_bp / jobs / job_blocks; steps repeat the same list "
                             "(small lists partly stay in the 126 MB L2: a kernel-side figure, not a cold one)"},
            "e2e": {"value": m["aligned_bp"] / (m["e2e_ms"] * 1e-3) / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms"], "call": "gat_score: 12-byte records, 24-byte jobs, pinned host buffers"},
            "gpu_launches": m["launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)",
                         "kernel": "scoreTilesKernel", "kernel_ms": m["kernel_ms"], "algorithmic_bytes_per_launch": m["alg_bytes"]},
            "cpu_baseline": cpu, "parity": parity, "aligned_bp": m["aligned_bp"]}
    if extra:
        line.update(extra)
    print(json.dumps(line))
    if parity.get("mismatches"):
        raise SystemExit("config %d: outputs differ from the reference" % cfg)


def chr1_inputs(d, blocks, qchroms):
    """chains of hg38 chr1 against mm10 chromosomes on synthetic .2bit at the real sizes."""
    from genomealignmenttools_b200 import synth
    import make_golden_helpers as helpers
    from cli_bench import write_chains_fast
    tn, ts = synth.read_chrom_sizes(os.path.join(EX, "hg38.chrom.sizes"))
    qn, qs = synth.read_chrom_sizes(os.path.join(EX, "mm10.chrom.sizes"))
    tn, ts, qn, qs = tn[:1], ts[:1], qn[:qchroms], qs[:qchroms]
    w = synth.make_workload(tn, ts, qn, qs, blocks, seed=0x5EED0010, telomere_n=10000, n_fraction=0.001)
    paths = helpers.write_genomes(w, d)
    heads, counts = helpers.chain_headers(w, tn, qn)
    write_chains_fast(paths["chain"], heads, w.blocks, w.jobs["firstBlock"], counts)
    for name, names, sizes in (("t.sizes", tn, ts), ("q.sizes", qn, qs)):
        with open(os.path.join(d, name), "w") as f:
            f.write("".join("%s\t%d\n" % p for p in zip(names, sizes)))
    log("inputs: %d chains, %d blocks, %.1f Mbp aligned" % (len(w.jobs), w.total, w.aligned_bp / 1e6))
    return w, paths, tn, qn


def load_dump(prefix, w, t_names, q_names):
    """The largest work-list a tool dumped (GAT_DUMP_WORKLIST) + the genomes restricted / ordered as the tool uploaded them."""
    from genomealignmenttools_b200.records import JOB_DTYPE, BLOCK_DTYPE
    from genomealignmenttools_b200.twobit import PackedGenome
    best, k, n_dumps, all_jobs = None, 0, 0, 0
    while os.path.exists("%s.%d.jobs" % (prefix, k)):
        sz = os.path.getsize("%s.%d.blocks" % (prefix, k))
        all_jobs += os.path.getsize("%s.%d.jobs" % (prefix, k)) // JOB_DTYPE.itemsize
        if best is None or sz > best[1]:
            best = (k, sz)
        k += 1; n_dumps += 1
    if best is None:
        raise SystemExit("the tool dumped no work-list")
    jobs = np.fromfile("%s.%d.jobs" % (prefix, best[0]), dtype=JOB_DTYPE)
    blocks = np.fromfile("%s.%d.blocks" % (prefix, best[0]), dtype=BLOCK_DTYPE)

    def ordered(g, names, path):
        use = [names.index(x) for x in open(path).read().split()]
        chunks, offs, cur = [], [], 0
        for i in use:
            nb = (int(g.sizes[i]) + 3) // 4
            chunks.append(g.packed[int(g.byte_offsets[i]):int(g.byte_offsets[i]) + nb]); offs.append(cur); cur += nb
        runs = g.n_runs[np.isin(g.n_runs["seq"], use)].copy()
        remap = {old: new for new, old in enumerate(use)}
        runs["seq"] = [remap[int(x)] for x in runs["seq"]]
        return PackedGenome([names[i] for i in use], g.sizes[use], np.concatenate(chunks), offs, runs)

    t = ordered(w.t, t_names, prefix + ".tseqs")
    q = ordered(w.q, q_names, prefix + ".qseqs")
    # total job-blocks: the CSR is closed by the record count of the last job
    return jobs, blocks, t, q, n_dumps, all_jobs


def run(args):
    from genomealignmenttools_b200 import Scoring, ScoreScheme
    cfg = args.config
    d = tempfile.mkdtemp(prefix="gat_cfg%d_" % cfg)
    if cfg in (1, 2, 3):
        w, paths, tn, qn = chr1_inputs(d, args.blocks if args.blocks != 10_000_000 else 2_000_000, 8)
        ref_score = os.path.join(d, "ref.chain")
        t_ref_sc, _ = timed([os.path.join(REFDIR, "scoreChain"), paths["chain"], paths["t"], paths["q"], ref_score, "-linearGap=medium"])
    if cfg == 1:
        ours = os.path.join(d, "ours.chain")
        t_ours, _ = timed([os.path.join(OURDIR, "scoreChain"), paths["chain"], paths["t"], paths["q"], ours, "-linearGap=medium"])
        bad = differing_lines(ref_score, ours)
        m = gpu_line(w.t, w.q, Scoring(None, "medium"), w.jobs, w.total, w.blocks, args.steps, args.warmup)
        # scoring-only CPU time of the reference: its getChainScore loop in memory (oracle/_ref/ref_driver), one core
        r = subprocess.run([os.path.join(REFDIR, "ref_driver"), paths["chain"], paths["t"], paths["q"], "medium", "-", "2"],
                           capture_output=True, text=True)
        rd = json.loads(r.stdout.strip().splitlines()[-1])
        cpu = {"value": rd["aligned_bp"] / rd["score_s_best"] / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": "the whole set: unmodified getChainScore loop of src/scoreChain/scoreChain.c in memory (ref_driver), best of 2; "
                         "the reference tool end to end took %.2f s, the drop-in %.2f s" % (t_ref_sc, t_ours),
               "reference_tool_wall_s": round(t_ref_sc, 2), "dropin_tool_wall_s": round(t_ours, 2)}
        emit(args, 1, "scoreChain: %d whole chains of hg38 chr1 x 8 mm10 chromosomes (%d blocks), default matrix, linearGap medium"
             % (len(w.jobs), w.total), m, cpu, {"checked": "out.chain of the drop-in vs the reference binary, line by line", "mismatches": bad})
        return
    if cfg in (2, 3):
        srt = os.path.join(d, "sorted.chain")
        subprocess.check_call([os.path.join(REFDIR, "chainSort"), ref_score, srt])
        sizes = [os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes")]
    if cfg == 2:
        common = ["-linearGap=medium", "-tNibDir=" + paths["t"], "-qNibDir=" + paths["q"], srt] + sizes
        nets = {k: [os.path.join(d, "%s.%s.net" % (k, x)) for x in "tq"] for k in ("ref", "our", "plain")}
        t_ref, _ = timed([os.path.join(REFDIR, "chainNet"), "-rescore"] + common + nets["ref"])
        t_plain, _ = timed([os.path.join(REFDIR, "chainNet")] + [c for c in common if not c.startswith(("-linearGap", "-tNib", "-qNib"))] + nets["plain"])
        prefix = os.path.join(d, "dump")
        t_ours, _ = timed([os.path.join(OURDIR, "chainNet"), "-rescore"] + common + nets["our"], env=dict(os.environ, GAT_DUMP_WORKLIST=prefix))
        bad = sum(differing_lines(a, b) for a, b in zip(nets["ref"], nets["our"]))
        jobs, blocks, t, q, n_dumps, _ = load_dump(prefix, w, tn, qn)
        k_best = max(range(n_dumps), key=lambda k: os.path.getsize("%s.%d.blocks" % (prefix, k)))
        total = int(json.load(open("%s.%d.meta" % (prefix, k_best)))["totalJobBlocks"])
        m = gpu_line(t, q, Scoring(None, "medium"), jobs, total, blocks, args.steps, args.warmup)
        rescoring_s = max(t_ref - t_plain, 1e-9)
        cpu = {"value": m["aligned_bp"] / rescoring_s / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": "unmodified chainNet -rescore (src/chainNet/chainNet.c:832-835 per fill) on the same files: %.2f s wall, "
                         "%.2f s without -rescore; the difference is charged to rescoring (sequence loading included). "
                         "Drop-in chainNet -rescore end to end: %.2f s" % (t_ref, t_plain, t_ours),
               "reference_tool_wall_s": round(t_ref, 2), "reference_tool_wall_s_without_rescore": round(t_plain, 2),
               "dropin_tool_wall_s": round(t_ours, 2)}
        emit(args, 2, "chainNet -rescore: %d partial fills of the target net as clipped jobs over %d job-blocks (%d shared records), "
             "linearGap medium" % (len(jobs), total, len(blocks)), m, cpu,
             {"checked": "t.net and q.net of the drop-in vs the reference binary, line by line", "mismatches": bad})
        return
    if cfg == 3:
        env = dict(os.environ, PATH=REFDIR + ":" + os.environ["PATH"])
        subprocess.check_call("chainNet -minScore=0 sorted.chain t.sizes q.sizes stdout /dev/null | NetFilterNonNested.perl /dev/stdin "
                              "-minScore1 3000 > in.net", shell=True, executable="/bin/bash", cwd=d, env=env, stderr=subprocess.DEVNULL)
        res = {}
        prefix = os.path.join(d, "dump")
        for k, bindir in (("ref", REFDIR), ("our", OURDIR)):
            env = dict(os.environ, PATH=bindir + ":" + os.environ["PATH"], GAT_DUMP_WORKLIST=prefix)
            cmd = [os.path.join(bindir, "chainCleaner"), "sorted.chain", "t.2bit", "q.2bit", k + ".clean.chain", k + ".clean.bed",
                   "-net=in.net", "-linearGap=loose"]
            res[k], _ = timed(cmd, cwd=d, env=env)
        bad = sum(differing_lines(os.path.join(d, "ref.clean." + x), os.path.join(d, "our.clean." + x)) for x in ("chain", "bed"))
        jobs, blocks, t, q, n_dumps, all_jobs = load_dump(prefix, w, tn, qn)
        k_best = max(range(n_dumps), key=lambda k: os.path.getsize("%s.%d.blocks" % (prefix, k)))
        total = int(json.load(open("%s.%d.meta" % (prefix, k_best)))["totalJobBlocks"])
        m = gpu_line(t, q, Scoring(None, "loose"), jobs, total, blocks, args.steps, args.warmup)
        cpu = {"value": None, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": "unmodified chainCleaner (src/chainCleaner/chainCleaner.c:1226-1229 per sub-chain) on the same files: %.2f s wall; "
                         "drop-in chainCleaner end to end: %.2f s; the reference does not separate its rescoring time, so no Gbp/s is "
                         "derived for it" % (res["ref"], res["our"]),
               "reference_tool_wall_s": round(res["ref"], 2), "dropin_tool_wall_s": round(res["our"], 2)}
        emit(args, 3, "chainCleaner: the largest batch of suspect sub-chains (%d of the %d jobs the run scored in %d batches), %d job-blocks, "
             "linearGap loose" % (len(jobs), all_jobs, n_dumps, total), m, cpu,
             {"checked": "out.chain and out.bed of the drop-in vs the reference binary, line by line", "mismatches": bad})
        return
    if cfg == 4:
        from genomealignmenttools_b200 import synth
        import make_golden_helpers as helpers
        from cli_bench import write_chains_fast
        tn, ts = synth.read_chrom_sizes(os.path.join(EX, "hg38.chrom.sizes"))
        qn, qs = synth.read_chrom_sizes(os.path.join(EX, "danRer10.chrom.sizes"))
        ti, qi = tn.index("chr2"), qn.index("chr22")
        t = synth.random_genome([tn[ti]], [ts[ti]], 0x5EED0003, telomere_n=10000)
        q = synth.random_genome([qn[qi]], [qs[qi]], 0x5EED0004, telomere_n=10000)
        n = args.blocks if args.blocks != 10_000_000 else 1_000_000
        jobs, total, blocks = synth.make_chains([ts[ti]], [qs[qi]], n, seed=0x5EED0040, mean_log_len=3.2, sigma_log_len=0.9,
                                                double_gap_fraction=0.14, max_chain_blocks=20000)
        synth.plant_homology(t, q, jobs, total, blocks, 0.35, 0x5EED0041)
        w = synth.Workload(t, q, jobs, total, blocks)
        paths = helpers.write_genomes(w, d)
        heads, counts = helpers.chain_headers(w, [tn[ti]], [qn[qi]])
        write_chains_fast(paths["chain"], heads, blocks, jobs["firstBlock"], counts)
        hox = os.path.join(EX, "HoxD55.q")
        m = gpu_line(t, q, Scoring(ScoreScheme.read(hox), "loose"), jobs, total, blocks, args.steps, args.warmup)
        scores = os.path.join(d, "ref.scores")
        r = subprocess.run([os.path.join(REFDIR, "ref_driver"), paths["chain"], paths["t"], paths["q"], "loose", hox, "2", scores],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("ref_driver failed: " + r.stderr[-500:])
        rd = json.loads(r.stdout.strip().splitlines()[-1])
        rows = np.loadtxt(scores, dtype=np.int64, ndmin=2)
        j = rows[:, 0] - 1
        bad = int((m["global"][j] != rows[:, 1]).sum() + (m["local"][j] != rows[:, 2]).sum())
        cpu = {"value": rd["aligned_bp"] / rd["score_s_best"] / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": "the whole set: unmodified getChainScore loop of src/scoreChain/scoreChain.c in memory (ref_driver), best of 2"}
        emit(args, 4, "distant species: %d whole chains, %d blocks of mean %.0f bp, hg38 chr2 x danRer10 chr22, HoxD55.q, linearGap loose"
             % (len(jobs), total, w.aligned_bp / total), m, cpu,
             {"checked": "global and local score of every chain vs the unmodified reference (%d chains)" % len(j), "mismatches": bad})
        return
    raise SystemExit("--config must be 1..5")
