"""Parity of the CUDA path (through the C ABI) with the oracle, the reference's golden vectors
and the fixtures produced by the compiled reference.  Bit-exact: these are integer scores."""
import os
import numpy as np
import pytest
from genomealignmenttools_b200 import ChainScorer, Scoring, ScoreScheme, GapCalc, GatError, chainio, synth
from genomealignmenttools_b200.twobit import PackedGenome
from genomealignmenttools_b200.records import job_block_counts
from genomealignmenttools_b200.records import (JOB_DTYPE, BLOCK_DTYPE, NRUN_DTYPE, NO_CLIP_START, NO_CLIP_END,
                                               BLOCK_JOINED, QSEQ_MINUS, ali_bases)
import make_golden_helpers as helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "stream", "list"])
def long_blocks_mode(request, monkeypatch):
    """Every test of this file runs three times: with the instantiation of the scoring kernel the library picks from the
    list's block sizes, and with each of the two forced (the one that streams blocks of more than 1056 bases, the one
    that lists every block).  Both must score every list exactly."""
    if request.param != "auto":
        monkeypatch.setenv("GAT_LONG_BLOCKS", request.param)
    else:
        monkeypatch.delenv("GAT_LONG_BLOCKS", raising=False)
    return request.param


def gpu_score(t, q, scoring, jobs, total, blocks):
    with ChainScorer(0) as sc:
        sc.load_genome("t", t)
        sc.load_genome("q", q)
        sc.set_scoring(scoring)
        return sc.score(jobs, total, blocks)


def scoring_of(golden, matrix, gap):
    return Scoring(ScoreScheme.read(os.path.join(golden, matrix)) if matrix else None, gap)


@pytest.mark.parametrize("name,want", [("newStyleLastz", 671823), ("oldStyleBlastz", 671644)])
def test_kent_axtchain_golden(golden, name, want):
    d = os.path.join(golden, "kent_chrM")
    t, q = PackedGenome.read_2bit(os.path.join(d, "hg19.chrM.2bit")), PackedGenome.read_2bit(os.path.join(d, "susScr3.chrM.2bit"))
    cs = chainio.ChainSet.read(os.path.join(d, name + ".chain"))
    jobs, total = cs.jobs(t, q)
    g, l = gpu_score(t, q, scoring_of(golden, "kent_chrM/%s.Q.txt" % name, "loose"), jobs, total, cs.blocks)
    assert g[0] == want == int(cs.score[0]) and l[0] == want


def test_chrM_known_answers(golden):
    d = os.path.join(golden, "kent_chrM")
    t, q = PackedGenome.read_2bit(os.path.join(d, "hg19.chrM.2bit")), PackedGenome.read_2bit(os.path.join(d, "susScr3.chrM.2bit"))
    cs = chainio.ChainSet.read(os.path.join(d, "newStyleLastz.chain"))
    jobs, total = cs.jobs(t, q)
    g, l = gpu_score(t, q, scoring_of(golden, "example/HoxD55.q", "loose"), jobs, total, cs.blocks)
    assert (g[0], l[0]) == (829041, 829041)
    g, l = gpu_score(t, q, Scoring(None, "medium"), jobs, total, cs.blocks)
    assert (g[0], l[0]) == (767508, 767508)
    assert ali_bases(jobs, total, cs.blocks)[0] == 15310


@pytest.mark.parametrize("tag,matrix,gap", [
    ("medium_default", None, "medium"), ("loose_hoxd55", "example/HoxD55.q", "loose"),
    ("loose_lastz", "kent_chrM/newStyleLastz.Q.txt", "loose"), ("medium_asym", "synth_small/asym.q", "medium")])
@pytest.mark.parametrize("tfile", ["t.2bit", "t.swapped.2bit"])
def test_synth_small_reference_scorechain(golden, tag, matrix, gap, tfile):
    d = os.path.join(golden, "synth_small")
    t, q = PackedGenome.read_2bit(os.path.join(d, tfile)), PackedGenome.read_2bit(os.path.join(d, "q.2bit"))
    cs = chainio.ChainSet.read(os.path.join(d, "in.chain"))
    jobs, total = cs.jobs(t, q)
    g, l = gpu_score(t, q, scoring_of(golden, matrix, gap), jobs, total, cs.blocks)
    rows = np.loadtxt(os.path.join(d, "scores_%s.tsv" % tag), dtype=np.int64)
    assert np.array_equal(g, rows[:, 1]) and np.array_equal(l, rows[:, 2])
    assert np.array_equal(ali_bases(jobs, total, cs.blocks), rows[:, 3])


def test_synth_small_reference_subchains(golden):
    d = os.path.join(golden, "synth_small")
    t, q = PackedGenome.read_2bit(os.path.join(d, "t.2bit")), PackedGenome.read_2bit(os.path.join(d, "q.2bit"))
    cs = chainio.ChainSet.read(os.path.join(d, "in.chain"))
    whole, _ = cs.jobs(t, q)
    rows = np.loadtxt(os.path.join(d, "sub_medium_default.tsv"), dtype=np.int64)
    jobs = np.zeros(len(rows), dtype=JOB_DTYPE); ptr = 0
    for k, (ix, s, e, is_null, *_rest) in enumerate(rows):
        fb, nb, c0, c1 = cs.subset_job(int(ix), int(s), int(e))
        jobs[k] = (whole[ix]["tSeq"], whole[ix]["qSeq"], fb, ptr, c0, c1); ptr += nb
    g, l = gpu_score(t, q, Scoring(None, "medium"), jobs, ptr, cs.blocks)
    live = rows[:, 3] == 0
    assert np.array_equal(g[live], rows[live, 4]) and np.array_equal(l[live], rows[live, 5])
    assert np.array_equal(ali_bases(jobs, ptr, cs.blocks)[live], rows[live, 6])
    assert np.all(g[~live] == 0) and np.all(l[~live] == 0)        # kent: NULL sub-chain


def oracle_scores(oracle, w, t_names, q_names, matrix, gap, tmp, jobs=None, total=None):
    paths = helpers.write_genomes(w, tmp)
    jobs = w.jobs if jobs is None else jobs
    total = w.total if total is None else total
    return oracle.score_jobs(oracle.scoring(matrix, gap), oracle.genome(paths["t"]), oracle.genome(paths["q"]),
                             jobs, total, w.blocks)


CASES = [
    # seed, n_blocks, kwargs
    (101, 30000, dict(max_chain_blocks=20000)),                       # jobs spanning many CTAs
    (102, 20000, dict(max_chain_blocks=3, n_fraction=0.2)),           # tiny jobs, lots of N
    (103, 8000, dict(mean_log_len=6.5, sigma_log_len=1.5, max_len=60000, max_chain_blocks=50)),   # long blocks
    (104, 25000, dict(mean_log_len=1.0, sigma_log_len=1.0, max_chain_blocks=2000, max_gap=3)),    # 1-10 bp blocks
    (105, 15000, dict(minus_fraction=1.0, max_chain_blocks=700)),
    (106, 15000, dict(minus_fraction=0.0, max_gap=2000000, gap_sigma=4.0, max_chain_blocks=100)), # huge gaps
]


@pytest.mark.parametrize("seed,n_blocks,kw", CASES)
@pytest.mark.parametrize("matrix,gap", [(None, "medium"), ("example/HoxD55.q", "loose"), ("synth_small/asym.q", "loose")])
def test_random_workloads_against_oracle(oracle, golden, tmp_path, seed, n_blocks, kw, matrix, gap):
    t_names, q_names = ["chrA", "chrB", "s3"], ["chrX", "chrY"]
    kw = dict(kw)
    nf = kw.pop("n_fraction", 0.01)
    w = synth.make_workload(t_names, [3000000, 800000, 70001], q_names, [2500000, 900003], n_blocks, seed=seed,
                            telomere_n=700, n_fraction=nf, **kw)
    m = os.path.join(golden, matrix) if matrix else None
    g, l = gpu_score(w.t, w.q, scoring_of(golden, matrix, gap), w.jobs, w.total, w.blocks)
    og, ol, oa = oracle_scores(oracle, w, t_names, q_names, m, gap, tmp_path)
    assert np.array_equal(g, og)
    assert np.array_equal(l, ol)
    assert np.array_equal(ali_bases(w.jobs, w.total, w.blocks), oa)


def small_world(seed=7, n_blocks=6000, **kw):
    t_names, q_names = ["chrA", "chrB"], ["chrX", "chrY"]
    w = synth.make_workload(t_names, [500000, 200000], q_names, [450000, 180000], n_blocks, seed=seed,
                            telomere_n=400, n_fraction=0.02, **kw)
    return w, t_names, q_names


def test_clipped_jobs_sharing_blocks(oracle, tmp_path):
    """chainNet / chainCleaner style: many sub-chain jobs pointing into the same chain records."""
    w, tn, qn = small_world(max_chain_blocks=3000)
    rng = np.random.default_rng(5)
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    jobs = []
    ptr = 0
    for _ in range(4000):
        j = int(rng.integers(0, len(w.jobs)))
        fb, nb = int(w.jobs[j]["firstBlock"]), int(counts[j])
        b = w.blocks[fb:fb + nb]
        lo, hi = int(b[0]["tStart"]), int(b[-1]["tStart"]) + int(b[-1]["size"])
        s = int(rng.integers(lo - 5, hi)); e = int(rng.integers(s, hi + 5))
        t_end = b["tStart"].astype(np.int64) + b["size"]
        keep = np.nonzero(t_end > s)[0]
        a = int(keep[0]) if len(keep) else nb
        stop = np.nonzero(b["tStart"][a:] >= e)[0]
        z = a + int(stop[0]) if len(stop) else nb
        jobs.append((w.jobs[j]["tSeq"], w.jobs[j]["qSeq"], fb + a, ptr, s, e)); ptr += z - a
    jobs = np.array(jobs, dtype=JOB_DTYPE)
    g, l = gpu_score(w.t, w.q, Scoring(None, "loose"), jobs, ptr, w.blocks)
    og, ol, oa = oracle_scores(oracle, w, tn, qn, None, "loose", tmp_path, jobs, ptr)
    assert np.array_equal(g, og) and np.array_equal(l, ol)
    assert np.array_equal(ali_bases(jobs, ptr, w.blocks), oa)


def test_joined_split_blocks_equal_unsplit(oracle, tmp_path):
    """GAT_BLOCK_JOINED: a long block cut into pieces scores exactly like the block."""
    w, tn, qn = small_world(seed=9, n_blocks=3000, mean_log_len=6.0, sigma_log_len=1.2, max_len=40000, max_chain_blocks=40)
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    nb, nj = [], []
    for j, job in enumerate(w.jobs):
        start = len(nb)
        for b in w.blocks[int(job["firstBlock"]):int(job["firstBlock"]) + int(counts[j])]:
            off, size, first = 0, int(b["size"]), True
            piece = 97
            while off < size:
                n = min(piece, size - off)
                nb.append((int(b["tStart"]) + off, int(b["qStart"]) + off, n | (0 if first else BLOCK_JOINED)))
                off += n; first = False; piece = piece * 3 + 1
        nj.append((job["tSeq"], job["qSeq"], start, start, NO_CLIP_START, NO_CLIP_END))
    blocks2 = np.array(nb, dtype=BLOCK_DTYPE); jobs2 = np.array(nj, dtype=JOB_DTYPE)
    g1, l1 = gpu_score(w.t, w.q, Scoring(None, "medium"), w.jobs, w.total, w.blocks)
    g2, l2 = gpu_score(w.t, w.q, Scoring(None, "medium"), jobs2, len(blocks2), blocks2)
    og, ol, _ = oracle_scores(oracle, w, tn, qn, None, "medium", tmp_path)
    assert np.array_equal(g1, og) and np.array_equal(l1, ol)
    assert np.array_equal(g2, og) and np.array_equal(l2, ol)


def test_edge_cases(oracle, tmp_path):
    rng = np.random.default_rng(1)
    sizes_t, sizes_q = [1000, 33, 4096], [999, 64, 5000]
    t = PackedGenome.from_codes(["t0", "t1", "t2"], [rng.integers(0, 4, n) for n in sizes_t],
                                n_runs=np.array([(0, 0, 3), (0, 500, 40), (0, 990, 10), (2, 31, 2)], dtype=NRUN_DTYPE))
    q = PackedGenome.from_codes(["q0", "q1", "q2"], [rng.integers(0, 4, n) for n in sizes_q],
                                n_runs=np.array([(0, 100, 1), (1, 0, 64), (2, 4990, 10)], dtype=NRUN_DTYPE))
    M = QSEQ_MINUS
    blocks = np.array([
        (0, 0, 1000 - 1), (0, 0, 999),            # whole sequences, + and -
        (999, 998, 1), (0, 0, 1),                 # last / first base
        (0, 0, 33), (0, 31, 33),                  # whole t1 against q1 (all N on the query)
        (0, 0, 32), (32, 32, 32), (64, 64, 64),   # exact word multiples
        (1, 7, 31), (40, 47, 33), (100, 200, 0),  # odd sizes and an empty block
        (0, 0, 4096), (0, 904, 4096),             # long block at both ends of q2
        (5, 5, 10), (30, 20, 10), (20, 40, 10),   # out-of-order blocks (negative gaps clamp to 0)
    ], dtype=BLOCK_DTYPE)
    jobs = np.array([
        (0, 0, 0, 0, NO_CLIP_START, NO_CLIP_END), (0, 0 | M, 1, 1, NO_CLIP_START, NO_CLIP_END),
        (0, 0, 2, 2, NO_CLIP_START, NO_CLIP_END), (0, 0 | M, 3, 3, NO_CLIP_START, NO_CLIP_END),
        (1, 1, 4, 4, NO_CLIP_START, NO_CLIP_END), (1, 1 | M, 5, 5, NO_CLIP_START, NO_CLIP_END),
        (0, 0, 6, 6, NO_CLIP_START, NO_CLIP_END), (0, 0 | M, 6, 9, NO_CLIP_START, NO_CLIP_END),
        (0, 0, 9, 12, NO_CLIP_START, NO_CLIP_END),
        (0, 0, 12, 15, NO_CLIP_START, NO_CLIP_END),                      # empty job (0 blocks)
        (2, 2, 12, 15, NO_CLIP_START, NO_CLIP_END), (2, 2 | M, 13, 16, NO_CLIP_START, NO_CLIP_END),
        (0, 0, 14, 17, NO_CLIP_START, NO_CLIP_END), (0, 0 | M, 14, 20, 8, 35),
        (0, 0, 6, 23, 10, 100),                                          # clip through word-multiple blocks
    ], dtype=JOB_DTYPE)
    total = 26
    w = synth.Workload(t, q, jobs, total, blocks)
    for matrix, gap in ((None, "medium"), (None, "loose")):
        g, l = gpu_score(t, q, Scoring(None, gap), jobs, total, blocks)
        og, ol, _ = oracle_scores(oracle, w, ["t0", "t1", "t2"], ["q0", "q1", "q2"], None, gap, tmp_path)
        assert np.array_equal(g, og), (g, og)
        assert np.array_equal(l, ol), (l, ol)
    assert g[4] == 0 and g[5] == 0 and g[9] == 0      # all-N query, empty job


@pytest.mark.parametrize("minus", [False, True])
def test_long_block_lengths_around_the_streaming_threshold(oracle, golden, tmp_path, minus):
    """Blocks of more than 1056 bases leave the item list and are streamed by the warp, 1024 bases a step and four steps a
    group: lengths on both sides of each of those edges, with N runs inside some, next to short blocks in the same tile."""
    rng = np.random.default_rng(11)
    lens = [1055, 1056, 1057, 1058, 40, 1087, 1088, 1089, 7, 2079, 2080, 2081, 4127, 4128, 4129, 4130, 33, 5152, 5153, 8224, 8225,
            12000, 1, 20001]
    size = sum(lens) + 3 * len(lens) + 500
    t_codes, q_codes = rng.integers(0, 4, size), rng.integers(0, 4, size + 77)
    blocks, tp, qp = [], 11, 60
    for n in lens:
        q_codes[qp:qp + n] = np.where(rng.random(n) < 0.8, t_codes[tp:tp + n], q_codes[qp:qp + n])
        blocks.append((tp, qp, n)); tp += n + 2; qp += n + 3
    t = PackedGenome.from_codes(["t"], [t_codes], n_runs=np.array([(0, 2000, 30), (0, 9000, 1), (0, 30000, 2500)], dtype=NRUN_DTYPE))
    if minus:                                                          # the chain runs on the query's minus strand
        q_codes = (q_codes[::-1] ^ 2).copy()                          # T=0 C=1 A=2 G=3
    q = PackedGenome.from_codes(["q"], [q_codes], n_runs=np.array([(0, 5000, 3), (0, 47000, 90)], dtype=NRUN_DTYPE))
    blocks = np.array(blocks, dtype=BLOCK_DTYPE)
    jobs = np.array([(0, QSEQ_MINUS if minus else 0, 0, 0, NO_CLIP_START, NO_CLIP_END),
                     (0, QSEQ_MINUS if minus else 0, 0, len(lens), 1500, int(blocks["tStart"][-1]) + 5000)], dtype=JOB_DTYPE)
    total = 2 * len(lens)
    w = synth.Workload(t, q, jobs, total, blocks)
    for matrix in (None, "synth_small/asym.q"):
        m = os.path.join(golden, matrix) if matrix else None
        g, l = gpu_score(t, q, scoring_of(golden, matrix, "loose"), jobs, total, blocks)
        og, ol, _ = oracle_scores(oracle, w, ["t"], ["q"], m, "loose", tmp_path)
        assert np.array_equal(g, og), (g, og)
        assert np.array_equal(l, ol), (l, ol)
    assert g[0] > 0


def test_the_library_picks_the_instantiation_from_the_block_sizes(long_blocks_mode):
    """gat_stats.long_streamed: short-block lists take the listing kernel, long-block lists the streaming one, for the
    one-shot, the compact and the resident call alike; GAT_LONG_BLOCKS overrides."""
    from genomealignmenttools_b200.records import pack_compact
    names_t, names_q = ["chrA"], ["chrX"]
    got = {}
    for tag, kw in (("short", {}), ("long", dict(mean_log_len=7.5, sigma_log_len=0.8, max_len=16000))):
        w = synth.make_workload(names_t, [4000000], names_q, [4000000], 6000, seed=21, telomere_n=100, **kw)
        with ChainScorer(0) as sc:
            sc.load_genome("t", w.t); sc.load_genome("q", w.q)
            sc.set_scoring(Scoring(None, "loose"))
            g0, l0 = sc.score(w.jobs, w.total, w.blocks)
            a = sc.stats()["long_streamed"]
            g1, l1 = sc.score_compact(*pack_compact(w.jobs, w.total, w.blocks))
            b = sc.stats()["long_streamed"]
            wl = sc.upload(w.jobs, w.total, w.blocks)
            wl.run()
            g2, l2 = wl.results()
            c = sc.stats()["long_streamed"]
            wl.free()
        assert np.array_equal(g0, g1) and np.array_equal(g0, g2) and np.array_equal(l0, l1) and np.array_equal(l0, l2)
        got[tag] = (a, b, c)
    want = {"auto": {"short": (0, 0, 0), "long": (1, 1, 1)}, "stream": {"short": (1, 1, 1), "long": (1, 1, 1)},
            "list": {"short": (0, 0, 0), "long": (0, 0, 0)}}[long_blocks_mode]
    assert got == want


@pytest.mark.parametrize("clip", [False, True])
def test_empty_jobs_anywhere_in_the_list(clip):
    """kent's NULL sub-chains (chain.c:535-539) as jobs without job-blocks: in front, at the end, in runs, on chunk
    boundaries.  They score 0 and every other job scores what it scores in the list without them -- through the one-shot
    call (the device reports them and the list is scored again where it lies), the resident list (seen on the host) and a
    second one-shot call on the same context (the scratch list forgets)."""
    w, _, _ = small_world(seed=5, n_blocks=40000, max_chain_blocks=3000)
    jobs = w.jobs.copy()
    if clip:                                            # clipped jobs take the general kernel
        jobs["clipStart"][::3] = 0
    rng = np.random.default_rng(8)
    n = len(jobs)
    where = np.sort(np.concatenate([[0, 0, 0, n, n], rng.integers(0, n + 1, 300), np.repeat(rng.integers(0, n + 1, 10), 7)]))
    ptr_of = np.append(jobs["blockPtr"], w.total)
    empties = np.zeros(len(where), dtype=JOB_DTYPE)
    empties["blockPtr"] = ptr_of[where]; empties["firstBlock"] = empties["blockPtr"]
    empties["tSeq"] = 0; empties["qSeq"] = 0; empties["clipStart"] = NO_CLIP_START; empties["clipEnd"] = NO_CLIP_END
    mixed = np.insert(jobs, where, empties)
    is_empty = np.insert(np.zeros(n, dtype=bool), where, True)
    assert np.all(np.diff(mixed["blockPtr"].astype(np.int64)) >= 0)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "medium"))
        g0, l0 = sc.score(jobs, w.total, w.blocks)
        g1, l1 = sc.score(mixed, w.total, w.blocks)
        wl = sc.upload(mixed, w.total, w.blocks)
        wl.run(); g2, l2 = wl.results()
        wl.run(); g3, l3 = wl.results()
        wl.free()
        g4, l4 = sc.score(jobs, w.total, w.blocks)
    for g, l in ((g1, l1), (g2, l2), (g3, l3)):
        assert np.all(g[is_empty] == 0) and np.all(l[is_empty] == 0)
        assert np.array_equal(g[~is_empty], g0) and np.array_equal(l[~is_empty], l0)
    assert np.array_equal(g4, g0) and np.array_equal(l4, l0)


def test_empty_and_invalid_worklists():
    w, _, _ = small_world(n_blocks=500)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "loose"))
        g, l = sc.score(np.zeros(0, dtype=JOB_DTYPE), 0, np.zeros(0, dtype=BLOCK_DTYPE))
        assert len(g) == 0
        empty = np.zeros(5, dtype=JOB_DTYPE)
        g, l = sc.score(empty, 0, w.blocks)
        assert np.all(g == 0) and np.all(l == 0)
        bad = w.blocks.copy(); bad["tStart"][3] = 2 ** 30
        with pytest.raises(GatError) as e:
            sc.score(w.jobs, w.total, bad)
        assert e.value.code == -4
        badj = w.jobs.copy(); badj["tSeq"][0] = 77
        with pytest.raises(GatError):
            sc.score(badj, w.total, w.blocks)
        badp = w.jobs.copy(); badp["blockPtr"] = badp["blockPtr"][::-1].copy()
        with pytest.raises(GatError) as e:
            sc.score(badp, w.total, w.blocks)
        assert e.value.code == -4
        toolong = w.blocks.copy(); toolong["size"][0] = 1 << 20
        with pytest.raises(GatError):
            sc.score(w.jobs, w.total, toolong)
        g, l = sc.score(w.jobs, w.total, w.blocks)      # the context survives a rejected work-list
        g2, l2 = sc.score(w.jobs, w.total, w.blocks)
        assert np.array_equal(g, g2) and np.array_equal(l, l2)
    with ChainScorer(0) as sc:
        with pytest.raises(GatError) as e:
            sc.score(w.jobs, w.total, w.blocks)
        assert e.value.code == -3


def test_record_length_limit_follows_the_matrix(oracle, tmp_path):
    """32-bit record sums: gat_max_record_bases() = min(2^20-1, (2^31-1)/max|M|), never below GAT_SPLIT_BASES; a record at the
    limit scores exactly, one past it is rejected, |M| beyond 2^18 is refused by gat_set_scoring."""
    rng = np.random.default_rng(4)
    n = 9000
    codes = rng.integers(0, 4, n)
    t = PackedGenome.from_codes(["t"], [codes])
    q = PackedGenome.from_codes(["q"], [codes.copy()])
    big = Scoring(ScoreScheme(np.array([[250000, -250000, -1, -2], [-250000, 250000, -3, -4], [-1, -3, 250000, -250000],
                                        [-2, -4, -250000, 250000]], dtype=np.int32)), "loose")
    with ChainScorer(0) as sc:
        sc.load_genome("t", t); sc.load_genome("q", q)
        sc.set_scoring(Scoring(None, "loose"))
        assert sc.max_record_bases() == (1 << 20) - 1
        sc.set_scoring(big)
        lim = sc.max_record_bases()
        assert lim == (2 ** 31 - 1) // 250000 and lim >= 4096
        jobs = np.array([(0, 0, 0, 0, NO_CLIP_START, NO_CLIP_END)], dtype=JOB_DTYPE)
        g, l = sc.score(jobs, 1, np.array([(0, 0, lim)], dtype=BLOCK_DTYPE))
        assert g[0] == 250000 * lim and l[0] == g[0]                    # identical sequences: every base a match
        with pytest.raises(GatError):
            sc.score(jobs, 1, np.array([(0, 0, lim + 1)], dtype=BLOCK_DTYPE))
        # the same block as JOINED records of GAT_SPLIT_BASES scores as one block and may exceed 32 bits
        pieces = [(o, o, min(4096, n - o) | (BLOCK_JOINED if o else 0)) for o in range(0, n, 4096)]
        g, l = sc.score(jobs, len(pieces), np.array(pieces, dtype=BLOCK_DTYPE))
        assert g[0] == 250000 * n and l[0] == g[0]
        too_big = Scoring(ScoreScheme(np.full((4, 4), 300000, dtype=np.int32)), "loose")
        with pytest.raises(GatError):
            sc.set_scoring(too_big)


def test_crossover_matches_oracle(oracle, tmp_path):
    """gat_crossover (cBlockFindCrossover, kent chainConnect.c:61-105) against the oracle: both strands, N runs, two
    matrices, overlaps of 0..150 bases at arbitrary offsets, sequence ends; out-of-range pairs are rejected."""
    from genomealignmenttools_b200.records import XPAIR_DTYPE
    from test_oracle import crossover_cases
    w, tn, qn = small_world(seed=21, n_blocks=3000)
    paths = helpers.write_case(w, tn, qn, tmp_path)
    tg, qg = oracle.genome(paths["t"]), oracle.genome(paths["q"])
    rng = np.random.default_rng(8)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q)
        for matrix in (None, "synth_small/asym.q"):
            golden = os.path.join(os.path.dirname(__file__), "golden")
            sc.set_scoring(scoring_of(golden, matrix, "loose"))
            s = oracle.scoring(os.path.join(golden, matrix) if matrix else None, "loose")
            for ti, qi, strand in ((0, 0, "+"), (1, 1, "-"), (0, 1, "-"), (1, 0, "+")):
                tsz, qsz = int(w.t.sizes[ti]), int(w.q.sizes[qi])
                cases = crossover_cases(rng, tsz, qsz, 600)
                cases += [(tsz, qsz, 0, 0, 64), (64, 64, tsz - 64, qsz - 64, 64), (1, 1, 0, 0, 1), (10, 10, 5, 5, 0)]
                # long overlaps (beyond XOVER_SHORT the warp takes a pair, lane = base): word boundaries and several words
                for ov in (65, 96, 97, 128, 500, 3000):
                    for _ in range(6):
                        lt, lq = int(rng.integers(ov, tsz)), int(rng.integers(ov, qsz))
                        cases.append((lt, lq, int(rng.integers(0, tsz - ov)), int(rng.integers(0, qsz - ov)), ov))
                # overlaps along planted homology: block starts of real chains, shifted against themselves
                for j in np.nonzero((w.jobs["tSeq"] == ti) & ((w.jobs["qSeq"] & 0x7FFFFFFF) == qi) &
                                    ((w.jobs["qSeq"] >> 31) == (1 if strand == "-" else 0)))[0][:150]:
                    b = w.blocks[int(w.jobs[j]["firstBlock"])]
                    ov = int(min(b["size"], 90))
                    if ov >= 4:
                        cases.append((int(b["tStart"]) + ov, int(b["qStart"]) + ov, int(b["tStart"]) + ov // 3, int(b["qStart"]) + ov // 3, ov - ov // 3))
                pairs = np.zeros(len(cases), dtype=XPAIR_DTYPE)
                pairs["tSeq"] = ti; pairs["qSeq"] = qi | (QSEQ_MINUS if strand == "-" else 0)
                for k, name in enumerate(("leftTEnd", "leftQEnd", "rightTStart", "rightQStart", "overlap")):
                    pairs[name] = [c[k] for c in cases]
                pos, adj = sc.crossover(pairs)
                opos, oadj = oracle.crossover(s, tg, qg, ti, qi, strand, cases)
                assert np.array_equal(pos, opos) and np.array_equal(adj, oadj)
                assert (pos > 0).sum() > 10
        bad = np.zeros(1, dtype=XPAIR_DTYPE); bad["leftTEnd"] = 5; bad["leftQEnd"] = 5; bad["overlap"] = 9
        with pytest.raises(GatError):
            sc.crossover(bad)
        pos, adj = sc.crossover(np.zeros(0, dtype=XPAIR_DTYPE))
        assert len(pos) == 0


def test_remove_partial_overlaps_matches_reference(kentref, tmp_path):
    """gathost::removePartialOverlaps (host loop + gat_crossover batches) against the unmodified
    chainRemovePartialOverlaps (kent chainConnect.c:255-344): chains whose neighbouring blocks overlap by a few bases,
    by most of a block, or swallow a block whole; both strands."""
    if kentref is None:
        pytest.skip("oracle/_ref not built")
    import ctypes
    import hostlib
    w, tn, qn = small_world(seed=33, n_blocks=4000, max_chain_blocks=60, max_len=1500)
    rng = np.random.default_rng(2)
    counts = job_block_counts(w.jobs, w.total)
    blocks = w.blocks.copy()
    for j in range(len(w.jobs)):
        fb, nb = int(w.jobs[j]["firstBlock"]), int(counts[j])
        for i in range(fb, fb + nb - 1):
            if rng.random() < 0.4:
                nxt = blocks[i + 1]
                gap = min(int(nxt["tStart"]) - (int(blocks[i]["tStart"]) + int(blocks[i]["size"])), int(nxt["qStart"]) - (int(blocks[i]["qStart"]) + int(blocks[i]["size"])))
                room = int(nxt["size"]) - 1 if rng.random() < 0.8 else int(nxt["size"]) + 3     # sometimes past the whole next block
                ext = gap + int(rng.integers(1, max(2, room)))
                # starts must keep increasing (checkChainIncreases): the extension moves only this block's end
                blocks[i]["size"] = int(blocks[i]["size"]) + max(0, ext)
    w2 = synth.Workload(w.t, w.q, w.jobs, w.total, blocks)
    paths = helpers.write_case(w2, tn, qn, tmp_path)
    kentref.set_scoring(None, "loose")
    n = kentref.open(paths["t"], paths["q"], paths["chain"])
    lib = hostlib.load()
    cs = lib.gathost_chains_read(paths["chain"].encode())
    assert cs, lib.gathost_last_error()
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "loose"))
        ct = np.ascontiguousarray(w.jobs["tSeq"], dtype=np.uint32); cq = np.ascontiguousarray(w.jobs["qSeq"] & 0x7FFFFFFF, dtype=np.uint32)
        lib.gathost_chains_remove_partial_overlaps.argtypes = [ctypes.c_void_p] * 4
        rc = lib.gathost_chains_remove_partial_overlaps(cs, sc.ctx, ct.ctypes.data, cq.ctypes.data)
        assert rc == 0, lib.gathost_last_error()
    heads = hostlib.chain_heads(lib, cs)
    ours = hostlib.chain_blocks(lib, cs)
    changed = 0
    for c in range(n):
        ref_blocks, bounds = kentref.remove_partial_overlaps(c)
        h = heads[c]
        mine = ours[h["firstBlock"]:h["firstBlock"] + h["nBlocks"]]
        got = np.stack([mine["tStart"], mine["qStart"], mine["size"].astype(np.int32)], axis=1)
        assert np.array_equal(got, ref_blocks), c
        assert [h["tStart"], h["tEnd"], h["qStart"], h["qEnd"]] == list(bounds)
        changed += int(len(ref_blocks) != counts[c])
    assert changed > 5          # some blocks dried up and were removed


@pytest.mark.parametrize("n_blocks,kw", [(9000, {}), (1024, dict(max_chain_blocks=5)), (3, dict(max_chain_blocks=2)),
                                         (30000, dict(gap_mu=6.0, gap_sigma=3.0, max_gap=900000, max_len=60000, mean_log_len=5.0))])
def test_compact_worklist_scores_like_plain(n_blocks, kw):
    """gat_score_compact (delta-coded work-list expanded on the device) gives the scores of gat_score on the same list."""
    from genomealignmenttools_b200.records import pack_compact, split_long_blocks
    w, _, _ = small_world(seed=44, n_blocks=n_blocks, **kw)
    jobs, total, blocks = split_long_blocks(w.jobs, w.total, w.blocks, 4096)
    if n_blocks == 9000:        # out-of-order blocks inside a chain: a negative gap takes the absolute escape
        blocks = blocks.copy()
        i = int(jobs["firstBlock"][np.argmax(job_block_counts(jobs, total))]) + 1
        blocks["qStart"][i + 1] = blocks["qStart"][i] - 3
    cj, cb, ab, an = pack_compact(jobs, total, blocks)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "medium"))
        g, l = sc.score(jobs, total, blocks)
        cg, cl = sc.score_compact(cj, cb, ab, an)
        assert np.array_equal(g, cg) and np.array_equal(l, cl)
        escapes = np.nonzero((cb["size"] & 0x8000) != 0)[0]
        escapes = escapes[escapes % 1024 != 0]                                   # a group's first record takes its anchor
        if len(escapes):
            bad = cb.copy(); bad["dt"][escapes[0]] = 0xFFFF; bad["dq"][escapes[0]] = 0xFFFF      # absolute index out of range
            with pytest.raises(GatError):
                sc.score_compact(cj, bad, ab, an)
            g2, l2 = sc.score_compact(cj, cb, ab, an)                            # the context survives
            assert np.array_equal(g, g2)


@pytest.mark.parametrize("n_blocks,kw", [(9000, {}), (1024, dict(max_chain_blocks=5)), (3, dict(max_chain_blocks=2)), (2049, {}),
                                         (30000, dict(gap_mu=6.0, gap_sigma=3.0, max_gap=900000, max_len=60000, mean_log_len=5.0)),
                                         (1_200_000, dict(max_chain_blocks=400000))])
def test_packed_worklist_scores_like_plain(n_blocks, kw):
    """gat_score_packed (4-byte words, absolute records in list order, expanded on the device; sliced above a million blocks)
    gives the scores of gat_score on the same chains."""
    from genomealignmenttools_b200.records import pack_packed, PBLOCK_ABS
    if n_blocks > 100000:
        w = synth.make_workload(["chrA", "chrB"], [60_000_000, 9_000_000], ["chrX", "chrY"], [50_000_000, 8_000_000], n_blocks, seed=45,
                                telomere_n=400, n_fraction=0.002, **kw)
    else:
        w, _, _ = small_world(seed=45, n_blocks=n_blocks, **kw)
    blocks = w.blocks
    if n_blocks == 9000:        # out-of-order blocks inside a chain: a negative gap is an absolute record
        blocks = blocks.copy()
        i = int(w.jobs["firstBlock"][np.argmax(job_block_counts(w.jobs, w.total))]) + 1
        blocks["qStart"][i + 1] = blocks["qStart"][i] - 3
    cj, pb, ab, an, base = pack_packed(w.jobs, w.total, blocks)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "medium"))
        g, l = sc.score(w.jobs, w.total, blocks)
        pg, pl = sc.score_packed(cj, pb, ab, an, base)
        assert np.array_equal(g, pg) and np.array_equal(l, pl)
        if n_blocks > 100000:   # pinned result arrays: the scores of a slice's jobs come back while later slices arrive
            from genomealignmenttools_b200.engine import PinnedArray
            from genomealignmenttools_b200.records import pack_compact, split_long_blocks
            hg, hl = PinnedArray(len(cj), np.int64), PinnedArray(len(cj), np.int64)
            hg.array[:] = -7; hl.array[:] = -7
            sc.score_packed(cj, pb, ab, an, base, hg.array, hl.array)
            assert np.array_equal(g, hg.array) and np.array_equal(l, hl.array)
            hg.array[:] = -7; hl.array[:] = -7
            sc.score_compact(*pack_compact(*split_long_blocks(w.jobs, w.total, blocks, 4096)), hg.array, hl.array)
            assert np.array_equal(g, hg.array) and np.array_equal(l, hl.array)
            hg.free(); hl.free()
        if len(ab) > 1:         # a table that is too short is rejected, and the context survives
            with pytest.raises(GatError):
                sc.score_packed(cj, pb, ab[:-1], an, base)
            g2, l2 = sc.score_packed(cj, pb, ab, an, base)
            assert np.array_equal(g, g2) and np.array_equal(l, l2)


def test_resident_worklist_matches_one_shot():
    w, _, _ = small_world(seed=12, n_blocks=9000)
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "medium"))
        g, l = sc.score(w.jobs, w.total, w.blocks)
        wl = sc.upload(w.jobs, w.total, w.blocks)
        sc.set_profiling(True)
        for _ in range(3):
            wl.run()
        g2, l2 = wl.results()
        st = sc.stats()
        wl.free()
    assert np.array_equal(g, g2) and np.array_equal(l, l2)
    assert st["kernel_launches"] == 3 and st["score_kernel_ms"] > 0


@pytest.mark.parametrize("pieces_share", [400_000, 1_500_000])
def test_chain_split_into_pieces_joins_to_the_unsplit_scores(oracle, tmp_path, pieces_share):
    """a giant chain cut at block boundaries, its pieces scored as jobs of their own (as if on different GPUs),
    joined on the host from their tuples (gat_request_tuples, gat_tuple_join): bit-identical to the unsplit chain and to
    the oracle, through gat_score, the resident work-list and gat_score_compact."""
    from genomealignmenttools_b200 import sharding
    from genomealignmenttools_b200.records import pack_compact
    w = synth.make_workload(["chrA", "chrB"], [40_000_000, 9_000_000], ["chrX", "chrY"], [36_000_000, 8_000_000], 90_000,
                            seed=21, telomere_n=2000, n_fraction=0.002, zipf_s=1.2, max_chain_blocks=30_000)
    scoring = Scoring(None, "medium")
    paths = helpers.write_case(w, ["chrA", "chrB"], ["chrX", "chrY"], str(tmp_path))
    og, ol, _ = oracle.score_jobs(oracle.scoring(None, "medium"), oracle.genome(paths["t"]), oracle.genome(paths["q"]),
                                  w.jobs, w.total, w.blocks)
    pj, origin, first_piece = sharding.split_giant_jobs(w.jobs, w.total, w.blocks, 2, share=pieces_share)
    cut = np.nonzero(np.diff(first_piece) > 1)[0]
    assert len(cut) >= 1
    part_ix = np.concatenate([np.arange(first_piece[j], first_piece[j + 1]) for j in cut])
    with ChainScorer(0) as sc:
        sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(scoring)
        wg, wl_ = sc.score(w.jobs, w.total, w.blocks)
        assert np.array_equal(wg, og) and np.array_equal(wl_, ol)
        results = []
        tup = sc.request_tuples(part_ix)
        g, l = sc.score(pj, w.total, w.blocks)
        results.append((g, l, tup.copy()))
        res = sc.upload(pj, w.total, w.blocks)
        tup = sc.request_tuples(part_ix)
        res.run()
        g, l = res.results()
        results.append((g, l, tup.copy()))
        res.free()
        cj, cb, ab, an = pack_compact(pj, w.total, w.blocks)
        tup = sc.request_tuples(part_ix)
        g, l = sc.score_compact(cj, cb, ab, an)
        results.append((g, l, tup.copy()))
        # a request names jobs that the fix-up kernel finishes: a short job is refused
        short = np.nonzero(job_block_counts(pj, w.total) < 200)[0][:1]
        sc.request_tuples(short)
        with pytest.raises(GatError):
            sc.score(pj, w.total, w.blocks)
    for g, l, tup in results:
        tuples = {int(p): tup[k] for k, p in enumerate(part_ix)}
        jg, jl = sharding.join_pieces(w.jobs, w.total, w.blocks, pj, origin, first_piece, g, l, tuples, scoring.gap.cost)
        assert np.array_equal(jg, og) and np.array_equal(jl, ol)


GAP_SMALL30 = """tableSize 9
smallSize 30
position 1 2 3 11 30 111 2111 12111 32111
qGap 325 360 400 450 600 900 2900 22900 57900
tGap 325 360 400 450 600 900 2900 22900 57900
bothGap 625 660 700 750 900 1400 4900 37900 97900
"""
# last knot beyond 2^22: the dense device table stops at 2^22, gaps between there and the knot take the exact routine
GAP_FAR_KNOT = """tableSize 12
smallSize 111
position 1 2 3 11 111 2111 12111 32111 72111 152111 252111 5000000
qGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900 4317900
tGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900 4317901
bothGap 750 825 850 1000 1300 3300 23300 58300 118300 218300 318300 6318300
"""


@pytest.mark.parametrize("name,text", [("small30", GAP_SMALL30), ("far_knot", GAP_FAR_KNOT)])
def test_custom_linear_gap_files_on_the_device(oracle, kentref, tmp_path, name, text):
    """-linearGap=<file> with smallSize != 111 and with a last knot beyond the dense table (gat_capi.cu: denseSize is capped at
    2^22): gap costs evaluated on the device (gat_gap_cost, the routines of the scoring kernel) and whole chains scored
    with them match the oracle and, where it is built, the unmodified reference's gapCalcCost."""
    gap_file = str(tmp_path / (name + ".gap"))
    open(gap_file, "w").write(text)
    t_names, q_names = ["chrA"], ["chrX"]
    w = synth.make_workload(t_names, [12_000_000], q_names, [11_000_000], 4000, seed=41, telomere_n=500, n_fraction=0.01,
                            max_chain_blocks=300, gap_mu=6.0, gap_sigma=3.5, max_gap=6_000_000)
    paths = helpers.write_case(w, t_names, q_names, str(tmp_path))
    sc = oracle.scoring(None, gap_file)
    rng = np.random.default_rng(8)
    dq = np.concatenate([np.arange(0, 140), rng.integers(0, 9_000_000, 3000), [29, 30, 31, 110, 111, 112, 4194303, 4194304, 4194305,
                                                                                  4999999, 5000000, 5000001, 2147483647, -5]])
    dt = np.concatenate([np.zeros(140, dtype=np.int64), rng.integers(0, 3, 3000) * rng.integers(0, 4_000_000, 3000), np.zeros(14, dtype=np.int64)])
    dq, dt = np.concatenate([dq, dt]), np.concatenate([dt, dq])
    want = np.array([oracle.lib.orc_gap_cost(sc, int(a), int(b)) for a, b in zip(dq, dt)], dtype=np.int64)
    if kentref is not None:
        kentref.set_scoring(None, gap_file)
        live = np.array([kentref.lib.ref_gap_cost(int(a), int(b)) for a, b in zip(dq, dt)], dtype=np.int64)
        ok = (dq.astype(np.int64) + dt.astype(np.int64)) < 2 ** 31          # dq + dt overflowing int is undefined in the reference
        assert np.array_equal(want[ok], live[ok])
    scoring = Scoring(None, GapCalc.from_file(gap_file))
    with ChainScorer(0) as s:
        s.load_genome("t", w.t); s.load_genome("q", w.q); s.set_scoring(scoring)
        got = s.gap_cost(dq, dt)
        ok = (dq.astype(np.int64) + dt.astype(np.int64)) < 2 ** 31
        assert np.array_equal(got[ok].astype(np.int64), want[ok])
        g, l = s.score(w.jobs, w.total, w.blocks)
    og, ol, _ = oracle.score_jobs(sc, oracle.genome(paths["t"]), oracle.genome(paths["q"]), w.jobs, w.total, w.blocks)
    assert np.array_equal(g, og) and np.array_equal(l, ol)
