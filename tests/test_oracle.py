"""Pins the CPU oracle (oracle/chain_oracle.c): against the reference's own golden chains, the
known-answer vectors of SURVEY.md section 4, fixtures produced by the compiled reference, and --
when oracle/_ref has been built in this container -- the unmodified reference in process."""
import os
import numpy as np
import pytest
from genomealignmenttools_b200 import chainio
from genomealignmenttools_b200.twobit import PackedGenome
from genomealignmenttools_b200.records import JOB_DTYPE, NO_CLIP_START, NO_CLIP_END

# SURVEY.md section 4 (captured from the compiled reference)
MEDIUM_KAT = {(0, 0): 0, (1, 0): 350, (0, 1): 350, (1, 1): 825, (2, 0): 425, (3, 0): 450, (5, 0): 487, (10, 0): 581,
              (11, 0): 600, (12, 0): 603, (50, 0): 717, (110, 0): 897, (111, 0): 900, (112, 0): 901, (500, 0): 1289,
              (2111, 0): 2900, (2112, 0): 2902, (5000, 0): 8678, (12111, 0): 22900, (20000, 0): 36705,
              (100000, 0): 152761, (252110, 0): 317899, (252111, 0): 317900, (252112, 0): 317901,
              (1000000, 0): 1065789, (1, 2): 850, (5, 6): 1000, (55, 55): 1297, (55, 56): 1300,
              (6548, 2467): 17108, (16472, 34876): 87155, (100000, 152111): 318300, (500000, 500000): 1066189,
              (84240, 6489540): 6639969, (-5, 3): 450, (-1, -1): 0}
LOOSE_KAT = {(1, 0): 325, (1, 1): 660, (2, 0): 360, (3, 0): 400, (5, 0): 412, (10, 0): 443, (11, 0): 450, (12, 0): 451,
             (50, 0): 508, (110, 0): 598, (111, 0): 600, (112, 0): 600, (500, 0): 697, (2111, 0): 1100,
             (2112, 0): 1100, (5000, 0): 1822, (12111, 0): 3600, (20000, 0): 5177, (100000, 0): 21177,
             (252110, 0): 56599, (252111, 0): 56600, (252112, 0): 56600, (1000000, 0): 243572, (1, 2): 700,
             (5, 6): 750, (55, 55): 898, (55, 56): 900, (6548, 2467): 3195, (16472, 34876): 11847,
             (100000, 152111): 57000, (500000, 500000): 243972, (84240, 6489540): 1637417, (-5, 3): 400}


def score_chain_file(oracle, chain_path, t2bit, q2bit, matrix, gap):
    tg, qg = oracle.genome(t2bit), oracle.genome(q2bit)
    t, q = PackedGenome.read_2bit(t2bit), PackedGenome.read_2bit(q2bit)
    cs = chainio.ChainSet.read(chain_path)
    jobs, total = cs.jobs(t, q)
    s = oracle.scoring(matrix, gap)
    return cs, oracle.score_jobs(s, tg, qg, jobs, total, cs.blocks)


@pytest.mark.parametrize("spec,kat", [("medium", MEDIUM_KAT), ("loose", LOOSE_KAT)])
def test_gap_cost_known_answers(oracle, spec, kat):
    s = oracle.scoring(None, spec)
    for (dq, dt), want in kat.items():
        assert oracle.lib.orc_gap_cost(s, dq, dt) == want, (spec, dq, dt)


def test_gap_cost_reference_fixture(oracle, golden):
    sc = {k: oracle.scoring(None, k) for k in ("medium", "loose")}
    n = 0
    for line in open(os.path.join(golden, "gap_kat.tsv")):
        if line.startswith("#"):
            continue
        spec, dq, dt, cost = line.split()
        assert oracle.lib.orc_gap_cost(sc[spec], int(dq), int(dt)) == int(cost), line
        n += 1
    assert n > 5000


@pytest.mark.parametrize("name,want", [("newStyleLastz", 671823), ("oldStyleBlastz", 671644)])
def test_kent_axtchain_golden(oracle, golden, name, want):
    """kent/src/hg/mouseStuff/axtChain/tests: the chain score in expected/*.chain was written by
    chainCalcScore (axtChain.c:290-293) with -linearGap=loose and the matching .Q.txt."""
    d = os.path.join(golden, "kent_chrM")
    cs, (g, l, a) = score_chain_file(oracle, os.path.join(d, name + ".chain"), os.path.join(d, "hg19.chrM.2bit"),
                                     os.path.join(d, "susScr3.chrM.2bit"), os.path.join(d, name + ".Q.txt"), "loose")
    assert len(cs) == 1 and int(cs.score[0]) == want
    assert g[0] == want and l[0] == want


def test_chrM_known_answers(oracle, golden):
    d = os.path.join(golden, "kent_chrM")
    args = (os.path.join(d, "newStyleLastz.chain"), os.path.join(d, "hg19.chrM.2bit"), os.path.join(d, "susScr3.chrM.2bit"))
    _, (g, l, a) = score_chain_file(oracle, *args, os.path.join(golden, "example", "HoxD55.q"), "loose")
    assert (g[0], l[0], a[0]) == (829041, 829041, 15310)
    _, (g, l, a) = score_chain_file(oracle, *args, None, "medium")
    assert (g[0], l[0], a[0]) == (767508, 767508, 15310)


def read_scores(path):
    rows = np.loadtxt(path, dtype=np.int64, ndmin=2)
    return rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3]


@pytest.mark.parametrize("tag,matrix,gap", [
    ("medium_default", None, "medium"), ("loose_hoxd55", "example/HoxD55.q", "loose"),
    ("loose_lastz", "kent_chrM/newStyleLastz.Q.txt", "loose"), ("medium_asym", "synth_small/asym.q", "medium")])
def test_synth_small_against_reference_scorechain(oracle, golden, tag, matrix, gap):
    d = os.path.join(golden, "synth_small")
    m = os.path.join(golden, matrix) if matrix else None
    cs, (g, l, a) = score_chain_file(oracle, os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"),
                                     os.path.join(d, "q.2bit"), m, gap)
    ids, rg, rl, ra = read_scores(os.path.join(d, "scores_%s.tsv" % tag))
    assert np.array_equal(ids, cs.id)
    assert np.array_equal(g, rg) and np.array_equal(l, rl) and np.array_equal(a, ra)
    assert (rg > 0).sum() > 20 and (rg < 0).sum() > 20      # the fixture covers both signs


def test_byte_swapped_2bit(oracle, golden):
    d = os.path.join(golden, "synth_small")
    a, b = oracle.genome(os.path.join(d, "t.2bit")), oracle.genome(os.path.join(d, "t.swapped.2bit"))
    for i in range(oracle.lib.orc_genome_count(a)):
        n = oracle.lib.orc_genome_size(a, i)
        import ctypes
        da = ctypes.string_at(oracle.lib.orc_genome_dna(a, i, b"+"), n)
        db = ctypes.string_at(oracle.lib.orc_genome_dna(b, i, b"+"), n)
        assert da == db and b"N" in da and any(c in da for c in (b"a", b"c", b"g", b"t"))


def test_subchain_clipping_against_reference(oracle, golden):
    """chainSubsetOnT + chainCalcScore(+Local) answers of the compiled reference for 600 ranges."""
    d = os.path.join(golden, "synth_small")
    rows = np.loadtxt(os.path.join(d, "sub_medium_default.tsv"), dtype=np.int64)
    tg, qg = oracle.genome(os.path.join(d, "t.2bit")), oracle.genome(os.path.join(d, "q.2bit"))
    t, q = PackedGenome.read_2bit(os.path.join(d, "t.2bit")), PackedGenome.read_2bit(os.path.join(d, "q.2bit"))
    ocs = oracle.chains(os.path.join(d, "in.chain"))
    cs = chainio.ChainSet.read(os.path.join(d, "in.chain"))
    whole, _ = cs.jobs(t, q)
    jobs = np.zeros(len(rows), dtype=JOB_DTYPE)
    ptr = 0
    for k, (ix, s, e, is_null, g, l, a) in enumerate(rows):
        ok, fb, nb, c0, c1 = oracle.subset(ocs, int(ix), int(s), int(e))
        assert (not ok) == bool(is_null)
        assert (fb, nb, c0, c1) == cs.subset_job(int(ix), int(s), int(e))   # host builder agrees with the oracle
        jobs[k] = (whole[ix]["tSeq"], whole[ix]["qSeq"], fb, ptr, c0, c1)
        ptr += nb
    g, l, a = oracle.score_jobs(oracle.scoring(None, "medium"), tg, qg, jobs, ptr, cs.blocks)
    live = rows[:, 3] == 0
    assert live.sum() > 400 and (~live).sum() > 20
    assert np.array_equal(g[live], rows[live, 4]) and np.array_equal(l[live], rows[live, 5])
    assert np.array_equal(a[live], rows[live, 6])
    assert np.all(g[~live] == 0) and np.all(l[~live] == 0)


def test_chain_reader_matches_oracle_reader(oracle, golden):
    for path in (os.path.join(golden, "synth_small", "in.chain"), os.path.join(golden, "example", "hg38.danRer10.chain")):
        cs = chainio.ChainSet.read(path)
        ocs = oracle.chains(path)
        heads = oracle.chain_headers(ocs)
        assert len(heads) == len(cs)
        for i, h in enumerate(heads):
            assert (h["tName"], h["tStart"], h["tEnd"], h["qName"], h["qStrand"], h["qStart"], h["qEnd"], h["id"]) == \
                (cs.tName[i], cs.tStart[i], cs.tEnd[i], cs.qName[i], cs.qStrand[i], cs.qStart[i], cs.qEnd[i], cs.id[i])
        assert np.array_equal(oracle.chain_blocks(ocs), cs.blocks)


# ---------------------------------------------------------------- live reference (build container only)
def test_oracle_vs_live_reference_random(oracle, kentref, tmp_path):
    if kentref is None:
        pytest.skip("oracle/_ref not built here")
    from genomealignmenttools_b200 import synth
    import make_golden_helpers as helpers
    rng = np.random.default_rng(99)
    for seed, gap, matrix in ((21, "loose", None), (22, "medium", os.path.join(os.path.dirname(__file__), "golden/example/HoxD55.q"))):
        w = synth.make_workload(["a", "b"], [50021, 20002], ["x", "y", "z"], [40003, 9001, 30000], 1500, seed=seed,
                                telomere_n=200, n_fraction=0.03, max_chain_blocks=300, max_gap=300000 if seed == 21 else 3000)
        paths = helpers.write_case(w, ["a", "b"], ["x", "y", "z"], tmp_path / ("case%d" % seed))
        kentref.set_scoring(matrix, gap)
        n = kentref.open(paths["t"], paths["q"], paths["chain"])
        rg, rl, ra = kentref.score_all(n)
        s = oracle.scoring(matrix, gap)
        g, l, a = oracle.score_jobs(s, oracle.genome(paths["t"]), oracle.genome(paths["q"]), w.jobs, w.total, w.blocks)
        assert np.array_equal(g, rg) and np.array_equal(l, rl) and np.array_equal(a, ra)
        # clipped sub-chains
        cs = chainio.ChainSet.read(paths["chain"])
        ix = rng.integers(0, n, 300)
        s0 = np.array([rng.integers(cs.tStart[i] - 50, cs.tEnd[i]) for i in ix])
        e0 = np.array([rng.integers(s0[k], cs.tEnd[i] + 50) for k, i in enumerate(ix)])
        rg, rl, ra, rz = kentref.score_sub(ix, s0, e0)
        jobs = np.zeros(len(ix), dtype=JOB_DTYPE); ptr = 0
        for k, i in enumerate(ix):
            fb, nb, c0, c1 = cs.subset_job(int(i), int(s0[k]), int(e0[k]))
            jobs[k] = (w.jobs[i]["tSeq"], w.jobs[i]["qSeq"], fb, ptr, c0, c1); ptr += nb
            assert (nb == 0) == bool(rz[k])
        g, l, a = oracle.score_jobs(oracle.scoring(matrix, gap), oracle.genome(paths["t"]), oracle.genome(paths["q"]),
                                    jobs, ptr, w.blocks)
        live = rz == 0
        assert np.array_equal(g[live], rg[live]) and np.array_equal(l[live], rl[live]) and np.array_equal(a[live], ra[live])


def crossover_cases(rng, t_size, q_size, n, max_overlap=150):
    """Random (leftTEnd, leftQEnd, rightTStart, rightQStart, overlap) with every overlap inside both sequences."""
    out = []
    for _ in range(n):
        ov = int(rng.integers(0, max_overlap))
        lt = int(rng.integers(ov, t_size + 1)); lq = int(rng.integers(ov, q_size + 1))
        rt = int(rng.integers(0, t_size - ov + 1)); rq = int(rng.integers(0, q_size - ov + 1))
        out.append((lt, lq, rt, rq, ov))
    return out


def test_crossover_against_live_reference(oracle, kentref, tmp_path):
    """orc_find_crossover vs the unmodified cBlockFindCrossover (kent chainConnect.c:61-105), both strands, with N."""
    if kentref is None:
        pytest.skip("oracle/_ref not built here")
    from genomealignmenttools_b200 import synth
    import make_golden_helpers as helpers
    rng = np.random.default_rng(123)
    w = synth.make_workload(["a", "b"], [30011, 8002], ["x", "y"], [25003, 9001], 800, seed=77, telomere_n=300, n_fraction=0.05,
                            max_chain_blocks=100)
    paths = helpers.write_case(w, ["a", "b"], ["x", "y"], tmp_path)
    for matrix in (None, os.path.join(os.path.dirname(__file__), "golden/synth_small/asym.q")):
        kentref.set_scoring(matrix, "loose")
        kentref.open(paths["t"], paths["q"], paths["chain"])
        s = oracle.scoring(matrix, "loose")
        tg, qg = oracle.genome(paths["t"]), oracle.genome(paths["q"])
        for (ti, tn, tsz), (qi, qn, qsz), strand in (((0, "a", 30011), (0, "x", 25003), "+"), ((1, "b", 8002), (1, "y", 9001), "-"),
                                                     ((0, "a", 30011), (1, "y", 9001), "-")):
            cases = crossover_cases(rng, tsz, qsz, 400)
            # overlaps of homologous stretches (planted along chain blocks) are where a crossover is not trivial
            cases += [(lt, lq, lt - ov // 2, lq - ov // 2, ov) for lt, lq, _, _, ov in cases[:200] if lt - ov // 2 + ov <= tsz and lq - ov // 2 + ov <= qsz]
            rp, ra = kentref.crossover(tn, qn, strand, cases)
            op, oa = oracle.crossover(s, tg, qg, ti, qi, strand, cases)
            assert np.array_equal(op, rp) and np.array_equal(oa, ra)
            assert (rp > 0).sum() > 20
