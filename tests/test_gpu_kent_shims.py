"""kent's own entry points (include/gat_kent.h, libgatkent.so) called from a C program that was compiled against
kent's headers (oracle/_ref/kent_shim_check, built by oracle/Makefile while /root/reference is present): chainCalcScore,
chainCalcScoreSubChain, chainScoreBlock, gapCalcCost, gapCalcFromFile, axtScoreSchemeRead / Default must return what the
oracle computes for the same chain -- on both strands (the caller hands in the reverse-complemented query, like
scoreChain.c:123-149 does), with N runs and soft-masked (lower-case) sequence."""
import ctypes
import os
import subprocess
import numpy as np
import pytest
from genomealignmenttools_b200 import synth
from genomealignmenttools_b200.records import job_block_counts
import make_golden_helpers as helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "oracle", "_ref", "kent_shim_check")


@pytest.mark.parametrize("matrix,gap", [(None, "medium"), ("example/HoxD55.q", "loose")])
def test_kent_named_entry_points_match_the_oracle(oracle, golden, tmp_path, matrix, gap):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/kent_shim_check not built (needs /root/reference at build time)")
    t_names, q_names = ["chrA"], ["chrX"]
    w = synth.make_workload(t_names, [900_000], q_names, [800_000], 3000, seed=17, telomere_n=300, n_fraction=0.02,
                            max_chain_blocks=400, max_len=9000)
    paths = helpers.write_case(w, t_names, q_names, str(tmp_path))
    mpath = os.path.join(golden, matrix) if matrix else None
    sc = oracle.scoring(mpath, gap)
    tg, qg = oracle.genome(paths["t"]), oracle.genome(paths["q"])
    og, ol, oa = oracle.score_jobs(sc, tg, qg, w.jobs, w.total, w.blocks)
    counts = job_block_counts(w.jobs, w.total)
    rng = np.random.default_rng(3)
    pairs = [(0, 0), (1, 0), (0, 1), (110, 0), (0, 111), (5, 7), (3000, 0), (0, 80000), (252110, 1), (252111, 0), (600000, 7)]
    pairs += [(int(a), int(b)) for a, b in zip(rng.integers(0, 300000, 20), rng.integers(0, 300, 20))]
    # the longest chain of each strand
    for minus in (0, 1):
        strand_jobs = np.nonzero((w.jobs["qSeq"] >> 31) == minus)[0]
        j = int(strand_jobs[np.argmax(counts[strand_jobs])])
        assert counts[j] > 20
        tdna = ctypes.string_at(oracle.lib.orc_genome_dna(tg, 0, b"+")).decode()
        qdna = ctypes.string_at(oracle.lib.orc_genome_dna(qg, 0, b"-" if minus else b"+")).decode()
        fb = int(w.jobs["firstBlock"][j])
        blocks = w.blocks[fb:fb + int(counts[j])]
        case = str(tmp_path / ("case%d.txt" % minus))
        with open(case, "w") as f:
            f.write("gap %s\nscheme %s\ntarget %s\nquery %s\n" % (gap, mpath or "-", tdna, qdna))
            f.write("chain %d %d\n" % (j + 1, len(blocks)))
            for b in blocks:
                f.write("%d %d %d %d\n" % (b["tStart"], b["tStart"] + b["size"], b["qStart"], b["qStart"] + b["size"]))
            f.write("pairs %d\n" % len(pairs))
            for dq, dt in pairs:
                f.write("%d %d\n" % (dq, dt))
        r = subprocess.run([CHECK, case], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        out = r.stdout.split("\n")
        assert out[0] == "layout ok"
        assert out[1] == "global %d" % og[j]
        assert out[2] == "local %d ali %d" % (ol[j], oa[j])
        assert out[3] == "sub %d" % og[j]
        b0 = blocks[0]
        want_b0 = oracle.lib.orc_score_block(sc, qdna[int(b0["qStart"]):].encode(), tdna[int(b0["tStart"]):].encode(), int(b0["size"]))
        assert out[4] == "block0 %d" % int(want_b0)
        for k, (dq, dt) in enumerate(pairs):
            assert out[5 + k] == "gap %d %d %d" % (dq, dt, oracle.lib.orc_gap_cost(sc, dq, dt))
