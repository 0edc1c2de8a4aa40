#!/usr/bin/env python
"""Regenerates tests/golden/ -- run in the build container, where /root/reference exists and
`make -C oracle ref` has produced oracle/_ref/.  The GPU box only ever reads the committed files.

  kent_chrM/   the reference's own golden test of the hot path, copied verbatim (data only):
               kent/src/hg/mouseStuff/axtChain/tests/{input,expected}  (tests/makefile:10-72)
  example/     example/HoxD55.q, example/hg38.danRer10.chain, example/*.chrom.sizes
  synth_small/ a seeded synthetic case (both strands, N runs, mask runs, long gaps) and the
               answers of the UNMODIFIED reference binaries / libkentref.so on it
  gap_kat.tsv  gapCalcCost(dq, dt) of the reference for medium and loose
"""
import os
import shutil
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
REFBIN = os.path.join(ROOT, "oracle", "_ref")

from genomealignmenttools_b200 import synth, chainio  # noqa: E402
from genomealignmenttools_b200.records import NRUN_DTYPE  # noqa: E402
import oracle_lib  # noqa: E402


def run(cmd, **kw):
    env = dict(os.environ, PATH=REFBIN + ":" + os.environ["PATH"])
    subprocess.check_call(cmd, env=env, **kw)


def copy_reference_data():
    src = os.path.join(REF, "kent/src/hg/mouseStuff/axtChain/tests")
    dst = os.path.join(HERE, "kent_chrM")
    os.makedirs(dst, exist_ok=True)
    for rel in ["input/hg19.chrM.2bit", "input/susScr3.chrM.2bit", "input/newStyleLastz.Q.txt",
                "input/oldStyleBlastz.Q.txt", "expected/newStyleLastz.chain", "expected/oldStyleBlastz.chain"]:
        shutil.copy(os.path.join(src, rel), os.path.join(dst, os.path.basename(rel)))
    dst = os.path.join(HERE, "example")
    os.makedirs(dst, exist_ok=True)
    for rel in ["HoxD55.q", "hg38.danRer10.chain", "hg38.chrom.sizes", "mm10.chrom.sizes", "danRer10.chrom.sizes"]:
        shutil.copy(os.path.join(REF, "example", rel), os.path.join(dst, rel))


def chain_headers(w, t_names, q_names):
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    b = w.blocks
    heads = []
    for j, job in enumerate(w.jobs):
        fb, nb = int(job["firstBlock"]), int(counts[j])
        last = b[fb + nb - 1]
        ts, qs = int(job["tSeq"]), int(job["qSeq"] & 0x7FFFFFFF)
        heads.append((0, t_names[ts], int(w.t.sizes[ts]), int(b[fb]["tStart"]), int(last["tStart"]) + int(last["size"]),
                      q_names[qs], int(w.q.sizes[qs]), "-" if job["qSeq"] >> 31 else "+",
                      int(b[fb]["qStart"]), int(last["qStart"]) + int(last["size"]), j + 1))
    return heads, counts


def make_synth_small():
    d = os.path.join(HERE, "synth_small")
    os.makedirs(d, exist_ok=True)
    t_names, t_sizes = ["chrT1", "chrT2", "scaffold_3"], [90007, 41011, 6003]
    q_names, q_sizes = ["chrQ1", "chrQ2", "contigQ"], [80021, 35002, 9001]
    w = synth.make_workload(t_names, t_sizes, q_names, q_sizes, 2500, seed=11, telomere_n=300, n_fraction=0.02,
                            max_chain_blocks=400, max_gap=20000, gap_sigma=2.5)
    # soft-mask runs: irrelevant to scores (axt.c:402-421) but they exercise the container parser
    rng = np.random.default_rng(3)
    for g in (w.t, w.q):
        runs = []
        for i, s in enumerate(g.sizes):
            starts = np.sort(rng.choice(int(s) - 200, size=5, replace=False))
            for st in starts:
                runs.append((i, int(st), int(rng.integers(10, 150))))
        g.mask_runs = np.array(runs, dtype=NRUN_DTYPE)
    w.t.write_2bit(os.path.join(d, "t.2bit"))
    w.q.write_2bit(os.path.join(d, "q.2bit"), version=1)
    w.t.write_2bit(os.path.join(d, "t.swapped.2bit"), swapped=True)
    heads, counts = chain_headers(w, t_names, q_names)
    chainio.write_chains(os.path.join(d, "in.chain"), heads, w.blocks, w.jobs["firstBlock"], counts)
    for name, sizes, names in (("t.sizes", t_sizes, t_names), ("q.sizes", q_sizes, q_names)):
        with open(os.path.join(d, name), "w") as f:
            for n, s in zip(names, sizes):
                f.write("%s\t%d\n" % (n, s))
    # ---- reference answers
    hox = os.path.join(REF, "example/HoxD55.q")
    for tag, opts in (("medium_default", ["-linearGap=medium"]),
                      ("loose_hoxd55", ["-linearGap=loose", "-scoreScheme=" + hox]),
                      ("loose_lastz", ["-linearGap=loose", "-scoreScheme=" + os.path.join(HERE, "kent_chrM/newStyleLastz.Q.txt")])):
        run([os.path.join(REFBIN, "scoreChain"), os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"),
             os.path.join(d, "q.2bit"), os.path.join(d, "scores_%s.tsv" % tag), "-returnOnlyScore"] + opts)
    # an asymmetric matrix exercises the general (non strand-symmetric) device path
    with open(os.path.join(d, "asym.q"), "w") as f:
        f.write("   A    C    G    T\n  90 -101  -33 -120\n -114   97 -125  -31\n  -29 -130  103 -111\n -123  -35 -104   88\n")
    run([os.path.join(REFBIN, "scoreChain"), os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"),
         os.path.join(d, "q.2bit"), os.path.join(d, "scores_medium_asym.tsv"), "-returnOnlyScore",
         "-linearGap=medium", "-scoreScheme=" + os.path.join(d, "asym.q")])
    # the other output modes of the tool, byte for byte
    for tag, opts in (("chain_medium", ["-linearGap=medium"]), ("chain_medium_doLocal", ["-linearGap=medium", "-doLocalScore"]),
                      ("chain_loose_forceLocal", ["-linearGap=loose", "-forceLocalScore"]),
                      ("coords_medium", ["-linearGap=medium", "-returnOnlyScoreAndCoords"])):
        run([os.path.join(REFBIN, "scoreChain"), os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"),
             os.path.join(d, "q.2bit"), os.path.join(d, "out_%s.txt" % tag)] + opts)
    # sub-chain (clip) answers straight from chainSubsetOnT + chainCalcScore + chainCalcScoreLocal
    ref = oracle_lib.load_ref()
    ref.set_scoring(None, "medium")
    n = ref.open(os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), os.path.join(d, "in.chain"))
    cs = chainio.ChainSet.read(os.path.join(d, "in.chain"))
    assert n == len(cs)
    ix, ss, ee = [], [], []
    for _ in range(600):
        c = int(rng.integers(0, n))
        a, b = cs.tStart[c], cs.tEnd[c]
        span = b - a
        s = int(rng.integers(a - span // 4 - 2, b + 2))
        e = int(rng.integers(s, b + span // 4 + 3))
        if rng.random() < 0.2:   # ranges that start/end exactly on block edges
            blk = cs.blocks[cs.firstBlock[c] + int(rng.integers(0, cs.nBlocks[c]))]
            s = int(blk["tStart"]) + int(rng.integers(0, 2)) * int(blk["size"])
        ix.append(c); ss.append(s); ee.append(max(e, s))
    g, l, a, z = ref.score_sub(ix, ss, ee)
    with open(os.path.join(d, "sub_medium_default.tsv"), "w") as f:
        f.write("#chainIx\tsubStart\tsubEnd\tisNull\tglobal\tlocal\taliBases\n")
        for k in range(len(ix)):
            f.write("%d\t%d\t%d\t%d\t%d\t%d\t%d\n" % (ix[k], ss[k], ee[k], z[k], g[k], l[k], a[k]))
    # chainNet -rescore on the score-sorted chains (config 2 in miniature)
    run([os.path.join(REFBIN, "scoreChain"), os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"),
         os.path.join(d, "q.2bit"), os.path.join(d, "rescored.chain"), "-linearGap=medium", "-forceLocalScore"])
    run([os.path.join(REFBIN, "chainSort"), os.path.join(d, "rescored.chain"), os.path.join(d, "sorted.chain")])
    os.remove(os.path.join(d, "rescored.chain"))
    run([os.path.join(REFBIN, "chainNet"), "-rescore", "-linearGap=medium", "-minSpace=5", "-minScore=0",
         "-tNibDir=" + os.path.join(d, "t.2bit"), "-qNibDir=" + os.path.join(d, "q.2bit"),
         os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"),
         os.path.join(d, "expected.t.net"), os.path.join(d, "expected.q.net")])
    # the same net without -rescore (approximate sub-scores): pure host logic, checkable without a GPU
    run([os.path.join(REFBIN, "chainNet"), "-minSpace=5", "-minScore=0", os.path.join(d, "sorted.chain"),
         os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"),
         os.path.join(d, "expected_plain.t.net"), os.path.join(d, "expected_plain.q.net")])
    run([os.path.join(REFBIN, "chainNet"), os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"),
         os.path.join(d, "expected_default.t.net"), os.path.join(d, "expected_default.q.net")])


def make_gap_kat():
    ref = oracle_lib.load_ref()
    rng = np.random.default_rng(7)
    pts = [(a, b) for a in (-3, 0, 1, 2, 3, 10, 11, 12, 50, 110, 111, 112) for b in (-1, 0, 1, 2, 55, 56, 110, 111, 112)]
    for scale in (300, 3000, 30000, 300000, 3000000, 240000000):
        for _ in range(150):
            pts.append((int(rng.integers(0, scale)), 0))
            pts.append((0, int(rng.integers(0, scale))))
            pts.append((int(rng.integers(0, scale)), int(rng.integers(0, scale))))
    pts += [(2111, 0), (2112, 0), (12111, 0), (32111, 0), (72111, 0), (152111, 0), (252110, 0), (252111, 0),
            (252112, 0), (100000, 152111), (500000, 500000), (84240, 6489540), (1075188, 2746361)]
    with open(os.path.join(HERE, "gap_kat.tsv"), "w") as f:
        f.write("#linearGap\tdq\tdt\tcost\n")
        for spec in ("medium", "loose"):
            ref.set_scoring(None, spec)
            for a, b in pts:
                f.write("%s\t%d\t%d\t%d\n" % (spec, a, b, ref.lib.ref_gap_cost(a, b)))


if __name__ == "__main__":
    if not os.path.isdir(REF) or oracle_lib.load_ref() is None:
        sys.exit("needs /root/reference and `make -C oracle ref`")
    copy_reference_data()
    make_gap_kat()
    make_synth_small()
    print("golden fixtures written under", HERE)
