"""bin/chainCleaner against the reference binary: same suspects, same decisions, same files.
Config 3 of BASELINE.json in the small: synthetic chains are scored, sorted, netted and filtered
with the reference tools, then cleaned by both implementations."""
import filecmp
import os
import subprocess
import pytest
import make_golden_helpers as helpers
from genomealignmenttools_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
REFBIN = os.path.join(ROOT, "oracle", "_ref")


def test_usage_and_errors(tmp_path):
    exe = os.path.join(BIN, "chainCleaner")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr.startswith("chainCleaner - Remove chain-breaking alignments")
    r = subprocess.run([exe, "a", "b", "c", "d", "e"], capture_output=True, text=True)
    assert r.returncode == 255 and "Must specify linear gap costs" in r.stderr
    r = subprocess.run([exe, "a", "b.2bit", "c.2bit", "d", "e", "-linearGap=loose"], capture_output=True, text=True)
    assert r.returncode == 255 and "ERROR: target 2bit file or nib directory b.2bit does not exist" in r.stderr


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    if not os.path.exists(os.path.join(REFBIN, "chainCleaner")):
        pytest.skip("oracle/_ref not built")
    d = tmp_path_factory.mktemp("cleaner")
    t_names, q_names = ["chr1", "chr2", "chrUn_1"], ["chrA", "chrB", "chrC"]
    t_sizes, q_sizes = [1500000, 600000, 40000], [1300000, 300000, 500000]
    w = synth.make_workload(t_names, t_sizes, q_names, q_sizes, 60000, seed=41, telomere_n=500, n_fraction=0.01,
                            max_chain_blocks=4000, subst=0.25)
    helpers.write_case(w, t_names, q_names, d)
    for name, names, sizes in (("t.sizes", t_names, t_sizes), ("q.sizes", q_names, q_sizes)):
        with open(d / name, "w") as f:
            f.write("".join("%s\t%d\n" % p for p in zip(names, sizes)))
    env = dict(os.environ, PATH=REFBIN + ":" + os.environ["PATH"])
    run = lambda cmd, **kw: subprocess.check_call(cmd, cwd=d, env=env, stderr=subprocess.DEVNULL, **kw)
    run([os.path.join(REFBIN, "scoreChain"), "in.chain", "t.2bit", "q.2bit", "scored.chain", "-linearGap=loose", "-forceLocalScore"])
    run([os.path.join(REFBIN, "chainSort"), "scored.chain", "sorted.chain"])
    run("chainNet -minScore=0 sorted.chain t.sizes q.sizes stdout /dev/null | NetFilterNonNested.perl /dev/stdin -minScore1 3000 > in.net",
        shell=True, executable="/bin/bash")
    return d


OPTION_SETS = {
    "suspects": ["-suspectDataFile=SIDE.suspects.bed"],
    "relaxed_pairs": ["-minBrokenChainScore=3000", "-LRfoldThreshold=1.2", "-doPairs", "-LRfoldThresholdPairs=1.5", "-newChainIDDict=SIDE.dict"],
    "relaxed_single": ["-minBrokenChainScore=1000", "-LRfoldThreshold=1.0", "-maxSuspectScore=50000", "-minLRGapSize=2"],
    "default": [],
}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(OPTION_SETS))
def test_matches_reference_files(case, tag):
    d = case
    outs = {}
    for side, bindir in (("ref", REFBIN), ("our", BIN)):
        env = dict(os.environ, PATH=bindir + ":" + os.environ["PATH"])
        opts = [o.replace("SIDE", "%s.%s" % (side, tag)) for o in OPTION_SETS[tag]]
        r = subprocess.run([os.path.join(bindir, "chainCleaner"), "sorted.chain", "t.2bit", "q.2bit", "%s.%s.chain" % (side, tag),
                            "%s.%s.bed" % (side, tag), "-net=in.net", "-linearGap=loose"] + opts, cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[side] = r
    for ext in ("chain", "bed") + (("suspects.bed",) if tag == "suspects" else ()) + (("dict",) if tag == "relaxed_pairs" else ()):
        a, b = d / ("ref.%s.%s" % (tag, ext)), d / ("our.%s.%s" % (tag, ext))
        assert filecmp.cmp(a, b, shallow=False), "%s differs for %s" % (ext, tag)
    if tag == "suspects":
        assert sum(1 for _ in open(d / "ref.suspects.suspects.bed")) > 50
    if tag.startswith("relaxed"):
        assert sum(1 for _ in open(d / ("ref.%s.bed" % tag))) > 20


@pytest.mark.gpu
def test_debug_side_files_match_reference(case):
    """-debug: chainsOfInterest.chain, the four sub-chain files and suspectsAndFills.bed (chainCleaner.c:591-616, 1312-1321),
    written to the working directory; each side runs in a directory of its own."""
    d = case
    names = ["chainsOfInterest.chain", "suspect.chain", "brokenChainLfill.chain", "brokenChainRfill.chain", "brokenChainfill.chain",
             "suspectsAndFills.bed", "out.chain", "out.bed"]
    for side, bindir in (("ref", REFBIN), ("our", BIN)):
        wd = d / ("debug_" + side)
        os.makedirs(wd, exist_ok=True)
        env = dict(os.environ, PATH=bindir + ":" + os.environ["PATH"])
        r = subprocess.run([os.path.join(bindir, "chainCleaner"), str(d / "sorted.chain"), str(d / "t.2bit"), str(d / "q.2bit"), "out.chain", "out.bed",
                            "-net=" + str(d / "in.net"), "-linearGap=loose", "-debug", "-minBrokenChainScore=3000", "-LRfoldThreshold=1.2", "-doPairs",
                            "-LRfoldThresholdPairs=1.5"], cwd=wd, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stdout == ""
    for n in names:
        assert filecmp.cmp(d / "debug_ref" / n, d / "debug_our" / n, shallow=False), n
    assert sum(1 for l in open(d / "debug_ref" / "suspectsAndFills.bed") if "REMOVED_" in l) > 10
