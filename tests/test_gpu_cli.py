"""The command-line drop-in on a GPU: bin/scoreChain must write the same bytes as the reference
binary did (fixtures under tests/golden/ were produced by the unmodified src/scoreChain)."""
import filecmp
import os
import subprocess
import pytest
from genomealignmenttools_b200 import _native

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "scoreChain")


def run(args, **kw):
    return subprocess.run([EXE] + args, capture_output=True, text=True, **kw)


@pytest.mark.parametrize("fixture,opts", [
    ("scores_medium_default.tsv", ["-linearGap=medium", "-returnOnlyScore"]),
    ("scores_loose_hoxd55.tsv", ["-linearGap=loose", "-returnOnlyScore", "-scoreScheme=GOLDEN/example/HoxD55.q"]),
    ("scores_loose_lastz.tsv", ["-returnOnlyScore", "-scoreScheme=GOLDEN/kent_chrM/newStyleLastz.Q.txt", "-linearGap=loose"]),
    ("scores_medium_asym.tsv", ["-linearGap=medium", "-returnOnlyScore", "-scoreScheme=GOLDEN/synth_small/asym.q"]),
    ("out_chain_medium.txt", ["-linearGap=medium"]),
    ("out_chain_medium_doLocal.txt", ["-linearGap=medium", "-doLocalScore"]),
    ("out_chain_loose_forceLocal.txt", ["-forceLocalScore", "-linearGap=loose"]),
    ("out_coords_medium.txt", ["-linearGap=medium", "-returnOnlyScoreAndCoords"]),
])
def test_scorechain_matches_reference_bytes(golden, tmp_path, fixture, opts):
    d = os.path.join(golden, "synth_small")
    out = str(tmp_path / "out.txt")
    opts = [o.replace("GOLDEN", golden) for o in opts]
    # options may sit anywhere on the command line (options.c:306)
    r = run([opts[0], os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), out] + opts[1:])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(out, os.path.join(d, fixture), shallow=False)


def test_scorechain_kent_golden_chain(golden, tmp_path):
    d = os.path.join(golden, "kent_chrM")
    out = str(tmp_path / "out.chain")
    r = run([os.path.join(d, "newStyleLastz.chain"), os.path.join(d, "hg19.chrM.2bit"), os.path.join(d, "susScr3.chrM.2bit"), out,
             "-linearGap=loose", "-scoreScheme=" + os.path.join(d, "newStyleLastz.Q.txt")])
    assert r.returncode == 0, r.stderr
    # rescoring the reference's expected chain with its own parameters reproduces it exactly; the '#'
    # header lines are dropped, as scoreChain does (kent/src/lib/linefile.c:907-922)
    want = "".join(l for l in open(os.path.join(d, "newStyleLastz.chain")) if not l.startswith("#"))
    assert open(out).read() == want


def test_scorechain_stdout_and_missing_sequence(golden, tmp_path):
    d = os.path.join(golden, "synth_small")
    r = run([os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), "stdout", "-linearGap=medium", "-returnOnlyScore"])
    assert r.returncode == 0 and r.stdout == open(os.path.join(d, "scores_medium_default.tsv")).read()
    k = os.path.join(golden, "kent_chrM")
    r = run([os.path.join(d, "in.chain"), os.path.join(k, "hg19.chrM.2bit"), os.path.join(d, "q.2bit"), str(tmp_path / "o"), "-linearGap=medium"])
    assert r.returncode == 255 and "is not in" in r.stderr


@pytest.mark.parametrize("gpus", [2, 3])
def test_scorechain_sharded_over_contexts_same_bytes(golden, tmp_path, gpus):
    """-gpus=N on a one-GPU box: GAT_DEVICES=0,0,... opens N contexts on device 0, so the sharded path (per-shard compacted
    records, pinned staging, gat_score_compact per shard) runs wherever the tests run."""
    d = os.path.join(golden, "synth_small")
    out = str(tmp_path / "out.txt")
    env = dict(os.environ, GAT_DEVICES=",".join(["0"] * gpus), GAT_TOOL_TIMING="1")
    r = run([os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), out, "-linearGap=medium", "-gpus=%d" % gpus], env=env)
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(out, os.path.join(d, "out_chain_medium.txt"), shallow=False)
    shards = [l for l in r.stderr.splitlines() if l.startswith("gpu shard")]
    assert len(shards) == gpus and all("(compact)" in l for l in shards)


def test_scorechain_cuts_a_giant_chain_over_contexts(tmp_path):
    """A chain with half of all aligned bases: -gpus=4 (four contexts on device 0) cuts it into pieces, scores them on
    different contexts and joins the tuples; the output file equals the one-context run byte for byte, every context
    receives about a quarter of the bytes."""
    import numpy as np
    import make_golden_helpers as helpers
    from genomealignmenttools_b200 import synth
    rng = np.random.default_rng(5)
    t_names, q_names = ["chrA", "chrB"], ["chrX", "chrY"]
    w = synth.make_workload(t_names, [30_000_000, 6_000_000], q_names, [28_000_000, 5_000_000], 60_000, seed=31,
                            telomere_n=1000, n_fraction=0.002, zipf_s=1.15, max_chain_blocks=40_000)
    paths = helpers.write_case(w, t_names, q_names, str(tmp_path))
    outs = {}
    for gpus in (1, 4):
        out = str(tmp_path / ("out%d.chain" % gpus))
        env = dict(os.environ, GAT_DEVICES=",".join(["0"] * gpus), GAT_TOOL_TIMING="1")
        r = run([paths["chain"], paths["t"], paths["q"], out, "-linearGap=loose", "-gpus=%d" % gpus], env=env)
        assert r.returncode == 0, r.stderr
        outs[gpus] = (out, [l for l in r.stderr.splitlines() if l.startswith("gpu shard")])
    assert filecmp.cmp(outs[1][0], outs[4][0], shallow=False)
    one = int(outs[1][1][0].split(" bytes")[0].split()[-1])
    four = [int(l.split(" bytes")[0].split()[-1]) for l in outs[4][1]]
    pieces = sum(int(l.split(" pieces")[0].split()[-1]) for l in outs[4][1])
    assert len(four) == 4 and pieces >= 2, outs[4][1]
    assert max(four) < 0.45 * one and abs(sum(four) - one) < 0.1 * one, (one, four)


def test_scorechain_two_gpus_same_bytes(golden, tmp_path):
    if _native.load().gat_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = os.path.join(golden, "synth_small")
    out = str(tmp_path / "out.txt")
    r = run([os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), out, "-linearGap=medium", "-gpus=2"])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(out, os.path.join(d, "out_chain_medium.txt"), shallow=False)
