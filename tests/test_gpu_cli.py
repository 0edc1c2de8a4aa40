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


def test_scorechain_two_gpus_same_bytes(golden, tmp_path):
    if _native.load().gat_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = os.path.join(golden, "synth_small")
    out = str(tmp_path / "out.txt")
    r = run([os.path.join(d, "in.chain"), os.path.join(d, "t.2bit"), os.path.join(d, "q.2bit"), out, "-linearGap=medium", "-gpus=2"])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(out, os.path.join(d, "out_chain_medium.txt"), shallow=False)
