"""bin/chainNet against the reference binary's .net files.  Without -rescore the tool is pure host
logic (runs anywhere); with -rescore the partial fills are scored on the GPU."""
import filecmp
import os
import subprocess
import numpy as np
import pytest
import make_golden_helpers as helpers
from genomealignmenttools_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "chainNet")
REFBIN = os.path.join(ROOT, "oracle", "_ref")


def run(exe, args):
    return subprocess.run([exe] + args, capture_output=True, text=True)


@pytest.mark.parametrize("tag,opts", [("expected_plain", ["-minSpace=5", "-minScore=0"]), ("expected_default", [])])
def test_plain_net_matches_reference_bytes(golden, tmp_path, tag, opts):
    d = os.path.join(golden, "synth_small")
    t, q = str(tmp_path / "t.net"), str(tmp_path / "q.net")
    r = run(EXE, opts + [os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"), t, q])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(t, os.path.join(d, tag + ".t.net"), shallow=False)
    assert filecmp.cmp(q, os.path.join(d, tag + ".q.net"), shallow=False)


def test_usage_and_errors(golden, tmp_path):
    d = os.path.join(golden, "synth_small")
    r = run(EXE, [])
    assert r.returncode == 255 and r.stderr.startswith("chainNet - Make alignment nets out of chains\nusage:")
    args = [os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"), str(tmp_path / "a"), str(tmp_path / "b")]
    r = run(EXE, args + ["-rescore"])
    assert r.returncode == 255 and "you must specify the target genome file" in r.stderr
    r = run(EXE, args + ["-rescore", "-tNibDir=x.2bit", "-qNibDir=y.2bit"])
    assert r.returncode == 255 and r.stderr.startswith("Must specify linear gap costs")
    r = run(EXE, ["-minScore=0", os.path.join(d, "out_chain_loose_forceLocal.txt")] + args[1:])         # unsorted input
    assert r.returncode == 255 and "must be sorted in order of score" in r.stderr
    r = run(EXE, [args[0], args[2], args[1]] + args[3:])           # sizes swapped
    assert r.returncode == 255 and ("not found" in r.stderr or " but " in r.stderr)


@pytest.mark.gpu
def test_rescore_matches_reference_bytes(golden, tmp_path):
    d = os.path.join(golden, "synth_small")
    t, q = str(tmp_path / "t.net"), str(tmp_path / "q.net")
    r = run(EXE, ["-rescore", "-linearGap=medium", "-minSpace=5", "-minScore=0", "-tNibDir=" + os.path.join(d, "t.2bit"),
                  "-qNibDir=" + os.path.join(d, "q.2bit"), os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"),
                  os.path.join(d, "q.sizes"), t, q])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(t, os.path.join(d, "expected.t.net"), shallow=False)
    assert filecmp.cmp(q, os.path.join(d, "expected.q.net"), shallow=False)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,gap,matrix", [(31, "medium", None), (32, "loose", "example/HoxD55.q")])
def test_rescore_against_live_reference_binaries(golden, tmp_path, seed, gap, matrix):
    """config 2 in the small: reference scoreChain -> chainSort -> chainNet -rescore vs bin/chainNet -rescore."""
    if not os.path.exists(os.path.join(REFBIN, "chainNet")):
        pytest.skip("oracle/_ref not built")
    t_names, q_names = ["chr1", "chr2", "chrUn_1"], ["chrA", "chrB_alt", "chrC"]
    t_sizes, q_sizes = [1500000, 600000, 40000], [1300000, 300000, 500000]
    w = synth.make_workload(t_names, t_sizes, q_names, q_sizes, 60000, seed=seed, telomere_n=500, n_fraction=0.01,
                            max_chain_blocks=4000, subst=0.25)
    paths = helpers.write_case(w, t_names, q_names, tmp_path)
    for name, names, sizes in (("t.sizes", t_names, t_sizes), ("q.sizes", q_names, q_sizes)):
        with open(tmp_path / name, "w") as f:
            f.write("".join("%s\t%d\n" % p for p in zip(names, sizes)))
    mopt = ["-scoreScheme=" + os.path.join(golden, matrix)] if matrix else []
    env = dict(os.environ, PATH=REFBIN + ":" + os.environ["PATH"])
    subprocess.check_call([os.path.join(REFBIN, "scoreChain"), paths["chain"], paths["t"], paths["q"], str(tmp_path / "scored.chain"),
                           "-linearGap=" + gap, "-forceLocalScore"] + mopt, env=env)
    subprocess.check_call([os.path.join(REFBIN, "chainSort"), str(tmp_path / "scored.chain"), str(tmp_path / "sorted.chain")], env=env)
    common = ["-rescore", "-linearGap=" + gap, "-tNibDir=" + paths["t"], "-qNibDir=" + paths["q"]] + mopt + \
             [str(tmp_path / "sorted.chain"), str(tmp_path / "t.sizes"), str(tmp_path / "q.sizes")]
    subprocess.check_call([os.path.join(REFBIN, "chainNet")] + common + [str(tmp_path / "ref.t.net"), str(tmp_path / "ref.q.net")],
                          env=env, stderr=subprocess.DEVNULL)
    r = run(EXE, common + [str(tmp_path / "our.t.net"), str(tmp_path / "our.q.net")])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(str(tmp_path / "our.t.net"), str(tmp_path / "ref.t.net"), shallow=False)
    assert filecmp.cmp(str(tmp_path / "our.q.net"), str(tmp_path / "ref.q.net"), shallow=False)
    n_fill = sum(1 for l in open(tmp_path / "ref.t.net") if l.lstrip().startswith("fill"))
    assert n_fill > 1000


def write_nib(path, genome, i):
    """kent nib: sig, size, then two bases per byte, first base in the high nibble; T=0 C=1 A=2 G=3 N=4, +8 = masked."""
    import struct
    codes = genome.codes(i).astype(np.uint8)
    for r in genome.n_runs[genome.n_runs["seq"] == i]:
        codes[int(r["start"]):int(r["start"]) + int(r["len"])] = 4
    for r in genome.mask_runs[genome.mask_runs["seq"] == i]:
        codes[int(r["start"]):int(r["start"]) + int(r["len"])] |= 8
    if len(codes) & 1:
        codes = np.append(codes, np.uint8(0))
    with open(path, "wb") as f:
        f.write(struct.pack("<II", 0x6BE93D3A, int(genome.sizes[i])))
        ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tofile(f)


@pytest.mark.gpu
def test_rescore_from_nib_directories(golden, tmp_path):
    """tNibDir / qNibDir may be directories of .nib files (chainNet.c:166-170): same nets as from .2bit."""
    from genomealignmenttools_b200.twobit import PackedGenome
    d = os.path.join(golden, "synth_small")
    for side in ("t", "q"):
        g = PackedGenome.read_2bit(os.path.join(d, side + ".2bit"))
        os.makedirs(tmp_path / (side + "nib"))
        for i, name in enumerate(g.names):
            write_nib(tmp_path / (side + "nib") / (name + ".nib"), g, i)
    t, q = str(tmp_path / "t.net"), str(tmp_path / "q.net")
    r = run(EXE, ["-rescore", "-linearGap=medium", "-minSpace=5", "-minScore=0", "-tNibDir=" + str(tmp_path / "tnib"),
                  "-qNibDir=" + str(tmp_path / "qnib"), os.path.join(d, "sorted.chain"), os.path.join(d, "t.sizes"),
                  os.path.join(d, "q.sizes"), t, q])
    assert r.returncode == 0, r.stderr
    assert filecmp.cmp(t, os.path.join(d, "expected.t.net"), shallow=False)
    assert filecmp.cmp(q, os.path.join(d, "expected.q.net"), shallow=False)
    if os.path.exists(os.path.join(REFBIN, "chainNet")):     # and the reference reads our nib files the same way
        rt, rq = str(tmp_path / "rt.net"), str(tmp_path / "rq.net")
        subprocess.check_call([os.path.join(REFBIN, "chainNet"), "-rescore", "-linearGap=medium", "-minSpace=5", "-minScore=0",
                               "-tNibDir=" + str(tmp_path / "tnib"), "-qNibDir=" + str(tmp_path / "qnib"), os.path.join(d, "sorted.chain"),
                               os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes"), rt, rq], stderr=subprocess.DEVNULL)
        assert filecmp.cmp(rt, t, shallow=False)
