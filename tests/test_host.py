"""Host logic (no GPU): the Python plumbing and the C++ host library against the oracle, the
reference fixtures and each other; the C ABI libraries load and export what the headers declare."""
import ctypes
import os
import re
import subprocess
import numpy as np
import pytest
from genomealignmenttools_b200 import _native, chainio, synth
from genomealignmenttools_b200.records import JOB_DTYPE, BLOCK_DTYPE, NRUN_DTYPE, ali_bases, jobs_from_counts
from genomealignmenttools_b200.scoring import GapCalc, ScoreScheme, ScoreSchemeError, GapCalcError
from genomealignmenttools_b200.twobit import PackedGenome
import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gat(?:host)?_[a-z0-9_]+)\s*\(", text)))


def test_libgat_exports_every_declared_symbol():
    lib = _native.load()
    names = declared("gat.h")
    assert set(names) == set(_native.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_libgathost_exports_every_declared_symbol():
    lib = hostlib.load()
    names = declared("gat_host.h")
    assert set(names) == set(hostlib.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_libgatkent_exports_kents_entry_points():
    """include/gat_kent.h: kent's names (chainConnect.h:34-44, gapCalc.h:11-35, axt.h:93-121) are exported by libgatkent.so."""
    lib = ctypes.CDLL(os.path.join(ROOT, "genomealignmenttools_b200", "libgatkent.so"))
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gat_kent.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined"})
    for n in ("chainScoreBlock", "chainCalcScore", "chainCalcScoreSubChain", "gapCalcCost", "gapCalcFromFile",
              "axtScoreSchemeRead", "axtScoreSchemeDefault"):
        assert n in names
    for n in names:
        assert hasattr(lib, n), n
    lib.gatKentLayout.restype = ctypes.c_long
    assert lib.gatKentLayout(3) == 8 + 256 * 256 * 4 + 4 + 4 + 8      # struct axtScoreScheme, kent/src/inc/axt.h:83-91


def test_record_layouts_match_header():
    assert BLOCK_DTYPE.itemsize == 12 and JOB_DTYPE.itemsize == 24 and NRUN_DTYPE.itemsize == 12
    assert ctypes.sizeof(_native.GatStats) == 48


def test_no_gpu_fails_loudly():
    lib = _native.load()
    if lib.gat_device_count() > 0:
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    assert lib.gat_create(ctypes.byref(ctx), 0, None) == -1
    assert b"no CPU path" in lib.gat_last_error()


@pytest.mark.parametrize("impl", ["python", "cxx"])
def test_gap_costs_match_reference_fixture(golden, impl):
    lib = hostlib.load()
    calc = {}
    for spec in ("medium", "loose"):
        if impl == "python":
            g = GapCalc.from_file(spec); calc[spec] = g.cost
        else:
            h = lib.gathost_gapcalc_open(spec.encode()); calc[spec] = lambda a, b, h=h: lib.gathost_gapcalc_cost(h, a, b)
    n = 0
    for line in open(os.path.join(golden, "gap_kat.tsv")):
        if line.startswith("#"):
            continue
        spec, dq, dt, cost = line.split()
        assert calc[spec](int(dq), int(dt)) == int(cost), line
        n += 1
    assert n > 5000


def test_gap_file_format_and_errors(tmp_path):
    lib = hostlib.load()
    p = tmp_path / "gap.txt"
    p.write_text("# comment\n\ntableSize 4\nsmallSize 11\nposition 1 2 11 1011\nqGap 10 20.5 30 1030\ntGap 11 21 31 1031\nbothGap 20 40 60 2060\n")
    g = GapCalc.from_file(str(p))
    h = lib.gathost_gapcalc_open(str(p).encode())
    for dq, dt in [(0, 0), (1, 0), (2, 0), (5, 0), (0, 7), (3, 3), (11, 0), (500, 0), (1011, 0), (5000, 0), (400, 700), (0, 100000)]:
        assert g.cost(dq, dt) == lib.gathost_gapcalc_cost(h, dq, dt)
    assert g.cost(2, 0) == 20 and g.cost(511, 0) == 530      # 20.5 truncates; halfway to the next knot
    bad = tmp_path / "bad.txt"
    bad.write_text("tableSize 2\nsmallSize 5\nposition 1 7\nqGap 1 2\ntGap 1 2\nbothGap 1 2\n")
    with pytest.raises(GapCalcError):
        GapCalc.from_file(str(bad))
    assert not lib.gathost_gapcalc_open(str(bad).encode())
    assert b"No position 5" in lib.gathost_last_error()


@pytest.mark.parametrize("rel", ["example/HoxD55.q", "kent_chrM/newStyleLastz.Q.txt", "kent_chrM/oldStyleBlastz.Q.txt", "synth_small/asym.q", None])
def test_score_scheme_readers_agree_with_oracle(oracle, golden, rel):
    path = os.path.join(golden, rel) if rel else None
    py = ScoreScheme.read(path).matrix if path else ScoreScheme.default().matrix
    m = ((ctypes.c_int32 * 4) * 4)()
    assert hostlib.load().gathost_scorescheme(path.encode() if path else None, m) == 0
    s = oracle.scoring(path, "loose")
    chars = "TCAG"      # kent codes 0..3
    for q in range(4):
        for t in range(4):
            want = oracle.lib.orc_matrix_at(s, ord(chars[q].lower()), ord(chars[t]))
            assert py[q, t] == want == m[q][t]
    assert oracle.lib.orc_matrix_at(s, ord("N"), ord("a")) == 0


def test_score_scheme_errors(tmp_path):
    lib = hostlib.load()
    m = ((ctypes.c_int32 * 4) * 4)()
    cases = {"short.q": "A C G T\n1 2 3 4\n", "notmatrix.q": "hello world foo bar\n", "noOE.q": "A C G T\n1 2 3 4\n1 2 3 4\n1 2 3 4\n1 2 3 4\nfoo\n"}
    for name, text in cases.items():
        p = tmp_path / name
        p.write_text(text)
        with pytest.raises(ScoreSchemeError):
            ScoreScheme.read(str(p))
        assert lib.gathost_scorescheme(str(p).encode(), m) == -1


@pytest.mark.parametrize("rel", ["synth_small/in.chain", "example/hg38.danRer10.chain", "kent_chrM/newStyleLastz.chain"])
def test_chain_readers_agree(oracle, golden, rel):
    path = os.path.join(golden, rel)
    lib = hostlib.load()
    cs = lib.gathost_chains_read(path.encode())
    assert cs, lib.gathost_last_error()
    heads = hostlib.chain_heads(lib, cs)
    py = chainio.ChainSet.read(path)
    oheads = oracle.chain_headers(oracle.chains(path))
    assert len(heads) == len(py) == len(oheads)
    for i, (h, o) in enumerate(zip(heads, oheads)):
        for k in ("tName", "tSize", "tStart", "tEnd", "qName", "qSize", "qStrand", "qStart", "qEnd", "id", "firstBlock", "nBlocks"):
            assert h[k] == o[k], (i, k)
        assert h["score"] == o["score"] == py.score[i]
    assert np.array_equal(hostlib.chain_blocks(lib, cs), py.blocks)


@pytest.mark.parametrize("piece", ["1", "300", "5000"])
def test_chain_reader_pieces_parsed_concurrently_equal_sequential(golden, tmp_path, monkeypatch, piece):
    """readChains cuts the file at chain headers and parses the pieces on host threads; chains, blocks, sequential
    ids of id-less chains, '#' lines and the first error must be those of a sequential read (GAT_PARSE_PIECE_BYTES
    forces cuts in small files)."""
    lib = hostlib.load()
    src = open(os.path.join(golden, "synth_small", "in.chain")).read().split("\n")
    lines, k = [], 0
    for ln in src:
        if ln.startswith("chain "):
            k += 1
            if k % 3 == 0:
                lines.append("# note before chain %d" % k)
            if k % 2 == 0:
                ln = " ".join(ln.split()[:12])          # drop the id column: chain.c:276-279 numbers those
        lines.append(ln)
    path = tmp_path / "mixed.chain"
    path.write_text("\n".join(lines))
    monkeypatch.delenv("GAT_PARSE_PIECE_BYTES", raising=False)
    a = lib.gathost_chains_read(str(path).encode())
    monkeypatch.setenv("GAT_PARSE_PIECE_BYTES", piece)
    b = lib.gathost_chains_read(str(path).encode())
    assert a and b
    assert hostlib.chain_heads(lib, a) == hostlib.chain_heads(lib, b)
    assert np.array_equal(hostlib.chain_blocks(lib, a), hostlib.chain_blocks(lib, b))
    # an interrupted chain is reported like a sequential read reports it, whichever piece meets it
    bad = lines[:]
    third = [i for i, ln in enumerate(bad) if ln.startswith("chain ")][2]
    last = max(i for i in range(third) if bad[i] and bad[i][0].isdigit())
    bad[last] = bad[last] + "\t3\t4"                     # the chain before now expects another block line
    (tmp_path / "bad.chain").write_text("\n".join(bad))
    msgs = []
    for env in (None, piece):
        if env is None:
            monkeypatch.delenv("GAT_PARSE_PIECE_BYTES", raising=False)
        else:
            monkeypatch.setenv("GAT_PARSE_PIECE_BYTES", env)
        assert not lib.gathost_chains_read(str(tmp_path / "bad.chain").encode())
        msgs.append(lib.gathost_last_error())
    assert msgs[0] == msgs[1] and b"line" in msgs[0]


def test_compact_worklist_roundtrip():
    """records.pack_compact (6-byte delta-coded blocks + absolute table + group anchors) expands back to the very records,
    including chain starts, gaps beyond 16 bits, negative gaps (out-of-order blocks) and JOINED pieces."""
    from genomealignmenttools_b200.records import pack_compact, unpack_compact, split_long_blocks, CBLOCK_ABS
    jobs, total, blocks = synth.make_chains([5_000_000, 800_000], [4_000_000, 900_000], 20000, seed=5, max_chain_blocks=3000,
                                            max_len=30000, gap_mu=6.0, gap_sigma=3.0, max_gap=900000)
    blocks["tStart"][100], blocks["tStart"][101] = blocks["tStart"][101], blocks["tStart"][100]      # a negative gap
    jobs, total, blocks = split_long_blocks(jobs, total, blocks, 4096)
    cj, cb, ab, an = pack_compact(jobs, total, blocks)
    j2, t2, b2 = unpack_compact(cj, cb, ab, an)
    assert np.array_equal(jobs, j2) and total == t2 and np.array_equal(blocks, b2)
    assert len(ab) > len(jobs)                                     # escapes beyond the chain starts
    assert (cj.nbytes + cb.nbytes + ab.nbytes + an.nbytes) < 0.6 * (jobs.nbytes + blocks.nbytes)
    with pytest.raises(ValueError):
        pack_compact(*synth.make_chains([5_000_000], [4_000_000], 2000, seed=6, max_len=30000, mean_log_len=9.0))   # unsplit long blocks


def test_packed_worklist_roundtrip():
    """records.pack_packed (one 32-bit word per block, absolute records take the next table entry in list order, absBase per
    group) expands back to the records cut at 4095 bases: chain starts, gaps beyond 9 bits, negative gaps, JOINED pieces,
    empty blocks, groups that open with an absolute record."""
    from genomealignmenttools_b200.records import pack_packed, unpack_packed, split_long_blocks, PBLOCK_ABS, CGROUP
    jobs, total, blocks = synth.make_chains([5_000_000, 800_000], [4_000_000, 900_000], 20000, seed=5, max_chain_blocks=3000,
                                            max_len=30000, gap_mu=6.0, gap_sigma=3.0, max_gap=900000)
    blocks["tStart"][100], blocks["tStart"][101] = blocks["tStart"][101], blocks["tStart"][100]      # a negative gap
    blocks["size"][200] = 0                                                                           # an empty block
    cj, pb, ab, an, base = pack_packed(jobs, total, blocks)
    j2, t2, b2 = unpack_packed(cj, pb, ab, an, base)
    js, ts, bs = split_long_blocks(jobs, total, blocks, 4095)
    assert np.array_equal(js, j2) and ts == t2 and np.array_equal(bs, b2)
    assert len(ab) > len(jobs) and int(((pb & PBLOCK_ABS) != 0).sum()) == len(ab)
    assert base[0] == 0 and np.all(np.diff(base.astype(np.int64)) >= 0) and len(base) == (t2 + CGROUP - 1) // CGROUP
    jobs, total, blocks = synth.make_chains([50_000_000], [40_000_000], 200000, seed=9)               # typical gaps: 4 bytes pay
    cj, pb, ab, an, base = pack_packed(jobs, total, blocks)
    assert (cj.nbytes + pb.nbytes + ab.nbytes + an.nbytes + base.nbytes) < 0.45 * (jobs.nbytes + blocks.nbytes)


def test_cpp_compact_packer_agrees_with_python(golden):
    """gathost::packCompact (what bin/scoreChain sends to gat_score_compact) == records.pack_compact on the same chains."""
    from genomealignmenttools_b200.records import (pack_compact, CJOB_DTYPE, CBLOCK_DTYPE, CABS_DTYPE, NO_CLIP_START, NO_CLIP_END,
                                                   QSEQ_MINUS)
    lib = hostlib.load()
    path = os.path.join(golden, "synth_small", "in.chain")
    cs = lib.gathost_chains_read(path.encode())
    assert cs
    heads = hostlib.chain_heads(lib, cs)
    blocks = hostlib.chain_blocks(lib, cs)                 # device records (long blocks already cut)
    n = len(heads)
    ct = (np.arange(n) % 3).astype(np.uint32); cq = (np.arange(n) % 5).astype(np.uint32)
    lib.gathost_chains_compact.restype = ctypes.c_void_p
    lib.gathost_chains_compact.argtypes = [ctypes.c_void_p] * 3
    h = lib.gathost_chains_compact(cs, ct.ctypes.data, cq.ctypes.data)
    assert h, lib.gathost_last_error()
    ptrs = [ctypes.c_void_p() for _ in range(4)]; counts = [ctypes.c_uint64() for _ in range(4)]
    lib.gathost_compact_view.argtypes = [ctypes.c_void_p] * 9
    args = []
    for p_, c_ in zip(ptrs, counts):
        args += [ctypes.byref(p_), ctypes.byref(c_)]
    assert lib.gathost_compact_view(h, *[ctypes.cast(a, ctypes.c_void_p) for a in args]) == 0
    got = []
    for p_, c_, dt in zip(ptrs, counts, (CJOB_DTYPE, CBLOCK_DTYPE, CABS_DTYPE, CABS_DTYPE)):
        k = int(c_.value)
        buf = (ctypes.c_uint8 * (dt.itemsize * k)).from_address(p_.value) if k else b""
        got.append(np.frombuffer(buf, dtype=dt, count=k).copy())
    jobs = np.zeros(n, dtype=JOB_DTYPE)
    jobs["tSeq"] = ct
    jobs["qSeq"] = cq | np.array([QSEQ_MINUS if hd["qStrand"] == "-" else 0 for hd in heads], dtype=np.uint32)
    jobs["firstBlock"] = [hd["firstBlock"] for hd in heads]; jobs["blockPtr"] = jobs["firstBlock"]
    jobs["clipStart"] = NO_CLIP_START; jobs["clipEnd"] = NO_CLIP_END
    want = pack_compact(jobs, len(blocks), blocks)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    lib.gathost_compact_free.argtypes = [ctypes.c_void_p]
    lib.gathost_compact_free(h)


def test_chain_reader_errors(tmp_path):
    lib = hostlib.load()
    good = "chain 100 chrA 1000 + 10 60 chrB 900 - 5 65 7\n20\t10\t20\n20\n\n"
    cases = {"ok": (good, None), "tmismatch": (good.replace(" 60 ", " 61 "), "t end mismatch"),
             "short": ("chain 100 chrA 1000 + 10 60 chrB 900 - 5\n", "at least 12 words"),
             "past": (good.replace("chrB 900", "chrB 60"), "Past end of sequence"),
             "notchain": (good.replace("chain", "chian"), "Expecting 'chain'")}
    for name, (text, err) in cases.items():
        p = tmp_path / (name + ".chain")
        p.write_text(text)
        cs = lib.gathost_chains_read(str(p).encode())
        if err is None:
            assert cs and hostlib.chain_heads(lib, cs)[0]["id"] == 7
            assert len(chainio.ChainSet.read(str(p))) == 1
        else:
            assert not cs and err.encode() in lib.gathost_last_error()
            with pytest.raises(chainio.ChainFormatError):
                chainio.ChainSet.read(str(p))
    # chains without an id are numbered 1, 2, ... (chain.c:276-279); '#' lines and gz input are accepted
    p = tmp_path / "noid.chain"
    p.write_text("#meta line\n" + good.replace(" 7\n", "\n") * 2)
    subprocess.check_call(["gzip", "-k", str(p)])
    for path in (str(p), str(p) + ".gz"):
        cs = lib.gathost_chains_read(path.encode())
        assert [h["id"] for h in hostlib.chain_heads(lib, cs)] == [1, 2]
        assert chainio.ChainSet.read(path).id == [1, 2]


def test_subchain_selection_agrees_with_oracle(oracle, golden):
    path = os.path.join(golden, "synth_small", "in.chain")
    lib = hostlib.load()
    cs = lib.gathost_chains_read(path.encode())
    ocs = oracle.chains(path)
    py = chainio.ChainSet.read(path)
    rows = np.loadtxt(os.path.join(golden, "synth_small", "sub_medium_default.tsv"), dtype=np.int64)
    for ix, s, e, is_null, g, l, a in rows:
        ok, fb, nb, c0, c1, ali = hostlib.subset(lib, cs, int(ix), int(s), int(e))
        ook, ofb, onb, oc0, oc1 = oracle.subset(ocs, int(ix), int(s), int(e))
        assert bool(ok) == bool(ook) == (not is_null)
        if ok:
            assert (fb, nb, c0, c1) == (ofb, onb, oc0, oc1) == py.subset_job(int(ix), int(s), int(e))
            assert ali == a


@pytest.mark.parametrize("name", ["t.2bit", "q.2bit", "t.swapped.2bit"])
def test_twobit_readers_agree(oracle, golden, name):
    path = os.path.join(golden, "synth_small", name)
    lib = hostlib.load()
    tb = lib.gathost_twobit_open(path.encode())
    assert tb, lib.gathost_last_error()
    py = PackedGenome.read_2bit(path)
    og = oracle.genome(path)
    assert lib.gathost_twobit_count(tb) == len(py.names) == oracle.lib.orc_genome_count(og)
    for i in range(len(py.names)):
        nm = ctypes.c_char_p(); size = ctypes.c_uint32(); packed = ctypes.c_void_p(); nr = ctypes.c_uint32()
        ns = ctypes.c_void_p(); nl = ctypes.c_void_p()
        assert lib.gathost_twobit_seq(ctypes.c_void_p(tb), i, ctypes.byref(nm), ctypes.byref(size), ctypes.byref(packed),
                                      ctypes.byref(nr), ctypes.byref(ns), ctypes.byref(nl)) == 0
        assert nm.value.decode() == py.names[i] and size.value == py.sizes[i] == oracle.lib.orc_genome_size(og, i)
        nbytes = (size.value + 3) // 4
        raw = ctypes.string_at(packed.value, nbytes)
        assert raw == py.packed[int(py.byte_offsets[i]):int(py.byte_offsets[i]) + nbytes].tobytes()
        runs = py.n_runs[py.n_runs["seq"] == i]
        assert nr.value == len(runs)
        if nr.value:
            assert np.array_equal(np.frombuffer(ctypes.string_at(ns.value, 4 * nr.value), dtype=np.uint32), runs["start"])
        # unpacked view: payload + N runs reproduce the oracle's characters
        dna = ctypes.string_at(oracle.lib.orc_genome_dna(og, i, b"+"), size.value).upper()
        codes = py.codes(i)
        want = np.frombuffer(b"TCAG", dtype=np.uint8)[codes].copy()
        for r in runs:
            want[int(r["start"]):int(r["start"]) + int(r["len"])] = ord("N")
        assert want.tobytes() == dna
    assert not lib.gathost_twobit_open(os.path.join(golden, "synth_small", "in.chain").encode())
    assert b"valid twoBitSig" in lib.gathost_last_error()
    with pytest.raises(ValueError):
        PackedGenome.read_2bit(os.path.join(golden, "synth_small", "in.chain"))


def test_twobit_roundtrip(tmp_path):
    rng = np.random.default_rng(4)
    g = PackedGenome.from_codes(["a", "bb", "c" * 40], [rng.integers(0, 4, n) for n in (1, 1023, 4096)],
                                n_runs=np.array([(1, 5, 10), (2, 0, 1)], dtype=NRUN_DTYPE),
                                mask_runs=np.array([(1, 100, 50)], dtype=NRUN_DTYPE))
    for version, swapped in ((0, False), (1, False), (0, True), (1, True)):
        p = tmp_path / ("g%d%d.2bit" % (version, swapped))
        g.write_2bit(str(p), version=version, swapped=swapped)
        h = PackedGenome.read_2bit(str(p))
        assert h.names == g.names and np.array_equal(h.sizes, g.sizes)
        assert np.array_equal(h.n_runs, g.n_runs) and np.array_equal(h.mask_runs, g.mask_runs)
        for i in range(3):
            assert np.array_equal(h.codes(i), g.codes(i))


def test_lpt_sharding_balances_aligned_bases():
    lib = hostlib.load()
    jobs, total, blocks = synth.make_chains([5_000_000, 2_000_000], [4_000_000, 3_000_000], 200_000, seed=8)
    ali = ali_bases(jobs, total, blocks)
    for parts in (2, 4, 8):
        part = np.zeros(len(jobs), dtype=np.uint32)
        assert lib.gathost_shard_jobs(jobs.ctypes.data, len(jobs), total, ali.ctypes.data, parts, part.ctypes.data) == 0
        load = np.bincount(part, weights=ali, minlength=parts)
        assert part.max() == parts - 1
        assert load.max() <= max(1.05 * load.mean(), load.mean() + ali.max())


def test_worklist_helpers():
    jobs, total = jobs_from_counts([0, 1, 0], [2, 0, 1], [0, 1, 0], [3, 0, 2])
    assert list(jobs["blockPtr"]) == [0, 3, 3] and total == 5
    assert jobs["qSeq"][1] == 0x80000000
    blocks = np.array([(0, 0, 10), (20, 25, 5), (30, 40, 7), (0, 0, 4), (9, 9, 1)], dtype=BLOCK_DTYPE)
    assert list(ali_bases(jobs, total, blocks)) == [22, 0, 5]
    jobs["clipStart"][0] = 5; jobs["clipEnd"][0] = 32
    assert ali_bases(jobs, total, blocks)[0] == 5 + 5 + 2


def test_cli_usage_and_errors():
    exe = os.path.join(ROOT, "bin", "scoreChain")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr.startswith("scoreChain - (re)score existing chains\nusage:")
    r = subprocess.run([exe, "a", "b", "c", "d"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr.startswith("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults")
    r = subprocess.run([exe, "a", "b", "c", "d", "-linearGap=loose", "-bogus"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr == "-bogus is not a valid option\n"
    r = subprocess.run([exe, "a", "b", "c", "d", "-linearGap=loose", "-returnOnlyScore", "-returnOnlyScoreAndCoords"],
                       capture_output=True, text=True)
    assert r.returncode == 255 and "cannot specify both" in r.stderr
    r = subprocess.run([exe, "a", "b.2bit", "c.2bit", "d", "-linearGap=loose"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr.startswith("ERROR: target 2bit file or nib directory b.2bit does not exist")


# ----------------------------------------------------------------------------- chains split over GPUs
def _local_global(a, g):
    """chainCalcScore / chainCalcScoreLocal on block scores a[i] and gap costs g[i] (gap in front of block i; g[0] unused):
    the loops of kent chainConnect.c:24-40 and src/scoreChain/scoreChain.c:181-195."""
    score = 0
    best = 0
    glob = 0
    for i in range(len(a)):
        if i:
            score -= g[i]
            glob -= g[i]
            if score < 0:
                score = 0
        score += a[i]
        glob += a[i]
        best = max(best, score)
    return glob, best


def _tuple_of(a, g):
    """(d, c, e, f) of a run of blocks scored as a job of its own (first block: no gap, no peak test)."""
    NEG = -(1 << 60)
    d, c, e, f = a[0], NEG, NEG, NEG
    for i in range(1, len(a)):
        dy = a[i] - g[i]
        d, c, e, f = d + dy, max(a[i], c + dy), max(e, d), max(f, c)
    return d, c, e, f


def test_tuple_join_reproduces_the_unsplit_chain():
    """gat_tuple_join / gat_tuple_scores (include/gat.h): pieces of a chain scored apart and joined with the gap cost at
    every cut give the scores of the whole chain, whatever the cuts."""
    import ctypes
    from genomealignmenttools_b200 import _native
    from genomealignmenttools_b200.sharding import TUPLE_DTYPE
    lib = _native.load()
    rng = np.random.default_rng(11)
    for trial in range(300):
        n = int(rng.integers(2, 60))
        a = [int(x) for x in rng.integers(-3000, 6000, n)]
        g = [0] + [int(x) for x in rng.integers(0, 5000, n - 1)]
        want = _local_global(a, g)
        k = int(rng.integers(1, min(6, n)))
        cuts = sorted(set(int(x) for x in rng.integers(1, n, k)))
        bounds = [0] + cuts + [n]
        acc = None
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            t = np.array([_tuple_of(a[lo:hi], [0] + g[lo + 1:hi])], dtype=TUPLE_DTYPE)
            if acc is None:
                acc = t.copy()
            else:
                lib.gat_tuple_join(acc.ctypes.data, g[lo], t.ctypes.data)
        gg, ll = ctypes.c_int64(), ctypes.c_int64()
        lib.gat_tuple_scores(acc.ctypes.data, ctypes.byref(gg), ctypes.byref(ll))
        assert (gg.value, ll.value) == want, (trial, a, g, cuts)


def test_split_giant_jobs_partitions_the_records():
    from genomealignmenttools_b200 import sharding, synth
    from genomealignmenttools_b200.records import job_block_counts, ali_bases, BLOCK_JOINED
    jobs, total, blocks = synth.make_chains([60_000_000, 20_000_000], [50_000_000, 18_000_000], 150_000, seed=9,
                                            zipf_s=1.3, max_chain_blocks=60_000)
    pj, origin, first_piece = sharding.split_giant_jobs(jobs, total, blocks, 4)
    counts = job_block_counts(pj, total)
    assert counts.sum() == total and len(pj) > len(jobs)
    # pieces tile the records of their job in order
    oc = job_block_counts(jobs, total)
    for j in np.nonzero(np.diff(first_piece) > 1)[0]:
        p = np.arange(first_piece[j], first_piece[j + 1])
        assert counts[p].sum() == oc[j] and counts[p].min() >= sharding.TUPLE_MIN_BLOCKS
        fb = pj["firstBlock"][p].astype(np.int64)
        assert fb[0] == jobs["firstBlock"][j] and np.array_equal(fb[1:], fb[:-1] + counts[p][:-1])
        assert not (blocks["size"][fb] & np.uint32(BLOCK_JOINED)).any()
    limit = ali_bases(jobs, total, blocks).sum() // 16
    assert ali_bases(pj, total, blocks).max() <= 2 * limit + 1


def test_gz_inputs_go_through_gzip_without_a_shell(golden, tmp_path):
    """linefile.c:40-53 reads .gz through a decompressor child: a quote in the file name is data, and a truncated file is an
    error (the child's exit status is checked), not a shorter chain set."""
    import shutil
    lib = hostlib.load()
    lib.gathost_chains_read.restype = ctypes.c_void_p
    lib.gathost_chains_read.argtypes = [ctypes.c_char_p]
    lib.gathost_chains_count.argtypes = [ctypes.c_void_p]
    lib.gathost_chains_count.restype = ctypes.c_uint64
    lib.gathost_last_error.restype = ctypes.c_char_p
    src = os.path.join(golden, "synth_small", "in.chain")
    odd = str(tmp_path / "it's a.chain")
    shutil.copy(src, odd)
    subprocess.check_call(["gzip", "-k", odd])
    c = lib.gathost_chains_read((odd + ".gz").encode())
    assert c and lib.gathost_chains_count(c) == 660
    cut = str(tmp_path / "cut.chain.gz")
    open(cut, "wb").write(open(odd + ".gz", "rb").read()[:3000])
    assert not lib.gathost_chains_read(cut.encode())
    assert b"gzip -dc" in lib.gathost_last_error() and b"failed" in lib.gathost_last_error()
