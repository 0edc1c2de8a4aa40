"""ctypes access to the checkers: oracle/_build/liboracle.so (our CPU restatement) and, when it
has been built, oracle/_ref/libkentref.so (the unmodified reference).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libkentref.so")

_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_last_error.restype = ctypes.c_char_p
        lib.orc_genome_open.restype = _vp; lib.orc_genome_open.argtypes = [ctypes.c_char_p]
        lib.orc_genome_close.argtypes = [_vp]
        lib.orc_genome_count.argtypes = [_vp]
        lib.orc_genome_name.restype = ctypes.c_char_p; lib.orc_genome_name.argtypes = [_vp, _i32]
        lib.orc_genome_size.restype = _i64; lib.orc_genome_size.argtypes = [_vp, _i32]
        lib.orc_genome_find.argtypes = [_vp, ctypes.c_char_p]
        lib.orc_genome_dna.restype = _vp; lib.orc_genome_dna.argtypes = [_vp, _i32, ctypes.c_char]
        lib.orc_scoring_new.restype = _vp; lib.orc_scoring_new.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.orc_scoring_free.argtypes = [_vp]
        lib.orc_matrix_at.argtypes = [_vp, _i32, _i32]
        lib.orc_gap_cost.argtypes = [_vp, _i32, _i32]
        lib.orc_score_block.restype = ctypes.c_double
        lib.orc_score_block.argtypes = [_vp, ctypes.c_char_p, ctypes.c_char_p, _i32]
        lib.orc_score_jobs.argtypes = [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp]
        lib.orc_chains_read.restype = _vp; lib.orc_chains_read.argtypes = [ctypes.c_char_p]
        lib.orc_chains_free.argtypes = [_vp]
        lib.orc_chains_count.restype = _i64; lib.orc_chains_count.argtypes = [_vp]
        lib.orc_chains_total_blocks.restype = _i64; lib.orc_chains_total_blocks.argtypes = [_vp]
        lib.orc_chains_blocks.restype = _vp; lib.orc_chains_blocks.argtypes = [_vp]
        lib.orc_chains_subset.argtypes = [_vp, _i64, _i32, _i32] + [_vp] * 4

    def err(self):
        return self.lib.orc_last_error().decode()

    def genome(self, path):
        g = self.lib.orc_genome_open(str(path).encode())
        if not g:
            raise RuntimeError(self.err())
        return g

    def scoring(self, matrix_file, linear_gap):
        s = self.lib.orc_scoring_new(matrix_file.encode() if matrix_file else None, str(linear_gap).encode())
        if not s:
            raise RuntimeError(self.err())
        return s

    def score_jobs(self, scoring, tg, qg, jobs, total, blocks):
        jobs = np.ascontiguousarray(jobs); blocks = np.ascontiguousarray(blocks)
        n = len(jobs)
        g = np.zeros(n, dtype=np.int64); l = np.zeros(n, dtype=np.int64); a = np.zeros(n, dtype=np.int64)
        rc = self.lib.orc_score_jobs(scoring, tg, qg, jobs.ctypes.data, n, int(total), blocks.ctypes.data,
                                     len(blocks), g.ctypes.data, l.ctypes.data, a.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.err())
        return g, l, a

    def crossover(self, scoring, tg, qg, t_ix, q_ix, strand, pairs):
        """pairs: rows of (leftTEnd, leftQEnd, rightTStart, rightQStart, overlap) on one sequence pair."""
        self.lib.orc_genome_dna.restype = ctypes.c_void_p
        t = self.lib.orc_genome_dna(tg, int(t_ix), b"+")
        q = self.lib.orc_genome_dna(qg, int(q_ix), strand.encode())
        self.lib.orc_find_crossover.argtypes = [_vp, _vp, _vp] + [_i32] * 5 + [_vp, _vp]
        pos, adj = [], []
        for lt, lq, rt, rq, ov in pairs:
            a, b = ctypes.c_int(), ctypes.c_int()
            self.lib.orc_find_crossover(scoring, q, t, int(lt), int(lq), int(rt), int(rq), int(ov), ctypes.byref(a), ctypes.byref(b))
            pos.append(a.value); adj.append(b.value)
        return np.array(pos, dtype=np.int32), np.array(adj, dtype=np.int32)

    def chains(self, path):
        cs = self.lib.orc_chains_read(str(path).encode())
        if not cs:
            raise RuntimeError(self.err())
        return cs

    def chain_headers(self, cs):
        n = self.lib.orc_chains_count(cs)
        out = []
        f = self.lib.orc_chains_header
        for i in range(n):
            score = ctypes.c_double(); tn = ctypes.c_char_p(); qn = ctypes.c_char_p()
            v = [ctypes.c_int() for _ in range(8)]
            strand = ctypes.c_char(); fb = ctypes.c_int64(); nb = ctypes.c_int64()
            f(_vp(cs), _i64(i), ctypes.byref(score), ctypes.byref(tn), ctypes.byref(v[0]), ctypes.byref(v[1]),
              ctypes.byref(v[2]), ctypes.byref(qn), ctypes.byref(v[3]), ctypes.byref(strand), ctypes.byref(v[4]),
              ctypes.byref(v[5]), ctypes.byref(v[6]), ctypes.byref(fb), ctypes.byref(nb))
            out.append(dict(score=score.value, tName=tn.value.decode(), tSize=v[0].value, tStart=v[1].value,
                            tEnd=v[2].value, qName=qn.value.decode(), qSize=v[3].value,
                            qStrand=strand.value.decode(), qStart=v[4].value, qEnd=v[5].value, id=v[6].value,
                            firstBlock=fb.value, nBlocks=nb.value))
        return out

    def chain_blocks(self, cs):
        from genomealignmenttools_b200.records import BLOCK_DTYPE
        n = self.lib.orc_chains_total_blocks(cs)
        p = self.lib.orc_chains_blocks(cs)
        buf = (ctypes.c_uint8 * (12 * n)).from_address(p)
        return np.frombuffer(buf, dtype=BLOCK_DTYPE, count=n).copy()

    def subset(self, cs, ix, s, e):
        fb = ctypes.c_int64(); nb = ctypes.c_int64(); cs_ = ctypes.c_int32(); ce = ctypes.c_int32()
        ok = self.lib.orc_chains_subset(cs, ix, s, e, ctypes.byref(fb), ctypes.byref(nb), ctypes.byref(cs_), ctypes.byref(ce))
        return ok, fb.value, nb.value, cs_.value, ce.value


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        src = os.path.join(ORACLE_DIR, "chain_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        _oracle = Oracle(ctypes.CDLL(ORACLE_SO))
    return _oracle


class KentRef:
    """The reference itself, in process (oracle/ref_shim.c)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_set_scoring.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.ref_open_genomes.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.ref_load_chains.argtypes = [ctypes.c_char_p]
        lib.ref_gap_cost.argtypes = [_i32, _i32]
        lib.ref_matrix.argtypes = [_i32, _i32]
        lib.ref_score_all.argtypes = [_vp, _vp, _vp]
        lib.ref_score_sub.argtypes = [_i32] + [_vp] * 7

    def set_scoring(self, matrix_file, linear_gap):
        self.lib.ref_set_scoring(matrix_file.encode() if matrix_file else None, str(linear_gap).encode())

    def open(self, t2bit, q2bit, chain):
        self.lib.ref_open_genomes(str(t2bit).encode(), str(q2bit).encode())
        return self.lib.ref_load_chains(str(chain).encode())

    def score_all(self, n):
        g = np.zeros(n); l = np.zeros(n); a = np.zeros(n, dtype=np.int32)
        self.lib.ref_score_all(g.ctypes.data, l.ctypes.data, a.ctypes.data)
        return g.astype(np.int64), l.astype(np.int64), a.astype(np.int64)

    def crossover(self, t_name, q_name, strand, pairs):
        cols = [np.ascontiguousarray([p[k] for p in pairs], dtype=np.int32) for k in range(5)]
        n = len(pairs)
        pos = np.zeros(n, dtype=np.int32); adj = np.zeros(n, dtype=np.int32)
        self.lib.ref_crossover.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char, _i32] + [_vp] * 7
        self.lib.ref_crossover(t_name.encode(), q_name.encode(), strand.encode(), n, *[c.ctypes.data for c in cols],
                               pos.ctypes.data, adj.ctypes.data)
        return pos, adj

    def remove_partial_overlaps(self, ix):
        """chainRemovePartialOverlaps on loaded chain ix; returns (blocks as (n,3) int array, bounds)."""
        n = self.lib.ref_remove_partial_overlaps(int(ix))
        t = np.zeros(n, dtype=np.int32); q = np.zeros(n, dtype=np.int32); sz = np.zeros(n, dtype=np.int32); b = np.zeros(4, dtype=np.int32)
        self.lib.ref_chain_blocks.argtypes = [_i32] + [_vp] * 4
        self.lib.ref_chain_blocks(int(ix), t.ctypes.data, q.ctypes.data, sz.ctypes.data, b.ctypes.data)
        return np.stack([t, q, sz], axis=1), b

    def score_sub(self, ix, s, e):
        ix = np.ascontiguousarray(ix, dtype=np.int32); s = np.ascontiguousarray(s, dtype=np.int32)
        e = np.ascontiguousarray(e, dtype=np.int32)
        n = len(ix)
        g = np.zeros(n); l = np.zeros(n); a = np.zeros(n, dtype=np.int32); z = np.zeros(n, dtype=np.int32)
        self.lib.ref_score_sub(n, ix.ctypes.data, s.ctypes.data, e.ctypes.data, g.ctypes.data, l.ctypes.data,
                               a.ctypes.data, z.ctypes.data)
        return g.astype(np.int64), l.astype(np.int64), a.astype(np.int64), z


def load_ref():
    if not os.path.exists(REF_SO):
        return None
    return KentRef(ctypes.CDLL(REF_SO))
