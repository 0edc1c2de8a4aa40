"""ctypes access to libgathost.so (include/gat_host.h): the C++ host code the CLI tools run."""
import ctypes
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "genomealignmenttools_b200", "libgathost.so")
SYMBOLS = ["gathost_last_error", "gathost_gapcalc_open", "gathost_gapcalc_close", "gathost_gapcalc_cost",
           "gathost_gapcalc_fill", "gathost_scorescheme", "gathost_chains_read", "gathost_chains_close",
           "gathost_chains_count", "gathost_chains_block_count", "gathost_chains_blocks", "gathost_chains_head",
           "gathost_chains_subset", "gathost_twobit_open", "gathost_twobit_close", "gathost_twobit_count",
           "gathost_twobit_seq", "gathost_shard_jobs", "gathost_chains_remove_partial_overlaps", "gathost_chains_compact",
           "gathost_compact_free", "gathost_compact_view"]
_vp, _u64 = ctypes.c_void_p, ctypes.c_uint64


def load():
    lib = ctypes.CDLL(PATH)
    lib.gathost_last_error.restype = ctypes.c_char_p
    lib.gathost_gapcalc_open.restype = _vp; lib.gathost_gapcalc_open.argtypes = [ctypes.c_char_p]
    lib.gathost_gapcalc_close.argtypes = [_vp]
    lib.gathost_gapcalc_cost.argtypes = [_vp, ctypes.c_int, ctypes.c_int]
    lib.gathost_scorescheme.argtypes = [ctypes.c_char_p, _vp]
    lib.gathost_chains_read.restype = _vp; lib.gathost_chains_read.argtypes = [ctypes.c_char_p]
    lib.gathost_chains_close.argtypes = [_vp]
    lib.gathost_chains_count.restype = _u64; lib.gathost_chains_count.argtypes = [_vp]
    lib.gathost_chains_block_count.restype = _u64; lib.gathost_chains_block_count.argtypes = [_vp]
    lib.gathost_chains_blocks.restype = _vp; lib.gathost_chains_blocks.argtypes = [_vp]
    lib.gathost_chains_subset.argtypes = [_vp, _u64, ctypes.c_int, ctypes.c_int] + [_vp] * 5
    lib.gathost_twobit_open.restype = _vp; lib.gathost_twobit_open.argtypes = [ctypes.c_char_p]
    lib.gathost_twobit_close.argtypes = [_vp]
    lib.gathost_twobit_count.restype = ctypes.c_uint32; lib.gathost_twobit_count.argtypes = [_vp]
    lib.gathost_shard_jobs.argtypes = [_vp, _u64, _u64, _vp, ctypes.c_int, _vp]
    return lib


def chain_heads(lib, cs):
    out = []
    for i in range(lib.gathost_chains_count(cs)):
        score = ctypes.c_double(); tn = ctypes.c_char_p(); qn = ctypes.c_char_p(); strand = ctypes.c_char()
        v = [ctypes.c_int() for _ in range(7)]
        fb = ctypes.c_uint64(); nb = ctypes.c_uint64()
        rc = lib.gathost_chains_head(_vp(cs), _u64(i), ctypes.byref(score), ctypes.byref(tn), ctypes.byref(v[0]),
                                     ctypes.byref(v[1]), ctypes.byref(v[2]), ctypes.byref(qn), ctypes.byref(v[3]),
                                     ctypes.byref(strand), ctypes.byref(v[4]), ctypes.byref(v[5]), ctypes.byref(v[6]),
                                     ctypes.byref(fb), ctypes.byref(nb))
        assert rc == 0
        out.append(dict(score=score.value, tName=tn.value.decode(), tSize=v[0].value, tStart=v[1].value, tEnd=v[2].value,
                        qName=qn.value.decode(), qSize=v[3].value, qStrand=strand.value.decode(), qStart=v[4].value,
                        qEnd=v[5].value, id=v[6].value, firstBlock=fb.value, nBlocks=nb.value))
    return out


def chain_blocks(lib, cs):
    from genomealignmenttools_b200.records import BLOCK_DTYPE
    n = lib.gathost_chains_block_count(cs)
    buf = (ctypes.c_uint8 * (12 * n)).from_address(lib.gathost_chains_blocks(cs))
    return np.frombuffer(buf, dtype=BLOCK_DTYPE, count=n).copy()


def subset(lib, cs, ix, s, e):
    fb = ctypes.c_uint64(); nb = ctypes.c_uint64(); c0 = ctypes.c_int32(); c1 = ctypes.c_int32(); ali = ctypes.c_int64()
    ok = lib.gathost_chains_subset(cs, ix, s, e, ctypes.byref(fb), ctypes.byref(nb), ctypes.byref(c0), ctypes.byref(c1), ctypes.byref(ali))
    return ok, fb.value, nb.value, c0.value, c1.value, ali.value
