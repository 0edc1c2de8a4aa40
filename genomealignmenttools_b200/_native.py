"""ctypes binding of the C ABI in include/gat.h.  Fails loudly when the CUDA library is missing:
there is no fallback implementation."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GAT_LIB_PATH") or os.path.join(_HERE, "libgat.so")  # env override: kernel-variant experiments
SYNTH_LIB_PATH = os.path.join(_HERE, "libgatsynth.so")

# every symbol include/gat.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "gat_last_error", "gat_device_count", "gat_create", "gat_destroy", "gat_load_genome",
    "gat_set_scoring", "gat_score", "gat_worklist_create", "gat_worklist_run", "gat_worklist_results",
    "gat_worklist_destroy", "gat_synchronize", "gat_get_stats", "gat_set_profiling",
    "gat_host_alloc", "gat_host_free", "gat_max_record_bases", "gat_crossover", "gat_score_compact", "gat_score_packed",
    "gat_request_tuples", "gat_tuple_join", "gat_tuple_scores", "gat_gap_cost",
]


class GatScoring(ctypes.Structure):
    _fields_ = [("matrix", (ctypes.c_int32 * 4) * 4), ("smallSize", ctypes.c_int32),
                ("qSmall", ctypes.c_void_p), ("tSmall", ctypes.c_void_p), ("bSmall", ctypes.c_void_p),
                ("longCount", ctypes.c_int32), ("longPos", ctypes.c_void_p),
                ("qLong", ctypes.c_void_p), ("tLong", ctypes.c_void_p), ("bLong", ctypes.c_void_p)]


class GatStats(ctypes.Structure):
    _fields_ = [("score_kernel_ms", ctypes.c_float), ("all_kernels_ms", ctypes.c_float),
                ("h2d_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("chunks", ctypes.c_uint32),
                ("long_streamed", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for chain scoring." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
    lib.gat_last_error.restype = ctypes.c_char_p
    lib.gat_device_count.restype = i32
    lib.gat_create.argtypes = [ctypes.POINTER(vp), i32, vp]
    lib.gat_destroy.argtypes = [vp]
    lib.gat_destroy.restype = None
    lib.gat_load_genome.argtypes = [vp, i32, vp, u64, vp, vp, ctypes.c_uint32, vp, u64]
    lib.gat_set_scoring.argtypes = [vp, ctypes.POINTER(GatScoring)]
    lib.gat_score.argtypes = [vp, vp, u64, u64, vp, u64, vp, vp]
    lib.gat_worklist_create.argtypes = [vp, vp, u64, u64, vp, u64, ctypes.POINTER(vp)]
    lib.gat_worklist_run.argtypes = [vp, vp]
    lib.gat_worklist_results.argtypes = [vp, vp, vp, vp]
    lib.gat_worklist_destroy.argtypes = [vp, vp]
    lib.gat_worklist_destroy.restype = None
    lib.gat_crossover.argtypes = [vp, vp, u64, vp, vp]
    lib.gat_score_compact.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, vp, vp]
    lib.gat_score_packed.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, vp, vp, vp]
    lib.gat_gap_cost.argtypes = [vp, vp, vp, u64, vp]
    lib.gat_request_tuples.argtypes = [vp, vp, u64, vp]
    lib.gat_tuple_join.argtypes = [vp, ctypes.c_int64, vp]
    lib.gat_tuple_join.restype = None
    lib.gat_tuple_scores.argtypes = [vp, vp, vp]
    lib.gat_tuple_scores.restype = None
    lib.gat_max_record_bases.argtypes = [vp]
    lib.gat_max_record_bases.restype = ctypes.c_uint32
    lib.gat_synchronize.argtypes = [vp]
    lib.gat_get_stats.argtypes = [vp, ctypes.POINTER(GatStats)]
    lib.gat_set_profiling.argtypes = [vp, i32]
    lib.gat_host_alloc.argtypes = [ctypes.c_size_t]
    lib.gat_host_alloc.restype = vp
    lib.gat_host_free.argtypes = [vp]
    lib.gat_host_free.restype = None
    _lib = lib
    return lib


_synth = None


def load_synth():
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError("%s is missing: run __graft_entry__.build()" % SYNTH_LIB_PATH)
        s = ctypes.CDLL(SYNTH_LIB_PATH)
        vp, u64 = ctypes.c_void_p, ctypes.c_uint64
        s.gat_synth_fill.argtypes = [vp, u64, u64]
        s.gat_synth_fill.restype = None
        s.gat_synth_plant.argtypes = [vp, vp, vp, vp, vp, vp, u64, u64, vp, ctypes.c_uint32, u64]
        s.gat_synth_plant.restype = None
        _synth = s
    return _synth
