"""ChainScorer: the host-side handle on one GPU's scoring context (gat_ctx).

Mirrors how the reference tools use the path: open the two genomes once, fix the scoring
scheme, then score chains -- except that chains arrive as one CSR work-list per call instead of
one chainCalcScore() call each (kent/src/lib/chainConnect.c:24)."""
import ctypes
import numpy as np
from . import _native
from .records import BLOCK_DTYPE, JOB_DTYPE, NRUN_DTYPE

TARGET, QUERY = 0, 1


class GatError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gat error %d: %s" % (code, message))
        self.code = code


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else ctypes.c_void_p(0)


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (gat_host_alloc)."""

    def __init__(self, shape, dtype):
        lib = _native.load()
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        self._p = lib.gat_host_alloc(max(n, 1))
        if not self._p:
            raise GatError(-5, lib.gat_last_error().decode())
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            _native.load().gat_host_free(self._p)
            self._p = None
            self.array = None


class ResidentWorklist:
    def __init__(self, scorer, handle, n_jobs):
        self.scorer, self.handle, self.n_jobs = scorer, handle, n_jobs

    def run(self):
        self.scorer._check(self.scorer.lib.gat_worklist_run(self.scorer.ctx, self.handle))

    def results(self, out_global=None, out_local=None):
        g = np.empty(self.n_jobs, dtype=np.int64) if out_global is None else out_global
        l = np.empty(self.n_jobs, dtype=np.int64) if out_local is None else out_local
        self.scorer._check(self.scorer.lib.gat_worklist_results(self.scorer.ctx, self.handle, _ptr(g), _ptr(l)))
        return g, l

    def free(self):
        if self.handle:
            self.scorer.lib.gat_worklist_destroy(self.scorer.ctx, self.handle)
            self.handle = None


class ChainScorer:
    def __init__(self, device=0, stream=None):
        self.lib = _native.load()
        ctx = ctypes.c_void_p()
        rc = self.lib.gat_create(ctypes.byref(ctx), int(device), ctypes.c_void_p(stream or 0))
        if rc != 0:
            raise GatError(rc, self.lib.gat_last_error().decode())
        self.ctx = ctx
        self._keep = []

    def _check(self, rc):
        if rc != 0:
            raise GatError(rc, self.lib.gat_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.gat_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def score_compact(self, cjobs, cblocks, ab, anchors, out_global=None, out_local=None):
        """gat_score_compact: the work-list as records.pack_compact() makes it (about half the bytes of score())."""
        from .records import CBLOCK_DTYPE, CJOB_DTYPE, CABS_DTYPE
        cjobs = np.ascontiguousarray(cjobs, dtype=CJOB_DTYPE); cblocks = np.ascontiguousarray(cblocks, dtype=CBLOCK_DTYPE)
        ab = np.ascontiguousarray(ab, dtype=CABS_DTYPE); anchors = np.ascontiguousarray(anchors, dtype=CABS_DTYPE)
        g = out_global if out_global is not None else np.zeros(len(cjobs), dtype=np.int64)
        l = out_local if out_local is not None else np.zeros(len(cjobs), dtype=np.int64)
        self._check(self.lib.gat_score_compact(self.ctx, _ptr(cjobs), len(cjobs), _ptr(cblocks), len(cblocks), _ptr(ab), len(ab),
                                               _ptr(anchors), _ptr(g), _ptr(l)))
        return g, l

    def score_packed(self, cjobs, pblocks, ab, anchors, abs_base, out_global=None, out_local=None):
        """gat_score_packed: the work-list as records.pack_packed() makes it (4 bytes per block)."""
        from .records import CJOB_DTYPE, CABS_DTYPE
        cjobs = np.ascontiguousarray(cjobs, dtype=CJOB_DTYPE); pblocks = np.ascontiguousarray(pblocks, dtype=np.uint32)
        ab = np.ascontiguousarray(ab, dtype=CABS_DTYPE); anchors = np.ascontiguousarray(anchors, dtype=CABS_DTYPE)
        abs_base = np.ascontiguousarray(abs_base, dtype=np.uint32)
        g = out_global if out_global is not None else np.zeros(len(cjobs), dtype=np.int64)
        l = out_local if out_local is not None else np.zeros(len(cjobs), dtype=np.int64)
        self._check(self.lib.gat_score_packed(self.ctx, _ptr(cjobs), len(cjobs), _ptr(pblocks), len(pblocks), _ptr(ab), len(ab),
                                              _ptr(anchors), _ptr(abs_base), _ptr(g), _ptr(l)))
        return g, l

    def crossover(self, pairs):
        """cBlockFindCrossover (kent chainConnect.c:61-105) for a batch of overlapping block pairs (XPAIR_DTYPE):
        returns (pos, adjust) int32 arrays."""
        from .records import XPAIR_DTYPE
        pairs = np.ascontiguousarray(pairs, dtype=XPAIR_DTYPE)
        pos = np.zeros(len(pairs), dtype=np.int32); adj = np.zeros(len(pairs), dtype=np.int32)
        self._check(self.lib.gat_crossover(self.ctx, _ptr(pairs), len(pairs), _ptr(pos), _ptr(adj)))
        return pos, adj

    def request_tuples(self, job_indices):
        """gat_request_tuples: ask the next scoring call for the (d, c, e, f) tuples of the listed jobs (parts of chains that
        were split over GPUs).  Returns the array the call will fill (sharding.TUPLE_DTYPE)."""
        from .sharding import TUPLE_DTYPE
        idx = np.ascontiguousarray(job_indices, dtype=np.uint32)
        out = np.zeros(len(idx), dtype=TUPLE_DTYPE)
        self._keep = [idx, out]          # must outlive the scoring call
        self._check(self.lib.gat_request_tuples(self.ctx, _ptr(idx), len(idx), _ptr(out)))
        return out

    def gap_cost(self, dq, dt):
        """gat_gap_cost: gapCalcCost for arrays of (dq, dt) on the device."""
        dq = np.ascontiguousarray(dq, dtype=np.int32); dt = np.ascontiguousarray(dt, dtype=np.int32)
        out = np.zeros(len(dq), dtype=np.int32)
        self._check(self.lib.gat_gap_cost(self.ctx, _ptr(dq), _ptr(dt), len(dq), _ptr(out)))
        return out

    def max_record_bases(self):
        """Longest record (gat_block.size) the device accepts under the current scoring parameters."""
        return int(self.lib.gat_max_record_bases(self.ctx))

    def load_genome(self, side, genome):
        """side: 0/'t' target, 1/'q' query; genome: PackedGenome (payload stays packed)."""
        side = {"t": TARGET, "q": QUERY}.get(side, side)
        runs = np.ascontiguousarray(genome.n_runs, dtype=NRUN_DTYPE)
        self._check(self.lib.gat_load_genome(
            self.ctx, side, _ptr(genome.packed), genome.packed.size, _ptr(genome.byte_offsets),
            _ptr(genome.sizes), len(genome.sizes), _ptr(runs), len(runs)))

    def set_scoring(self, scoring):
        s = _native.GatScoring()
        m = np.ascontiguousarray(scoring.scheme.matrix, dtype=np.int32)
        for q in range(4):
            for t in range(4):
                s.matrix[q][t] = int(m[q, t])
        g = scoring.gap
        keep = [np.ascontiguousarray(a) for a in (g.q_small, g.t_small, g.b_small, g.long_pos, g.q_long, g.t_long, g.b_long)]
        s.smallSize = g.small_size
        s.qSmall, s.tSmall, s.bSmall = keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data
        s.longCount = len(g.long_pos)
        s.longPos, s.qLong, s.tLong, s.bLong = (k.ctypes.data for k in keep[3:])
        self._check(self.lib.gat_set_scoring(self.ctx, ctypes.byref(s)))

    @staticmethod
    def _norm(jobs, blocks):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        return jobs, blocks

    def score(self, jobs, total_job_blocks, blocks, out_global=None, out_local=None):
        """gat_score: host arrays in, (global, local) int64 arrays out."""
        jobs, blocks = self._norm(jobs, blocks)
        g = np.empty(len(jobs), dtype=np.int64) if out_global is None else out_global
        l = np.empty(len(jobs), dtype=np.int64) if out_local is None else out_local
        self._check(self.lib.gat_score(self.ctx, _ptr(jobs), len(jobs), int(total_job_blocks),
                                       _ptr(blocks), len(blocks), _ptr(g), _ptr(l)))
        return g, l

    def upload(self, jobs, total_job_blocks, blocks):
        jobs, blocks = self._norm(jobs, blocks)
        h = ctypes.c_void_p()
        self._check(self.lib.gat_worklist_create(self.ctx, _ptr(jobs), len(jobs), int(total_job_blocks),
                                                 _ptr(blocks), len(blocks), ctypes.byref(h)))
        return ResidentWorklist(self, h, len(jobs))

    def synchronize(self):
        self._check(self.lib.gat_synchronize(self.ctx))

    def set_profiling(self, on):
        self._check(self.lib.gat_set_profiling(self.ctx, 1 if on else 0))

    def stats(self):
        st = _native.GatStats()
        self._check(self.lib.gat_get_stats(self.ctx, ctypes.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}
