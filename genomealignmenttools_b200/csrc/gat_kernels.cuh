// gat_kernels.cuh -- sm_100a kernels of the chain-rescoring path.
//
// Replaces, per job, the CPU loops of kent chainCalcScore / chainScoreBlock
// (kent/src/lib/chainConnect.c:14-40), gapCalcCost (kent/src/lib/gapCalc.c:298-331) and
// hillerlab chainCalcScoreLocal (src/scoreChain/scoreChain.c:176-198), including the clip of
// chainFastSubsetOnT (kent/src/lib/chain.c:510-522).  See DESIGN.md for the data layout.
//
// HBM layout of a genome ("bit-sliced 2-bit"): bases are grouped by 128; a group is one 32-byte
// DRAM sector = 4 words of high bits (bit1 of the kent base code T=0 C=1 A=2 G=3) followed by
// 4 words of low bits (bit0); base p of a 32-base word sits at bit p%32.  So 32 aligned bases of
// one plane are ONE 32-bit word, complement is "flip the high plane", reverse is __brev, and an
// unaligned 32-base window is a funnel shift of two neighbouring words per plane.  A third,
// separate plane holds N (1 bit/base) and is only read for blocks whose 1 kb windows contain N.
//
// Work decomposition: the job-blocks of all jobs form one virtual array; a CTA owns CHUNK
// consecutive job-blocks.  Inside a warp 32 blocks are expanded into 32-base "items" and the
// items -- not the blocks -- are dealt to lanes, so a 30 kb block and a 5 bp block cost what
// their bases cost.  Per-block sums come back through one warp scan; per-job global and local
// scores are a segmented, ordered reduction of a 4-number max-plus tuple (see Tup).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gat.h"

namespace gat {

constexpr int TPB = 256;               // threads per CTA
constexpr int WARPS = TPB / 32;
constexpr int BPT = 4;                 // job-blocks per thread = 32-block tiles per warp
constexpr int CHUNK = TPB * BPT;       // job-blocks per CTA
constexpr unsigned FULL = 0xffffffffu;
constexpr int GROUP_BASES = 128;       // bases per 32-byte sector group
constexpr int PAD_FRONT_GROUPS = 1;    // '-' strand windows may start up to 31 bases early
constexpr int PAD_BACK_GROUPS = 2;     // funnel shifts read one word past the last
constexpr int NWIN_SHIFT = 10;         // N summary: one bit per 1024 bases

constexpr int ERR_SEQ = 1, ERR_BLOCKIDX = 2, ERR_COORD = 4;

struct GenomeView {
    const uint32_t *planes;   // groups of 8 words: hi0..hi3 lo0..lo3
    const uint32_t *nplane;   // word n = N bits of bases [32n, 32n+32)
    const uint32_t *nwin;     // bit w = window w (1024 bases) contains an N
    const int64_t *seqBase;   // first base of each sequence in the padded coordinate (multiple of 128)
    const uint32_t *seqSize;
    uint32_t nSeq;
};

// (d, c, e, f): effect of a run of blocks on the local-score state.  Entering with running
// score s (and best-so-far M) the run leaves   s' = max(c, s + d)   and   M' = max(M, s + e, f).
// d alone is the global score contribution (sum of blocks - sum of gaps).  Runs compose
// associatively (not commutatively), which is what lets jobs span lanes, warps, CTAs.
struct Tup { long long d, c, e, f; };
constexpr long long NEG = -(1LL << 60);

__device__ __forceinline__ Tup tupIdentity() { return Tup{0, NEG, NEG, NEG}; }
__device__ __forceinline__ long long max64(long long a, long long b) { return a > b ? a : b; }
__device__ __forceinline__ Tup tupCombine(const Tup &x, const Tup &y)   // x first, then y
{
    Tup r;
    r.d = x.d + y.d;
    r.c = max64(y.c, x.c + y.d);
    r.e = max64(x.e, x.d + y.e);
    r.f = max64(max64(x.f, y.f), x.c + y.e);
    return r;
}
__device__ __forceinline__ long long shfl64(long long v, int src)
{
    int lo = __shfl_sync(FULL, (int)(unsigned long long)v, src);
    int hi = __shfl_sync(FULL, (int)((unsigned long long)v >> 32), src);
    return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
}
__device__ __forceinline__ Tup tupShfl(const Tup &t, int src)
{
    return Tup{shfl64(t.d, src), shfl64(t.c, src), shfl64(t.e, src), shfl64(t.f, src)};
}

struct GapView {           // tables of struct gapCalc (gapCalc.c:12-37) as the device sees them
    int smallSize, longCount, lastPos;
    double lastVal[3], lastSlope[3];   // q, t, both
};

struct ScoreParams {
    const gat_job *jobs;
    const gat_block *blocks;
    unsigned long long nJobs, totalJobBlocks, nBlocks;
    const uint32_t *chunkJob;   // job containing the first job-block of each chunk
    uint32_t nChunks;
    GenomeView t, q;
    int coef[16];               // SYM: 6 coefficients, general: 16 Moebius coefficients
    GapView gap;
    const int *gapSmall;        // [3][smallSize] in global; staged to shared
    const int *gapLongPos;      // [longCount]
    const double *gapLongVal;   // [3][longCount]
    long long *outGlobal, *outLocal;
    Tup *chunkHead, *chunkTail;
    int *chunkTailJob;
    int *err;
};

// ------------------------------------------------------------------ gap cost
__device__ __forceinline__ int truncToInt(double d)
{   // C's (int)double on x86-64 (cvttsd2si): toward zero, 0x80000000 when out of range
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(d);
}

// gapCalcCost, gapCalc.c:298-331 with interpolate() :82-104.  IEEE ops in the reference's order,
// spelled with the _rn intrinsics so nvcc can never contract them into an FMA.
__device__ __noinline__ int gapCostLong(const GapView &g, const int *longPos, const double *longVal,
                                        int which, int v)
{
    if (v >= g.lastPos)
        return truncToInt(__dadd_rn(g.lastVal[which], __dmul_rn(g.lastSlope[which], (double)(v - g.lastPos))));
    const double *val = longVal + which * g.longCount;
    for (int i = 0; i < g.longCount; i++) {
        int p = longPos[i];
        if (v == p) return truncToInt(val[i]);
        if (v < p) {
            int ds = p - longPos[i - 1];
            double dv = __dsub_rn(val[i], val[i - 1]);
            double prod = __dmul_rn(dv, (double)(v - longPos[i - 1]));
            return truncToInt(__dadd_rn(val[i - 1], __ddiv_rn(prod, (double)ds)));
        }
    }
    return INT32_MIN;   // unreachable: v < lastPos == longPos[longCount-1]
}

__device__ __forceinline__ int gapCost(const GapView &g, const int *small, const int *longPos,
                                       const double *longVal, int dq, int dt)
{
    if (dt < 0) dt = 0;
    if (dq < 0) dq = 0;
    int which, v;
    if (dt == 0) { which = 0; v = dq; }
    else if (dq == 0) { which = 1; v = dt; }
    else { which = 2; v = (int)((unsigned)dq + (unsigned)dt); }
    if (v < 0) return INT32_MIN;                      // dq+dt overflowed int (undefined in the reference)
    if (v < g.smallSize) return small[which * g.smallSize + v];
    return gapCostLong(g, longPos, longVal, which, v);
}

// ------------------------------------------------------------------ base windows
__device__ __forceinline__ uint32_t planeIdx(uint32_t n) { return ((n >> 2) << 3) | (n & 3u); }

// 32 bases starting `sh` bits into word n of both planes
__device__ __forceinline__ void loadWindow(const uint32_t *__restrict__ planes, uint32_t n, uint32_t sh,
                                           uint32_t &hi, uint32_t &lo)
{
    uint32_t i0 = planeIdx(n), i1 = planeIdx(n + 1);
    uint32_t h0 = __ldg(planes + i0), h1 = __ldg(planes + i1);
    uint32_t l0 = __ldg(planes + i0 + 4), l1 = __ldg(planes + i1 + 4);
    hi = __funnelshift_r(h0, h1, sh);
    lo = __funnelshift_r(l0, l1, sh);
}
__device__ __forceinline__ uint32_t loadNWindow(const uint32_t *__restrict__ np, uint32_t n, uint32_t sh)
{
    return __funnelshift_r(__ldg(np + n), __ldg(np + n + 1), sh);
}

// does [g0, g0+len) touch a 1 kb window that contains N?
__device__ __forceinline__ bool mayTouchN(const uint32_t *__restrict__ nwin, long long g0, int len)
{
    if (len <= 0) return false;
    unsigned long long w0 = (unsigned long long)g0 >> NWIN_SHIFT;
    unsigned long long w1 = (unsigned long long)(g0 + len - 1) >> NWIN_SHIFT;
    for (unsigned long long word = w0 >> 5; word <= (w1 >> 5); word++) {
        uint32_t bits = __ldg(nwin + word);
        unsigned lo = (word == (w0 >> 5)) ? (unsigned)(w0 & 31) : 0u;
        unsigned hi = (word == (w1 >> 5)) ? (unsigned)(w1 & 31) : 31u;
        uint32_t mask = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
        if (bits & mask) return true;
    }
    return false;
}

// ------------------------------------------------------------------ 32 base pairs -> score
// SYM: the matrix is strand-symmetric (M[a][b] == M[b][a] == M[comp a][comp b], true for the
// blastz default, HoxD55 and every lastz-inferred matrix): with X = q xor t the score depends on
// (X1, X0, q0 & ~X0) only and is  c0*n + c1|X1| + c2|X0| + c3|X1X0| + c4|Q'| + c5|X1Q'|.
// General: Moebius expansion over (q1,q0,t1,t0): sum of 16 coef * popc(product of planes).
template <bool SYM>
__device__ __forceinline__ int scoreWindow(const int *coef, uint32_t t1, uint32_t t0, uint32_t q1, uint32_t q0,
                                           uint32_t v, int nv)
{
    if (SYM) {
        uint32_t x1 = (q1 ^ t1) & v;
        uint32_t x0 = (q0 ^ t0) & v;
        uint32_t qp = q0 & ~x0 & v;
        return coef[0] * nv + coef[1] * __popc(x1) + coef[2] * __popc(x0) + coef[3] * __popc(x1 & x0) +
               coef[4] * __popc(qp) + coef[5] * __popc(x1 & qp);
    } else {
        t1 &= v; t0 &= v; q1 &= v; q0 &= v;
        uint32_t tt = t1 & t0, qq = q1 & q0;
        int s = coef[0] * nv;
        s += coef[1] * __popc(t0) + coef[2] * __popc(t1) + coef[3] * __popc(tt);
        s += coef[4] * __popc(q0) + coef[5] * __popc(q0 & t0) + coef[6] * __popc(q0 & t1) + coef[7] * __popc(q0 & tt);
        s += coef[8] * __popc(q1) + coef[9] * __popc(q1 & t0) + coef[10] * __popc(q1 & t1) + coef[11] * __popc(q1 & tt);
        s += coef[12] * __popc(qq) + coef[13] * __popc(qq & t0) + coef[14] * __popc(qq & t1) + coef[15] * __popc(qq & tt);
        return s;
    }
}

// ------------------------------------------------------------------ chunk index
// chunkJob[c] = the job that owns job-block c*CHUNK = (first j with blockPtr[j] > v) - 1.
// One warp per chunk, 32-ary search over the strided blockPtr column.
__global__ void chunkIndexKernel(const gat_job *__restrict__ jobs, unsigned long long nJobs,
                                 uint32_t *__restrict__ chunkJob, uint32_t nChunks)
{
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nChunks) return;
    unsigned long long target = (unsigned long long)warp * CHUNK;
    unsigned long long lo = 0, hi = nJobs;   // invariant: blockPtr[lo] <= target, answer in [lo, hi)
    while (hi - lo > 1) {
        unsigned long long span = hi - lo, step = (span + 31) / 32;
        unsigned long long idx = lo + lane * step;
        bool le = idx < hi && (unsigned long long)__ldg(&jobs[idx].blockPtr) <= target;
        unsigned m = __ballot_sync(FULL, le);
        int k = __popc(m);                  // lanes 0..k-1 are <= target (blockPtr is monotone)
        unsigned long long nlo = lo + (unsigned long long)(k - 1) * step;
        unsigned long long nhi = nlo + step;
        lo = nlo;
        hi = nhi < hi ? nhi : hi;
    }
    if (lane == 0) chunkJob[warp] = (uint32_t)lo;
}

// ------------------------------------------------------------------ the scoring kernel
struct __align__(16) StageRec { uint32_t tW, qW, n, misc; };   // misc: tSh | qSh<<5 | minus<<10 | mayN<<11

#ifndef GAT_MIN_CTAS
#define GAT_MIN_CTAS 3
#endif
template <bool SYM>
__global__ void __launch_bounds__(TPB, GAT_MIN_CTAS)
scoreChunksKernel(const __grid_constant__ ScoreParams P)
{
    __shared__ uint32_t sJob[CHUNK];            // job index + 1 of every job-block of the chunk
    __shared__ int sTs[CHUNK + 1], sQs[CHUNK + 1], sLen[CHUNK + 1];   // clipped block, +1 halo
    __shared__ long long sScore[CHUNK];
    __shared__ unsigned char sFlag[CHUNK + 1];  // 1 head of job, 2 end of job, 4 joined to previous, 8 valid
    __shared__ StageRec sStage[WARPS][32];
    __shared__ uint32_t sExcl[WARPS][32];
    __shared__ uint32_t sWarpMax[WARPS];
    __shared__ Tup sWarpAgg[WARPS];
    __shared__ int sWarpHead[WARPS];
    extern __shared__ unsigned char sDyn[];     // gap tables

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long vb0 = (unsigned long long)blockIdx.x * CHUNK;
    const unsigned long long total = P.totalJobBlocks;

    // ---- stage gap tables (gapCalc.c:12-37) in shared memory
    double *gLongVal = reinterpret_cast<double *>(sDyn);
    int *gLongPos = reinterpret_cast<int *>(gLongVal + 3 * P.gap.longCount);
    int *gSmall = gLongPos + P.gap.longCount;
    for (int i = tid; i < 3 * P.gap.longCount; i += TPB) gLongVal[i] = P.gapLongVal[i];
    for (int i = tid; i < P.gap.longCount; i += TPB) gLongPos[i] = P.gapLongPos[i];
    for (int i = tid; i < 3 * P.gap.smallSize; i += TPB) gSmall[i] = P.gapSmall[i];

    // ---- phase 0: which job owns each job-block of this chunk
    for (int i = tid; i < CHUNK; i += TPB) sJob[i] = 0;
    __syncthreads();
    {
        const uint32_t j0 = P.chunkJob[blockIdx.x];
        const uint32_t jEnd = (blockIdx.x + 1 < P.nChunks) ? P.chunkJob[blockIdx.x + 1] : (uint32_t)(P.nJobs - 1);
        const uint32_t jStart = blockIdx.x == 0 ? 0u : j0;   // chunk 0 also sweeps leading empty jobs
        for (uint32_t j = jStart + tid; j <= jEnd; j += TPB) {
            unsigned long long bp = P.jobs[j].blockPtr;
            unsigned long long np = (j + 1 < P.nJobs) ? (unsigned long long)P.jobs[j + 1].blockPtr : total;
            if (np > bp) {                                  // non-empty job
                if (bp >= vb0 && bp < vb0 + CHUNK) sJob[bp - vb0] = j + 1;
                else if (bp < vb0 && j == j0) sJob[0] = j + 1;
            } else {                                        // empty job (kent: NULL sub-chain): scores 0
                P.outGlobal[j] = 0;
                P.outLocal[j] = 0;
            }
        }
    }
    __syncthreads();
    {   // inclusive max-scan: job indices grow with position, so max = nearest head at or before
        uint32_t m0 = sJob[4 * tid], m1 = max(m0, sJob[4 * tid + 1]), m2 = max(m1, sJob[4 * tid + 2]),
                 m3 = max(m2, sJob[4 * tid + 3]);
        uint32_t run = m3;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, run, off);
            if (lane >= off) run = max(run, o);
        }
        if (lane == 31) sWarpMax[warp] = run;
        uint32_t before = __shfl_up_sync(FULL, run, 1);
        if (lane == 0) before = 0;
        __syncthreads();
        for (int w = 0; w < warp; w++) before = max(before, sWarpMax[w]);
        sJob[4 * tid] = max(m0, before);
        sJob[4 * tid + 1] = max(m1, before);
        sJob[4 * tid + 2] = max(m2, before);
        sJob[4 * tid + 3] = max(m3, before);
    }
    __syncthreads();

    // ---- phases 1+2: per 32-block tile, load + clip descriptors, expand to items, score
    for (int tile = 0; tile < BPT; tile++) {
        const int v = warp * (32 * BPT) + tile * 32 + lane;
        const unsigned long long gv = vb0 + v;
        const bool valid = gv < total;
        uint32_t tW = 0, qW = 0, n = 0, misc = 0;
        unsigned char flag = 0;
        int ts = 0, qs = 0, len = 0;
        if (valid) {
            const uint32_t j = sJob[v] - 1;
            const gat_job job = P.jobs[j];
            const unsigned long long np = (j + 1 < P.nJobs) ? (unsigned long long)P.jobs[j + 1].blockPtr : total;
            flag = 8;
            if (gv == job.blockPtr) flag |= 1;
            if (gv + 1 == np) flag |= 2;
            const unsigned long long bi = (unsigned long long)job.firstBlock + (gv - job.blockPtr);
            const uint32_t qSeq = job.qSeq & 0x7fffffffu;
            const bool minus = (job.qSeq >> 31) != 0;
            if (bi >= P.nBlocks) { atomicOr(P.err, ERR_BLOCKIDX); }
            else if (job.tSeq >= P.t.nSeq || qSeq >= P.q.nSeq) { atomicOr(P.err, ERR_SEQ); }
            else {
                const gat_block b = P.blocks[bi];
                if (b.size & GAT_BLOCK_JOINED) flag |= 4;
                const int size = (int)(b.size & 0x7fffffffu);
                // chainFastSubsetOnT clip, chain.c:513-522
                ts = b.tStart; qs = b.qStart;
                int te = ts + size;
                if (ts < job.clipStart) { qs += job.clipStart - ts; ts = job.clipStart; }
                if (te > job.clipEnd) te = job.clipEnd;
                len = te - ts;
                const int nn = len > 0 ? len : 0;
                const uint32_t tSize = P.t.seqSize[job.tSeq], qSize = P.q.seqSize[qSeq];
                if (nn > 0 && (ts < 0 || qs < 0 || (unsigned)ts + (unsigned)nn > tSize || (unsigned)qs + (unsigned)nn > qSize)) {
                    atomicOr(P.err, ERR_COORD);
                } else if (nn > 0) {
                    n = (uint32_t)nn;
                    const long long tG = P.t.seqBase[job.tSeq] + ts;
                    // '+': first base of the block.  '-': one past the block's last base in forward
                    // coordinates; rc position p is forward position qSize-1-p (dnautil.c:466-470).
                    const long long qBase = P.q.seqBase[qSeq];
                    const long long qG = minus ? qBase + ((long long)qSize - qs) : qBase + qs;
                    const long long qLo = minus ? qG - nn : qG;
                    const bool mayN = mayTouchN(P.t.nwin, tG, nn) || mayTouchN(P.q.nwin, qLo, nn);
                    tW = (uint32_t)(tG >> 5);
                    qW = minus ? (uint32_t)((qG - 32) >> 5) : (uint32_t)(qG >> 5);
                    misc = (uint32_t)(tG & 31) | ((uint32_t)(qG & 31) << 5) | (minus ? 1u << 10 : 0u) | (mayN ? 1u << 11 : 0u);
                }
            }
        }
        sTs[v] = ts; sQs[v] = qs; sLen[v] = len; sFlag[v] = flag;
        if (v == CHUNK - 1 && valid && !(flag & 2)) {
            // halo: the next job-block of the same job, needed for the gap after the chunk's last block
            const uint32_t j = sJob[v] - 1;
            const gat_job job = P.jobs[j];
            const unsigned long long bi = (unsigned long long)job.firstBlock + (gv + 1 - job.blockPtr);
            int hts = 0, hqs = 0, hlen = 0; unsigned char hflag = 8;
            if (bi < P.nBlocks) {
                const gat_block b = P.blocks[bi];
                if (b.size & GAT_BLOCK_JOINED) hflag |= 4;
                const int size = (int)(b.size & 0x7fffffffu);
                hts = b.tStart; hqs = b.qStart;
                int te = hts + size;
                if (hts < job.clipStart) { hqs += job.clipStart - hts; hts = job.clipStart; }
                if (te > job.clipEnd) te = job.clipEnd;
                hlen = te - hts;
            }
            sTs[CHUNK] = hts; sQs[CHUNK] = hqs; sLen[CHUNK] = hlen; sFlag[CHUNK] = hflag;
        }

        // expand: every block contributes max(1, ceil(n/32)) items; items are dealt to lanes
        const uint32_t items = n ? (n + 31) >> 5 : 1u;
        uint32_t incl = items;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += o;
        }
        const uint32_t excl = incl - items;
        const uint32_t totalItems = __shfl_sync(FULL, incl, 31);
        sStage[warp][lane] = StageRec{tW, qW, n, misc};
        sExcl[warp][lane] = excl;
        const bool anyN = __any_sync(FULL, (misc >> 11) & 1u);
        __syncwarp();

        long long acc = 0;
        for (uint32_t base = 0; base < totalItems; base += 32) {
            // owner of item base+lane: blocks started before `base` + heads at or before this lane
            const uint32_t rel = excl - base;
            const unsigned heads = __reduce_or_sync(FULL, rel < 32u ? 1u << rel : 0u);
            const int before = __popc(__ballot_sync(FULL, excl < base));
            const int owner = before - 1 + __popc(heads & (0xffffffffu >> (31 - lane)));
            const uint32_t x = base + lane;
            int s = 0;
            if (x < totalItems) {
                const StageRec r = sStage[warp][owner];
                const uint32_t k = x - sExcl[warp][owner];
                const int left = (int)r.n - (int)(k << 5);
                const int nv = left >= 32 ? 32 : (left > 0 ? left : 0);
                if (nv > 0) {
                    uint32_t vmask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
                    const uint32_t tSh = r.misc & 31u, qSh = (r.misc >> 5) & 31u;
                    const bool minus = (r.misc >> 10) & 1u;
                    uint32_t t1, t0, q1, q0;
                    loadWindow(P.t.planes, r.tW + k, tSh, t1, t0);
                    const uint32_t qn = minus ? r.qW - k : r.qW + k;
                    loadWindow(P.q.planes, qn, qSh, q1, q0);
                    if (minus) { q1 = ~__brev(q1); q0 = __brev(q0); }   // reverse, complement = flip bit1
                    if (anyN && ((r.misc >> 11) & 1u)) {
                        uint32_t nt = loadNWindow(P.t.nplane, r.tW + k, tSh);
                        uint32_t nq = loadNWindow(P.q.nplane, qn, qSh);
                        if (minus) nq = __brev(nq);
                        vmask &= ~(nt | nq);            // N scores 0 against everything (axt.c:431-454)
                    }
                    s = scoreWindow<SYM>(P.coef, t1, t0, q1, q0, vmask, __popc(vmask));
                }
            }
            // per-block sums: inclusive scan over lanes, then each block's owner lane differences it
            int S = s;
            for (int off = 1; off < 32; off <<= 1) {
                int o = __shfl_up_sync(FULL, S, off);
                if (lane >= off) S += o;
            }
            const uint32_t a = excl > base ? excl - base : 0u;            // my first item in this round
            const uint32_t bEnd = (incl - base) < 32u ? incl - base : 32u;  // one past my last (if any)
            const bool has = excl < base + 32u && incl > base;
            const int Sb = __shfl_sync(FULL, S, (int)((bEnd - 1u) & 31u));
            const int Sa = __shfl_sync(FULL, S, (int)((a - 1u) & 31u));
            if (has) acc += (long long)(Sb - (a ? Sa : 0));
        }
        sScore[v] = acc;
        __syncwarp();
    }
    __syncthreads();    // phase 3 reads the first block of the next warp's range

    // ---- phase 3: ordered segmented reduction of tuples over the chunk
    Tup cur = tupIdentity();      // open segment at the end of my run
    bool runHasHead = false;
    bool pend = false;            // an END reached before any HEAD of my run: needs the carry
    Tup pendTup = tupIdentity();
    uint32_t pendJob = 0;
    int lastValidK = -1;
    bool lastIsEnd = false;
    uint32_t lastJob = 0;
#pragma unroll
    for (int k = 0; k < BPT; k++) {
        const int v = 4 * tid + k;
        const unsigned char fl = sFlag[v];
        if (!(fl & 8)) break;
        const long long a = sScore[v];
        const bool isEnd = fl & 2;
        Tup e;
        if (isEnd) e = Tup{a, NEG, a, NEG};
        else if (sFlag[v + 1] & 4) e = Tup{a, NEG, NEG, NEG};       // next record continues this block
        else {
            const int qe = sQs[v] + sLen[v], te = sTs[v] + sLen[v];
            const int g = gapCost(P.gap, gSmall, gLongPos, gLongVal, sQs[v + 1] - qe, sTs[v + 1] - te);
            e = Tup{a - g, 0, a, NEG};
        }
        if (fl & 1) { cur = e; runHasHead = true; }
        else cur = tupCombine(cur, e);
        lastValidK = k; lastIsEnd = isEnd; lastJob = sJob[v] - 1;
        if (isEnd) {
            if (runHasHead) {   // job lies inside my run: done
                P.outGlobal[lastJob] = cur.d;
                P.outLocal[lastJob] = max64(0, max64(cur.e, cur.f));
            } else { pend = true; pendTup = cur; pendJob = lastJob; }
        }
    }
    // warp-level segmented inclusive scan of (cur, runHasHead)
    Tup inc = cur;
    bool incHead = runHasHead;
    for (int off = 1; off < 32; off <<= 1) {
        Tup o = tupShfl(inc, lane >= off ? lane - off : lane);
        bool oh = __shfl_sync(FULL, (int)incHead, lane >= off ? lane - off : lane);
        if (lane >= off && !incHead) { inc = tupCombine(o, inc); incHead = oh; }
    }
    if (lane == 31) { sWarpAgg[warp] = inc; sWarpHead[warp] = incHead; }
    Tup carry = tupShfl(inc, lane ? lane - 1 : 0);
    bool carryHead = __shfl_sync(FULL, (int)incHead, lane ? lane - 1 : 0);
    __syncthreads();
    {
        Tup wc = tupIdentity();
        bool wh = false;
        for (int w = 0; w < warp; w++) {
            if (sWarpHead[w]) { wc = sWarpAgg[w]; wh = true; }
            else wc = tupCombine(wc, sWarpAgg[w]);
        }
        if (lane == 0) { carry = wc; carryHead = wh; }
        else if (!carryHead) { carry = tupCombine(wc, carry); carryHead = wh; }
    }
    if (pend) {
        const Tup fin = tupCombine(carry, pendTup);
        if (carryHead) {
            P.outGlobal[pendJob] = fin.d;
            P.outLocal[pendJob] = max64(0, max64(fin.e, fin.f));
        } else P.chunkHead[blockIdx.x] = fin;       // job began in an earlier chunk and ends here
    }
    // the chunk's last valid job-block: does its job run on into the next chunk?
    {
        const unsigned long long lastV = (total - vb0 < (unsigned long long)CHUNK ? total - vb0 : (unsigned long long)CHUNK) - 1;
        if (lastValidK >= 0 && (unsigned long long)(4 * tid + lastValidK) == lastV) {
            if (lastIsEnd) P.chunkTailJob[blockIdx.x] = -1;
            else {
                const Tup open = runHasHead ? cur : tupCombine(carry, cur);
                if (runHasHead || carryHead) { P.chunkTail[blockIdx.x] = open; P.chunkTailJob[blockIdx.x] = (int)lastJob; }
                else { P.chunkHead[blockIdx.x] = open; P.chunkTailJob[blockIdx.x] = -1; }
            }
        }
    }
}

// ------------------------------------------------------------------ cross-chunk fix-up
// A job that starts in chunk c and ends in chunk c' > c:  tail(c) + head(c+1) + ... + head(c').
// One warp per chunk that has such a tail; lanes fold contiguous slices, then an ordered fold.
__global__ void fixupKernel(const gat_job *__restrict__ jobs, unsigned long long nJobs, unsigned long long total,
                            const Tup *__restrict__ chunkHead, const Tup *__restrict__ chunkTail,
                            const int *__restrict__ chunkTailJob, uint32_t nChunks,
                            long long *__restrict__ outGlobal, long long *__restrict__ outLocal)
{
    uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= nChunks) return;
    const int j = chunkTailJob[c];
    if (j < 0) return;
    const unsigned long long np = ((unsigned long long)j + 1 < nJobs) ? (unsigned long long)jobs[j + 1].blockPtr : total;
    const uint32_t cLast = (uint32_t)((np - 1) / CHUNK);
    const uint32_t count = cLast - c;                  // heads to fold: chunks c+1 .. cLast
    const uint32_t per = (count + 31) / 32;
    Tup mine = tupIdentity();
    for (uint32_t i = 0; i < per; i++) {
        uint32_t idx = lane * per + i;
        if (idx < count) mine = tupCombine(mine, chunkHead[c + 1 + idx]);
    }
    Tup all = chunkTail[c];
    for (int l = 0; l < 32; l++) {
        Tup o = tupShfl(mine, l);
        all = tupCombine(all, o);
    }
    if (lane == 0) {
        outGlobal[j] = all.d;
        outLocal[j] = max64(0, max64(all.e, all.f));
    }
}

// ------------------------------------------------------------------ genome ingest
// .2bit payload (4 bases/byte, first base in bits 7..6, twoBit.c:811-818) -> bit-sliced groups.
// One thread per 32-base word of one sequence.
__global__ void repackKernel(const uint8_t *__restrict__ raw, const unsigned long long *__restrict__ seqByteOffset,
                             const uint32_t *__restrict__ seqSize, const long long *__restrict__ seqBase,
                             const unsigned long long *__restrict__ seqWordStart, uint32_t nSeq,
                             unsigned long long totalWords, uint32_t *__restrict__ planes)
{
    unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= totalWords) return;
    uint32_t lo = 0, hi = nSeq;          // last sequence with seqWordStart <= w
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (seqWordStart[mid] <= w) lo = mid; else hi = mid;
    }
    const uint32_t s = lo;
    const unsigned long long wl = w - seqWordStart[s];
    const uint32_t size = seqSize[s];
    const unsigned long long firstBase = wl * 32;
    const uint8_t *src = raw + seqByteOffset[s] + wl * 8;
    uint32_t hiW = 0, loW = 0;
    for (int b = 0; b < 8; b++) {
        unsigned long long base = firstBase + 4ull * b;
        if (base >= size) break;
        uint32_t byte = src[b];
        for (int k = 0; k < 4; k++) {
            if (base + k >= size) break;
            uint32_t code = (byte >> (6 - 2 * k)) & 3u;
            hiW |= (code >> 1) << (4 * b + k);
            loW |= (code & 1u) << (4 * b + k);
        }
    }
    const unsigned long long n = (unsigned long long)(seqBase[s] >> 5) + wl;
    const uint32_t idx = planeIdx((uint32_t)n);
    planes[idx] = hiW;
    planes[idx + 4] = loW;
}

// N runs (twoBit.c:835-851) -> N plane + 1 kb window summary.  One warp per run.
__global__ void nRunKernel(const gat_nrun *__restrict__ runs, unsigned long long nRuns,
                           const long long *__restrict__ seqBase, uint32_t *__restrict__ nplane,
                           uint32_t *__restrict__ nwin)
{
    unsigned long long r = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (r >= nRuns) return;
    const gat_nrun run = runs[r];
    if (run.len == 0) return;
    const unsigned long long g0 = (unsigned long long)seqBase[run.seq] + run.start, g1 = g0 + run.len;  // [g0, g1)
    for (unsigned long long w = (g0 >> 5) + lane; w <= ((g1 - 1) >> 5); w += 32) {
        unsigned lo = (w == (g0 >> 5)) ? (unsigned)(g0 & 31) : 0u;
        unsigned hi = (w == ((g1 - 1) >> 5)) ? (unsigned)((g1 - 1) & 31) : 31u;
        atomicOr(&nplane[w], (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
    }
    for (unsigned long long win = (g0 >> NWIN_SHIFT) + lane; win <= ((g1 - 1) >> NWIN_SHIFT); win += 32)
        atomicOr(&nwin[win >> 5], 1u << (win & 31));
}

}  // namespace gat
