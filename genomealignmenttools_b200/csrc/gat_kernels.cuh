// gat_kernels.cuh -- sm_100a kernels of the chain-rescoring path.
//
// Replaces, per job, the CPU loops of kent chainCalcScore / chainScoreBlock
// (kent/src/lib/chainConnect.c:14-40), gapCalcCost (kent/src/lib/gapCalc.c:298-331) and
// hillerlab chainCalcScoreLocal (src/scoreChain/scoreChain.c:176-198), including the clip of
// chainFastSubsetOnT (kent/src/lib/chain.c:510-522).  See DESIGN.md for the data layout.
//
// HBM layout of a genome ("bit-sliced 2-bit"): 32 consecutive bases are one uint2 = {word of high
// bits, word of low bits} of the kent base code T=0 C=1 A=2 G=3; base p sits at bit p%32.  So 32
// aligned bases are ONE 8-byte load, 128 bases are one 32-byte DRAM sector, complement is "flip
// the high word", reverse is __brev, and an unaligned 32-base window is a funnel shift over two
// neighbouring uint2.  A third, separate plane holds N (1 bit/base) and is only read for blocks
// whose 256-base windows contain N.
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
constexpr int RUN = 16;                // job-blocks per thread in the job reduction (phase 3)
constexpr int P3_THREADS = CHUNK / RUN;
constexpr int GROUP_BASES = 128;       // sequences start on a 32-byte sector boundary
constexpr int PAD_FRONT_GROUPS = 1;    // '-' strand windows may start up to 31 bases early
constexpr int PAD_BACK_GROUPS = 2;     // funnel shifts read one word past the last
constexpr int LONG_ROUND_BASES = 1024;  // one warp-wide round of the long-block path (32 lanes x 32 bases)
constexpr int NWIN_SHIFT = 8;          // N summary: one bit per 256 bases

constexpr int ERR_SEQ = 1, ERR_BLOCKIDX = 2, ERR_COORD = 4;

// shared-memory index of job-block v: one pad word per RUN so that phase 3 (thread t walks
// v = RUN*t .. RUN*t+RUN-1) is bank-conflict free
__device__ __forceinline__ int padIdx(int v) { return v + (v >> 4); }
constexpr int PADDED = CHUNK + CHUNK / RUN + 2;
static_assert(RUN == 16, "padIdx assumes RUN == 16");

struct GenomeView {
    const uint2 *planes;      // word n = {high bits, low bits} of bases [32n, 32n+32)
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

// The same tuple in 32 bits, used by a warp whose 128 blocks are provably small enough
// (sum of |block score| + gap cost below 2^28): a third of the instructions of the 64-bit form.
template <typename T> struct TupT { T d, c, e, f; };
template <typename T> __device__ __forceinline__ T negInf();
template <> __device__ __forceinline__ long long negInf<long long>() { return NEG; }
template <> __device__ __forceinline__ int negInf<int>() { return -(1 << 30); }
template <typename T> __device__ __forceinline__ T maxT(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ TupT<T> tIdentity() { return TupT<T>{0, negInf<T>(), negInf<T>(), negInf<T>()}; }
template <typename T> __device__ __forceinline__ TupT<T> tCombine(const TupT<T> &x, const TupT<T> &y)
{
    TupT<T> r;
    r.d = x.d + y.d;
    r.c = maxT<T>(y.c, x.c + y.d);
    r.e = maxT<T>(x.e, x.d + y.e);
    r.f = maxT<T>(maxT<T>(x.f, y.f), x.c + y.e);
    return r;
}
__device__ __forceinline__ int shflT(int v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ long long shflT(long long v, int src) { return shfl64(v, src); }
template <typename T> __device__ __forceinline__ TupT<T> tShfl(const TupT<T> &t, int src)
{
    return TupT<T>{shflT(t.d, src), shflT(t.c, src), shflT(t.e, src), shflT(t.f, src)};
}
// to the 64-bit tuple that crosses warps / CTAs; anything at or below -2^29 is "minus infinity"
__device__ __forceinline__ long long widen(long long v) { return v; }
__device__ __forceinline__ long long widen(int v) { return v < -(1 << 29) ? NEG : (long long)v; }
template <typename T> __device__ __forceinline__ Tup tWiden(const TupT<T> &t) { return Tup{(long long)t.d, widen(t.c), widen(t.e), widen(t.f)}; }

struct GapView {           // tables of struct gapCalc (gapCalc.c:12-37) as the device sees them
    int smallSize, longCount, lastPos;
    int denseSize;                     // gapDense[which][v] holds gapCalcCost for every v < denseSize
    double lastVal[3], lastSlope[3];   // q, t, both
};

struct ScoreParams {
    const gat_job *jobs;
    const gat_block *blocks;
    unsigned long long nJobs, totalJobBlocks, nBlocks;
    const uint32_t *chunkJob;   // job containing the first job-block of each chunk
    uint32_t nChunks;
    uint32_t prefetchChunks;    // how many chunks ahead a CTA prefetches work-list records (= resident CTAs)
    GenomeView t, q;
    int coef[16];               // SYM: 6 coefficients, general: 16 Moebius coefficients
    GapView gap;
    const int *gapSmall;        // [3][smallSize] in global; staged to shared
    const int *gapDense;        // [3][denseSize] in global (L2-resident)
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
// spelled with the _rn intrinsics so nvcc can never contract them into an FMA.  This exact routine
// runs once per table entry when gat_set_scoring builds the dense cost table (gapDenseKernel) and,
// in the scoring kernel, only for gaps beyond that table.
__device__ __noinline__ int gapCostExact(const GapView &g, const int *small, const int *longPos,
                                         const double *longVal, int which, int v)
{
    if (v < g.smallSize) return small[which * g.smallSize + v];
    if (v >= g.lastPos)
        return truncToInt(__dadd_rn(g.lastVal[which], __dmul_rn(g.lastSlope[which], (double)(v - g.lastPos))));
    const double *val = longVal + which * g.longCount;
    for (int i = 0; i < g.longCount; i++) {
        const int p = longPos[i];
        if (v == p) return truncToInt(val[i]);
        if (v < p) {
            const int ds = p - longPos[i - 1];
            const double dv = __dsub_rn(val[i], val[i - 1]);
            const double prod = __dmul_rn(dv, (double)(v - longPos[i - 1]));
            return truncToInt(__dadd_rn(val[i - 1], __ddiv_rn(prod, (double)ds)));
        }
    }
    return INT32_MIN;   // unreachable: v < lastPos == longPos[longCount-1]
}

// Small gaps come from shared memory (the reference's qSmall/tSmall/bSmall), everything up to
// denseSize (normally the last knot, 252111) from a dense table that lives in L2, and only gaps
// beyond it evaluate the extrapolation -- one multiply and one add, no division.
__device__ __forceinline__ int gapCost(const GapView &g, const int *small, const int *__restrict__ dense,
                                       const int *longPos, const double *longVal, int dq, int dt)
{
    if (dt < 0) dt = 0;
    if (dq < 0) dq = 0;
    int which, v;
    if (dt == 0) { which = 0; v = dq; }
    else if (dq == 0) { which = 1; v = dt; }
    else { which = 2; v = (int)((unsigned)dq + (unsigned)dt); }
    if (v < 0) return INT32_MIN;                      // dq+dt overflowed int (undefined in the reference)
    if (v < g.smallSize) return small[which * g.smallSize + v];
    if (v < g.denseSize) return __ldg(dense + (size_t)which * g.denseSize + v);
    return gapCostExact(g, small, longPos, longVal, which, v);
}

__global__ void gapDenseKernel(GapView g, const int *__restrict__ small, const int *__restrict__ longPos,
                               const double *__restrict__ longVal, int *__restrict__ dense)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x, which = blockIdx.y;
    if (v < g.denseSize) dense[(size_t)which * g.denseSize + v] = gapCostExact(g, small, longPos, longVal, which, v);
}

// ------------------------------------------------------------------ base windows
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 1 << r for r < 32, else 0 (PTX shl clamps the shift amount; C++ << would be undefined)
__device__ __forceinline__ uint32_t shl1(uint32_t r)
{
    uint32_t out;
    asm("shl.b32 %0, 1, %1;" : "=r"(out) : "r"(r));
    return out;
}

// 32 bases starting `sh` bits into word n
__device__ __forceinline__ void loadWindow(const uint2 *__restrict__ planes, uint32_t n, uint32_t sh,
                                           uint32_t &hi, uint32_t &lo)
{
    const uint2 a = __ldg(planes + n), b = __ldg(planes + n + 1);
    hi = __funnelshift_r(a.x, b.x, sh);
    lo = __funnelshift_r(a.y, b.y, sh);
}
__device__ __forceinline__ uint32_t loadNWindow(const uint32_t *__restrict__ np, uint32_t n, uint32_t sh)
{
    return __funnelshift_r(__ldg(np + n), __ldg(np + n + 1), sh);
}

// does [g0, g0+len) touch a 256-base window that contains N?
__device__ __forceinline__ bool mayTouchN(const uint32_t *__restrict__ nwin, long long g0, int len)
{
    const uint32_t w0 = (uint32_t)((unsigned long long)g0 >> NWIN_SHIFT);          // genomes are < 2^36 bases
    const uint32_t w1 = (uint32_t)((unsigned long long)(g0 + len - 1) >> NWIN_SHIFT);
    const uint32_t word0 = w0 >> 5, word1 = w1 >> 5;
    const uint32_t loMask = 0xffffffffu << (w0 & 31), hiMask = 0xffffffffu >> (31 - (w1 & 31));
    if (word0 == word1) return (__ldg(nwin + word0) & loMask & hiMask) != 0;     // blocks under 8 kb
    if (__ldg(nwin + word0) & loMask) return true;
    for (uint32_t word = word0 + 1; word < word1; word++)
        if (__ldg(nwin + word)) return true;
    return (__ldg(nwin + word1) & hiMask) != 0;
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
// n < 2^20 (GAT_MAX_BLOCK_BASES); misc: tSh | qSh<<5 | minus<<10 | mayN<<11; excl: items of the warp before this block
struct __align__(16) StageRec { uint32_t tW, qW, nMisc, excl; };
constexpr int ERR_TOOLONG = 8, ERR_CSR = 16;

#ifndef GAT_MIN_CTAS
#define GAT_MIN_CTAS 5
#endif
#ifndef GAT_P1_UNROLL
#define GAT_P1_UNROLL 2
#endif
#ifndef GAT_PREFETCH
#define GAT_PREFETCH 1      // bit 0: genome windows from phase 1, bit 1: work-list records of a later chunk
#endif
constexpr int P1_UNROLL = GAT_P1_UNROLL;   // sub-tiles of phase 1 in flight per warp

// chainFastSubsetOnT clip (chain.c:513-522) of one record for one job
__device__ __forceinline__ void clipBlock(const gat_block &b, const gat_job &job, int &ts, int &qs, int &len, bool &joined)
{
    joined = (b.size & GAT_BLOCK_JOINED) != 0;
    const int size = (int)(b.size & 0x7fffffffu);
    ts = b.tStart; qs = b.qStart;
    int te = ts + size;
    if (ts < job.clipStart) { qs += job.clipStart - ts; ts = job.clipStart; }
    if (te > job.clipEnd) te = job.clipEnd;
    len = te - ts;
}

__device__ __forceinline__ gat_job loadJob(const gat_job *__restrict__ jobs, uint32_t j)
{
    const uint2 *p = reinterpret_cast<const uint2 *>(jobs + j);     // 24-byte records, 8-byte aligned
    const uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    gat_job r;
    r.tSeq = a.x; r.qSeq = a.y; r.firstBlock = b.x; r.blockPtr = b.y; r.clipStart = (int)c.x; r.clipEnd = (int)c.y;
    return r;
}
__device__ __forceinline__ gat_block loadBlock(const gat_block *__restrict__ blocks, unsigned long long i)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(blocks + i);
    gat_block r;
    r.tStart = (int)__ldg(p); r.qStart = (int)__ldg(p + 1); r.size = __ldg(p + 2);
    return r;
}

// Phase 3 of scoreChunksKernel for one warp (see there), in 32- or 64-bit tuples.
template <typename T>
__device__ __forceinline__ void warpJobReduce(const ScoreParams &P, const long long *sScore, const int *sGap,
                                              const unsigned char *sFlag, const uint32_t *sJob, int warpV0,
                                              unsigned long long vb0, unsigned long long chunkEnd, int warp, int lane,
                                              Tup *sWarpAgg, Tup *sWarpPend, int *sWarpHead, int *sWarpPendJob,
                                              int *sLastIsEnd, uint32_t *sLastJob)
{
    const T NEGT = negInf<T>();
    TupT<T> cur = tIdentity<T>();     // open segment at the end of my run
    bool runHasHead = false;
    bool pend = false;                // an END reached before any HEAD of my run: needs the carry
    TupT<T> pendTup = tIdentity<T>();
    uint32_t pendJob = 0;
    {
        const int o0 = 4 * lane;
        int p = padIdx(warpV0 + o0);
#pragma unroll
        for (int k = 0; k < 4; k++, p++) {
            const unsigned char fl = sFlag[p];
            if (fl & 8) {
                const T a = (T)sScore[p];
                const bool isEnd = fl & 2, joinedNext = fl & 4;
                T dY = a, cY = NEGT;
                if (!isEnd && !joinedNext) { dY = a - (T)sGap[p]; cY = 0; }
                if (fl & 1) { cur = tIdentity<T>(); runHasHead = true; }
                // cur = cur (+) element, specialised for a single block (its f is -inf)
                TupT<T> r;
                r.d = cur.d + dY;
                r.c = maxT<T>(cY, cur.c + dY);
                r.e = joinedNext ? cur.e : maxT<T>(cur.e, cur.d + a);
                r.f = joinedNext ? cur.f : maxT<T>(cur.f, cur.c + a);
                cur = r;
                if (isEnd) {
                    const uint32_t job = sJob[p] - 1;
                    if (runHasHead) {   // job lies inside my run: done
                        P.outGlobal[job] = (long long)cur.d;
                        P.outLocal[job] = max64(0, max64(widen(cur.e), widen(cur.f)));
                    } else { pend = true; pendTup = cur; pendJob = job; }
                }
                if (vb0 + (unsigned long long)(warpV0 + o0 + k) + 1 == chunkEnd) {
                    *sLastIsEnd = isEnd;                // the chunk's last valid job-block
                    *sLastJob = sJob[p] - 1;
                }
            }
        }
    }
    // warp-level segmented inclusive scan of (cur, runHasHead)
    TupT<T> inc = cur;
    bool incHead = runHasHead;
    for (int off = 1; off < 32; off <<= 1) {
        TupT<T> o = tShfl<T>(inc, lane >= off ? lane - off : lane);
        bool oh = __shfl_sync(FULL, (int)incHead, lane >= off ? lane - off : lane);
        if (lane >= off && !incHead) { inc = tCombine<T>(o, inc); incHead = oh; }
    }
    TupT<T> carry = tShfl<T>(inc, lane ? lane - 1 : 0);
    bool carryHead = __shfl_sync(FULL, (int)incHead, lane ? lane - 1 : 0);
    if (lane == 0) { carry = tIdentity<T>(); carryHead = false; }
    if (pend) {
        const TupT<T> fin = tCombine<T>(carry, pendTup);
        if (carryHead) {            // the job started inside this warp
            P.outGlobal[pendJob] = (long long)fin.d;
            P.outLocal[pendJob] = max64(0, max64(widen(fin.e), widen(fin.f)));
        } else {                    // it started before this warp: at most one such lane per warp
            sWarpPend[warp] = tWiden<T>(fin); sWarpPendJob[warp] = (int)pendJob;
        }
    }
    const bool anyCross = __any_sync(FULL, pend && !carryHead);
    if (lane == 31) {
        sWarpAgg[warp] = tWiden<T>(inc); sWarpHead[warp] = incHead;
        if (!anyCross) sWarpPendJob[warp] = -1;
    }
}

template <bool SYM>
__global__ void __launch_bounds__(TPB, GAT_MIN_CTAS)
scoreChunksKernel(const __grid_constant__ ScoreParams P)
{
    // per job-block of the chunk, index padIdx(v)
    __shared__ uint32_t sJob[PADDED];           // job index + 1
    __shared__ long long sScore[PADDED];        // block score (accumulated by the item loop)
    __shared__ int sGap[PADDED];                // cost of the gap that follows the block
    __shared__ unsigned char sFlag[PADDED];     // 1 head of job, 2 end of job, 4 next record joined, 8 valid
    // per warp: its 128 blocks expanded to items
    __shared__ StageRec sStage[WARPS][32 * BPT];
    __shared__ int sAcc[WARPS][32 * BPT];
    __shared__ uint32_t sWarpMax[WARPS];
    __shared__ Tup sWarpAgg[WARPS], sWarpPend[WARPS];
    __shared__ int sWarpHead[WARPS], sWarpPendJob[WARPS];
    __shared__ int sArrived, sLastIsEnd;
    __shared__ uint32_t sLastJob;
    extern __shared__ unsigned char sDyn[];     // gap tables

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long vb0 = (unsigned long long)blockIdx.x * CHUNK;
    const unsigned long long total = P.totalJobBlocks;
    if (tid == 0) sArrived = 0;

#if GAT_PREFETCH & 2
    {   // the records of the chunk that will run when this one retires: stream them into L2 now
        const unsigned long long ahead = vb0 + (unsigned long long)P.prefetchChunks * CHUNK;
        if (ahead < total) {
            const char *b0 = reinterpret_cast<const char *>(P.blocks + ahead);        // whole-chain work-lists: record index == job-block index
            const unsigned long long bytes = ((total - ahead < (unsigned long long)CHUNK) ? total - ahead : (unsigned long long)CHUNK) * sizeof(gat_block);
            for (unsigned long long off = (unsigned long long)tid * 32; off < bytes; off += TPB * 32) prefetchL2(b0 + off);
            const uint32_t ca = blockIdx.x + P.prefetchChunks;
            if (warp == 0 && ca + 1 < P.nChunks) {
                const uint32_t ja = __ldg(P.chunkJob + ca), jb = __ldg(P.chunkJob + ca + 1);
                const char *j0p = reinterpret_cast<const char *>(P.jobs + ja);
                const unsigned long long jbytes = (unsigned long long)(jb - ja + 1) * sizeof(gat_job);
                for (unsigned long long off = (unsigned long long)lane * 32; off < jbytes; off += 32 * 32) prefetchL2(j0p + off);
            }
        }
    }
#endif
    // ---- stage gap tables (gapCalc.c:12-37) in shared memory
    double *gLongVal = reinterpret_cast<double *>(sDyn);
    int *gLongPos = reinterpret_cast<int *>(gLongVal + 3 * P.gap.longCount);
    int *gSmall = gLongPos + P.gap.longCount;
#pragma unroll 1
    for (int i = tid; i < 3 * P.gap.longCount; i += TPB) gLongVal[i] = P.gapLongVal[i];
#pragma unroll 1
    for (int i = tid; i < P.gap.longCount; i += TPB) gLongPos[i] = P.gapLongPos[i];
#pragma unroll 1
    for (int i = tid; i < 3 * P.gap.smallSize; i += TPB) gSmall[i] = P.gapSmall[i];

    // ---- phase 0: which job owns each job-block of this chunk
#pragma unroll 1
    for (int i = tid; i < PADDED; i += TPB) sJob[i] = 0;
    __syncthreads();
    {
        const uint32_t j0 = P.chunkJob[blockIdx.x];
        const uint32_t jEnd = (blockIdx.x + 1 < P.nChunks) ? P.chunkJob[blockIdx.x + 1] : (uint32_t)(P.nJobs - 1);
        const uint32_t jStart = blockIdx.x == 0 ? 0u : j0;   // chunk 0 also sweeps leading empty jobs
        for (uint32_t j = jStart + tid; j <= jEnd; j += TPB) {
            unsigned long long bp = __ldg(&P.jobs[j].blockPtr);
            unsigned long long np = (j + 1 < P.nJobs) ? (unsigned long long)__ldg(&P.jobs[j + 1].blockPtr) : total;
            if (np > bp) {                                  // non-empty job
                if (bp >= vb0 && bp < vb0 + CHUNK) sJob[padIdx((int)(bp - vb0))] = j + 1;
                else if (bp < vb0 && j == j0) sJob[0] = j + 1;
            } else {                                        // empty job (kent: NULL sub-chain): scores 0
                P.outGlobal[j] = 0;
                P.outLocal[j] = 0;
            }
        }
    }
    __syncthreads();
    {   // inclusive max-scan: job indices grow with position, so max = nearest head at or before
        const int i0 = padIdx(4 * tid);                     // 4 consecutive job-blocks never straddle a pad
        uint32_t m0 = sJob[i0], m1 = max(m0, sJob[i0 + 1]), m2 = max(m1, sJob[i0 + 2]), m3 = max(m2, sJob[i0 + 3]);
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
        sJob[i0] = max(m0, before);
        sJob[i0 + 1] = max(m1, before);
        sJob[i0 + 2] = max(m2, before);
        sJob[i0 + 3] = max(m3, before);
    }
    __syncthreads();

    // ---- phase 1: this warp's 128 job-blocks, 32 at a time: load + clip the records, gap costs,
    // item counts.  Block v = warp*128 + sub*32 + lane.
    bool anyN = false;
#pragma unroll P1_UNROLL
    for (int sub = 0; sub < BPT; sub++) {
        const int v = warp * (32 * BPT) + sub * 32 + lane;
        const int pv = padIdx(v);
        const unsigned long long gv = vb0 + v;
        bool valid = gv < total;
        if (valid && sJob[pv] == 0) {       // blockPtr is not a non-decreasing CSR row pointer
            atomicOr(P.err, ERR_CSR);
            valid = false;
        }
        uint32_t tW = 0, qW = 0, n = 0, misc = 0;
        unsigned char flag = 0;
        int ts = 0, qs = 0, len = 0;
        bool joined = false;
        gat_job job;
        job.tSeq = job.qSeq = job.firstBlock = job.blockPtr = 0; job.clipStart = job.clipEnd = 0;
        if (valid) {
            const uint32_t j = sJob[pv] - 1;
            job = loadJob(P.jobs, j);
            flag = 8;
            if (gv == job.blockPtr) flag |= 1;
            // end of job <=> the next job-block belongs to another job (or there is none)
            if (v + 1 < CHUNK) { if (gv + 1 >= total || sJob[padIdx(v + 1)] - 1 != j) flag |= 2; }
            else {
                const unsigned long long np = (j + 1 < P.nJobs) ? (unsigned long long)__ldg(&P.jobs[j + 1].blockPtr) : total;
                if (gv + 1 == np) flag |= 2;
            }
            const unsigned long long bi = (unsigned long long)job.firstBlock + (gv - job.blockPtr);
            const uint32_t qSeq = job.qSeq & 0x7fffffffu;
            const bool minus = (job.qSeq >> 31) != 0;
            if (bi >= P.nBlocks) { atomicOr(P.err, ERR_BLOCKIDX); }
            else if (job.tSeq >= P.t.nSeq || qSeq >= P.q.nSeq) { atomicOr(P.err, ERR_SEQ); }
            else {
                clipBlock(loadBlock(P.blocks, bi), job, ts, qs, len, joined);
                const int nn = len > 0 ? len : 0;
                const uint32_t tSize = __ldg(P.t.seqSize + job.tSeq), qSize = __ldg(P.q.seqSize + qSeq);
                if (nn > 0 && (ts < 0 || qs < 0 || (unsigned)ts + (unsigned)nn > tSize || (unsigned)qs + (unsigned)nn > qSize)) {
                    atomicOr(P.err, ERR_COORD);
                } else if (nn >= (1 << 20)) {
                    atomicOr(P.err, ERR_TOOLONG);
                } else if (nn > 0) {
                    n = (uint32_t)nn;
                    const long long tG = __ldg(P.t.seqBase + job.tSeq) + ts;
                    // '+': first base of the block.  '-': one past the block's last base in forward
                    // coordinates; rc position p is forward position qSize-1-p (dnautil.c:466-470).
                    const long long qBase = __ldg(P.q.seqBase + qSeq);
                    const long long qG = minus ? qBase + ((long long)qSize - qs) : qBase + qs;
                    const long long qLo = minus ? qG - nn : qG;
                    const bool mayN = mayTouchN(P.t.nwin, tG, nn) || mayTouchN(P.q.nwin, qLo, nn);
                    tW = (uint32_t)(tG >> 5);
                    qW = minus ? (uint32_t)((qG - 32) >> 5) : (uint32_t)(qG >> 5);
                    misc = (uint32_t)(tG & 31) | ((uint32_t)(qG & 31) << 5) | (minus ? 1u << 10 : 0u) | (mayN ? 1u << 11 : 0u);
                    anyN |= mayN;
#if GAT_PREFETCH & 1
                    // phase 2 reads these windows a few microseconds from now: pull the first sectors into L2
                    prefetchL2(P.t.planes + tW);
                    prefetchL2(P.q.planes + (minus ? qW + 1 : qW));
#endif
                }
            }
        }
        // the block after mine (same job): lane+1 holds it; lane 31 fetches it itself
        int nts = __shfl_down_sync(FULL, ts, 1), nqs = __shfl_down_sync(FULL, qs, 1);
        bool njoined = __shfl_down_sync(FULL, (int)joined, 1);
        if (lane == 31 && valid && !(flag & 2)) {
            const unsigned long long bi = (unsigned long long)job.firstBlock + (gv + 1 - job.blockPtr);
            nts = nqs = 0; njoined = false;
            if (bi < P.nBlocks) {
                int nlen;
                clipBlock(loadBlock(P.blocks, bi), job, nts, nqs, nlen, njoined);
            }
        }
        int gap = 0;
        if (valid && !(flag & 2)) {
            if (njoined) flag |= 4;
            else gap = gapCost(P.gap, gSmall, P.gapDense, gLongPos, gLongVal, nqs - (qs + len), nts - (ts + len));
        }
        sGap[pv] = gap;
        sFlag[pv] = flag;
        sScore[pv] = 0;
        sStage[warp][sub * 32 + lane] = StageRec{tW, qW, n | (misc << 20), n ? (n + 31) >> 5 : 1u};   // excl = item count for now
        sAcc[warp][sub * 32 + lane] = 0;
    }
    anyN = __any_sync(FULL, anyN);

    // ---- phase 2: the warp's 128 blocks as one list of 32-base items, dealt to lanes.
    // For the owner search lane l speaks for blocks 4l..4l+3 of the warp (consecutive), so their
    // exclusive item prefixes are a local prefix plus one warp scan.
    const int warpV0 = warp * (32 * BPT);
    // Long blocks first: whole rounds of 1024 bases of ONE block need no item bookkeeping at all --
    // every lane takes one 32-base window per round, shifts and strand are warp-uniform, partial sums
    // stay in a register until the block's bulk is done.  What is left (< 1024 bases) joins the item list.
    {
        __syncwarp();
        const StageRec *mine = &sStage[warp][4 * lane];
        unsigned longBits = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) longBits |= ((mine[k].nMisc & 0xfffffu) >= 2u * LONG_ROUND_BASES ? 1u : 0u) << k;
        for (int k = 0; k < 4; k++) {
            unsigned m = __ballot_sync(FULL, (longBits >> k) & 1u);
            while (m) {
                const int o = 4 * (__ffs(m) - 1) + k;           // block index inside the warp (uniform)
                m &= m - 1;
                const StageRec r = sStage[warp][o];
                const uint32_t n = r.nMisc & 0xfffffu, misc = r.nMisc >> 20;
                const uint32_t rounds = n / LONG_ROUND_BASES;
                const uint32_t tSh = misc & 31u, qSh = (misc >> 5) & 31u;
                const bool minus = (misc >> 10) & 1u, mayN = (misc >> 11) & 1u;
                long long acc = 0;
                for (uint32_t rd = 0; rd < rounds; rd++) {
                    const uint32_t w = rd * 32 + lane;
                    uint32_t t1, t0, q1, q0;
                    loadWindow(P.t.planes, r.tW + w, tSh, t1, t0);
                    const uint32_t qn = minus ? r.qW - w : r.qW + w;
                    loadWindow(P.q.planes, qn, qSh, q1, q0);
                    if (minus) { q1 = ~__brev(q1); q0 = __brev(q0); }
                    uint32_t vmask = 0xffffffffu;
                    int nv = 32;
                    if (mayN) {
                        uint32_t nt = loadNWindow(P.t.nplane, r.tW + w, tSh);
                        uint32_t nq = loadNWindow(P.q.nplane, qn, qSh);
                        if (minus) nq = __brev(nq);
                        vmask = ~(nt | nq);
                        nv = __popc(vmask);
                    }
                    acc += scoreWindow<SYM>(P.coef, t1, t0, q1, q0, vmask, nv);
                }
                for (int off = 16; off; off >>= 1) acc += shfl64(acc, lane ^ off);
                if (lane == 0) {
                    sScore[padIdx(warpV0 + o)] += acc;
                    const uint32_t done = rounds * LONG_ROUND_BASES, rem = n - done;
                    sStage[warp][o] = StageRec{r.tW + rounds * 32, minus ? r.qW - rounds * 32 : r.qW + rounds * 32, rem | (misc << 20),
                                               rem ? (rem + 31) >> 5 : 1u};
                }
                __syncwarp();
            }
        }
    }
    uint32_t ex0, ex1, ex2, ex3, totalItems;
    {
        __syncwarp();
        StageRec *st = &sStage[warp][4 * lane];
        const uint32_t c0 = st[0].excl, c1 = st[1].excl, c2 = st[2].excl, c3 = st[3].excl;
        uint32_t incl = c0 + c1 + c2 + c3;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += o;
        }
        totalItems = __shfl_sync(FULL, incl, 31);
        ex0 = incl - (c0 + c1 + c2 + c3); ex1 = ex0 + c0; ex2 = ex1 + c1; ex3 = ex2 + c2;
        st[0].excl = ex0; st[1].excl = ex1; st[2].excl = ex2; st[3].excl = ex3;
        __syncwarp();
    }
    {
        // Software pipeline: the loads of round r+1 are issued before round r is scored, so two
        // rounds of genome windows are in flight per warp.
        // A round's state: heads (bit i = a block starts at lane i), owner (block of this lane's
        // item), meta = valid bases (6 bits) | misc << 6, and the four uint2 window halves.
        auto fetch = [&](uint32_t base, int before, unsigned &heads, int &owner, uint32_t &meta,
                         uint2 &ta, uint2 &tb, uint2 &qa, uint2 &qb) {
            heads = __reduce_or_sync(FULL, shl1(ex0 - base) | shl1(ex1 - base) | shl1(ex2 - base) | shl1(ex3 - base));
            owner = before - 1 + __popc(heads & (0xffffffffu >> (31 - lane)));
            meta = 0;
            ta = tb = qa = qb = make_uint2(0, 0);
            if (base + lane < totalItems) {
                const StageRec r = sStage[warp][owner];
                const uint32_t k = base + lane - r.excl, misc = r.nMisc >> 20;
                const int left = (int)(r.nMisc & 0xfffffu) - (int)(k << 5);
                if (left > 0) {
                    meta = (uint32_t)(left >= 32 ? 32 : left) | (misc << 6);
                    const uint2 *tp = P.t.planes + (r.tW + k);
                    const uint2 *qp = P.q.planes + ((misc >> 10) & 1u ? r.qW - k : r.qW + k);
                    ta = __ldg(tp); tb = __ldg(tp + 1); qa = __ldg(qp); qb = __ldg(qp + 1);
                }
            }
        };
        int before = 0;                         // blocks that start before the current round
        unsigned hA; int oA; uint32_t mA; uint2 a0, a1, a2, a3;
        if (totalItems) fetch(0, 0, hA, oA, mA, a0, a1, a2, a3);
        for (uint32_t base = 0; base < totalItems; base += 32) {
            unsigned hB = 0; int oB = 0; uint32_t mB = 0; uint2 b0, b1, b2, b3;
            b0 = b1 = b2 = b3 = make_uint2(0, 0);
            const int beforeNext = before + __popc(hA);
            if (base + 32 < totalItems) fetch(base + 32, beforeNext, hB, oB, mB, b0, b1, b2, b3);
            int s = 0;
            int nv = (int)(mA & 63u);
            if (nv) {
                const uint32_t misc = mA >> 6, tSh = misc & 31u, qSh = (misc >> 5) & 31u;
                const bool minus = (misc >> 10) & 1u;
                uint32_t vmask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
                const uint32_t t1 = __funnelshift_r(a0.x, a1.x, tSh), t0 = __funnelshift_r(a0.y, a1.y, tSh);
                uint32_t q1 = __funnelshift_r(a2.x, a3.x, qSh), q0 = __funnelshift_r(a2.y, a3.y, qSh);
                if (minus) { q1 = ~__brev(q1); q0 = __brev(q0); }   // reverse, complement = flip bit1
                if (anyN && ((misc >> 11) & 1u)) {
                    const StageRec r = sStage[warp][oA];
                    const uint32_t k = base + lane - r.excl;
                    uint32_t nt = loadNWindow(P.t.nplane, r.tW + k, tSh);
                    uint32_t nq = loadNWindow(P.q.nplane, minus ? r.qW - k : r.qW + k, qSh);
                    if (minus) nq = __brev(nq);
                    vmask &= ~(nt | nq);            // N scores 0 against everything (axt.c:431-454)
                    nv = __popc(vmask);
                }
                s = scoreWindow<SYM>(P.coef, t1, t0, q1, q0, vmask, nv);
            }
            // hand the 32 partial sums to the blocks that own them
            if ((hA >> 1) == 0) {           // one block owns this whole round (long block): warp reduce
                const int tot = __reduce_add_sync(FULL, s);
                if (lane == 0) sScore[padIdx(warpV0 + before - 1 + (int)(hA & 1u))] += tot;
            } else if (base + lane < totalItems) {
                atomicAdd(&sAcc[warp][oA], s);      // <= 2 such rounds per block: no 32-bit overflow (|M| <= 2^19)
            }
            before = beforeNext;
            hA = hB; oA = oB; mA = mB; a0 = b0; a1 = b1; a2 = b2; a3 = b3;
        }
    }
    __syncwarp();

    // ---- phase 3: ordered segmented reduction of tuples, per warp: lane l walks job-blocks 4l..4l+3
    // of the warp, one warp scan joins the lanes; what crosses warps is resolved by whichever warp
    // of the CTA finishes last (no CTA-wide barrier: warps retire at their own pace).
    {
        // 32-bit tuples if every partial sum of this warp's blocks stays below 2^28
        long long mag = 0;
        const int p0 = padIdx(warpV0 + 4 * lane);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long a = sScore[p0 + k] + sAcc[warp][4 * lane + k];
            sScore[p0 + k] = a;
            const long long g = sGap[p0 + k];
            mag += (a < 0 ? -a : a) + (g < 0 ? -g : g);
        }
        const bool small = __all_sync(FULL, mag < (1LL << 23));
        const unsigned long long chunkEnd = total < vb0 + CHUNK ? total : vb0 + CHUNK;
        if (small) warpJobReduce<int>(P, sScore, sGap, sFlag, sJob, warpV0, vb0, chunkEnd, warp, lane,
                                      sWarpAgg, sWarpPend, sWarpHead, sWarpPendJob, &sLastIsEnd, &sLastJob);
        else warpJobReduce<long long>(P, sScore, sGap, sFlag, sJob, warpV0, vb0, chunkEnd, warp, lane,
                                      sWarpAgg, sWarpPend, sWarpHead, sWarpPendJob, &sLastIsEnd, &sLastJob);
    }
    // last warp of the CTA to get here stitches the warps together
    __threadfence_block();
    __syncwarp();
    int arrived = 0;
    if (lane == 0) arrived = atomicAdd(&sArrived, 1);
    arrived = __shfl_sync(FULL, arrived, 0);
    if (arrived != WARPS - 1 || lane != 0) return;
    __threadfence_block();
    Tup c = tupIdentity();
    bool ch = false;
    for (int w = 0; w < WARPS; w++) {
        if (sWarpPendJob[w] >= 0) {
            const Tup fin = tupCombine(c, sWarpPend[w]);
            if (ch) {
                P.outGlobal[sWarpPendJob[w]] = fin.d;
                P.outLocal[sWarpPendJob[w]] = max64(0, max64(fin.e, fin.f));
            } else P.chunkHead[blockIdx.x] = fin;       // job began in an earlier chunk and ends here
        }
        if (sWarpHead[w]) { c = sWarpAgg[w]; ch = true; }
        else c = tupCombine(c, sWarpAgg[w]);
    }
    // the chunk's last valid job-block: does its job run on into the next chunk?
    if (sLastIsEnd) P.chunkTailJob[blockIdx.x] = -1;
    else if (ch) { P.chunkTail[blockIdx.x] = c; P.chunkTailJob[blockIdx.x] = (int)sLastJob; }
    else { P.chunkHead[blockIdx.x] = c; P.chunkTailJob[blockIdx.x] = -1; }
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
// .2bit payload (4 bases/byte, first base in bits 7..6, twoBit.c:811-818) -> bit-sliced words.
// One thread per 32-base word of one sequence.
__global__ void repackKernel(const uint8_t *__restrict__ raw, const unsigned long long *__restrict__ seqByteOffset,
                             const uint32_t *__restrict__ seqSize, const long long *__restrict__ seqBase,
                             const unsigned long long *__restrict__ seqWordStart, uint32_t nSeq,
                             unsigned long long totalWords, uint2 *__restrict__ planes)
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
    planes[n] = make_uint2(hiW, loW);
}

// N runs (twoBit.c:835-851) -> N plane + 256-base window summary.  One warp per run.
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
