// gat_kernels.cuh -- sm_100a kernels of the chain-rescoring path.
//
// Replaces, per job, the CPU loops of kent chainCalcScore / chainScoreBlock
// (kent/src/lib/chainConnect.c:14-40), gapCalcCost (kent/src/lib/gapCalc.c:298-331) and
// hillerlab chainCalcScoreLocal (src/scoreChain/scoreChain.c:176-198), including the clip of
// chainFastSubsetOnT (kent/src/lib/chain.c:510-522).  See DESIGN.md for the data layout.
//
// HBM layout of a genome ("bit-sliced 2-bit"): 32 consecutive bases are one uint2 = {word of high
// bits, word of low bits} of the kent base code T=0 C=1 A=2 G=3; base p sits at bit p%32.  So 32
// aligned bases are ONE 8-byte load, 128 bases are one 32-byte DRAM sector, and an unaligned 32-base
// window is a funnel shift over two neighbouring uint2.  The query is resident twice: forward and
// reverse-complemented (revCompKernel), so a '-' chain reads forward in its own coordinates and the
// scoring kernel has no strand logic.  A third, separate plane holds N (1 bit/base) and is only read for
// blocks whose 256-base windows contain N.
//
// Work decomposition (scoreTilesKernel, gat_tiles.cuh): the job-blocks of all jobs form one virtual array; a CTA owns CHUNK consecutive
// job-blocks, a warp 128 of them.  The first 32 bases of every block are scored lane = block; what is
// left is expanded into 32-base "items" and the items -- not the blocks -- are dealt to lanes, so a
// 30 kb block and a 40 bp block cost what their bases cost.  Per-job global and local scores are a
// segmented, ordered reduction of a 4-number max-plus tuple (see Tup).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gat.h"

namespace gat {

#ifndef GAT_TPB
#define GAT_TPB 64
#endif
constexpr int TPB = GAT_TPB;           // threads per CTA
constexpr int WARPS = TPB / 32;
constexpr int BPT = 4;                 // job-blocks per thread = 32-block tiles per warp
constexpr int CHUNK = TPB * BPT;       // job-blocks per CTA
constexpr unsigned FULL = 0xffffffffu;
constexpr int GROUP_BASES = 128;       // sequences start on a 32-byte sector boundary
// Idle lanes of the last item round read up to 31 words beyond (or, on '-', before) the last block of a warp,
// and funnel shifts read one word past a window: 36 words of slack at both ends of the genome buffers.
constexpr int PAD_FRONT_GROUPS = 9;
constexpr int PAD_BACK_GROUPS = 26;
constexpr int NWIN_SHIFT = 8;          // N summary: one bit per 256 bases

constexpr int ERR_SEQ = 1, ERR_BLOCKIDX = 2, ERR_COORD = 4, ERR_TOOLONG = 8, ERR_CSR = 16;
constexpr int ERR_EMPTYJOB = 32;   // not an error: a job without blocks; the host re-runs the list without such jobs

struct GenomeView {
    const uint2 *planes;      // word n = {high bits, low bits} of bases [32n, 32n+32)
    const uint32_t *nplane;   // word n = N bits of bases [32n, 32n+32)
    const uint32_t *nwin;     // bit w = window w (256 bases) contains an N
    const uint2 *nwin2;       // nwin2[i] = {nwin[i], nwin[i + 1]}: one aligned load covers the 32 windows from any window on
    const int64_t *seqBase;   // first base of each sequence in the padded coordinate (multiple of 128)
    const uint32_t *seqSize;
    uint32_t nSeq;
    uint32_t words;           // 32-base words in `planes`
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
// (sum of |block score| + gap cost below 2^27): a third of the instructions of the 64-bit form.
template <typename T> struct TupT { T d, c, e, f; };
template <typename T> __device__ __forceinline__ T negInf();
template <> __device__ __forceinline__ long long negInf<long long>() { return NEG; }
template <> __device__ __forceinline__ int negInf<int>() { return -(1 << 29); }   // two of them still add without overflow
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
// to the 64-bit tuple that crosses warps / CTAs; anything below -2^28 is "minus infinity"
__device__ __forceinline__ long long widen(long long v) { return v; }
__device__ __forceinline__ long long widen(int v) { return v < -(1 << 28) ? NEG : (long long)v; }
template <typename T> __device__ __forceinline__ Tup tWiden(const TupT<T> &t) { return Tup{(long long)t.d, widen(t.c), widen(t.e), widen(t.f)}; }

struct GapView {           // tables of struct gapCalc (gapCalc.c:12-37) as the device sees them
    int smallSize, longCount, lastPos;
    int denseSize;                     // gapDense[which][v] holds gapCalcCost for every v < denseSize
    double lastVal[3], lastSlope[3];   // q, t, both
};

// A job as the scoring kernel reads it (written by jobPrepKernel): one 32-byte sector.
struct __align__(16) JobInfo {
    uint32_t tBaseW, qBaseW;    // index of the 32-base word that holds base 0 of the target / query sequence
    uint32_t tSize, qSize;      // sequence sizes.  A '-' job's qBaseW is that of the reverse-complement copy of the query sequence
    int32_t clipStart, clipEnd;
    uint32_t delta;             // firstBlock - blockPtr (mod 2^32): record index = job-block index + delta
    uint32_t blockPtr;
};

struct ScoreParams {
    const JobInfo *info;        // [nJobs + 1], the last one closes the CSR
    const gat_block *blocks;
    unsigned long long nJobs, totalJobBlocks, nBlocks;
    const uint32_t *chunkJob;   // job containing the first job-block of each chunk
    const uint32_t *headBits;   // bit b: a (non-empty) job starts at job-block b; bit totalJobBlocks closes the list
    uint32_t nChunks;
    uint32_t chunkBase;         // first chunk of this launch (gat_score_compact scores slices as their records arrive)
    uint32_t maxBlockBases;     // records longer than this are rejected (32-bit block sums)
    uint32_t smallBases;        // a block of up to this many bases scores below 2^19 in absolute value (32-bit job tuples)
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
    const int *modeFlags;       // verdicts of jobPrepKernel about the whole list (MODE_*); NULL: the host knows the mode
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

__device__ __forceinline__ int gapCostOf(const GapView &g, const int *small, const int *__restrict__ dense,
                                         const int *longPos, const double *longVal, uint32_t which, uint32_t v)
{
    if ((int)v < 0) return INT32_MIN;                 // dq+dt overflowed int (undefined in the reference)
    if (v < (uint32_t)g.smallSize) return small[which * g.smallSize + v];
    if (v < (uint32_t)g.denseSize) return __ldg(dense + (size_t)which * g.denseSize + v);
    return gapCostExact(g, small, longPos, longVal, (int)which, (int)v);
}

// gapCalcCost for a batch of (dq, dt) pairs, by the routines the scoring kernel uses (dense table, then the exact one)
__global__ void gapBatchKernel(GapView g, const int *__restrict__ small, const int *__restrict__ dense, const int *__restrict__ longPos,
                               const double *__restrict__ longVal, const int *__restrict__ dq, const int *__restrict__ dt,
                               unsigned long long n, int *__restrict__ out);
__global__ void gapDenseKernel(GapView g, const int *__restrict__ small, const int *__restrict__ longPos,
                               const double *__restrict__ longVal, int *__restrict__ dense)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x, which = blockIdx.y;
    if (v < g.denseSize) dense[(size_t)which * g.denseSize + v] = gapCostExact(g, small, longPos, longVal, which, v);
}
__global__ void gapBatchKernel(GapView g, const int *__restrict__ small, const int *__restrict__ dense, const int *__restrict__ longPos,
                               const double *__restrict__ longVal, const int *__restrict__ dq, const int *__restrict__ dt,
                               unsigned long long n, int *__restrict__ out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gapCost(g, small, dense, longPos, longVal, dq[i], dt[i]);
}

// ------------------------------------------------------------------ base windows
// 1 << r for r < 32, else 0 (PTX shl clamps the shift amount; C++ << would be undefined)
__device__ __forceinline__ uint32_t shl1(uint32_t r)
{
    uint32_t out;
    asm("shl.b32 %0, 1, %1;" : "=r"(out) : "r"(r));
    return out;
}

// 0xffffffff >> r for r < 32, else 0
__device__ __forceinline__ uint32_t shrOnes(uint32_t r)
{
    uint32_t out;
    asm("shr.b32 %0, 0xffffffff, %1;" : "=r"(out) : "r"(r));
    return out;
}
// one step of an inclusive warp scan: x += (x of lane - off) where that lane exists
__device__ __forceinline__ int scanStep(int x, int off)
{
    asm volatile("{ .reg .pred p; .reg .s32 y; shfl.sync.up.b32 y|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, y; }" : "+r"(x) : "r"(off));
    return x;
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

// valid-base mask of one 32-base window pair when N may be present (cold)
__device__ __noinline__ uint32_t nFreeMask(const uint32_t *__restrict__ tn, uint32_t tW, uint32_t tSh,
                                           const uint32_t *__restrict__ qn, uint32_t qW, uint32_t qSh)
{
    return ~(loadNWindow(tn, tW, tSh) | loadNWindow(qn, qW, qSh));      // N scores 0 against everything (axt.c:431-454)
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

// ------------------------------------------------------------------ programmatic dependent launch
// The kernels of one scoring pass (jobPrepKernel -> scoreTilesKernel -> fixupKernel) are launched with programmatic
// stream serialization: a kernel's CTAs may be scheduled while the kernel in front of it in the stream drains, and
// dependsWait() is where it stops until that kernel has completed and its writes are visible.  Whatever a kernel does
// before dependsWait() must not touch anything the kernel in front of it (or any kernel that one waits for) writes.
// In a kernel launched without the attribute both are no-ops.
__device__ __forceinline__ void dependsWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void dependentsMayLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }


// ------------------------------------------------------------------ job preparation
// One thread per job: gat_job -> JobInfo (sequence bases and sizes resolved once per job instead of once
// per block), zero scores for empty jobs (kent: NULL sub-chain), CSR validation, and chunkJob[c] = the job
// that owns job-block c*CHUNK, written by the job itself for every chunk boundary it covers.
__global__ void jobPrepKernel(const gat_job *__restrict__ jobs, unsigned long long nJobs, unsigned long long total,
                              const int64_t *__restrict__ tSeqBase, const uint32_t *__restrict__ tSeqSize, uint32_t tNSeq,
                              const int64_t *__restrict__ qSeqBase, const uint32_t *__restrict__ qSeqSize, uint32_t qNSeq,
                              JobInfo *__restrict__ info, uint32_t *__restrict__ chunkJob, uint32_t nChunks,
                              uint32_t *__restrict__ headBits, int *__restrict__ modeFlags,
                              long long *__restrict__ outGlobal, long long *__restrict__ outLocal, int *__restrict__ err)
{
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    dependsWait();                  // (the fix-up kernel of the pass before this one reads what this kernel writes)
    dependentsMayLaunch();          // the scoring kernel starts on its record copies meanwhile
    JobInfo o;
    o.tBaseW = o.qBaseW = o.tSize = o.qSize = 0; o.clipStart = o.clipEnd = 0; o.delta = 0;
    o.blockPtr = (uint32_t)total;                       // sentinel record nJobs closes the CSR
    unsigned long long c0 = 0, c1 = 0;                  // chunks [c0, c1) start inside this job
    bool general = false;                               // clips, or does not start at its own blockPtr
    if (j < nJobs) {
        const gat_job job = loadJob(jobs, (uint32_t)j);
        general = job.firstBlock != job.blockPtr || job.clipStart != GAT_NO_CLIP_START || job.clipEnd != GAT_NO_CLIP_END;
        const unsigned long long bp = job.blockPtr;
        const unsigned long long np = (j + 1 < nJobs) ? (unsigned long long)__ldg(&jobs[j + 1].blockPtr) : total;
        int e = 0;
        if (np < bp || np > total || (j == 0 && bp != 0)) e |= ERR_CSR;
        const uint32_t qSeq = job.qSeq & 0x7fffffffu;
        o.blockPtr = job.blockPtr;
        o.delta = job.firstBlock - job.blockPtr;
        o.clipStart = job.clipStart; o.clipEnd = job.clipEnd;
        if (job.tSeq >= tNSeq || qSeq >= qNSeq) {
            if (np > bp) e |= ERR_SEQ;                  // sizes stay 0: every block of the job fails its range check
        } else {
            o.tBaseW = (uint32_t)(__ldg(tSeqBase + job.tSeq) >> 5);
            // the query buffer holds every sequence twice: forward (index s) and reverse-complemented (qNSeq + s)
            o.qBaseW = (uint32_t)(__ldg(qSeqBase + ((job.qSeq >> 31) ? qNSeq + qSeq : qSeq)) >> 5);
            o.tSize = __ldg(tSeqSize + job.tSeq);
            o.qSize = __ldg(qSeqSize + qSeq);
        }
        if (np <= bp) { outGlobal[j] = 0; outLocal[j] = 0; e |= ERR_EMPTYJOB; }
        if (e) atomicOr(err, e);
        if (np > bp && !(e & ERR_CSR)) {
            atomicOr(&headBits[bp >> 5], 1u << (bp & 31));
            c0 = (bp + CHUNK - 1) / CHUNK;
            c1 = (np + CHUNK - 1) / CHUNK;
            if (c1 > nChunks) c1 = nChunks;
        }
    }
    const unsigned lane = threadIdx.x & 31;
    if (__any_sync(FULL, general) && lane == 0) atomicOr(modeFlags, 1 /* MODE_GENERAL */);
    // chunkJob: a job that covers a few chunk boundaries writes them itself, the warp shares the long ones
    const bool wide = c1 > c0 + 4;
    if (!wide) for (unsigned long long c = c0; c < c1; c++) chunkJob[c] = (uint32_t)j;
    for (unsigned m = __ballot_sync(FULL, wide); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const unsigned long long b0 = __shfl_sync(FULL, c0, src), b1 = __shfl_sync(FULL, c1, src);
        const uint32_t jj = (uint32_t)__shfl_sync(FULL, j, src);
        for (unsigned long long c = b0 + lane; c < b1; c += 32) chunkJob[c] = jj;
    }
    if (j > nJobs) return;
    if (j == nJobs) atomicOr(&headBits[total >> 5], 1u << (total & 31));
    uint4 *dst = reinterpret_cast<uint4 *>(info + j);
    dst[0] = make_uint4(o.tBaseW, o.qBaseW, o.tSize, o.qSize);
    dst[1] = make_uint4((uint32_t)o.clipStart, (uint32_t)o.clipEnd, o.delta, o.blockPtr);
}

// ------------------------------------------------------------------ pieces of the scoring kernel (gat_tiles.cuh)
#ifndef GAT_MIN_CTAS
#define GAT_MIN_CTAS 14     // 72 registers, 28 warps per SM: measured best (16 CTAs at 64 registers rematerialise too much)
#endif
constexpr int TILE = 32 * BPT;             // job-blocks per warp

__device__ __forceinline__ JobInfo loadInfo(const JobInfo *__restrict__ info, uint32_t j)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(info + j);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    JobInfo r;
    r.tBaseW = a.x; r.qBaseW = a.y; r.tSize = a.z; r.qSize = a.w;
    r.clipStart = (int)b.x; r.clipEnd = (int)b.y; r.delta = b.z; r.blockPtr = b.w;
    return r;
}

// SEARCH instantiations (work-lists with empty jobs: kent's NULL sub-chains, chain.c:535-539): the job of job-block v is
// looked up in the CSR -- the last job whose blockPtr is <= v, which is the one that owns v because the empty jobs in
// front of it share its blockPtr -- instead of counted from the job-start bitmap, which presumes that consecutive job
// starts belong to consecutive jobs.  info[nJobs] is the sentinel (blockPtr = total).
__device__ __noinline__ uint32_t searchJob(const JobInfo *__restrict__ info, unsigned long long nJobs, uint32_t v)
{
    uint32_t lo = 0, hi = (uint32_t)nJobs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&info[mid].blockPtr) <= v) lo = mid; else hi = mid;
    }
    return lo;
}
template <bool SEARCH>
__device__ __forceinline__ uint32_t jobOf(const ScoreParams &P, uint32_t counted, uint32_t v)
{
    return SEARCH ? searchJob(P.info, P.nJobs, v) : counted;
}

template <typename T> __device__ __forceinline__ long long finalLocal(const TupT<T> &t)
{   // a job's running score ends at max(c, d) (entered with 0) and its last peak test is still due
    return max64(0, max64(max64(widen(t.c), (long long)t.d), max64(widen(t.e), widen(t.f))));
}
__device__ __forceinline__ long long finalLocal(const Tup &t) { return max64(0, max64(max64(t.c, t.d), max64(t.e, t.f))); }

// Phase 3 of scoreTilesKernel for one warp (see gat_tiles.cuh), in 32- or 64-bit tuples.  Lane l holds job-blocks
// 4l..4l+3 of the warp: scores a[], gap costs g[] (the gap BEFORE the block), flags fl4 (one byte each).
template <typename T, bool SEARCH>
__device__ __forceinline__ void warpJobReduce(const ScoreParams &P, const int (&a)[BPT], const int (&g)[BPT], uint32_t fl4,
                                              uint32_t myWr, uint32_t myHw, uint32_t vb0, int warpV0, int vEnd,
                                              int warp, int lane, Tup *sWarpAgg, Tup *sWarpPend, int *sWarpHead,
                                              int *sWarpPendJob, int *sLastIsEnd, uint32_t *sLastJob)
{
    TupT<T> cur = tIdentity<T>();     // open segment at the end of my run
    bool runHasHead = false;
    bool pend = false;                // an END reached before any HEAD of my run: needs the carry
    TupT<T> pendTup = tIdentity<T>();
    uint32_t pendJob = 0;
#pragma unroll
    for (int k = 0; k < BPT; k++) {
        const uint32_t fl = (fl4 >> (8 * k)) & 0xffu;
        if (fl & 8) {
            const T av = (T)a[k];
            if (fl & 1) { cur = tIdentity<T>(); runHasHead = true; }
            if (fl & 5) {             // first block of the job, or the continuation of a split block: plain add
                cur.d += av; cur.c += av;
            } else {                  // peak test of the previous block, gap, clamp at 0, add (scoreChain.c:181-195)
                const T dY = av - (T)g[k];
                TupT<T> r;
                r.e = maxT<T>(cur.e, cur.d);
                r.f = maxT<T>(cur.f, cur.c);
                r.c = maxT<T>(av, cur.c + dY);
                r.d = cur.d + dY;
                cur = r;
            }
            const int v = warpV0 + BPT * lane + k;
            if (fl & 2) {
                const uint32_t job = jobOf<SEARCH>(P, myWr + __popc(myHw & (0xffffffffu >> (31 - (v & 31)))), vb0 + (uint32_t)v);
                if (runHasHead) {   // job lies inside my run: done
                    P.outGlobal[job] = (long long)cur.d;
                    P.outLocal[job] = finalLocal(cur);
                } else { pend = true; pendTup = cur; pendJob = job; }
                if (v + 1 == vEnd) { *sLastIsEnd = 1; *sLastJob = job; }
            } else if (v + 1 == vEnd) {          // the chunk's last valid job-block: its job runs on
                *sLastIsEnd = 0; *sLastJob = jobOf<SEARCH>(P, myWr + __popc(myHw & (0xffffffffu >> (31 - (v & 31)))), vb0 + (uint32_t)v);
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
            P.outLocal[pendJob] = finalLocal(fin);
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

// the 64-bit form is rare (a warp whose 128 blocks sum past 2^27): keep it out of the hot instruction stream
template <bool SEARCH>
__device__ __noinline__ void warpJobReduceWide(const ScoreParams &P, const int (&a)[BPT], const int (&g)[BPT], uint32_t fl4,
                                               uint32_t myWr, uint32_t myHw, uint32_t vb0, int warpV0, int vEnd,
                                               int warp, int lane, Tup *sWarpAgg, Tup *sWarpPend, int *sWarpHead,
                                               int *sWarpPendJob, int *sLastIsEnd, uint32_t *sLastJob)
{
    warpJobReduce<long long, SEARCH>(P, a, g, fl4, myWr, myHw, vb0, warpV0, vEnd, warp, lane, sWarpAgg, sWarpPend, sWarpHead,
                                     sWarpPendJob, sLastIsEnd, sLastJob);
}

// ------------------------------------------------------------------ cross-chunk fix-up
// A job that starts in chunk c and ends in chunk c' > c:  tail(c) + head(c+1) + ... + head(c').
// (nChunks = chunks this launch looks at: all of them, or those of the slices that have been scored so far.)
// One thread per chunk.  Most open tails end within a few chunks: the thread folds those itself.  A job
// that runs over many chunks is folded by the whole CTA, one contiguous slice of heads per thread.
constexpr int FIX_TPB = 256;
constexpr uint32_t FIX_SERIAL = 16;
__device__ __forceinline__ Tup orderedWarpFold(Tup x, int lane)
{
    for (int off = 1; off < 32; off <<= 1) {
        const Tup o = tupShfl(x, lane + off < 32 ? lane + off : lane);
        if (lane + off < 32) x = tupCombine(x, o);
    }
    return x;       // lane 0: the fold of all 32, in lane order
}
__device__ __forceinline__ Tup loadTup(const Tup *p)
{
    const longlong2 a = *reinterpret_cast<const longlong2 *>(p), b = *(reinterpret_cast<const longlong2 *>(p) + 1);
    return Tup{a.x, a.y, b.x, b.y};
}
// heads[first .. first+n) folded in order; four loads in flight at a time
__device__ __forceinline__ Tup foldSlice(const Tup *__restrict__ heads, uint32_t first, uint32_t n)
{
    Tup acc = tupIdentity();
    for (uint32_t i = 0; i < n; i += 4) {
        Tup t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = loadTup(heads + first + (i + k < n ? i + k : i));
#pragma unroll
        for (int k = 0; k < 4; k++) if (i + k < n) acc = tupCombine(acc, t[k]);
    }
    return acc;
}
__global__ void __launch_bounds__(FIX_TPB)
fixupKernel(const JobInfo *__restrict__ info, unsigned long long nJobs, unsigned long long total,
            const Tup *__restrict__ chunkHead, const Tup *__restrict__ chunkTail,
            const int *__restrict__ chunkTailJob, uint32_t nChunks,
            long long *__restrict__ outGlobal, long long *__restrict__ outLocal, Tup *__restrict__ outTuple, const int *__restrict__ err,
            uint32_t *__restrict__ headBits, uint32_t headBitsWords, int fatal, uint32_t endLo, uint32_t endHi, int clearBits)
{
    __shared__ uint32_t sLongC[FIX_TPB], sLongN[FIX_TPB];
    __shared__ int sLongJ[FIX_TPB];
    __shared__ int sNLong;
    __shared__ Tup sPart[FIX_TPB / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = blockIdx.x * FIX_TPB + tid;
    dependsWait();
    dependentsMayLaunch();
    // the job-start bitmap goes back to all-zero for the next pass (jobPrepKernel sets bits with atomicOr; the words
    // behind the last chunk hold the closing bit, the slack and the mode word)
    if (clearBits && c < nChunks) {
        uint4 *w = reinterpret_cast<uint4 *>(headBits + (size_t)c * (CHUNK / 32));
        w[0] = make_uint4(0u, 0u, 0u, 0u); w[1] = make_uint4(0u, 0u, 0u, 0u);
        if (c + 1 == nChunks)
            for (size_t k = (size_t)nChunks * (CHUNK / 32); k < headBitsWords; k++) headBits[k] = 0u;
    }
    if (*err & fatal) return;       // (lists with empty jobs scored by the SEARCH instantiation: ERR_EMPTYJOB is not an error)
    if (tid == 0) sNLong = 0;
    __syncthreads();
    int j = -1;
    uint32_t count = 0;                                 // heads to fold: chunks c+1 .. c+count
    if (c < nChunks) {
        j = chunkTailJob[c];
        if (j >= 0) {
            const uint32_t endChunk = (uint32_t)((info[j + 1].blockPtr - 1) / CHUNK);
            count = endChunk - c;
            // a work-list that arrives in slices is fixed up slice by slice: this launch finishes the jobs that end in chunks
            // [endLo, endHi) and leaves the others to the launch that covers their last chunk
            if (endChunk < endLo || endChunk >= endHi) j = -1;
        }
    }
    if (j >= 0) {
        if (count <= FIX_SERIAL) {
            const Tup all = tupCombine(loadTup(chunkTail + c), foldSlice(chunkHead, c + 1, count));
            outGlobal[j] = all.d;
            outLocal[j] = finalLocal(all);
            if (outTuple) outTuple[j] = all;        // the job as a map on the local-score state (parts of a chain split over GPUs)
        } else {
            const int k = atomicAdd(&sNLong, 1);
            sLongC[k] = c; sLongN[k] = count; sLongJ[k] = j;
        }
    }
    __syncthreads();
    const int nLong = sNLong;
    for (int e = 0; e < nLong; e++) {
        const uint32_t cc = sLongC[e], n = sLongN[e];
        const uint32_t per = (n + FIX_TPB - 1) / FIX_TPB;
        const uint32_t first = (uint32_t)tid * per;
        Tup mine = foldSlice(chunkHead, cc + 1 + first, first < n ? (n - first < per ? n - first : per) : 0u);
        mine = orderedWarpFold(mine, lane);
        if (lane == 0) sPart[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            Tup all = loadTup(chunkTail + cc);
            for (int w = 0; w < FIX_TPB / 32; w++) all = tupCombine(all, sPart[w]);
            outGlobal[sLongJ[e]] = all.d;
            outLocal[sLongJ[e]] = finalLocal(all);
            if (outTuple) outTuple[sLongJ[e]] = all;
        }
        __syncthreads();
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

// Reverse-complement copy of the query (the reference keeps one per '-' chromosome, scoreChain.c:123-149):
// sequence s of size n gets a second image at seqBase[nSeq + s] whose base p is the complement of forward
// base n-1-p, so '-' chains read forward in their own coordinates.  One thread per 32-base word of a copy.
__global__ void revCompKernel(const uint32_t *__restrict__ seqSize, const long long *__restrict__ seqBase,
                              const unsigned long long *__restrict__ seqWordStart, uint32_t nSeq,
                              unsigned long long totalWords, uint2 *__restrict__ planes, uint32_t *__restrict__ nplane,
                              uint32_t *__restrict__ nwin)
{
    unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= totalWords) return;
    uint32_t lo = 0, hi = nSeq;          // last sequence with seqWordStart <= w
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (seqWordStart[mid] <= w) lo = mid; else hi = mid;
    }
    const uint32_t s = lo;
    const long long wl = (long long)(w - seqWordStart[s]);
    const long long size = seqSize[s];
    // rc bases [32wl, 32wl+32) are forward bases (e-32 .. e-1] read downwards, e = size - 32wl
    const long long e = size - 32 * wl;                       // > 0
    const long long fBase = seqBase[s] + e - 32;              // forward padded coordinate of the window start (may precede the sequence)
    const long long fw = fBase >> 5;                          // floor: seqBase >= 1152
    const uint32_t sh = (uint32_t)(fBase & 31);
    const uint2 a = planes[fw], b = planes[fw + 1];
    const uint32_t hiW = __funnelshift_r(a.x, b.x, sh), loW = __funnelshift_r(a.y, b.y, sh);
    const uint32_t nW = __funnelshift_r(nplane[fw], nplane[fw + 1], sh);
    const uint32_t keep = e >= 32 ? 0xffffffffu : (0xffffffffu >> (32 - (int)e));   // rc positions that exist
    const unsigned long long n = (unsigned long long)(seqBase[nSeq + s] >> 5) + (unsigned long long)wl;
    planes[n] = make_uint2(~__brev(hiW) & keep, __brev(loW) & keep);
    const uint32_t rn = __brev(nW) & keep;
    nplane[n] = rn;
    if (rn) atomicOr(&nwin[(n >> 3) >> 5], 1u << ((n >> 3) & 31));
}

// ------------------------------------------------------------------ compact work-list -> records
// gat_score_compact: 6-byte delta-coded blocks become gat_block records.  One CTA per group of GAT_CGROUP records;
// a record's start is its predecessor's end plus the gap in front of it, unless it is flagged absolute (chain start,
// oversized gap) or opens the group (anchor): a segmented prefix sum, four records per thread, one warp scan, one
// pass over the warps.
constexpr int CX_TPB = GAT_CGROUP / 4;
struct CxSeg { int t, q; bool abs; };       // start of the last record seen (abs) or sum of the steps so far
__device__ __forceinline__ CxSeg cxCombine(const CxSeg &l, const CxSeg &r) { return r.abs ? r : CxSeg{l.t + r.t, l.q + r.q, l.abs}; }

// PACKED (gat_score_packed): the records are gat_pblock words (4 bytes: size 12 bits, JOINED, ABS, dt and dq 9 bits each) and
// an absolute record takes the next entry of the absolute table in list order: its index is absBase[group] + the number of
// absolute records in front of it in the group (one more prefix count), so no index is stored.
template <bool PACKED>
__global__ void __launch_bounds__(CX_TPB)
expandBlocksKernel(const void *__restrict__ recs, unsigned long long nBlocks, const gat_cabs *__restrict__ absTab,
                   unsigned long long nAbs, const gat_cabs *__restrict__ anchors, const uint32_t *__restrict__ absBase,
                   gat_block *__restrict__ out, unsigned firstGroup, int *__restrict__ err)
{
    __shared__ CxSeg sWarp[CX_TPB / 32];
    __shared__ uint32_t sWarpAbs[CX_TPB / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned group = blockIdx.x + firstGroup;                 // a launch may cover a slice of the groups
    const unsigned long long g0 = (unsigned long long)group * GAT_CGROUP, r0 = g0 + 4ull * tid;
    // my four records, and the size of the record in front of them (its end is where my first step starts)
    uint32_t size[4], dt[4], dq[4];     // size keeps the flags of gat_cblock (GAT_CBLOCK_ABS / JOINED) in both formats
    uint32_t prevSize = 0;
    unsigned long long absIx = 0;       // PACKED: index of my first absolute record
    if (PACKED) {
        const uint32_t *pb = static_cast<const uint32_t *>(recs);
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + 3 < nBlocks) w = __ldg(reinterpret_cast<const uint4 *>(pb + r0));      // groups start on multiples of 1024 records
        else { if (r0 < nBlocks) w.x = __ldg(pb + r0); if (r0 + 1 < nBlocks) w.y = __ldg(pb + r0 + 1); if (r0 + 2 < nBlocks) w.z = __ldg(pb + r0 + 2); }
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            size[k] = (ws[k] & GAT_PBLOCK_MAX_SIZE) | ((ws[k] & GAT_PBLOCK_JOINED) ? GAT_CBLOCK_JOINED : 0u) | ((ws[k] & GAT_PBLOCK_ABS) ? GAT_CBLOCK_ABS : 0u);
            dt[k] = (ws[k] >> 14) & 0x1ffu; dq[k] = ws[k] >> 23;
            mine += (ws[k] & GAT_PBLOCK_ABS) ? 1u : 0u;
        }
        if (tid > 0 && r0 - 1 < nBlocks) prevSize = __ldg(pb + r0 - 1) & GAT_PBLOCK_MAX_SIZE;
        uint32_t inc = mine;
        for (int off = 1; off < 32; off <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, off); if (lane >= off) inc += o; }
        if (lane == 31) sWarpAbs[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
        for (int w2 = 0; w2 < warp; w2++) before += sWarpAbs[w2];
        absIx = (unsigned long long)__ldg(absBase + group) + before + inc - mine;
    } else {
        const gat_cblock *cb = static_cast<const gat_cblock *>(recs);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in = r0 + k < nBlocks;
            const uint16_t *p = reinterpret_cast<const uint16_t *>(cb + (in ? r0 + k : 0));
            size[k] = in ? __ldg(p) : 0u; dt[k] = in ? __ldg(p + 1) : 0u; dq[k] = in ? __ldg(p + 2) : 0u;
        }
        if (tid > 0 && r0 - 1 < nBlocks) prevSize = __ldg(reinterpret_cast<const uint16_t *>(cb + r0 - 1)) & GAT_CBLOCK_MAX_SIZE;
    }
    // per record: step from the previous record's start (or an absolute start); running fold inside the thread
    CxSeg rec[4];
    CxSeg acc{0, 0, false};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        CxSeg e;
        const bool first = tid == 0 && k == 0;
        if (first) { const gat_cabs a = anchors[group]; e = CxSeg{a.tStart, a.qStart, true}; }
        else if (size[k] & GAT_CBLOCK_ABS) {
            const unsigned long long ix = PACKED ? absIx : ((unsigned long long)dt[k] | ((unsigned long long)dq[k] << 16));
            if (ix < nAbs) { const gat_cabs a = absTab[ix]; e = CxSeg{a.tStart, a.qStart, true}; }
            else { atomicOr(err, ERR_BLOCKIDX); e = CxSeg{0, 0, true}; }
        } else {
            const int before = (int)(k == 0 ? prevSize : (size[k - 1] & GAT_CBLOCK_MAX_SIZE));
            e = CxSeg{before + (int)dt[k], before + (int)dq[k], false};
        }
        if (PACKED && (size[k] & GAT_CBLOCK_ABS)) absIx++;        // (the group's first record skips its entry but owns one)
        acc = k == 0 ? e : cxCombine(acc, e);
        rec[k] = acc;                               // relative to whatever precedes the thread
    }
    // exclusive scan of the threads' folds: warp, then across warps
    CxSeg inc = acc;
    for (int off = 1; off < 32; off <<= 1) {
        CxSeg o{__shfl_up_sync(FULL, inc.t, off), __shfl_up_sync(FULL, inc.q, off), __shfl_up_sync(FULL, (int)inc.abs, off) != 0};
        if (lane >= off) inc = cxCombine(o, inc);
    }
    if (lane == 31) sWarp[warp] = inc;
    CxSeg carry{__shfl_up_sync(FULL, inc.t, 1), __shfl_up_sync(FULL, inc.q, 1), __shfl_up_sync(FULL, (int)inc.abs, 1) != 0};
    if (lane == 0) carry = CxSeg{0, 0, false};
    __syncthreads();
    CxSeg before{0, 0, false};
    for (int w = 0; w < warp; w++) before = cxCombine(before, sWarp[w]);
    carry = cxCombine(before, carry);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (r0 + k >= nBlocks) break;
        const CxSeg s = cxCombine(carry, rec[k]);   // absolute: the group's first record is
        uint32_t *o = reinterpret_cast<uint32_t *>(out + r0 + k);
        o[0] = (uint32_t)s.t; o[1] = (uint32_t)s.q;
        o[2] = (size[k] & GAT_CBLOCK_MAX_SIZE) | ((size[k] & GAT_CBLOCK_JOINED) ? GAT_BLOCK_JOINED : 0u);
    }
}

__global__ void expandJobsKernel(const gat_cjob *__restrict__ cj, unsigned long long nJobs, gat_job *__restrict__ out)
{
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nJobs) return;
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(cj + j));
    const uint32_t tSeq = v.y & 0xffffu, q = v.y >> 16;
    uint2 *o = reinterpret_cast<uint2 *>(out + j);
    o[0] = make_uint2(tSeq, (q & 0x7fffu) | ((q & GAT_CJOB_MINUS) ? GAT_QSEQ_MINUS : 0u));
    o[1] = make_uint2(v.x, v.x);                                        // firstBlock = blockPtr: whole chains
    o[2] = make_uint2((uint32_t)GAT_NO_CLIP_START, (uint32_t)GAT_NO_CLIP_END);
}

// ------------------------------------------------------------------ crossover of overlapping blocks
// cBlockFindCrossover (kent/src/lib/chainConnect.c:61-105).  With L[i], R[i] the scores of base i of the overlap in the
// left / right block, the reference starts from sum(R), adds L[i] - R[i] base by base and keeps the first strictly better
// position: pos = first argmax of the prefix sums P (0 if none is positive) and
// retScoreAdjustment = sum(R) + sum(L) - (sum(R) + max(0, max P)) = sum(L) - max(0, max P).
// One thread per pair for overlaps of up to XOVER_SHORT bases (what chainRemovePartialOverlaps mostly meets); longer
// overlaps are taken by the whole warp one after the other, lane = base: prefix sums by a warp scan, the first arg-max of
// a 32-base word by REDUX.MAX + ballot.
constexpr int XOVER_SHORT = 64;
struct XoverParams {
    const gat_xpair *pairs;
    unsigned long long nPairs;
    GenomeView t, q;            // q.seqBase holds forward images then reverse-complement images
    int matrix[16];             // [q][t], kent base codes
    int *pos, *adjust;
    int *err;
};

__device__ __forceinline__ void xoverWindow(const GenomeView &g, long long base, uint32_t &hi, uint32_t &lo, uint32_t &n)
{   // 32 bases starting at padded coordinate `base`
    const uint32_t w = (uint32_t)(base >> 5), sh = (uint32_t)(base & 31);
    loadWindow(g.planes, w, sh, hi, lo);
    n = loadNWindow(g.nplane, w, sh);
}
// score of base b of a window pair: 0 if either side is N (axt.c:431-454)
__device__ __forceinline__ int xoverScore(const int *matrix, uint32_t q1, uint32_t q0, uint32_t t1, uint32_t t0, uint32_t n, int b)
{
    const int code = (int)((((q1 >> b) & 1u) << 3) | (((q0 >> b) & 1u) << 2) | (((t1 >> b) & 1u) << 1) | ((t0 >> b) & 1u));
    return ((n >> b) & 1u) ? 0 : matrix[code];
}

__global__ void crossoverKernel(const __grid_constant__ XoverParams P)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool ok = i < P.nPairs;
    gat_xpair p{0, 0, 0, 0, 0, 0, 0};
    if (ok) p = P.pairs[i];
    const uint32_t qSeq = p.qSeq & 0x7fffffffu;
    long long lt = 0, lq = 0, rt = 0, rq = 0;
    const long long ov = p.overlap;
    if (ok) {
        if (p.tSeq >= P.t.nSeq || qSeq >= P.q.nSeq) { atomicOr(P.err, ERR_SEQ); ok = false; }
        else {
            const long long tSize = P.t.seqSize[p.tSeq], qSize = P.q.seqSize[qSeq];
            if (ov < 0 || p.leftTEnd - ov < 0 || p.leftQEnd - ov < 0 || p.leftTEnd > tSize || p.leftQEnd > qSize ||
                p.rightTStart < 0 || p.rightQStart < 0 || p.rightTStart + ov > tSize || p.rightQStart + ov > qSize) {
                atomicOr(P.err, ERR_COORD);
                ok = false;
            } else {
                const long long tBase = P.t.seqBase[p.tSeq], qBase = P.q.seqBase[(p.qSeq >> 31) ? P.q.nSeq + qSeq : qSeq];
                lt = tBase + p.leftTEnd - ov; lq = qBase + p.leftQEnd - ov; rt = tBase + p.rightTStart; rq = qBase + p.rightQStart;
            }
        }
    }
    if (ok && ov <= XOVER_SHORT) {      // my own loop
        long long sumL = 0, prefix = 0, best = 0;
        int bestPos = 0;
        for (long long off = 0; off < ov; off += 32) {
            uint32_t lt1, lt0, ltn, lq1, lq0, lqn, rt1, rt0, rtn, rq1, rq0, rqn;
            xoverWindow(P.t, lt + off, lt1, lt0, ltn);
            xoverWindow(P.q, lq + off, lq1, lq0, lqn);
            xoverWindow(P.t, rt + off, rt1, rt0, rtn);
            xoverWindow(P.q, rq + off, rq1, rq0, rqn);
            const int nb = ov - off < 32 ? (int)(ov - off) : 32;
            for (int b = 0; b < nb; b++) {
                const int L = xoverScore(P.matrix, lq1, lq0, lt1, lt0, ltn | lqn, b), R = xoverScore(P.matrix, rq1, rq0, rt1, rt0, rtn | rqn, b);
                sumL += L;
                prefix += L - R;
                if (prefix > best) { best = prefix; bestPos = (int)off + b + 1; }
            }
        }
        P.pos[i] = bestPos;
        P.adjust[i] = (int)(sumL - best);
    }
    // long overlaps, one after the other, the warp's lanes = the bases of a 32-base word
    for (unsigned m = __ballot_sync(FULL, ok && ov > XOVER_SHORT); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const long long wlt = shfl64(lt, src), wlq = shfl64(lq, src), wrt = shfl64(rt, src), wrq = shfl64(rq, src), wov = shfl64(ov, src);
        long long sumL = 0, base = 0, best = 0;
        long long bestPos = 0;
        for (long long off = 0; off < wov; off += 32) {
            uint32_t lt1, lt0, ltn, lq1, lq0, lqn, rt1, rt0, rtn, rq1, rq0, rqn;
            xoverWindow(P.t, wlt + off, lt1, lt0, ltn);
            xoverWindow(P.q, wlq + off, lq1, lq0, lqn);
            xoverWindow(P.t, wrt + off, rt1, rt0, rtn);
            xoverWindow(P.q, wrq + off, rq1, rq0, rqn);
            const bool live = off + lane < wov;
            const int L = live ? xoverScore(P.matrix, lq1, lq0, lt1, lt0, ltn | lqn, lane) : 0;
            const int R = live ? xoverScore(P.matrix, rq1, rq0, rt1, rt0, rtn | rqn, lane) : 0;
            int pre = L - R;
            pre = scanStep(pre, 1); pre = scanStep(pre, 2); pre = scanStep(pre, 4); pre = scanStep(pre, 8); pre = scanStep(pre, 16);
            const int wordMax = __reduce_max_sync(FULL, live ? pre : INT32_MIN);
            if (base + wordMax > best) {        // strictly better: the first position that reaches it
                best = base + wordMax;
                bestPos = off + __ffs(__ballot_sync(FULL, live && pre == wordMax));
            }
            base += __shfl_sync(FULL, pre, 31);
            sumL += __reduce_add_sync(FULL, L);
        }
        if (lane == src) { P.pos[i] = (int)bestPos; P.adjust[i] = (int)(sumL - best); }
    }
}

}  // namespace gat
