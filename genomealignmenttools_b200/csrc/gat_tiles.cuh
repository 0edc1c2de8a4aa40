// gat_tiles.cuh -- the scoring kernel (sm_100a): scoreTilesKernel.
//
// Same job as round 1's scoreChunksKernel (gone) -- kent chainCalcScore / chainScoreBlock
// (kent/src/lib/chainConnect.c:14-40), gapCalcCost (kent/src/lib/gapCalc.c:298-331), hillerlab
// chainCalcScoreLocal (src/scoreChain/scoreChain.c:176-198) and the clip of chainFastSubsetOnT
// (kent/src/lib/chain.c:510-522) -- and the same decomposition (a warp owns a tile of 128 job-blocks,
// a CTA a chunk of 256; first 32 bases lane = block, the rest as a list of 32-base items dealt to
// lanes; jobs reduced as max-plus tuples), rebuilt around what the round-1 profile showed: the old kernel
// was bound by integer-ALU issue, L1 data-pipe wavefronts (shared memory + shuffles) and dependent
// load chains, not by HBM.  What changed:
//
//  * the tile's 128 block records (1536 contiguous bytes whenever jobs tile the record array) arrive by
//    ONE bulk asynchronous copy per warp (cp.async.bulk + mbarrier, SASS UBLKCP) instead of 12 per-lane
//    loads that each depended on the job descriptor; the per-lane gather remains for work-lists whose
//    jobs share records (net fills, sub-chains);
//  * PLAIN work-lists (whole chains, no clip) read a 16-byte job descriptor and skip the clip arithmetic;
//  * every list slot is stored pre-biased (word index minus its position in the list), so an item is
//    `slot + i`: no per-item subtraction, no second shared-memory look-up;
//  * the slots that start inside a round of the item list come from one warp-wide OR (REDUX) of "my slot starts at
//    position p": no item-head bitmap in shared memory, no atomics;
//  * one dense gap table (small gaps included) and an N summary stored as overlapping word pairs: one
//    load and a funnel shift per test, no branches in front of the loads;
//  * all loads of a sub-tile (genome windows, N summary, gap cost) are issued before any is consumed;
//  * blocks of more than 1056 bases are streamed by the whole warp outside the list: no owner search, no scan.
#pragma once
#include "gat_kernels.cuh"

namespace gat {

// job-preparation verdicts about the whole list (one word behind the job-start bitmap)
constexpr int MODE_GENERAL = 1;     // some job clips or does not start at its own blockPtr: the PLAIN kernel must not run

// ------------------------------------------------------------------ async-copy plumbing (PTX)
__device__ __forceinline__ uint32_t smemU32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// x << r for r < 32, else 0
__device__ __forceinline__ uint32_t shlClamp(uint32_t x, uint32_t r)
{
    uint32_t out;
    asm("shl.b32 %0, %1, %2;" : "=r"(out) : "r"(x), "r"(r));
    return out;
}

// ------------------------------------------------------------------ explicit shared-memory accesses
// The kernel addresses its shared memory through 32-bit shared-window addresses and ld/st.shared: one base register
// per warp and immediate offsets, instead of generic pointers the compiler rebuilds from %cluster_ctaid.
__device__ __forceinline__ uint32_t lds(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.b8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Shared memory of one warp's tile, byte offsets from the warp's base.
//   PREV   the record in front of the tile, so that every block finds its predecessor at "my record - 12"
//   REC    phase 1: the tile's 128 gat_block records (the bulk copy lands here);
//   END    phase 2: running item-score sum at each list slot's last item (mod 2^32) -- same bytes as REC
//   SLOT   a block with more than 32 bases, as the item loop wants it: item i of the warp's list (i counts over the
//          whole tile) that belongs to the slot reads words tW0 + i / qW0 + i and has nEnd - 32 i bases left;
//          misc = tSh | qSh << 5 | block (index inside the tile) << 12
//   EX     list position of each slot's first item
//   GAP, SCORE, FLAG   per block: gap cost in front of it, its score, 1 head of job | 2 end of job | 4 continues the
//          previous record | 8 valid | 16 may contain N
constexpr uint32_t SM_PREV = 4, SM_REC = 16, SM_END = 16;
constexpr uint32_t SM_SLOT = SM_REC + TILE * 12;
constexpr uint32_t SM_EX = SM_SLOT + TILE * 16;
constexpr uint32_t SM_GAP = SM_EX + TILE * 4;
constexpr uint32_t SM_SCORE = SM_GAP + TILE * 4;
constexpr uint32_t SM_FLAG = SM_SCORE + TILE * 4;
constexpr uint32_t SM_BAR = SM_FLAG + TILE;
constexpr uint32_t SM_NLONG = SM_BAR + 8;           // long blocks of this tile (streamLongBlocks)
constexpr uint32_t SM_HEAD = SM_BAR + 16;            // this warp's 4 words of the job-start bitmap and the one behind them
constexpr uint32_t SM_RANK = SM_HEAD + 32;           // jobs that start in front of each of the 4 words (+ the chunk's first job)
constexpr uint32_t SM_WARP_BYTES = SM_RANK + 16;
static_assert(SM_SLOT % 16 == 0 && SM_GAP % 16 == 0 && SM_SCORE % 16 == 0 && SM_FLAG % 16 == 0 && SM_BAR % 8 == 0 && SM_WARP_BYTES % 16 == 0, "alignment");

// read-only global loads as volatile asm: they are issued where they are written (the compiler would sink them to
// their first use to save registers, which is exactly the latency the schedule wants to overlap)
__device__ __forceinline__ uint2 ldgPair(const uint2 *p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldgQuad(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldgWord(const int *p)
{
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// nn if the block [ts, ts + nn) x [qs, qs + nn) lies inside both sequences, is not longer than maxBases and `live`,
// else 0: one predicate chain (unsigned compares: a negative start is a huge one)
__device__ __forceinline__ uint32_t keepIfInside(uint32_t nn, uint32_t ts, uint32_t qs, uint32_t tSize, uint32_t qSize, uint32_t maxBases, bool live)
{
    uint32_t n;
    asm("{\n\t.reg .pred p;\n\t.reg .u32 rt, rq;\n\t"
        "sub.u32 rt, %4, %1;\n\tsub.u32 rq, %5, %1;\n\t"
        "setp.ne.u32 p, %7, 0;\n\t"
        "setp.le.and.u32 p, %1, %4, p;\n\tsetp.le.and.u32 p, %2, rt, p;\n\t"
        "setp.le.and.u32 p, %1, %5, p;\n\tsetp.le.and.u32 p, %3, rq, p;\n\t"
        "setp.le.and.u32 p, %1, %6, p;\n\t"
        "selp.u32 %0, %1, 0, p;\n\t}"
        : "=r"(n) : "r"(nn), "r"(ts), "r"(qs), "r"(tSize), "r"(qSize), "r"(maxBases), "r"((uint32_t)live));
    return n;
}
template <int N> struct IntC { static constexpr int value = N; __host__ __device__ constexpr operator int() const { return N; } };
// a sub-tile between "loads issued" and "loads consumed"
struct Front {
    uint2 ta, tb, qa, qb, tnw, qnw;
    int gap;
    uint32_t n, tW, qW;
    uint32_t misc;      // tSh | qSh << 5 | flags as stored << 16
};

// which blocks of a sub-tile broke a rule (cold)
__device__ __noinline__ void reportBlockErrors(int *err, bool live, uint32_t nn, int ts, int qs, uint32_t tSize, uint32_t qSize,
                                               uint32_t maxBases)
{
    if (!live || nn == 0u) return;
    const bool okCoord = nn <= tSize && (uint32_t)ts <= tSize - nn && nn <= qSize && (uint32_t)qs <= qSize - nn;
    if (!okCoord) atomicOr(err, ERR_COORD);
    else if (nn > maxBases) atomicOr(err, ERR_TOOLONG);
}

// gap beyond the dense table (cold)
__device__ __noinline__ int gapCostBeyond(const ScoreParams &P, int dq, int dt)
{
    const uint32_t v = (uint32_t)dq + (uint32_t)dt;   // dq, dt >= 0, one of them 0 unless both sequences gap
    if ((int)v < 0) return INT32_MIN;                 // dq+dt overflowed int (undefined in the reference)
    return gapCostExact(P.gap, P.gapSmall, P.gapLongPos, P.gapLongVal, dt == 0 ? 0 : (dq == 0 ? 1 : 2), (int)v);
}

// Phase 3 of scoreTilesKernel in 32 bits, for tiles whose block scores and gap costs all stay below 2^19 (every partial
// sum of the tile then stays below 2^27).  Lane l holds job-blocks 4l..4l+3 of the tile: scores a4, gap costs g4 (the gap
// BEFORE the block; 0 for a job's first block and for continuations), flags fl4 (one byte each; a position without a
// block has flags 0, score 0, gap 0 and passes through as a continuation).  Job scores are the max-plus tuples of Tup
// (gat_kernels.cuh), reduced in order: four blocks per lane, then one segmented warp scan in which "the nearest lane at
// or below me whose blocks contain a job start" comes from a ballot, so only the four tuple words are shuffled.
template <bool SEARCH>
__device__ __forceinline__ void warpJobReduce32(const ScoreParams &P, const uint4 a4, const uint4 g4, const uint32_t fl4,
                                                const uint32_t myWr, const uint32_t myHw, const uint32_t tileBase, const int lane, const uint32_t leMask,
                                                const int lastIdx, Tup *warpAgg, Tup *warpPend, int *warpHead, int *warpPendJob,
                                                int *sLastIsEnd, uint32_t *sLastJob)
{
    constexpr int NEGI = -(1 << 29);
    const int a[BPT] = {(int)a4.x, (int)a4.y, (int)a4.z, (int)a4.w};
    const int g[BPT] = {(int)g4.x, (int)g4.y, (int)g4.z, (int)g4.w};
    int d = 0, c = NEGI, e = NEGI, f = NEGI;      // the open job at the end of my blocks so far (since its start, or since my first block)
    bool seenHead = false, pend = false;          // pend: a job that started in front of my blocks ends inside them
    int pd = 0, pc = NEGI, pe = NEGI, pf = NEGI;
    uint32_t pendJob = 0;
    uint32_t job = myWr + __popc(myHw & (0xffffffffu >> (31 - ((4 * lane) & 31))));     // job of my first block
#pragma unroll
    for (int k = 0; k < BPT; k++) {
        const uint32_t fl = fl4 >> (8 * k);
        const bool head = (fl & 1u) != 0, end = (fl & 2u) != 0, plain = (fl & 12u) == 8u;     // plain: a block that opens with a gap
        if (k > 0) job += fl & 1u;
        const int av = a[k], dY = av - g[k];
        if (head) { d = 0; c = NEGI; e = NEGI; f = NEGI; seenHead = true; }
        // a gapped block: peak test of the block in front, gap, clamp at 0, add (scoreChain.c:181-195); else: plain add
        const int te = max(e, d), tf = max(f, c), tc = max(av, c + dY);
        e = plain && !head ? te : e;
        f = plain && !head ? tf : f;
        c = plain && !head ? tc : c + av;
        d += dY;
        if (end) {
            const uint32_t jb = jobOf<SEARCH>(P, job, tileBase + 4u * (uint32_t)lane + (uint32_t)k);
            if (seenHead) {             // the job lies inside my blocks: done
                P.outGlobal[jb] = (long long)d;
                P.outLocal[jb] = (long long)max(max(0, max(c, d)), max(e, f));
            } else { pend = true; pd = d; pc = c; pe = e; pf = f; pendJob = jb; }
        }
    }
    // the chunk's last job-block: does its job run on into the next chunk?  (the fix-up kernel wants to know)
    if ((lastIdx >> 2) == lane && lastIdx >= 0) {
        *sLastIsEnd = ((fl4 >> (8 * (lastIdx & 3))) & 2u) != 0;
        *sLastJob = jobOf<SEARCH>(P, myWr + __popc(myHw & (0xffffffffu >> (31 - (lastIdx & 31)))), tileBase + (uint32_t)lastIdx);
    }
    // segmented inclusive scan over the lanes: lane i takes in lanes down to the nearest one whose blocks contain a job start
    const uint32_t headMask = __ballot_sync(FULL, seenHead);
    const uint32_t below = headMask & leMask;
    const int dist = below ? __clz(below) - (31 - lane) : lane;       // lanes I may take in
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int od = __shfl_up_sync(FULL, d, off), oc = __shfl_up_sync(FULL, c, off);
        const int oe = __shfl_up_sync(FULL, e, off), of = __shfl_up_sync(FULL, f, off);
        if (off <= dist) {              // (od, oc, oe, of) first, then mine
            const int nf = max(max(of, f), oc + e), ne = max(oe, od + e), nc = max(c, oc + d);
            f = nf; e = ne; c = nc; d += od;
        }
    }
    int cd = __shfl_up_sync(FULL, d, 1), cc = __shfl_up_sync(FULL, c, 1), ce = __shfl_up_sync(FULL, e, 1), cf = __shfl_up_sync(FULL, f, 1);
    if (lane == 0) { cd = 0; cc = NEGI; ce = NEGI; cf = NEGI; }
    const bool carryHead = (headMask & (leMask >> 1)) != 0;     // the job that runs into my blocks started inside this warp
    if (pend) {
        const TupT<int> fin{cd + pd, max(pc, cc + pd), max(ce, cd + pe), max(max(cf, pf), cc + pe)};
        if (carryHead) {
            P.outGlobal[pendJob] = (long long)fin.d;
            P.outLocal[pendJob] = finalLocal(fin);
        } else { *warpPend = tWiden<int>(fin); *warpPendJob = (int)pendJob; }     // it started before this warp: at most one such lane
    }
    const bool anyCross = __any_sync(FULL, pend && !carryHead);
    if (lane == 31) {
        *warpAgg = tWiden<int>(TupT<int>{d, c, e, f}); *warpHead = headMask != 0;
        if (!anyCross) *warpPendJob = -1;
    }
}

// LONG instantiations: blocks of more than LONG_BASES bases are not listed.  After the item rounds the warp streams them
// one after the other, lane = word, 1024 bases per step and LONG_STEPS steps (4 loads per lane each) in flight; every lane
// keeps its own sum and one REDUX per block closes it: no owner search, no scan, no slot look-ups.  Descriptors (words of
// the block's 33rd base, bases behind the first window, shifts and block index) sit at the far end of the slot array.
//
// Why a template parameter and not one kernel: the per-block path (phase 1 unrolled four times) is as large as the SM's
// instruction cache takes -- with this loop behind it warps wait for instruction fetches (no_instruction 0.22 -> 1.84 per
// issue, +17 % kernel time at 67-base blocks, profiles/README.md).  So the LONG instantiations run phase 1 as a loop (+3 to 7 %
// at 67-base blocks) and stream long blocks, four steps in flight (9-kb blocks: 0.42 -> 0.67 of the HBM roofline; two steps: 0.58); the others list every block as
// before.  The host picks per work-list from a sample of the block sizes (gat_capi.cu: pickLong).
constexpr uint32_t LONG_BASES = 32 + 1024;
constexpr int LONG_STEPS = 4;           // steps (of 1024 bases) a warp keeps in flight per long block
template <bool SYM>
__device__ __forceinline__ void streamLongBlocks(const ScoreParams &P, uint32_t sm, int nLong, int lane, bool anyN)
{
    const uint2 *__restrict__ tPlanes = P.t.planes, *__restrict__ qPlanes = P.q.planes;
    for (int k = 0; k < nLong; k++) {
        const uint4 d = lds128(sm + SM_SLOT + 16u * (uint32_t)(TILE - 1 - k));
        const uint32_t v = (d.w >> 12) & 127u;
        const bool mayN = anyN && (lds8(sm + SM_FLAG + v) & 16u) != 0u;
        const int bases = (int)d.z;
        int acc = 0;
        for (int first = 0; first < bases; first += 1024 * LONG_STEPS) {         // LONG_STEPS steps of 1024 bases at a time
            uint2 w[LONG_STEPS][4];
#pragma unroll
            for (int j = 0; j < LONG_STEPS; j++) {
                const uint32_t word = (uint32_t)(first >> 5) + 32u * j + (uint32_t)lane;
                w[j][0] = w[j][1] = w[j][2] = w[j][3] = make_uint2(0u, 0u);
                if ((int)(word << 5) < bases) {
                    w[j][0] = ldgPair(tPlanes + d.x + word); w[j][1] = ldgPair(tPlanes + d.x + word + 1);
                    w[j][2] = ldgPair(qPlanes + d.y + word); w[j][3] = ldgPair(qPlanes + d.y + word + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < LONG_STEPS; j++) {
                if (first + 1024 * j < bases) {
                    const uint32_t word = (uint32_t)(first >> 5) + 32u * j + (uint32_t)lane;
                    const int left = bases - (int)(word << 5);
                    uint32_t m = shrOnes(32u - (uint32_t)(left > 32 ? 32 : (left < 0 ? 0 : left)));
                    if (mayN && left > 0) m &= nFreeMask(P.t.nplane, d.x + word, d.w & 31u, P.q.nplane, d.y + word, (d.w >> 5) & 31u);
                    acc += scoreWindow<SYM>(P.coef, __funnelshift_r(w[j][0].x, w[j][1].x, d.w), __funnelshift_r(w[j][0].y, w[j][1].y, d.w),
                                            __funnelshift_r(w[j][2].x, w[j][3].x, d.w >> 5), __funnelshift_r(w[j][2].y, w[j][3].y, d.w >> 5), m, __popc(m));
                }
            }
        }
        const int tot = __reduce_add_sync(FULL, acc);
        if (lane == 0) { const uint32_t a = sm + SM_SCORE + 4u * v; sts(a, lds(a) + (uint32_t)tot); }
    }
}

template <bool SYM, bool PLAIN, bool LONG, bool SEARCH>
__global__ void __launch_bounds__(TPB, GAT_MIN_CTAS)
scoreTilesKernel(const __grid_constant__ ScoreParams P)
{
    __shared__ __align__(16) unsigned char sRaw[WARPS * SM_WARP_BYTES];
    __shared__ Tup sWarpAgg[WARPS], sWarpPend[WARPS];
    __shared__ int sWarpHead[WARPS], sWarpPendJob[WARPS];
    __shared__ int sArrived, sLastIsEnd;
    __shared__ uint32_t sLastJob;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sArrived = 0;
    __syncthreads();                    // the only CTA-wide barrier: from here on warps run on their own
    const uint32_t chunk = blockIdx.x + P.chunkBase;
    const uint32_t vb0 = chunk * (uint32_t)CHUNK;
    const uint32_t total = (uint32_t)P.totalJobBlocks;
    const int vEnd = (int)(total - vb0 < (uint32_t)CHUNK ? total - vb0 : (uint32_t)CHUNK);   // valid job-blocks of this chunk
    const int warpV0 = warp * TILE;
    const int nValid = vEnd - warpV0 < 0 ? 0 : (vEnd - warpV0 > TILE ? TILE : vEnd - warpV0);  // of this warp's tile
    const uint32_t tileBase = vb0 + (uint32_t)warpV0;
    const uint32_t sm = smemU32(sRaw) + (uint32_t)warp * SM_WARP_BYTES;

    // ---- phase 0: start the record copy, then the chunk's words of the job-start bitmap, the job that owns the
    // chunk's first block, the verdicts of jobPrepKernel.
    int nLive = nValid;                 // job-blocks of this tile whose record exists
    if (PLAIN) {                        // record index = job-block index
        const long long left = (long long)P.nBlocks - (long long)tileBase;
        if (left < (long long)nValid) nLive = left < 0 ? 0 : (int)left;
        if (lane == 0 && nLive > 0) {
            mbarInit(sm + SM_BAR, 1);
            const uint32_t bytes = ((uint32_t)nLive * 12u + 15u) & ~15u;        // the record array has slack behind it
            mbarExpectTx(sm + SM_BAR, bytes);
            bulkLoad(sm + SM_REC, P.blocks + tileBase, bytes, sm + SM_BAR);
        }
    }
    if (LONG && lane == 0) sts(sm + SM_NLONG, 0u);
    // (up to here nothing was read that jobPrepKernel writes: the records are the caller's, or expandBlocksKernel's, which
    //  jobPrepKernel waited for)
    dependsWait();
    dependentsMayLaunch();
    constexpr int WORDS = CHUNK / 32;   // bitmap words per chunk; lane WORDS sees the next chunk's first word
    const uint32_t myHeadWord = __ldg(P.headBits + (size_t)chunk * WORDS + lane);
    const uint32_t j0 = __ldg(P.chunkJob + chunk);
    const int verdict = *reinterpret_cast<volatile const int *>(P.err);
    const int mode = P.modeFlags ? *reinterpret_cast<volatile const int *>(P.modeFlags) : (PLAIN ? 0 : MODE_GENERAL);
    if (PLAIN && lane == 0 && tileBase > 0 && nLive > 0) {      // the record in front of the tile's first block
        const gat_block pb = loadBlock(P.blocks, tileBase - 1);
        sts(sm + SM_PREV, (uint32_t)pb.tStart); sts(sm + SM_PREV + 4, (uint32_t)pb.qStart); sts(sm + SM_PREV + 8, pb.size);
    }
    // rejected by jobPrepKernel (or an empty job), or the other instantiation's list: leave (as a whole warp, and never
    // while a copy into this CTA's shared memory is in flight)
    if (__any_sync(FULL, (SEARCH ? verdict & ~ERR_EMPTYJOB : verdict) != 0 || ((mode & MODE_GENERAL) != 0) == PLAIN)) {
        if (PLAIN && nLive > 0) mbarWait(sm + SM_BAR, 0);
        return;
    }
    if (PLAIN && lane == 0 && nLive < nValid) atomicOr(P.err, ERR_BLOCKIDX);
    {   // rank[k] + popc(head word k & "lanes up to mine") = job of a block of sub-tile k: the jobs that start in the chunk's
        // words in front of word k, plus the chunk's first job, minus one if that job starts exactly at the chunk's first block
        const uint32_t pc = lane < WORDS ? __popc(myHeadWord) : 0u;
        uint32_t inc = pc;
#pragma unroll
        for (int off = 1; off < WORDS; off <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, off);
            if (lane >= off) inc += o;
        }
        const uint32_t rank = j0 - (__shfl_sync(FULL, myHeadWord, 0) & 1u) + inc - pc;
        const uint32_t k = (uint32_t)(lane - warp * BPT);
        if (k <= (uint32_t)BPT) sts(sm + SM_HEAD + 4u * k, myHeadWord);
        if (k < (uint32_t)BPT) sts(sm + SM_RANK + 4u * k, rank);
    }
    const uint32_t leMask = 0xffffffffu >> (31 - lane);

    if (!PLAIN) {
        // work-lists whose jobs point into shared records: every lane fetches its own record (record = job-block + delta)
        // into the tile's staging area, which the loop below reads like a bulk-copied tile.
        __syncwarp();
        int errAny = 0;
#pragma unroll 1
        for (int sub = 0; sub < BPT; sub++) {
            const uint32_t hw = lds(sm + SM_HEAD + 4u * (uint32_t)sub);
            const uint32_t job = jobOf<SEARCH>(P, lds(sm + SM_RANK + 4u * (uint32_t)sub) + __popc(hw & leMask), tileBase + 32u * (uint32_t)sub + (uint32_t)lane);
            const bool isHead = ((hw >> lane) & 1u) != 0;
            const int v = sub * 32 + lane;
            const bool valid = v < nValid;
            const JobInfo info = loadInfo(P.info, job);
            const uint32_t bi = tileBase + (uint32_t)v + info.delta;
            const bool ok = valid && (unsigned long long)bi < P.nBlocks;
            gat_block rec = loadBlock(P.blocks, ok ? bi : 0u);
            if (!ok) { rec.size = 0; rec.tStart = rec.qStart = 0; }
            if (valid && !ok) errAny |= ERR_BLOCKIDX;
            const uint32_t a = sm + SM_REC + 12u * (uint32_t)v;
            sts(a, (uint32_t)rec.tStart); sts(a + 4, (uint32_t)rec.qStart); sts(a + 8, rec.size);
            if (sub == 0 && lane == 0 && ok && !isHead && bi > 0) {     // the record in front of the tile's first block
                const gat_block pb = loadBlock(P.blocks, bi - 1);
                sts(sm + SM_PREV, (uint32_t)pb.tStart); sts(sm + SM_PREV + 4, (uint32_t)pb.qStart); sts(sm + SM_PREV + 8, pb.size);
            }
        }
        if (errAny) atomicOr(P.err, errAny);
    } else if (nLive > 0) {
        mbarWait(sm + SM_BAR, 0);
    }
    __syncwarp();

    // ---- phase 1: the tile's job-blocks, 32 at a time, lane = block.  Each sub-tile is handled in two halves: front()
    // reads the records, validates them and issues every global load the sub-tile needs (first window of both genomes,
    // N summaries, gap cost); back() builds the item list and only then consumes the loads, so the shuffle chain of the
    // list prefix runs under the memory latency.  Sub-tiles go one after the other (front s, back s; the next sub-tile's job
    // descriptor is fetched a sub-tile ahead): keeping two sub-tiles' loads in flight needs 80 registers and was slower
    // (profiles/README.md).  The four sub-tiles are unrolled (every shared-memory offset an immediate) unless the
    // instantiation streams long blocks (see streamLongBlocks).
    uint32_t seen = 0;                  // 1: some block of mine may contain N, 2: some block of mine is big (see back())
    int nSlots = 0;                     // blocks of this tile with more than 32 bases: they get a slot in the item list
    uint32_t itemBase = 0;              // items of the list so far
    const uint2 *__restrict__ tPlanes = P.t.planes, *__restrict__ qPlanes = P.q.planes;
    {
        const uint32_t recA = sm + SM_REC + 12u * (uint32_t)lane;     // my record of sub-tile 0
        const uint32_t outA = sm + 4u * (uint32_t)lane;
        const uint32_t flagA = sm + SM_FLAG + (uint32_t)lane;
        const uint32_t D = (uint32_t)P.gap.denseSize;
        uint32_t hwC = lds(sm + SM_HEAD);
        uint32_t jobC = jobOf<SEARCH>(P, lds(sm + SM_RANK) + __popc(hwC & leMask), tileBase + (uint32_t)lane);
        // (the descriptor array has slack behind it: lanes without a block read whatever index they compute)
        uint4 infoN = ldgQuad(reinterpret_cast<const uint4 *>(P.info + jobC));

        auto front = [&](auto subC) -> Front {          // IntC<n> (unrolled) or int (loop)
            const int sub = subC;
            Front F;
            const int v = sub * 32 + lane;
            const uint32_t hn = lds(sm + SM_HEAD + 4u * (uint32_t)(sub + 1));
            const uint4 ia = infoN;                     // tBaseW, qBaseW, tSize, qSize
            const uint32_t headBit = (hwC >> lane) & 1u, endBit = (__funnelshift_r(hwC, hn, 1) >> lane) & 1u;
            const bool live = v < nLive;
            int clipStart = 0, clipEnd = 0;
            if (!PLAIN) {
                const uint4 ib = ldgQuad(reinterpret_cast<const uint4 *>(P.info + jobC) + 1);
                clipStart = (int)ib.x; clipEnd = (int)ib.y;
            }
            hwC = hn;
            if (sub + 1 < BPT) {                        // the next sub-tile's job descriptor is fetched a sub-tile ahead
                jobC = jobOf<SEARCH>(P, lds(sm + SM_RANK + 4u * (uint32_t)(sub + 1)) + __popc(hn & leMask), tileBase + 32u * (uint32_t)(sub + 1) + (uint32_t)lane);
                infoN = ldgQuad(reinterpret_cast<const uint4 *>(P.info + jobC));
            }
            // my record and the one in front of it
            const uint32_t r0 = lds(recA + 384u * sub), r1 = lds(recA + 384u * sub + 4), r2 = lds(recA + 384u * sub + 8);
            const uint32_t p0 = lds(recA + 384u * sub - 12), p1 = lds(recA + 384u * sub - 8), p2 = lds(recA + 384u * sub - 4);
            int ts = (int)r0, qs = (int)r1;
            uint32_t nn = r2 & 0x7fffffffu;
            const uint32_t joinedBit = r2 >> 31;
            int pte = (int)p0 + (int)(p2 & 0x7fffffffu), pqe;
            if (!PLAIN) {                               // chainFastSubsetOnT clip (chain.c:513-522), of both records
                int te = ts + (int)nn;
                const int cut = clipStart > ts ? clipStart - ts : 0;
                ts += cut; qs += cut;
                te = te > clipEnd ? clipEnd : te;
                nn = te > ts ? (uint32_t)(te - ts) : 0u;
                pte = pte > clipEnd ? clipEnd : pte;
                pqe = (int)p1 + (pte - (int)p0);
            } else pqe = (int)p1 + (int)(p2 & 0x7fffffffu);
            // inside both sequences (unsigned: a negative start is a huge one) and short enough for 32-bit sums?
            const uint32_t n = keepIfInside(nn, (uint32_t)ts, (uint32_t)qs, ia.z, ia.w, P.maxBlockBases, live);
            if (__any_sync(FULL, live & (n != nn))) reportBlockErrors(P.err, live, nn, ts, qs, ia.z, ia.w, P.maxBlockBases);
            const uint32_t tSh = (uint32_t)ts & 31u, qSh = (uint32_t)qs & 31u;
            F.n = n;
            F.tW = n ? ia.x + ((uint32_t)ts >> 5) : 0u;       // blocks without bases read the front padding
            F.qW = n ? ia.y + ((uint32_t)qs >> 5) : 0u;
            F.ta = ldgPair(tPlanes + F.tW); F.tb = ldgPair(tPlanes + F.tW + 1);
            F.qa = ldgPair(qPlanes + F.qW); F.qb = ldgPair(qPlanes + F.qW + 1);
            F.tnw = ldgPair(P.t.nwin2 + (F.tW >> 8)); F.qnw = ldgPair(P.q.nwin2 + (F.qW >> 8));
            // the gap in front of the block (gapCalcCost, gapCalc.c:298-331): one dense table, the three gap kinds behind each other
            int dq = qs - pqe, dt = ts - pte;
            dq = dq < 0 ? 0 : dq; dt = dt < 0 ? 0 : dt;
            const uint32_t gv = (uint32_t)dq + (uint32_t)dt;                     // one of them is 0 unless both sequences gap
            const uint32_t goff = dt == 0 ? 0u : (dq == 0 ? D : 2u * D);
            const bool gapped = live & ((headBit | joinedBit) == 0u);
            F.gap = 0;
            if (gapped & (gv < D)) F.gap = ldgWord(P.gapDense + (goff + gv));
            if (__any_sync(FULL, gapped & (gv >= D))) {
                if (gapped & (gv >= D)) F.gap = gapCostBeyond(P, dq, dt);
            }
            F.misc = tSh | (qSh << 5) | ((live ? (8u | headBit | (endBit << 1) | ((joinedBit & ~headBit) << 2)) : 0u) << 16);
            return F;
        };

        auto back = [&](auto subC, const Front &F) {
            const int sub = subC;
            const int v = sub * 32 + lane;
            const uint32_t n = F.n;
            // what is left of the block joins the warp's item list (nothing here waits for the loads) -- unless the block is long:
            // those are streamed after the rounds (streamLongBlocks), their descriptors fill the slot array from its far end
            const bool isLong = LONG && n > LONG_BASES;
            const bool listed = n > 32u && !isLong;
            const uint32_t cnt = listed ? (n - 1u) >> 5 : 0u;
            uint32_t inc = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) inc = (uint32_t)scanStep((int)inc, off);
            const uint32_t lb = __ballot_sync(FULL, listed);
            if (listed) {
                const uint32_t e = itemBase + inc - cnt;
                const uint32_t slot = (uint32_t)nSlots + __popc(lb & (leMask >> 1));
                sts128(sm + SM_SLOT + 16u * slot, F.tW + 1u - e, F.qW + 1u - e, n - 32u + (e << 5), (F.misc & 0x3ffu) | ((uint32_t)v << 12));
                sts(sm + SM_EX + 4u * slot, e);
            }
            itemBase += __shfl_sync(FULL, inc, 31);
            nSlots += __popc(lb);
            const uint32_t lbLong = LONG ? __ballot_sync(FULL, isLong) : 0u;
            if (LONG && lbLong) {               // rare: the count lives in shared memory, not in a register of the hot path
                const uint32_t have = lds(sm + SM_NLONG);
                if (isLong)
                    sts128(sm + SM_SLOT + 16u * ((uint32_t)(TILE - 1) - have - __popc(lbLong & (leMask >> 1))), F.tW + 1u, F.qW + 1u, n - 32u,
                           (F.misc & 0x3ffu) | ((uint32_t)v << 12));
                __syncwarp();
                if (lane == 0) sts(sm + SM_NLONG, have + __popc(lbLong));
                __syncwarp();
            }
            // first 32 bases
            const uint32_t tSh = F.misc, qSh = F.misc >> 5;
            const uint32_t t1 = __funnelshift_r(F.ta.x, F.tb.x, tSh), t0 = __funnelshift_r(F.ta.y, F.tb.y, tSh);
            const uint32_t q1 = __funnelshift_r(F.qa.x, F.qb.x, qSh), q0 = __funnelshift_r(F.qa.y, F.qb.y, qSh);
            uint32_t vmask = shrOnes(32u - (n >= 32u ? 32u : n));
            // N: a block of n bases touches at most (n + 254) / 256 + 1 groups of 256 bases, whatever its offset: one mask for
            // both genomes over the summary bits from the block's first group on
            const uint32_t dMax = (n + 254u) >> 8;
            const uint32_t nbits = __funnelshift_r(F.tnw.x, F.tnw.y, F.tW >> 3) | __funnelshift_r(F.qnw.x, F.qnw.y, F.qW >> 3);
            const bool mayN = ((nbits & ~shlClamp(~1u, dMax)) != 0u) | (dMax > 31u);
            if (__any_sync(FULL, mayN)) {
                if (mayN) vmask &= nFreeMask(P.t.nplane, F.tW, tSh & 31u, P.q.nplane, F.qW, qSh & 31u);
            }
            sts(outA + SM_SCORE + 128u * sub, (uint32_t)scoreWindow<SYM>(P.coef, t1, t0, q1, q0, vmask, __popc(vmask)));
            sts(outA + SM_GAP + 128u * sub, (uint32_t)F.gap);
            sts8(flagA + 32u * sub, (F.misc >> 16) | (mayN ? 16u : 0u));
            // 2: the block's score or gap cost may reach 2^19: this tile's jobs are reduced in 64 bits
            seen |= (mayN ? 1u : 0u) | ((n > P.smallBases) | ((uint32_t)(F.gap + (1 << 19)) >= (1u << 20)) ? 2u : 0u);
        };

        if (LONG) {
#pragma unroll 1
            for (int sub = 0; sub < BPT; sub++) { const Front F = front(sub); back(sub, F); }
        } else {
            { const Front F = front(IntC<0>()); back(IntC<0>(), F); }
            { const Front F = front(IntC<1>()); back(IntC<1>(), F); }
            { const Front F = front(IntC<2>()); back(IntC<2>(), F); }
            { const Front F = front(IntC<3>()); back(IntC<3>(), F); }
        }
    }
    const bool anyN = __any_sync(FULL, (seen & 1u) != 0u), small = !__any_sync(FULL, (seen & 2u) != 0u);
    __syncwarp();

    // ---- phase 2: the item list, 32 items per round, adjacent lanes = adjacent words of a block (coalesced).  The slots
    // that start inside a round are the next ones in list order: lane L looks at slot before + L and one warp-wide OR
    // (REDUX) of "my slot starts at position p of this round" is the round's head mask, so the owner of an item is two
    // popcounts; no bitmap in shared memory, no atomics.  Scores leave the loop as a running prefix sum stored at each
    // slot's last item; a slot's sum is the difference of two such records.
    if (nSlots) {
        const int nRounds = (int)((itemBase + 31u) >> 5);
        int before = 0;                         // list slots that start before the round being fetched
        int sRun = 0;                           // sum of all item scores of earlier rounds (mod 2^32)
        // A round's state: owner slot, bases left in its block from this item on (<= 0: idle lane of the list's last
        // round, which lands on the last slot and reads words behind it: harmless, it scores 0), misc, four words.
#define GAT_FETCH(R, OW, LEFT, MISC, W0, W1, W2, W3)                                                        \
        {                                                                                                   \
            const uint32_t cand = (uint32_t)(before + lane);                                                \
            const uint32_t eL = cand < (uint32_t)nSlots ? lds(sm + SM_EX + 4u * cand) : 0xffffffffu;        \
            const uint32_t heads = __reduce_or_sync(FULL, shl1(eL - ((uint32_t)(R) << 5)));                 \
            OW = before - 1 + __popc(heads & leMask);                                                       \
            before += __popc(heads);                                                                        \
            const uint4 sl = lds128(sm + SM_SLOT + 16u * (uint32_t)OW);                                     \
            const uint32_t i = ((uint32_t)(R) << 5) + (uint32_t)lane;                                       \
            MISC = sl.w;                                                                                    \
            LEFT = (int)sl.z - (int)(i << 5);                                                               \
            const uint2 *tp = tPlanes + (sl.x + i);                                                         \
            const uint2 *qp = qPlanes + (sl.y + i);                                                         \
            W0 = ldgPair(tp); W1 = ldgPair(tp + 1); W2 = ldgPair(qp); W3 = ldgPair(qp + 1);                         \
        }
#define GAT_CONSUME(R, OW, LEFT, MISC, W0, W1, W2, W3)                                                      \
        {                                                                                                   \
            const uint32_t tSh = MISC, qSh = MISC >> 5;                                                     \
            const uint32_t t1 = __funnelshift_r(W0.x, W1.x, tSh), t0 = __funnelshift_r(W0.y, W1.y, tSh);    \
            const uint32_t q1 = __funnelshift_r(W2.x, W3.x, qSh), q0 = __funnelshift_r(W2.y, W3.y, qSh);    \
            uint32_t vmask = shrOnes(32u - (uint32_t)(LEFT > 32 ? 32 : (LEFT < 0 ? 0 : LEFT)));             \
            if (anyN) {                                                                                     \
                if (LEFT > 0 && (lds8(sm + SM_FLAG + (MISC >> 12)) & 16u)) {                                \
                    const uint4 sl = lds128(sm + SM_SLOT + 16u * (uint32_t)OW);                             \
                    const uint32_t i = ((uint32_t)(R) << 5) + (uint32_t)lane;                               \
                    vmask &= nFreeMask(P.t.nplane, sl.x + i, tSh & 31u, P.q.nplane, sl.y + i, qSh & 31u);   \
                }                                                                                           \
            }                                                                                               \
            int x = scoreWindow<SYM>(P.coef, t1, t0, q1, q0, vmask, __popc(vmask));                         \
            /* item scores leave the loop as a running prefix sum stored at each slot's last item (END) */  \
            x = scanStep(x, 1); x = scanStep(x, 2); x = scanStep(x, 4); x = scanStep(x, 8); x = scanStep(x, 16); \
            if ((uint32_t)(LEFT - 1) < 32u) sts(sm + SM_END + 4u * (uint32_t)OW, (uint32_t)(sRun + x));     \
            sRun += __shfl_sync(FULL, x, 31);                                                               \
        }
        int oA, lA, oB, lB; uint32_t mA, mB; uint2 a0, a1, a2, a3, b0, b1, b2, b3;
        GAT_FETCH(0, oA, lA, mA, a0, a1, a2, a3)
        for (int r = 0;; r += 2) {          // software pipeline: round r+1 is in flight while round r is scored
            if (r + 1 < nRounds) GAT_FETCH(r + 1, oB, lB, mB, b0, b1, b2, b3)
            GAT_CONSUME(r, oA, lA, mA, a0, a1, a2, a3)
            if (r + 1 >= nRounds) break;
            if (r + 2 < nRounds) GAT_FETCH(r + 2, oA, lA, mA, a0, a1, a2, a3)
            GAT_CONSUME(r + 1, oB, lB, mB, b0, b1, b2, b3)
            if (r + 2 >= nRounds) break;
        }
#undef GAT_FETCH
#undef GAT_CONSUME
        __syncwarp();
        // a slot's items summed to END[slot] - END[slot - 1]: add that to its block's score
        for (int s = lane; s < nSlots; s += 32) {
            const uint32_t d = lds(sm + SM_END + 4u * (uint32_t)s) - (s ? lds(sm + SM_END + 4u * (uint32_t)s - 4u) : 0u);
            const uint32_t a = sm + SM_SCORE + 4u * (lds(sm + SM_SLOT + 16u * (uint32_t)s + 12u) >> 12);
            sts(a, lds(a) + d);
        }
    }
    {
        const uint32_t nLong = LONG ? lds(sm + SM_NLONG) : 0u;
        if (nLong) streamLongBlocks<SYM>(P, sm, (int)nLong, lane, anyN);
    }
    __syncwarp();

    // ---- phase 3: ordered segmented reduction of tuples, per warp: lane l walks job-blocks 4l..4l+3 of the tile, one
    // warp scan joins the lanes; what crosses warps is resolved by whichever warp of the CTA finishes last.
    {
        const uint4 a4 = lds128(sm + SM_SCORE + 16u * (uint32_t)lane);
        const uint4 g4 = lds128(sm + SM_GAP + 16u * (uint32_t)lane);
        const uint32_t fl4 = lds(sm + SM_FLAG + 4u * (uint32_t)lane);
        // bitmap word of my four blocks and the job rank in front of it (job = rank + popc(word & lanes up to the block))
        const uint32_t myWr = lds(sm + SM_RANK + 4u * (uint32_t)(lane >> 3)), myHw = lds(sm + SM_HEAD + 4u * (uint32_t)(lane >> 3));
        if (small) {
            warpJobReduce32<SEARCH>(P, a4, g4, fl4, myWr, myHw, tileBase, lane, leMask, vEnd - 1 - warpV0,
                            &sWarpAgg[warp], &sWarpPend[warp], &sWarpHead[warp], &sWarpPendJob[warp], &sLastIsEnd, &sLastJob);
        } else {
            const int a[BPT] = {(int)a4.x, (int)a4.y, (int)a4.z, (int)a4.w};
            const int g[BPT] = {(int)g4.x, (int)g4.y, (int)g4.z, (int)g4.w};
            warpJobReduceWide<SEARCH>(P, a, g, fl4, myWr, myHw, vb0, warpV0, vEnd, warp, lane,
                              sWarpAgg, sWarpPend, sWarpHead, sWarpPendJob, &sLastIsEnd, &sLastJob);
        }
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
                P.outLocal[sWarpPendJob[w]] = finalLocal(fin);
            } else P.chunkHead[chunk] = fin;       // job began in an earlier chunk and ends here
        }
        if (sWarpHead[w]) { c = sWarpAgg[w]; ch = true; }
        else c = tupCombine(c, sWarpAgg[w]);
    }
    if (sLastIsEnd) P.chunkTailJob[chunk] = -1;
    else if (ch) { P.chunkTail[chunk] = c; P.chunkTailJob[chunk] = (int)sLastJob; }
    else { P.chunkHead[chunk] = c; P.chunkTailJob[chunk] = -1; }
}

// nwin2[i] = {nwin[i], nwin[i + 1]}: the N summary as overlapping word pairs (one load per test in the scoring kernel)
__global__ void nwinPairKernel(const uint32_t *__restrict__ nwin, size_t words, uint2 *__restrict__ nwin2)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) nwin2[i] = make_uint2(nwin[i], i + 1 < words ? nwin[i + 1] : 0u);
}

}  // namespace gat
