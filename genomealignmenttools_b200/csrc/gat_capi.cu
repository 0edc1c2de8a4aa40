// gat_capi.cu -- the C ABI of include/gat.h on top of the kernels in gat_kernels.cuh.
// No CPU scoring path exists in this library: every score comes out of scoreTilesKernel (gat_tiles.cuh).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "gat.h"
#include "gat_kernels.cuh"
#include "gat_tiles.cuh"
#include <stdlib.h>

using namespace gat;

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(GAT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

#ifndef GAT_COMPACT_SLICES
#define GAT_COMPACT_SLICES 8     // measured: 3 / 4 / 6 / 8 / 12 slices = 423 / 428 / 435 / 436 / 436 Gbp/s end to end at 10 M blocks
#endif
constexpr int COMPACT_SLICES = GAT_COMPACT_SLICES;
static_assert(GAT_CGROUP % gat::CHUNK == 0, "a slice of record groups must be a whole number of chunks");

struct GenomeDev {
    uint2 *planes = nullptr;
    uint32_t *nplane = nullptr, *nwin = nullptr;
    uint2 *nwin2 = nullptr;
    long long *seqBase = nullptr;
    uint32_t *seqSize = nullptr;
    uint32_t nSeq = 0;
    uint32_t words = 0;
    bool loaded = false;
    void release()
    {
        cudaFree(planes); cudaFree(nplane); cudaFree(nwin); cudaFree(nwin2); cudaFree(seqBase); cudaFree(seqSize);
        planes = nullptr; nplane = nwin = nullptr; nwin2 = nullptr; seqBase = nullptr; seqSize = nullptr; nSeq = 0; loaded = false;
    }
    GenomeView view() const { return GenomeView{planes, nplane, nwin, nwin2, (const int64_t *)seqBase, seqSize, nSeq, words}; }
};

struct gat_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    GenomeDev genome[2];
    // scoring
    bool scoringSet = false, sym = false;
    int coef[16];
    int matrix[16];                     // as given, [q][t]: the crossover kernel looks entries up
    GapView gap;
    int *gapSmall = nullptr, *gapLongPos = nullptr, *gapDense = nullptr;
    double *gapLongVal = nullptr;
    size_t dynSmem = 0;
    int *err = nullptr;
    // profiling
    bool profiling = false;
    cudaEvent_t ev[6];
    gat_stats stats;
    gat_worklist *scratch = nullptr;   // device buffers of gat_score(), grown on demand and reused
    void *xoverBuf = nullptr;          // device scratch of gat_crossover(): pairs in, positions and adjustments out
    size_t xoverCap = 0;
    void *compactBuf = nullptr;        // device staging of gat_score_compact(): jobs, blocks, abs, anchors
    size_t compactCap = 0;
    cudaStream_t copyStream = nullptr; // gat_score_compact(): slices are copied here while earlier ones are scored
    cudaEvent_t sliceEv[COMPACT_SLICES + 1] = {};
    cudaStream_t backStream = nullptr;  // ... and the scores of the jobs that end in a slice are copied back here meanwhile
    cudaEvent_t fixEv[COMPACT_SLICES] = {};
    uint32_t residentCtas = 0;         // scoring-kernel CTAs the device holds at once
    uint32_t maxBlockBases = GAT_MAX_BLOCK_BASES;   // longest record whose score surely fits 32 bits
    uint32_t smallBases = 1;
    bool pdl = true;                    // programmatic dependent launch of the pass's kernels (GAT_NO_PDL=1 turns it off)
    int forceLong = -1;                 // GAT_LONG_BLOCKS=stream / list in the environment at gat_create (tests, measurements): see pickLong
    std::vector<uint32_t> partJobs;    // gat_request_tuples(): jobs of the next scoring call whose tuple the caller wants
    gat_tuple *partOut = nullptr;
};

struct gat_worklist {
    uint64_t capJobs = 0, capBlocks = 0, capChunks = 0;   // allocated capacities (scratch reuse)
    gat_job *jobs = nullptr;
    JobInfo *info = nullptr;            // jobs as the scoring kernel reads them (jobPrepKernel)
    gat_block *blocks = nullptr;
    uint64_t nJobs = 0, totalJobBlocks = 0, nBlocks = 0;
    uint32_t nChunks = 0;
    uint32_t *chunkJob = nullptr;
    uint32_t *headBits = nullptr;       // job-start bitmap over job-blocks (jobPrepKernel), CHUNK/32 words per chunk + slack
    bool borrowedBlocks = false;        // blocks belong to another work-list (re-run without empty jobs)
    int plain = -1;                     // 1: whole chains tiling the record array (PLAIN kernel), 0: clips / shared records,
                                        // -1: not known on the host, jobPrepKernel decides (both instantiations are launched)
    bool searchJobs = false;            // the list has empty jobs: SEARCH instantiation (job of a block looked up in the CSR, gat_kernels.cuh)
    bool streamLong = true;             // LONG instantiation (blocks of more than 1056 bases are streamed, gat_tiles.cuh): pickLong()
    Tup *chunkHead = nullptr, *chunkTail = nullptr;
    int *chunkTailJob = nullptr;
    long long *outGlobal = nullptr, *outLocal = nullptr;
    Tup *outTuple = nullptr;            // per job, only when tuples were requested (written by the fix-up kernel)
    uint64_t capTuples = 0;
};

static void freeWorklistBuffers(gat_worklist *wl)
{
    cudaFree(wl->jobs); cudaFree(wl->info); if (!wl->borrowedBlocks) cudaFree(wl->blocks); cudaFree(wl->chunkJob); cudaFree(wl->headBits); cudaFree(wl->chunkHead);
    cudaFree(wl->chunkTail); cudaFree(wl->chunkTailJob); cudaFree(wl->outGlobal); cudaFree(wl->outLocal); cudaFree(wl->outTuple);
    wl->outTuple = nullptr; wl->capTuples = 0;
    wl->jobs = nullptr; wl->info = nullptr; wl->blocks = nullptr; wl->chunkJob = nullptr; wl->headBits = nullptr; wl->chunkHead = wl->chunkTail = nullptr;
    wl->chunkTailJob = nullptr; wl->outGlobal = wl->outLocal = nullptr;
    wl->capJobs = wl->capBlocks = wl->capChunks = 0;
}

extern "C" const char *gat_last_error(void) { return g_err; }

extern "C" int gat_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int gat_create(gat_ctx **out, int device, void *stream)
{
    if (!out) return fail(GAT_EINVAL, "gat_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(GAT_ECUDA, "gat_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(GAT_EINVAL, "gat_create: device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GAT_ECUDA, "gat_create: device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor);
    gat_ctx *ctx = new gat_ctx();
    ctx->device = device;
    if (const char *e = getenv("GAT_NO_PDL")) ctx->pdl = !(e[0] == '1');
    if (const char *e = getenv("GAT_LONG_BLOCKS")) {
        if (!strcmp(e, "stream")) ctx->forceLong = 1;
        else if (!strcmp(e, "list")) ctx->forceLong = 0;
    }
    for (auto &ev : ctx->ev) ev = nullptr;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    cudaError_t ce = cudaSuccess;
    if (stream) ctx->stream = (cudaStream_t)stream;
    else { ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); ctx->ownStream = ce == cudaSuccess; }
    if (ce == cudaSuccess) ce = cudaMalloc(&ctx->err, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(ctx->err, 0, sizeof(int));
    for (auto &ev : ctx->ev) if (ce == cudaSuccess) ce = cudaEventCreate(&ev);
    if (ce != cudaSuccess) {            // release whatever exists
        const bool haveStream = ctx->stream != nullptr;
        if (!haveStream) ctx->ownStream = false;
        gat_destroy(ctx);
        return fail(GAT_ECUDA, "gat_create failed: %s", cudaGetErrorString(ce));
    }
    *out = ctx;
    return GAT_OK;
}

extern "C" void gat_destroy(gat_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->genome[0].release();
    ctx->genome[1].release();
    if (ctx->scratch) { freeWorklistBuffers(ctx->scratch); delete ctx->scratch; }
    cudaFree(ctx->compactBuf);
    cudaFree(ctx->xoverBuf);
    if (ctx->copyStream) { cudaStreamDestroy(ctx->copyStream); for (auto &e : ctx->sliceEv) cudaEventDestroy(e); }
    if (ctx->backStream) { cudaStreamDestroy(ctx->backStream); for (auto &e : ctx->fixEv) cudaEventDestroy(e); }
    cudaFree(ctx->gapSmall); cudaFree(ctx->gapLongPos); cudaFree(ctx->gapLongVal); cudaFree(ctx->gapDense); cudaFree(ctx->err);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int gat_load_genome(gat_ctx *ctx, int side, const uint8_t *packed, uint64_t packedBytes,
                               const uint64_t *seqByteOffset, const uint32_t *seqSize, uint32_t nSeq,
                               const gat_nrun *nRuns, uint64_t nNRuns)
{
    if (!ctx || (side != GAT_TARGET && side != GAT_QUERY)) return fail(GAT_EINVAL, "gat_load_genome: bad ctx/side");
    if (nSeq == 0 || !seqByteOffset || !seqSize) return fail(GAT_EINVAL, "gat_load_genome: no sequences");
    if (packedBytes && !packed) return fail(GAT_EINVAL, "gat_load_genome: packed is NULL");
    CU(cudaSetDevice(ctx->device));
    GenomeDev &g = ctx->genome[side];
    g.release();

    // host-side layout: every sequence starts on a 128-base group boundary; the query is followed by a
    // reverse-complement image of each of its sequences (index nSeq + i), built on the device below
    const uint32_t nImages = side == GAT_QUERY ? 2 * nSeq : nSeq;
    std::vector<long long> base(nImages);
    std::vector<unsigned long long> wordStart(nSeq + 1);
    long long cursor = (long long)PAD_FRONT_GROUPS * GROUP_BASES;
    unsigned long long words = 0;
    for (uint32_t i = 0; i < nSeq; i++) {
        uint64_t bytes = ((uint64_t)seqSize[i] + 3) / 4;
        if (seqByteOffset[i] + bytes > packedBytes)
            return fail(GAT_EINVAL, "gat_load_genome: sequence %u payload runs past packedBytes", i);
        base[i] = cursor;
        wordStart[i] = words;
        words += ((unsigned long long)seqSize[i] + 31) / 32;
        cursor += (long long)(((unsigned long long)seqSize[i] + GROUP_BASES - 1) / GROUP_BASES) * GROUP_BASES;
    }
    wordStart[nSeq] = words;
    for (uint32_t i = nSeq; i < nImages; i++) {
        base[i] = cursor;
        cursor += (long long)(((unsigned long long)seqSize[i - nSeq] + GROUP_BASES - 1) / GROUP_BASES) * GROUP_BASES;
    }
    const long long totalBases = cursor + (long long)PAD_BACK_GROUPS * GROUP_BASES;
    if ((unsigned long long)totalBases >> 5 >= 0x7fffffffull)
        return fail(GAT_EINVAL, "gat_load_genome: genome of %lld bases exceeds the 2^36-base layout limit", totalBases);
    const size_t planeWords = (size_t)(totalBases / 32) * 2;    // one uint2 per 32 bases
    const size_t nWords = (size_t)(totalBases / 32) + 2;
    const size_t winWords = (size_t)((totalBases >> NWIN_SHIFT) / 32) + 2;

    // N runs: the reference stops at the first run that starts at/after the sequence end and clips
    // the others to it (twoBit.c:838-850)
    std::vector<gat_nrun> runs;
    runs.reserve(nNRuns);
    {
        uint32_t curSeq = 0xffffffffu;
        bool stopped = false;
        for (uint64_t i = 0; i < nNRuns; i++) {
            gat_nrun r = nRuns[i];
            if (r.seq >= nSeq) return fail(GAT_EINVAL, "gat_load_genome: N run %llu names sequence %u", (unsigned long long)i, r.seq);
            if (r.seq != curSeq) { curSeq = r.seq; stopped = false; }
            if (stopped) continue;
            if (r.start >= seqSize[r.seq]) { stopped = true; continue; }
            if ((uint64_t)r.start + r.len > seqSize[r.seq]) r.len = seqSize[r.seq] - r.start;
            if (r.len) runs.push_back(r);
        }
    }

    uint8_t *dRaw = nullptr;
    unsigned long long *dByteOff = nullptr, *dWordStart = nullptr;
    gat_nrun *dRuns = nullptr;
    auto cleanup = [&]() { cudaFree(dRaw); cudaFree(dByteOff); cudaFree(dWordStart); cudaFree(dRuns); };
#define CUG(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { cleanup(); g.release();                                            \
            return fail(GAT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); }               \
    } while (0)
    CUG(cudaMalloc(&g.planes, planeWords * 4));
    CUG(cudaMalloc(&g.nplane, nWords * 4));
    CUG(cudaMalloc(&g.nwin, winWords * 4));
    CUG(cudaMalloc(&g.nwin2, winWords * sizeof(uint2)));
    CUG(cudaMalloc(&g.seqBase, nImages * sizeof(long long)));
    CUG(cudaMalloc(&g.seqSize, nSeq * sizeof(uint32_t)));
    CUG(cudaMalloc(&dRaw, packedBytes + 16));
    CUG(cudaMalloc(&dByteOff, nSeq * sizeof(unsigned long long)));
    CUG(cudaMalloc(&dWordStart, (nSeq + 1) * sizeof(unsigned long long)));
    cudaStream_t st = ctx->stream;
    CUG(cudaMemsetAsync(g.planes, 0, planeWords * 4, st));
    CUG(cudaMemsetAsync(g.nplane, 0, nWords * 4, st));
    CUG(cudaMemsetAsync(g.nwin, 0, winWords * 4, st));
    CUG(cudaMemcpyAsync(dRaw, packed, packedBytes, cudaMemcpyHostToDevice, st));
    CUG(cudaMemcpyAsync(dByteOff, seqByteOffset, nSeq * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUG(cudaMemcpyAsync(dWordStart, wordStart.data(), (nSeq + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    CUG(cudaMemcpyAsync(g.seqBase, base.data(), nImages * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUG(cudaMemcpyAsync(g.seqSize, seqSize, nSeq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (words) {
        unsigned long long grid = (words + 255) / 256;
        repackKernel<<<(unsigned)grid, 256, 0, st>>>(dRaw, dByteOff, g.seqSize, g.seqBase, dWordStart, nSeq, words, g.planes);
        CUG(cudaGetLastError());
    }
    if (!runs.empty()) {
        CUG(cudaMalloc(&dRuns, runs.size() * sizeof(gat_nrun)));
        CUG(cudaMemcpyAsync(dRuns, runs.data(), runs.size() * sizeof(gat_nrun), cudaMemcpyHostToDevice, st));
        unsigned long long grid = (runs.size() * 32 + 255) / 256;
        nRunKernel<<<(unsigned)grid, 256, 0, st>>>(dRuns, runs.size(), g.seqBase, g.nplane, g.nwin);
        CUG(cudaGetLastError());
    }
    if (side == GAT_QUERY && words) {
        unsigned long long grid = (words + 255) / 256;
        revCompKernel<<<(unsigned)grid, 256, 0, st>>>(g.seqSize, g.seqBase, dWordStart, nSeq, words, g.planes, g.nplane, g.nwin);
        CUG(cudaGetLastError());
    }
    nwinPairKernel<<<(unsigned)((winWords + 255) / 256), 256, 0, st>>>(g.nwin, winWords, g.nwin2);
    CUG(cudaGetLastError());
    CUG(cudaStreamSynchronize(st));
#undef CUG
    cleanup();
    g.nSeq = nSeq;
    g.words = (uint32_t)(totalBases / 32);
    g.loaded = true;
    return GAT_OK;
}

// Is M the 6-value strand-symmetric form?  (see scoreWindow<true>)
static bool symmetricCoefs(const int32_t m[4][4], int coef[16])
{
    const int mTT = m[0][0], mCC = m[1][1], ts = m[0][1], tvTA = m[0][2], tvCG = m[1][3], tvX = m[0][3];
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) {
            int x1 = ((q ^ t) >> 1) & 1, x0 = (q ^ t) & 1, q0 = q & 1, want;
            if (!x1 && !x0) want = q0 ? mCC : mTT;
            else if (!x1) want = ts;
            else if (!x0) want = q0 ? tvCG : tvTA;
            else want = tvX;
            if (m[q][t] != want) return false;
        }
    coef[0] = mTT;
    coef[1] = tvTA - mTT;
    coef[2] = ts - mTT;
    coef[3] = tvX - tvTA - ts + mTT;
    coef[4] = mCC - mTT;
    coef[5] = tvCG - tvTA - mCC + mTT;
    for (int i = 6; i < 16; i++) coef[i] = 0;
    return true;
}

static void moebiusCoefs(const int32_t m[4][4], int coef[16])
{   // index bits: 8=q1 4=q0 2=t1 1=t0
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) coef[(q << 2) | t] = m[q][t];
    for (int bit = 1; bit < 16; bit <<= 1)
        for (int i = 0; i < 16; i++)
            if (i & bit) coef[i] -= coef[i ^ bit];
}

extern "C" int gat_set_scoring(gat_ctx *ctx, const gat_scoring *s)
{
    if (!ctx || !s) return fail(GAT_EINVAL, "gat_set_scoring: NULL argument");
    if (s->smallSize < 1 || s->smallSize > 8192 || s->longCount < 2 || s->longCount > 64)
        return fail(GAT_EINVAL, "gat_set_scoring: smallSize %d / longCount %d outside supported range (1..8192, 2..64)",
                    s->smallSize, s->longCount);
    if (!s->qSmall || !s->tSmall || !s->bSmall || !s->longPos || !s->qLong || !s->tLong || !s->bLong)
        return fail(GAT_EINVAL, "gat_set_scoring: NULL table");
    if (s->longPos[0] != s->smallSize)
        return fail(GAT_EINVAL, "gat_set_scoring: longPos[0] must equal smallSize (gapCalc.c:185-195)");
    for (int i = 1; i < s->longCount; i++)
        if (s->longPos[i] <= s->longPos[i - 1]) return fail(GAT_EINVAL, "gat_set_scoring: longPos not increasing");
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++)
            if (s->matrix[q][t] > (1 << 18) || s->matrix[q][t] < -(1 << 18))
                return fail(GAT_EINVAL, "gat_set_scoring: |matrix| above 2^18 would overflow the 32-bit sums of a GAT_SPLIT_BASES record");
    CU(cudaSetDevice(ctx->device));
    ctx->scoringSet = false;            // stays false if anything below fails: no scoring on half-built tables
    {   // block sums are 32-bit on the device: records may hold at most (2^31-1)/max|M| bases
        int maxAbs = 1;
        for (int q = 0; q < 4; q++)
            for (int t = 0; t < 4; t++) { const int a = s->matrix[q][t] < 0 ? -s->matrix[q][t] : s->matrix[q][t]; if (a > maxAbs) maxAbs = a; }
        const uint32_t lim = (uint32_t)(0x7fffffff / maxAbs);
        ctx->maxBlockBases = lim < GAT_MAX_BLOCK_BASES ? lim : GAT_MAX_BLOCK_BASES;
        ctx->smallBases = (uint32_t)((1 << 19) / maxAbs);
    }
    ctx->sym = symmetricCoefs(s->matrix, ctx->coef);
    if (!ctx->sym) moebiusCoefs(s->matrix, ctx->coef);
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) ctx->matrix[q * 4 + t] = s->matrix[q][t];
    const int S = s->smallSize, L = s->longCount;
    std::vector<int> small(3 * (size_t)S);
    std::vector<double> longVal(3 * (size_t)L);
    for (int i = 0; i < S; i++) { small[i] = s->qSmall[i]; small[S + i] = s->tSmall[i]; small[2 * S + i] = s->bSmall[i]; }
    for (int i = 0; i < L; i++) { longVal[i] = s->qLong[i]; longVal[L + i] = s->tLong[i]; longVal[2 * L + i] = s->bLong[i]; }
    ctx->gap.smallSize = S;
    ctx->gap.longCount = L;
    ctx->gap.lastPos = s->longPos[L - 1];
    for (int w = 0; w < 3; w++) {
        const double *v = &longVal[(size_t)w * L];
        ctx->gap.lastVal[w] = v[L - 1];
        // calcSlope, gapCalc.c:106-110, evaluated on the host in IEEE double like the reference
        volatile double dy = v[L - 1] - v[L - 2];
        volatile double dx = (double)s->longPos[L - 1] - (double)s->longPos[L - 2];
        ctx->gap.lastSlope[w] = dy / dx;
    }
    cudaFree(ctx->gapSmall); cudaFree(ctx->gapLongPos); cudaFree(ctx->gapLongVal); cudaFree(ctx->gapDense);
    ctx->gapSmall = ctx->gapLongPos = ctx->gapDense = nullptr; ctx->gapLongVal = nullptr;
    // dense cost table up to the last knot (capped at 4 M entries per gap kind = 48 MB)
    ctx->gap.denseSize = ctx->gap.lastPos < (1 << 22) ? ctx->gap.lastPos : (1 << 22);
    CU(cudaMalloc(&ctx->gapSmall, small.size() * sizeof(int)));
    CU(cudaMalloc(&ctx->gapLongPos, L * sizeof(int)));
    CU(cudaMalloc(&ctx->gapLongVal, longVal.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->gapSmall, small.data(), small.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->gapLongPos, s->longPos, L * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->gapLongVal, longVal.data(), longVal.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMalloc(&ctx->gapDense, (size_t)3 * ctx->gap.denseSize * sizeof(int)));
    {
        dim3 grid((unsigned)((ctx->gap.denseSize + 255) / 256), 3);
        gapDenseKernel<<<grid, 256, 0, ctx->stream>>>(ctx->gap, ctx->gapSmall, ctx->gapLongPos, ctx->gapLongVal, ctx->gapDense);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->dynSmem = 0;       // the small gap tables are read through L1
    {
        int perSm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, scoreTilesKernel<true, true, false, false>, TPB, ctx->dynSmem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        ctx->residentCtas = (uint32_t)(perSm > 0 ? perSm : 1) * (uint32_t)sms;
    }
    ctx->scoringSet = true;
    return GAT_OK;
}

// Words of the job-start bitmap for n chunks: the chunks' own, one chunk of look-ahead (is the block after a chunk's last
// a job start?), the closing bit, and the 32 words every warp of the last chunk reads (one per lane).
static size_t headWords(uint64_t nChunks) { return (size_t)(nChunks + 2) * (CHUNK / 32) + 32; }

// Size (or re-size) a work-list's device buffers.  Buffers only ever grow, so a scratch work-list
// that has seen the largest batch allocates nothing on later calls.
static int shapeWorklist(gat_ctx *ctx, gat_worklist *wl, uint64_t nJobs, uint64_t totalJobBlocks, uint64_t nBlocks)
{
    if (nJobs >= 0x7fffffffull || nBlocks > 0xffffffffull || totalJobBlocks > 0xffffffffull)
        return fail(GAT_EINVAL, "work-list too large for 32-bit indices (jobs %llu, blocks %llu, job-blocks %llu)",
                    (unsigned long long)nJobs, (unsigned long long)nBlocks, (unsigned long long)totalJobBlocks);
    const uint64_t nChunks = (totalJobBlocks + CHUNK - 1) / CHUNK;
    if (nJobs > wl->capJobs || nBlocks > wl->capBlocks || nChunks > wl->capChunks) {
        CU(cudaStreamSynchronize(ctx->stream));
        const uint64_t cj = nJobs > wl->capJobs ? nJobs + nJobs / 8 : wl->capJobs;
        const uint64_t cb = nBlocks > wl->capBlocks ? nBlocks + nBlocks / 8 : wl->capBlocks;
        const uint64_t cc = nChunks > wl->capChunks ? nChunks + nChunks / 8 : wl->capChunks;
        freeWorklistBuffers(wl);
        CU(cudaMalloc(&wl->jobs, (cj + 1) * sizeof(gat_job)));
        CU(cudaMalloc(&wl->info, (cj + 1 + 2 * CHUNK) * sizeof(JobInfo)));     // slack: lanes without a block read past the last job
        if (!wl->borrowedBlocks) CU(cudaMalloc(&wl->blocks, (cb + 3) * sizeof(gat_block)));   // slack: bulk copies round up to 16 bytes
        CU(cudaMalloc(&wl->chunkJob, (cc + 1) * sizeof(uint32_t)));
        CU(cudaMalloc(&wl->headBits, (headWords(cc) + 1) * sizeof(uint32_t)));     // + the mode word behind the bitmap
        CU(cudaMemsetAsync(wl->headBits, 0, (headWords(cc) + 1) * sizeof(uint32_t), ctx->stream));
        CU(cudaMalloc(&wl->chunkHead, (cc + 1) * sizeof(Tup)));
        CU(cudaMalloc(&wl->chunkTail, (cc + 1) * sizeof(Tup)));
        CU(cudaMalloc(&wl->chunkTailJob, (cc + 1) * sizeof(int)));
        CU(cudaMalloc(&wl->outGlobal, (cj + 1) * sizeof(long long)));
        CU(cudaMalloc(&wl->outLocal, (cj + 1) * sizeof(long long)));
        wl->capJobs = cj; wl->capBlocks = cb; wl->capChunks = cc;
    }
    wl->nJobs = nJobs; wl->totalJobBlocks = totalJobBlocks; wl->nBlocks = nBlocks;
    wl->nChunks = (uint32_t)nChunks;
    return GAT_OK;
}

static int allocWorklist(gat_ctx *ctx, uint64_t nJobs, uint64_t totalJobBlocks, uint64_t nBlocks, gat_worklist **out)
{
    gat_worklist *wl = new gat_worklist();
    *out = wl;
    return shapeWorklist(ctx, wl, nJobs, totalJobBlocks, nBlocks);
}

extern "C" void gat_worklist_destroy(gat_ctx *ctx, gat_worklist *wl)
{
    if (!wl) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    freeWorklistBuffers(wl);
    delete wl;
}

// Which instantiation scores the list: the one that streams long blocks pays 3 to 7 % on short-block lists and gains up
// to 60 % on long-block ones (gat_tiles.cuh, streamLongBlocks).  A sample of at most 4096 record sizes, evenly spread,
// decides: stream when blocks of more than 1056 bases hold at least a fifth of the sampled bases (about where the gain on
// the long blocks meets the cost on the short ones).  Both instantiations
// score every list exactly; this is about speed only.
template <typename Rec>
static bool pickLong(const Rec *recs, uint64_t n, uint32_t sizeMask)
{
    if (!recs || n == 0) return false;
    const uint64_t step = n > 4096 ? n / 4096 : 1;
    uint64_t all = 0, inLong = 0;
    for (uint64_t i = 0; i < n; i += step) {
        const uint32_t size = recs[i].size & sizeMask;
        all += size;
        if (size > 1056u) inLong += size;
    }
    return inLong * 5 >= all && inLong > 0;
}

static bool pickLongWords(const gat_pblock *recs, uint64_t n)
{
    if (!recs || n == 0) return false;
    const uint64_t step = n > 4096 ? n / 4096 : 1;
    uint64_t all = 0, inLong = 0;
    for (uint64_t i = 0; i < n; i += step) {
        const uint32_t size = recs[i] & GAT_PBLOCK_MAX_SIZE;
        all += size;
        if (size > 1056u) inLong += size;
    }
    return inLong * 5 >= all && inLong > 0;
}

static int uploadWorklist(gat_ctx *ctx, gat_worklist *wl, const gat_job *jobs, const gat_block *blocks)
{
    if (wl->nJobs) CU(cudaMemcpyAsync(wl->jobs, jobs, wl->nJobs * sizeof(gat_job), cudaMemcpyHostToDevice, ctx->stream));
    if (wl->nBlocks) CU(cudaMemcpyAsync(wl->blocks, blocks, wl->nBlocks * sizeof(gat_block), cudaMemcpyHostToDevice, ctx->stream));
    return GAT_OK;
}

extern "C" int gat_worklist_create(gat_ctx *ctx, const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks,
                                   const gat_block *blocks, uint64_t nBlocks, gat_worklist **out)
{
    if (!ctx || !out) return fail(GAT_EINVAL, "gat_worklist_create: NULL argument");
    if ((nJobs && !jobs) || (nBlocks && !blocks)) return fail(GAT_EINVAL, "gat_worklist_create: NULL array");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    gat_worklist *wl = nullptr;
    int rc = allocWorklist(ctx, nJobs, totalJobBlocks, nBlocks, &wl);
    if (rc == GAT_OK) {     // a resident list is looked at once: whole chains tiling the record array take the PLAIN kernel
        wl->plain = totalJobBlocks <= nBlocks ? 1 : 0;
        for (uint64_t j = 0; j < nJobs; j++) {
            if (jobs[j].firstBlock != jobs[j].blockPtr || jobs[j].clipStart != GAT_NO_CLIP_START || jobs[j].clipEnd != GAT_NO_CLIP_END) wl->plain = 0;
            if ((j + 1 < nJobs ? jobs[j + 1].blockPtr : totalJobBlocks) <= jobs[j].blockPtr) wl->searchJobs = true;      // an empty job
        }
    }
    if (rc == GAT_OK) wl->streamLong = ctx->forceLong >= 0 ? ctx->forceLong != 0 : pickLong(blocks, nBlocks, 0x7fffffffu);
    if (rc == GAT_OK) rc = uploadWorklist(ctx, wl, jobs, blocks);
    if (rc == GAT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(GAT_ECUDA, "work-list upload failed");
    if (rc != GAT_OK) { gat_worklist_destroy(ctx, wl); return rc; }
    *out = wl;
    return GAT_OK;
}

static ScoreParams scoreParams(gat_ctx *ctx, gat_worklist *wl)
{
    ScoreParams P;
    P.info = wl->info; P.blocks = wl->blocks;
    P.nJobs = wl->nJobs; P.totalJobBlocks = wl->totalJobBlocks; P.nBlocks = wl->nBlocks;
    P.chunkJob = wl->chunkJob; P.nChunks = wl->nChunks; P.headBits = wl->headBits; P.chunkBase = 0;
    P.maxBlockBases = ctx->maxBlockBases; P.smallBases = ctx->smallBases;
    P.t = ctx->genome[GAT_TARGET].view(); P.q = ctx->genome[GAT_QUERY].view();
    memcpy(P.coef, ctx->coef, sizeof P.coef);
    P.gap = ctx->gap;
    P.gapSmall = ctx->gapSmall; P.gapDense = ctx->gapDense; P.gapLongPos = ctx->gapLongPos; P.gapLongVal = ctx->gapLongVal;
    P.outGlobal = wl->outGlobal; P.outLocal = wl->outLocal;
    P.chunkHead = wl->chunkHead; P.chunkTail = wl->chunkTail; P.chunkTailJob = wl->chunkTailJob;
    P.err = ctx->err;
    P.modeFlags = wl->plain < 0 ? reinterpret_cast<const int *>(wl->headBits + headWords(wl->nChunks)) : nullptr;
    return P;
}

// gat_request_tuples(): the fix-up kernel also stores the tuple of every job it finishes; the buffer exists only then
static int ensureTupleBuffer(gat_ctx *ctx, gat_worklist *wl)
{
    if (ctx->partJobs.empty()) return GAT_OK;
    if (wl->capTuples < wl->nJobs) {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(wl->outTuple);
        wl->outTuple = nullptr; wl->capTuples = 0;
        CU(cudaMalloc(&wl->outTuple, (wl->nJobs + 1) * sizeof(Tup)));
        wl->capTuples = wl->nJobs;
    }
    for (uint32_t j : ctx->partJobs) {
        if (j >= wl->nJobs) return fail(GAT_EINVAL, "gat_request_tuples: job %u out of range", j);
        CU(cudaMemsetAsync(wl->outTuple + j, 0x80, sizeof(Tup), ctx->stream));     // recognisable if the fix-up kernel never writes it
    }
    return GAT_OK;
}

// bitmap reset + jobPrepKernel: needs the jobs only, not the block records
// The three kernels of a pass are launched with programmatic stream serialization (gat_kernels.cuh, dependsWait): each may
// be scheduled while the kernel in front of it drains and waits on the device for it to complete.  GAT_NO_PDL=1 in the
// environment at gat_create turns the attribute off (measurements).
template <typename... KArgs, typename... Args>
static cudaError_t launchDependent(gat_ctx *ctx, bool early, void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (ctx->pdl && early) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// (the job-start bitmap is all-zero here: zeroed when allocated, and fixupKernel clears what jobPrepKernel set)
static int launchPrep(gat_ctx *ctx, gat_worklist *wl, cudaStream_t st)
{
    const GenomeDev &t = ctx->genome[GAT_TARGET], &q = ctx->genome[GAT_QUERY];
    unsigned grid = (unsigned)((wl->nJobs + 1 + 255) / 256);
    CU(launchDependent(ctx, true, jobPrepKernel, grid, 256, st, wl->jobs, wl->nJobs, wl->totalJobBlocks, (const int64_t *)t.seqBase, t.seqSize, t.nSeq,
                       (const int64_t *)q.seqBase, q.seqSize, q.nSeq, wl->info, wl->chunkJob, wl->nChunks,
                       wl->headBits, reinterpret_cast<int *>(wl->headBits + headWords(wl->nChunks)),
                       wl->outGlobal, wl->outLocal, ctx->err));
    return GAT_OK;
}

// chunks [first, first + count)
// `plain`: 1 / 0 = the list's mode as the host knows it, -1 = jobPrepKernel's verdict decides on the device: both
// instantiations are launched and the one that does not apply returns at once.
// `early`: the kernel in front of this launch in the stream is jobPrepKernel (the scoring kernel copies its records before
// it waits for that kernel, so whatever wrote the records must have completed before jobPrepKernel did).
template <bool PLAIN>
static void launchTiles(gat_ctx *ctx, const ScoreParams &P, uint32_t count, bool streamLong, bool searchJobs, bool early, cudaStream_t st)
{
    if (searchJobs) {           // lists with empty jobs (rare): one instantiation per matrix kind
        if (ctx->sym) launchDependent(ctx, early, scoreTilesKernel<true, PLAIN, true, true>, count, TPB, st, P);
        else launchDependent(ctx, early, scoreTilesKernel<false, PLAIN, true, true>, count, TPB, st, P);
    } else if (ctx->sym) {
        if (streamLong) launchDependent(ctx, early, scoreTilesKernel<true, PLAIN, true, false>, count, TPB, st, P);
        else launchDependent(ctx, early, scoreTilesKernel<true, PLAIN, false, false>, count, TPB, st, P);
    } else {
        if (streamLong) launchDependent(ctx, early, scoreTilesKernel<false, PLAIN, true, false>, count, TPB, st, P);
        else launchDependent(ctx, early, scoreTilesKernel<false, PLAIN, false, false>, count, TPB, st, P);
    }
}

static int launchScoring(gat_ctx *ctx, ScoreParams P, uint32_t first, uint32_t count, int plain, bool streamLong, bool searchJobs, bool early, cudaStream_t st)
{
    if (count == 0) return 0;
    P.chunkBase = first;
    int launches = 0;
    if (plain != 0) { launchTiles<true>(ctx, P, count, streamLong, searchJobs, early, st); launches++; }
    if (plain <= 0) { launchTiles<false>(ctx, P, count, streamLong, searchJobs, early, st); launches++; }
    return launches;
}

// chunks [0, upTo) are looked at, jobs that end in chunks [endLo, endHi) are finished; `last`: also clear the job-start bitmap
static void launchFixupRange(gat_ctx *ctx, gat_worklist *wl, cudaStream_t st, uint32_t upTo, uint32_t endLo, uint32_t endHi, bool last)
{
    const uint32_t look = last ? wl->nChunks : upTo;
    launchDependent(ctx, true, fixupKernel, (look + FIX_TPB - 1) / FIX_TPB, FIX_TPB, st, wl->info, wl->nJobs, wl->totalJobBlocks, wl->chunkHead,
                    wl->chunkTail, wl->chunkTailJob, look, wl->outGlobal, wl->outLocal, ctx->partJobs.empty() ? nullptr : wl->outTuple,
                    ctx->err, wl->headBits, (uint32_t)(headWords(wl->nChunks) + 1), wl->searchJobs ? ~ERR_EMPTYJOB : ~0, endLo, endHi, last ? 1 : 0);
}

static void launchFixup(gat_ctx *ctx, gat_worklist *wl, cudaStream_t st) { launchFixupRange(ctx, wl, st, wl->nChunks, 0u, 0xffffffffu, true); }

extern "C" int gat_worklist_run(gat_ctx *ctx, gat_worklist *wl)
{
    if (!ctx || !wl) return fail(GAT_EINVAL, "gat_worklist_run: NULL argument");
    if (!ctx->genome[0].loaded || !ctx->genome[1].loaded) return fail(GAT_ESTATE, "gat_worklist_run: load both genomes first");
    if (!ctx->scoringSet) return fail(GAT_ESTATE, "gat_worklist_run: call gat_set_scoring first");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->stats.kernel_launches = 0;
    ctx->stats.chunks = wl->nChunks;
    ctx->stats.long_streamed = wl->streamLong ? 1u : 0u;
    if (wl->nJobs == 0) return GAT_OK;
    if (wl->nChunks == 0) {     // every job is empty: scores are 0
        CU(cudaMemsetAsync(wl->outGlobal, 0, wl->nJobs * sizeof(long long), st));
        CU(cudaMemsetAsync(wl->outLocal, 0, wl->nJobs * sizeof(long long), st));
        return GAT_OK;
    }
    int rc = ensureTupleBuffer(ctx, wl);
    if (rc != GAT_OK) return rc;
    ScoreParams P = scoreParams(ctx, wl);
    const bool prof = ctx->profiling;
    if (prof) CU(cudaEventRecord(ctx->ev[0], st));
    rc = launchPrep(ctx, wl, st);
    if (rc != GAT_OK) return rc;
    if (prof) CU(cudaEventRecord(ctx->ev[1], st));
    const int scoreLaunches = launchScoring(ctx, P, 0, wl->nChunks, wl->plain, wl->streamLong, wl->searchJobs, true, st);
    if (prof) CU(cudaEventRecord(ctx->ev[2], st));
    launchFixup(ctx, wl, st);
    if (prof) CU(cudaEventRecord(ctx->ev[3], st));
    CU(cudaGetLastError());
    ctx->stats.kernel_launches = 2 + scoreLaunches;
    return GAT_OK;
}

static int finishStats(gat_ctx *ctx)
{
    if (ctx->profiling && ctx->stats.kernel_launches) {
        CU(cudaEventSynchronize(ctx->ev[3]));
        CU(cudaEventElapsedTime(&ctx->stats.score_kernel_ms, ctx->ev[1], ctx->ev[2]));
        CU(cudaEventElapsedTime(&ctx->stats.all_kernels_ms, ctx->ev[0], ctx->ev[3]));
    }
    return GAT_OK;
}

static int readDeviceError(gat_ctx *ctx, int *err)
{
    *err = 0;
    CU(cudaMemcpyAsync(err, ctx->err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (*err) CU(cudaMemsetAsync(ctx->err, 0, sizeof(int), ctx->stream));
    return GAT_OK;
}

static int rejectWorklist(int err)
{
    return fail(GAT_EWORKLIST, "work-list rejected by the device:%s%s%s%s%s",
                (err & ERR_SEQ) ? " sequence index out of range;" : "",
                (err & ERR_BLOCKIDX) ? " block index out of range;" : "",
                (err & ERR_COORD) ? " block coordinates outside their sequence;" : "",
                (err & ERR_TOOLONG) ? " a record longer than gat_max_record_bases() (split it with GAT_BLOCK_JOINED);" : "",
                (err & ERR_CSR) ? " blockPtr is not a non-decreasing CSR row pointer starting at 0;" : "");
}

// The scoring kernel numbers jobs by counting job starts, which presumes that every job owns at least one job-block.
// Work-lists with empty jobs (a sub-chain that clips to nothing: kent's NULL sub-chain, score 0) take the SEARCH
// instantiation, which looks the job of a block up in the CSR.  gat_worklist_create sees them on the host; for a
// one-shot list jobPrepKernel reports them (ERR_EMPTYJOB: the ordinary kernel leaves at once), and the list, still on
// the device, is scored again here with the SEARCH instantiation: no copy back, no host filter, no buffers.
// scores of jobs [firstJob, nJobs) (the delta-coded calls copy the jobs of a slice back while later slices arrive)
static int fetchResults(gat_ctx *ctx, gat_worklist *wl, int64_t *global, int64_t *local, uint64_t firstJob)
{
    for (int pass = 0; pass < 2; pass++) {
        if (wl->nJobs > firstJob) {
            const uint64_t n = wl->nJobs - firstJob;
            if (global) CU(cudaMemcpyAsync(global + firstJob, wl->outGlobal + firstJob, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            if (local) CU(cudaMemcpyAsync(local + firstJob, wl->outLocal + firstJob, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        }
        const std::vector<uint32_t> &parts = ctx->partJobs;          // a request covers one scoring call
        if (!parts.empty()) {
            if (!wl->outTuple || wl->capTuples < wl->nJobs) { ctx->partJobs.clear(); return fail(GAT_ESTATE, "gat_request_tuples must precede the scoring call"); }
            for (size_t k = 0; k < parts.size(); k++)
                CU(cudaMemcpyAsync(ctx->partOut + k, wl->outTuple + parts[k], sizeof(gat_tuple), cudaMemcpyDeviceToHost, ctx->stream));
        }
        int err = 0;
        int rc = readDeviceError(ctx, &err);
        if (rc != GAT_OK) { ctx->partJobs.clear(); return rc; }
        if (wl->searchJobs) err &= ~ERR_EMPTYJOB;
        if (err & ~ERR_EMPTYJOB) { ctx->partJobs.clear(); return rejectWorklist(err); }
        if (err & ERR_EMPTYJOB) {           // first sight of an empty job in this list: once more, looking jobs up
            wl->searchJobs = true;
            firstJob = 0;
            rc = gat_worklist_run(ctx, wl);
            if (rc != GAT_OK) { ctx->partJobs.clear(); return rc; }
            continue;
        }
        break;
    }
    std::vector<uint32_t> parts;
    parts.swap(ctx->partJobs);
    for (size_t k = 0; k < parts.size(); k++)       // still the fill pattern: the job was not finished by the fix-up kernel
        if ((uint64_t)ctx->partOut[k].c == 0x8080808080808080ull)
            return fail(GAT_EINVAL, "gat_request_tuples: job %u owns fewer than GAT_TUPLE_MIN_BLOCKS job-blocks", parts[k]);
    return finishStats(ctx);
}

extern "C" int gat_worklist_results(gat_ctx *ctx, gat_worklist *wl, int64_t *global, int64_t *local)
{
    if (!ctx || !wl) return fail(GAT_EINVAL, "gat_worklist_results: NULL argument");
    CU(cudaSetDevice(ctx->device));
    return fetchResults(ctx, wl, global, local, 0);
}

extern "C" int gat_score(gat_ctx *ctx, const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks,
                         const gat_block *blocks, uint64_t nBlocks, int64_t *global, int64_t *local)
{
    if (!ctx) return fail(GAT_EINVAL, "gat_score: NULL ctx");
    if ((nJobs && (!jobs || !global || !local)) || (nBlocks && !blocks)) return fail(GAT_EINVAL, "gat_score: NULL array");
    if (!ctx->genome[0].loaded || !ctx->genome[1].loaded) return fail(GAT_ESTATE, "gat_score: load both genomes first");
    if (!ctx->scoringSet) return fail(GAT_ESTATE, "gat_score: call gat_set_scoring first");
    if (nJobs == 0) return GAT_OK;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->scratch) ctx->scratch = new gat_worklist();
    gat_worklist *wl = ctx->scratch;
    int rc = shapeWorklist(ctx, wl, nJobs, totalJobBlocks, nBlocks);
    wl->plain = -1;
    wl->searchJobs = false;
    wl->streamLong = ctx->forceLong >= 0 ? ctx->forceLong != 0 : pickLong(blocks, nBlocks, 0x7fffffffu);
    cudaStream_t st = ctx->stream;
    const bool prof = ctx->profiling;
    if (rc == GAT_OK && prof) cudaEventRecord(ctx->ev[4], st);
    if (rc == GAT_OK) rc = uploadWorklist(ctx, wl, jobs, blocks);
    if (rc == GAT_OK) rc = gat_worklist_run(ctx, wl);
    if (rc == GAT_OK && prof) cudaEventRecord(ctx->ev[5], st);
    if (rc == GAT_OK) rc = gat_worklist_results(ctx, wl, global, local);
    if (rc == GAT_OK && prof) {
        float total = 0;
        cudaEventElapsedTime(&total, ctx->ev[4], ctx->ev[0]);
        ctx->stats.h2d_ms = total;
        ctx->stats.h2d_bytes = nJobs * sizeof(gat_job) + nBlocks * sizeof(gat_block);
        ctx->stats.d2h_bytes = 2 * nJobs * sizeof(long long);
    }
    return rc;
}

// gat_score_compact (6-byte gat_cblock records, absBase == NULL) and gat_score_packed (4-byte gat_pblock words)
static int scoreDeltaCoded(gat_ctx *ctx, const char *who, const gat_cjob *jobs, uint64_t nJobs, const void *blocksV, size_t recBytes, uint64_t nBlocks,
                           const gat_cabs *abs, uint64_t nAbs, const gat_cabs *anchors, const uint32_t *absBase, int64_t *global, int64_t *local)
{
    const bool packed = recBytes == sizeof(gat_pblock);
    const char *blocks = static_cast<const char *>(blocksV);
    if (!ctx) return fail(GAT_EINVAL, "%s: NULL ctx", who);
    if ((nJobs && (!jobs || !global || !local)) || (nBlocks && (!blocks || !anchors || (packed && !absBase))) || (nAbs && !abs))
        return fail(GAT_EINVAL, "%s: NULL array", who);
    if (!ctx->genome[0].loaded || !ctx->genome[1].loaded) return fail(GAT_ESTATE, "%s: load both genomes first", who);
    if (!ctx->scoringSet) return fail(GAT_ESTATE, "%s: call gat_set_scoring first", who);
    if (nJobs == 0) return GAT_OK;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->scratch) ctx->scratch = new gat_worklist();
    gat_worklist *wl = ctx->scratch;
    int rc = shapeWorklist(ctx, wl, nJobs, nBlocks, nBlocks);
    if (rc != GAT_OK) return rc;
    wl->plain = 1;              // whole chains by construction
    wl->searchJobs = false;
    wl->streamLong = ctx->forceLong >= 0 ? ctx->forceLong != 0
                     : packed ? pickLongWords(reinterpret_cast<const gat_pblock *>(blocks), nBlocks)
                              : pickLong(reinterpret_cast<const gat_cblock *>(blocks), nBlocks, GAT_CBLOCK_MAX_SIZE);
    cudaStream_t st = ctx->stream;
    const uint64_t nGroups = (nBlocks + GAT_CGROUP - 1) / GAT_CGROUP;
    auto up8 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t oJobs = 0, oBlocks = oJobs + up8(nJobs * sizeof(gat_cjob)), oAbs = oBlocks + up8(nBlocks * recBytes + 8),
                 oAnch = oAbs + up8(nAbs * sizeof(gat_cabs)), oBase = oAnch + up8(nGroups * sizeof(gat_cabs)),
                 need = oBase + up8(packed ? nGroups * sizeof(uint32_t) : 0) + 16;
    if (need > ctx->compactCap) {
        CU(cudaStreamSynchronize(st));
        cudaFree(ctx->compactBuf);
        ctx->compactBuf = nullptr; ctx->compactCap = 0;
        CU(cudaMalloc(&ctx->compactBuf, need + need / 8));
        ctx->compactCap = need + need / 8;
    }
    char *base = static_cast<char *>(ctx->compactBuf);
    const gat_cjob *dJobs = reinterpret_cast<const gat_cjob *>(base + oJobs);
    const void *dBlocks = base + oBlocks;
    const gat_cabs *dAbs = reinterpret_cast<const gat_cabs *>(base + oAbs), *dAnch = reinterpret_cast<const gat_cabs *>(base + oAnch);
    const uint32_t *dBase = packed ? reinterpret_cast<const uint32_t *>(base + oBase) : nullptr;
    auto expand = [&](unsigned groups, unsigned firstGroup) {
        if (packed) expandBlocksKernel<true><<<groups, CX_TPB, 0, st>>>(dBlocks, nBlocks, dAbs, nAbs, dAnch, dBase, wl->blocks, firstGroup, ctx->err);
        else expandBlocksKernel<false><<<groups, CX_TPB, 0, st>>>(dBlocks, nBlocks, dAbs, nAbs, dAnch, dBase, wl->blocks, firstGroup, ctx->err);
    };
    uint64_t doneJobs = 0;              // jobs whose scores are already on their way back (sliced lists)
    bool fixedUp = false;
    const bool prof = ctx->profiling;
    ctx->stats.chunks = wl->nChunks;
    ctx->stats.long_streamed = wl->streamLong ? 1u : 0u;
    // Big lists go over in slices on a copy stream while the slices that have arrived are expanded and scored: a group
    // of GAT_CGROUP records expands on its own and a chunk needs no record beyond its own.  (With profiling on, one
    // slice on one stream, so that the events of gat_get_stats mean what they say.)
    const uint64_t slices = (!prof && wl->nChunks && nBlocks >= (1u << 20)) ? COMPACT_SLICES : 1;
    if (slices > 1 && !ctx->copyStream) {
        CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        for (auto &e : ctx->sliceEv) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&ctx->backStream, cudaStreamNonBlocking));
        for (auto &e : ctx->fixEv) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t cp = slices > 1 ? ctx->copyStream : st;
    if (prof) cudaEventRecord(ctx->ev[4], st);
    CU(cudaMemcpyAsync(base + oJobs, jobs, nJobs * sizeof(gat_cjob), cudaMemcpyHostToDevice, cp));
    if (nAbs) CU(cudaMemcpyAsync(base + oAbs, abs, nAbs * sizeof(gat_cabs), cudaMemcpyHostToDevice, cp));
    if (nGroups) CU(cudaMemcpyAsync(base + oAnch, anchors, nGroups * sizeof(gat_cabs), cudaMemcpyHostToDevice, cp));
    if (packed && nGroups) CU(cudaMemcpyAsync(base + oBase, absBase, nGroups * sizeof(uint32_t), cudaMemcpyHostToDevice, cp));
    if (slices > 1) { CU(cudaEventRecord(ctx->sliceEv[COMPACT_SLICES], cp)); CU(cudaStreamWaitEvent(st, ctx->sliceEv[COMPACT_SLICES], 0)); }
    expandJobsKernel<<<(unsigned)((nJobs + 255) / 256), 256, 0, st>>>(dJobs, nJobs, wl->jobs);
    ctx->stats.kernel_launches = 0;
    if (wl->nChunks == 0) {     // every job is empty: scores are 0
        CU(cudaMemsetAsync(wl->outGlobal, 0, nJobs * sizeof(long long), st));
        CU(cudaMemsetAsync(wl->outLocal, 0, nJobs * sizeof(long long), st));
    } else {
        rc = ensureTupleBuffer(ctx, wl);
        if (rc != GAT_OK) return rc;
        const ScoreParams P = scoreParams(ctx, wl);
        if (slices == 1) {
            CU(cudaMemcpyAsync(base + oBlocks, blocks, nBlocks * recBytes, cudaMemcpyHostToDevice, st));
            expand((unsigned)nGroups, 0);
            if (prof) CU(cudaEventRecord(ctx->ev[0], st));
            rc = launchPrep(ctx, wl, st);
            if (rc != GAT_OK) return rc;
            if (prof) CU(cudaEventRecord(ctx->ev[1], st));
            launchScoring(ctx, P, 0, wl->nChunks, wl->plain, wl->streamLong, wl->searchJobs, true, st);
            if (prof) CU(cudaEventRecord(ctx->ev[2], st));
        } else {
            rc = launchPrep(ctx, wl, st);
            if (rc != GAT_OK) return rc;
            const uint64_t groupsPerSlice = (nGroups + slices - 1) / slices;
            uint32_t sliceLaunches = 0;
            // The fix-up runs per slice and finishes the jobs that end in it, so their scores go back over PCIe (the other
            // direction, a stream of its own) while later slices still arrive -- when the caller's arrays are pinned: a copy
            // into pageable memory would hold this thread up instead.
            cudaPointerAttributes pa;
            const bool early = global && local && cudaPointerGetAttributes(&pa, global) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                               cudaPointerGetAttributes(&pa, local) == cudaSuccess && pa.type == cudaMemoryTypeHost;
            cudaGetLastError();
            uint32_t fixedChunks = 0;
            for (uint64_t s = 0; s < slices; s++) {
                const uint64_t g0 = s * groupsPerSlice, g1 = std::min<uint64_t>(nGroups, g0 + groupsPerSlice);
                if (g0 >= g1) break;
                const uint64_t r0 = g0 * GAT_CGROUP, r1 = std::min<uint64_t>(nBlocks, g1 * GAT_CGROUP);
                CU(cudaMemcpyAsync(base + oBlocks + r0 * recBytes, blocks + r0 * recBytes, (r1 - r0) * recBytes, cudaMemcpyHostToDevice, cp));
                CU(cudaEventRecord(ctx->sliceEv[s], cp));
                CU(cudaStreamWaitEvent(st, ctx->sliceEv[s], 0));
                expand((unsigned)(g1 - g0), (unsigned)g0);
                const uint32_t c0 = (uint32_t)(r0 / CHUNK), c1 = (uint32_t)((r1 + CHUNK - 1) / CHUNK);       // GAT_CGROUP is a multiple of CHUNK
                launchScoring(ctx, P, c0, c1 - c0, wl->plain, wl->streamLong, wl->searchJobs, false, st);      // (behind expandBlocksKernel: no early start)
                sliceLaunches += 2;
                const bool lastSlice = g1 == nGroups;
                launchFixupRange(ctx, wl, st, c1, fixedChunks, lastSlice ? 0xffffffffu : c1, lastSlice);
                fixedChunks = c1;
                sliceLaunches += 1;
                if (early && !lastSlice && !wl->searchJobs) {
                    // jobs [doneJobs, upTo) end in chunks below c1, i.e. own no record at or behind record c1 * CHUNK
                    const uint64_t limit = (uint64_t)c1 * CHUNK;
                    uint64_t lo = doneJobs, hi = nJobs;         // first job j with end(j) > limit; end(j) = blockPtr[j + 1] (or nBlocks)
                    while (lo < hi) {
                        const uint64_t mid = (lo + hi) / 2, end = mid + 1 < nJobs ? jobs[mid + 1].blockPtr : nBlocks;
                        if (end <= limit) lo = mid + 1; else hi = mid;
                    }
                    if (lo > doneJobs) {
                        CU(cudaEventRecord(ctx->fixEv[s], st));
                        CU(cudaStreamWaitEvent(ctx->backStream, ctx->fixEv[s], 0));
                        CU(cudaMemcpyAsync(global + doneJobs, wl->outGlobal + doneJobs, (lo - doneJobs) * sizeof(long long), cudaMemcpyDeviceToHost, ctx->backStream));
                        CU(cudaMemcpyAsync(local + doneJobs, wl->outLocal + doneJobs, (lo - doneJobs) * sizeof(long long), cudaMemcpyDeviceToHost, ctx->backStream));
                        doneJobs = lo;
                    }
                }
            }
            ctx->stats.kernel_launches = sliceLaunches - 3;
            fixedUp = true;
        }
        if (!fixedUp) launchFixup(ctx, wl, st);
        if (prof) CU(cudaEventRecord(ctx->ev[3], st));
        ctx->stats.kernel_launches += 5;        // expandJobs, expandBlocks, jobPrep, scoreTiles, fixup (+ 3 per further slice)
    }
    CU(cudaGetLastError());
    if (prof) cudaEventRecord(ctx->ev[5], st);
    if (doneJobs) CU(cudaStreamSynchronize(ctx->backStream));      // (before anything else may write the caller's arrays)
    rc = fetchResults(ctx, wl, global, local, doneJobs);
    if (rc == GAT_OK && prof && wl->nChunks) {
        float total = 0;
        cudaEventElapsedTime(&total, ctx->ev[4], ctx->ev[0]);
        ctx->stats.h2d_ms = total;          // copies + the two expansion kernels
        ctx->stats.h2d_bytes = nJobs * sizeof(gat_cjob) + nBlocks * recBytes + nAbs * sizeof(gat_cabs) + nGroups * (sizeof(gat_cabs) + (packed ? sizeof(uint32_t) : 0));
        ctx->stats.d2h_bytes = 2 * nJobs * sizeof(long long);
    }
    return rc;
}

extern "C" int gat_score_compact(gat_ctx *ctx, const gat_cjob *jobs, uint64_t nJobs, const gat_cblock *blocks, uint64_t nBlocks,
                                 const gat_cabs *abs, uint64_t nAbs, const gat_cabs *anchors, int64_t *global, int64_t *local)
{
    return scoreDeltaCoded(ctx, "gat_score_compact", jobs, nJobs, blocks, sizeof(gat_cblock), nBlocks, abs, nAbs, anchors, nullptr, global, local);
}

extern "C" int gat_score_packed(gat_ctx *ctx, const gat_cjob *jobs, uint64_t nJobs, const gat_pblock *blocks, uint64_t nBlocks,
                                const gat_cabs *abs, uint64_t nAbs, const gat_cabs *anchors, const uint32_t *absBase, int64_t *global, int64_t *local)
{
    return scoreDeltaCoded(ctx, "gat_score_packed", jobs, nJobs, blocks, sizeof(gat_pblock), nBlocks, abs, nAbs, anchors, absBase, global, local);
}

extern "C" int gat_crossover(gat_ctx *ctx, const gat_xpair *pairs, uint64_t nPairs, int32_t *pos, int32_t *adjust)
{
    if (!ctx) return fail(GAT_EINVAL, "gat_crossover: NULL ctx");
    if (nPairs && (!pairs || !pos || !adjust)) return fail(GAT_EINVAL, "gat_crossover: NULL array");
    if (!ctx->genome[0].loaded || !ctx->genome[1].loaded) return fail(GAT_ESTATE, "gat_crossover: load both genomes first");
    if (!ctx->scoringSet) return fail(GAT_ESTATE, "gat_crossover: call gat_set_scoring first");
    if (nPairs == 0) return GAT_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // scratch of the context, grown on demand and reused: no allocation per call
    const size_t need = nPairs * sizeof(gat_xpair) + 2 * nPairs * sizeof(int) + 64;
    if (need > ctx->xoverCap) {
        CU(cudaStreamSynchronize(st));
        cudaFree(ctx->xoverBuf);
        ctx->xoverBuf = nullptr; ctx->xoverCap = 0;
        if (cudaMalloc(&ctx->xoverBuf, need + need / 4) != cudaSuccess) return fail(GAT_ENOMEM, "gat_crossover: out of device memory");
        ctx->xoverCap = need + need / 4;
    }
    gat_xpair *dPairs = static_cast<gat_xpair *>(ctx->xoverBuf);
    int *dOut = reinterpret_cast<int *>(static_cast<char *>(ctx->xoverBuf) + ((nPairs * sizeof(gat_xpair) + 15) & ~(size_t)15));
    XoverParams P;
    P.pairs = dPairs; P.nPairs = nPairs;
    P.t = ctx->genome[GAT_TARGET].view(); P.q = ctx->genome[GAT_QUERY].view();
    memcpy(P.matrix, ctx->matrix, sizeof P.matrix);
    P.pos = dOut; P.adjust = dOut + nPairs; P.err = ctx->err;
    int rc = GAT_OK, err = 0;
    cudaError_t e = cudaMemcpyAsync(dPairs, pairs, nPairs * sizeof(gat_xpair), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        crossoverKernel<<<(unsigned)((nPairs + 127) / 128), 128, 0, st>>>(P);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos, P.pos, nPairs * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(adjust, P.adjust, nPairs * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = fail(GAT_ECUDA, "gat_crossover failed: %s", cudaGetErrorString(e));
    if (rc == GAT_OK) rc = readDeviceError(ctx, &err);
    cudaStreamSynchronize(st);
    if (rc == GAT_OK && err)
        rc = fail(GAT_EWORKLIST, "crossover pairs rejected by the device:%s%s", (err & ERR_SEQ) ? " sequence index out of range;" : "",
                  (err & ERR_COORD) ? " overlap outside its sequence;" : "");
    return rc;
}

extern "C" int gat_gap_cost(gat_ctx *ctx, const int32_t *dq, const int32_t *dt, uint64_t n, int32_t *out)
{
    if (!ctx || (n && (!dq || !dt || !out))) return fail(GAT_EINVAL, "gat_gap_cost: NULL argument");
    if (!ctx->scoringSet) return fail(GAT_ESTATE, "gat_gap_cost: call gat_set_scoring first");
    if (n == 0) return GAT_OK;
    CU(cudaSetDevice(ctx->device));
    int *d = nullptr;
    CU(cudaMalloc(&d, 3 * n * sizeof(int)));
    cudaError_t e = cudaMemcpyAsync(d, dq, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, dt, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        gapBatchKernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->gap, ctx->gapSmall, ctx->gapDense, ctx->gapLongPos, ctx->gapLongVal,
                                                                           d, d + n, n, d + 2 * n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(GAT_ECUDA, "gat_gap_cost failed: %s", cudaGetErrorString(e));
    return GAT_OK;
}

extern "C" uint32_t gat_max_record_bases(const gat_ctx *ctx) { return ctx ? ctx->maxBlockBases : 0; }


extern "C" int gat_synchronize(gat_ctx *ctx)
{
    if (!ctx) return fail(GAT_EINVAL, "gat_synchronize: NULL ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return GAT_OK;
}

extern "C" int gat_get_stats(gat_ctx *ctx, gat_stats *out)
{
    if (!ctx || !out) return fail(GAT_EINVAL, "gat_get_stats: NULL argument");
    int rc = finishStats(ctx);
    *out = ctx->stats;
    return rc;
}

extern "C" int gat_set_profiling(gat_ctx *ctx, int on)
{
    if (!ctx) return fail(GAT_EINVAL, "gat_set_profiling: NULL ctx");
    ctx->profiling = on != 0;
    return GAT_OK;
}

extern "C" int gat_request_tuples(gat_ctx *ctx, const uint32_t *jobIx, uint64_t n, gat_tuple *out)
{
    if (!ctx || (n && (!jobIx || !out))) return fail(GAT_EINVAL, "gat_request_tuples: NULL argument");
    ctx->partJobs.assign(jobIx, jobIx + n);
    ctx->partOut = out;
    return GAT_OK;
}

// acc = acc (+) [peak test, gap, clamp at 0] (+) next: what chainCalcScoreLocal does between the last block of one part
// and the first block of the next (scoreChain.c:181-195), as tuples (see Tup in gat_kernels.cuh)
extern "C" void gat_tuple_join(gat_tuple *acc, int64_t gapCost, const gat_tuple *next)
{
    const Tup a{acc->d, acc->c, acc->e, acc->f}, x{-gapCost, 0, 0, NEG}, b{next->d, next->c, next->e, next->f};
    auto comb = [](const Tup &p, const Tup &q) {
        Tup r;
        r.d = p.d + q.d;
        r.c = std::max(q.c, p.c + q.d);
        r.e = std::max(p.e, p.d + q.e);
        r.f = std::max(std::max(p.f, q.f), p.c + q.e);
        return r;
    };
    const Tup r = comb(comb(a, x), b);
    acc->d = r.d; acc->c = r.c; acc->e = r.e; acc->f = r.f;
}

extern "C" void gat_tuple_scores(const gat_tuple *t, int64_t *global, int64_t *local)
{
    if (global) *global = t->d;
    if (local) *local = std::max<int64_t>(0, std::max(std::max(t->c, t->d), std::max(t->e, t->f)));
}

extern "C" void *gat_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(GAT_ENOMEM, "gat_host_alloc: cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

extern "C" void gat_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
