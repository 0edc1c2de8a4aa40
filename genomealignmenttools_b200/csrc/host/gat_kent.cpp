// gat_kent.cpp -- libgatkent.so: kent's per-chain entry points (include/gat_kent.h) as one-job batches of the GPU ABI.
// Nothing in here scores on the CPU: chainCalcScore & co. pack the caller's sequences to the .2bit payload, upload them
// once (cached), turn the chain into a one-job work-list and call gat_score(); gapCalcCost calls gat_gap_cost().
#include "gat_kent.h"
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "gat.h"
#include "gat_host.hpp"

using gathost::errAbort;

// ---- kent's types as the shims see them (layouts of kent/src/inc/chain.h:17-25, 48-63, dnaseq.h:18-26, axt.h:83-91;
// verified against kent's own headers by oracle/kent_shim_check.c through gatKentLayout)
struct cBlock {
    struct cBlock *next;
    int tStart, tEnd;
    int qStart, qEnd;
    int score;
    void *data;
};
struct chain {
    struct chain *next;
    struct cBlock *blockList;
    double score;
    char *tName;
    int tSize;
    int tStart, tEnd;
    char *qName;
    int qSize;
    char qStrand;
    int qStart, qEnd;
    int id;
};
struct dnaSeq {
    struct dnaSeq *next;
    char *name;
    char *dna;
    int size;
    void *mask;
};
struct axtScoreScheme {
    void *next;
    int matrix[256][256];
    int gapOpen;
    int gapExtend;
    char *extra;
};
struct gapCalc {
    gathost::GapCalc g;
};

namespace {

std::mutex gLock;
gat_ctx *gCtx = nullptr;

gat_ctx *context()
{
    if (!gCtx) {
        const char *d = getenv("GAT_DEVICE");
        if (gat_create(&gCtx, d ? atoi(d) : 0, nullptr) != GAT_OK) errAbort("%s", gat_last_error());
    }
    return gCtx;
}

// kent base code (T=0 C=1 A=2 G=3, dnautil.h:23-27) of a sequence character, -1 for anything else (scores 0 like N)
inline int baseCode(char c)
{
    switch (c) {
    case 't': case 'T': return 0;
    case 'c': case 'C': return 1;
    case 'a': case 'A': return 2;
    case 'g': case 'G': return 3;
    default: return -1;
    }
}

struct PackedSeq {
    std::vector<uint8_t> packed;        // .2bit payload: 4 bases per byte, first base in bits 7..6 (twoBit.c:811-818)
    std::vector<gat_nrun> nRuns;
};

void packDna(const char *dna, size_t size, PackedSeq &out)
{
    out.packed.assign((size + 3) / 4, 0);
    out.nRuns.clear();
    unsigned T = std::thread::hardware_concurrency();
    T = std::max(1u, std::min(T ? T : 1u, std::min(16u, (unsigned)(size >> 22) + 1u)));
    std::vector<std::vector<gat_nrun>> runs(T);
    auto body = [&](unsigned k) {       // whole bytes [lo, hi) of the payload
        const size_t nBytes = out.packed.size(), lo = nBytes / T * k, hi = k + 1 == T ? nBytes : nBytes / T * (k + 1);
        uint32_t runStart = 0, runLen = 0;
        for (size_t b = lo; b < hi; b++) {
            uint8_t byte = 0;
            for (int j = 0; j < 4; j++) {
                const size_t p = 4 * b + j;
                int code = 0;
                if (p < size) {
                    code = baseCode(dna[p]);
                    if (code < 0) {
                        if (runLen && runStart + runLen == p) runLen++;
                        else { if (runLen) runs[k].push_back(gat_nrun{0, runStart, runLen}); runStart = (uint32_t)p; runLen = 1; }
                        code = 0;       // stored as T, like twoBitFromDnaSeq does
                    }
                }
                byte |= (uint8_t)(code << (6 - 2 * j));
            }
            out.packed[b] = byte;
        }
        if (runLen) runs[k].push_back(gat_nrun{0, runStart, runLen});
    };
    std::vector<std::thread> th;
    for (unsigned k = 1; k < T; k++) th.emplace_back(body, k);
    body(0);
    for (auto &t : th) t.join();
    for (unsigned k = 0; k < T; k++)
        for (const gat_nrun &r : runs[k]) {
            if (!out.nRuns.empty() && out.nRuns.back().start + out.nRuns.back().len == r.start) out.nRuns.back().len += r.len;   // a run across two slices
            else out.nRuns.push_back(r);
        }
}

// what is resident on the device for one side
struct Resident {
    const char *dna = nullptr;
    size_t size = 0;
    uint64_t stamp = 0;
    bool valid = false;
} gSide[2];

uint64_t sampleStamp(const char *dna, size_t size)
{   // FNV over the first / last 256 characters and 4096 evenly spread ones: catches reuse of a buffer for another sequence
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](size_t i) { h = (h ^ (uint8_t)dna[i]) * 1099511628211ull; };
    const size_t edge = std::min<size_t>(256, size);
    for (size_t i = 0; i < edge; i++) mix(i);
    for (size_t i = size - edge; i < size; i++) mix(i);
    if (size > 4096) for (size_t k = 0; k < 4096; k++) mix(size / 4096 * k);
    return h ^ size;
}

void ensureResident(int side, const char *dna, size_t size)
{
    const uint64_t stamp = sampleStamp(dna, size);
    Resident &r = gSide[side];
    if (r.valid && r.dna == dna && r.size == size && r.stamp == stamp) return;
    PackedSeq ps;
    packDna(dna, size, ps);
    const uint64_t off = 0;
    const uint32_t sz = (uint32_t)size;
    if (gat_load_genome(context(), side, ps.packed.data(), ps.packed.size(), &off, &sz, 1, ps.nRuns.data(), ps.nRuns.size()) != GAT_OK)
        errAbort("%s", gat_last_error());
    r.dna = dna; r.size = size; r.stamp = stamp; r.valid = true;
}

// the live 4x4 of a 256x256 matrix, [q][t] in kent codes; everything else must be what axtScoreSchemeRead leaves there
void liveMatrix(int matrix[256][256], int32_t out[4][4])
{
    static const char lower[4] = {'t', 'c', 'a', 'g'}, upper[4] = {'T', 'C', 'A', 'G'};
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) {
            const int v = matrix[(int)lower[q]][(int)lower[t]];
            if (matrix[(int)upper[q]][(int)upper[t]] != v || matrix[(int)lower[q]][(int)upper[t]] != v || matrix[(int)upper[q]][(int)lower[t]] != v)
                errAbort("gat_kent: score scheme differs between upper and lower case (not what propagateCase leaves, axt.c:402-421)");
            out[q][t] = v;
        }
    for (int i = 0; i < 256; i++)
        for (int j = 0; j < 256; j++)
            if (matrix[i][j] != 0 && (baseCode((char)i) < 0 || baseCode((char)j) < 0))
                errAbort("gat_kent: score scheme has an entry for '%c' x '%c'; only a/c/g/t are scored on the GPU path", i, j);
}

struct ScoringKey {
    int32_t m[4][4];
    const struct gapCalc *gc = nullptr;
    bool valid = false;
} gScoring;

void ensureScoring4(const int32_t m[4][4], const struct gapCalc *gc)
{
    if (gScoring.valid && gScoring.gc == gc && memcmp(gScoring.m, m, sizeof gScoring.m) == 0) return;
    gathost::ScoreScheme ss;
    memcpy(ss.matrix, m, sizeof ss.matrix);
    static const gathost::GapCalc none = gathost::GapCalc::fromFile("loose");        // chainScoreBlock needs no gap costs
    try { gathost::setScoring(context(), ss, gc ? gc->g : none); } catch (const gathost::Error &e) { errAbort("%s", e.message.c_str()); }
    memcpy(gScoring.m, m, sizeof gScoring.m);
    gScoring.gc = gc; gScoring.valid = true;
}

void ensureScoring(int matrix[256][256], const struct gapCalc *gc)
{
    static int (*checked)[256] = nullptr;       // the last matrix that passed the full 256x256 inspection
    int32_t m[4][4];
    if (checked == matrix) {
        static const char lower[4] = {'t', 'c', 'a', 'g'};
        for (int q = 0; q < 4; q++)
            for (int t = 0; t < 4; t++) m[q][t] = matrix[(int)lower[q]][(int)lower[t]];
    } else { liveMatrix(matrix, m); checked = matrix; }
    ensureScoring4(m, gc);
}

// blocks of a chain as device records (long blocks cut into JOINED records), coordinates shifted by (tOff, qOff)
void chainRecords(const struct chain *chain, int tOff, int qOff, std::vector<gat_block> &out, int *aliBases)
{
    const uint32_t piece = GAT_SPLIT_BASES;
    long long ali = 0;
    for (const struct cBlock *b = chain->blockList; b; b = b->next) {
        const int size = b->tEnd - b->tStart;           // chainCalcScore takes the size from the target (chainConnect.c:32)
        if (size < 0) errAbort("gat_kent: block with tEnd < tStart in chain %d", chain->id);
        ali += size;
        uint32_t done = 0;
        do {
            const uint32_t n = std::min<uint32_t>(piece, (uint32_t)size - done);
            out.push_back(gat_block{b->tStart - tOff + (int)done, b->qStart - qOff + (int)done, n | (done ? GAT_BLOCK_JOINED : 0u)});
            done += n;
        } while (done < (uint32_t)size);
    }
    if (aliBases) *aliBases = (int)ali;
}

void scoreOne(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gc, struct dnaSeq *query, struct dnaSeq *target,
              int tOff, int qOff, int64_t *global, int64_t *local, int *aliBases)
{
    std::lock_guard<std::mutex> hold(gLock);
    if (!chain || !ss || !gc || !query || !target) errAbort("gat_kent: NULL argument");
    *global = *local = 0;
    if (aliBases) *aliBases = 0;
    if (!chain->blockList) return;
    ensureResident(GAT_TARGET, target->dna, (size_t)target->size);
    ensureResident(GAT_QUERY, query->dna, (size_t)query->size);
    ensureScoring(ss->matrix, gc);
    std::vector<gat_block> blocks;
    chainRecords(chain, tOff, qOff, blocks, aliBases);
    // the caller's query is already in the chain's coordinates (reverse-complemented by the caller on '-',
    // scoreChain.c:123-149): always the forward image here
    const gat_job job{0, 0, 0, 0, GAT_NO_CLIP_START, GAT_NO_CLIP_END};
    if (gat_score(context(), &job, 1, blocks.size(), blocks.data(), blocks.size(), global, local) != GAT_OK) errAbort("%s", gat_last_error());
}

}  // namespace

extern "C" {

double chainCalcScore(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query, struct dnaSeq *target)
{
    int64_t g, l;
    scoreOne(chain, ss, gapCalc, query, target, 0, 0, &g, &l, nullptr);
    return (double)g;
}

double chainCalcScoreSubChain(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query, struct dnaSeq *target)
{   // query / target hold only the chain's span (chainConnect.c:42-59)
    int64_t g, l;
    scoreOne(chain, ss, gapCalc, query, target, chain ? chain->tStart : 0, chain ? chain->qStart : 0, &g, &l, nullptr);
    return (double)g;
}

double chainCalcScoreLocal(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query, struct dnaSeq *target,
                           int *retAliBases)
{
    int64_t g, l;
    scoreOne(chain, ss, gapCalc, query, target, 0, 0, &g, &l, retAliBases);
    return (double)l;
}

double chainScoreBlock(char *q, char *t, int size, int matrix[256][256])
{
    std::lock_guard<std::mutex> hold(gLock);
    if (size <= 0) return 0;
    ensureResident(GAT_TARGET, t, (size_t)size);
    ensureResident(GAT_QUERY, q, (size_t)size);
    ensureScoring(matrix, gScoring.valid ? gScoring.gc : nullptr);
    std::vector<gat_block> blocks;
    for (uint32_t done = 0; done < (uint32_t)size; done += GAT_SPLIT_BASES)
        blocks.push_back(gat_block{(int)done, (int)done, std::min<uint32_t>(GAT_SPLIT_BASES, (uint32_t)size - done) | (done ? GAT_BLOCK_JOINED : 0u)});
    const gat_job job{0, 0, 0, 0, GAT_NO_CLIP_START, GAT_NO_CLIP_END};
    int64_t g = 0, l = 0;
    if (gat_score(context(), &job, 1, blocks.size(), blocks.data(), blocks.size(), &g, &l) != GAT_OK) errAbort("%s", gat_last_error());
    return (double)g;
}

static struct gapCalc *newGapCalc(const gathost::GapCalc &g)
{
    struct gapCalc *gc = new gapCalc();
    gc->g = g;
    return gc;
}
struct gapCalc *gapCalcFromFile(char *fileName)
{
    try { return newGapCalc(gathost::GapCalc::fromFile(fileName)); } catch (const gathost::Error &e) { errAbort("%s", e.message.c_str()); }
}
struct gapCalc *gapCalcFromString(char *s)
{
    try { return newGapCalc(gathost::GapCalc::fromString(s ? s : "")); } catch (const gathost::Error &e) { errAbort("%s", e.message.c_str()); }
}
struct gapCalc *gapCalcDefault(void) { static char name[] = "loose"; return gapCalcFromFile(name); }      // defaultGapCosts, gapCalc.c:50-56, 224-228
struct gapCalc *gapCalcOriginal(void) { static char name[] = "medium"; return gapCalcFromFile(name); }    // originalGapCosts, gapCalc.c:40-46, 271-275
void gapCalcFree(struct gapCalc **pGapCalc)
{
    if (!pGapCalc || !*pGapCalc) return;
    std::lock_guard<std::mutex> hold(gLock);
    if (gScoring.gc == *pGapCalc) gScoring.valid = false;
    delete *pGapCalc;
    *pGapCalc = nullptr;
}
char *gapCalcSampleFileContents(void) { return const_cast<char *>(gathost::GapCalc::sampleFileContents()); }

int gapCalcCost(struct gapCalc *gapCalc, int dq, int dt)
{
    std::lock_guard<std::mutex> hold(gLock);
    if (!gapCalc) errAbort("gat_kent: NULL gapCalc");
    if (!gScoring.valid || gScoring.gc != gapCalc) {        // the gap tables travel with a matrix: keep the current one
        const gathost::ScoreScheme d = gathost::ScoreScheme::defaultScheme();
        int32_t m[4][4];
        memcpy(m, gScoring.valid ? gScoring.m : d.matrix, sizeof m);
        ensureScoring4(m, gapCalc);
    }
    int32_t out = 0;
    if (gat_gap_cost(context(), &dq, &dt, 1, &out) != GAT_OK) errAbort("%s", gat_last_error());
    return out;
}

static void fillScheme(struct axtScoreScheme *ss, const gathost::ScoreScheme &s)
{
    static const char lower[4] = {'t', 'c', 'a', 'g'}, upper[4] = {'T', 'C', 'A', 'G'};
    memset(ss, 0, sizeof *ss);
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) {
            const int v = s.matrix[q][t];
            ss->matrix[(int)lower[q]][(int)lower[t]] = ss->matrix[(int)upper[q]][(int)upper[t]] = v;
            ss->matrix[(int)lower[q]][(int)upper[t]] = ss->matrix[(int)upper[q]][(int)lower[t]] = v;      // propagateCase, axt.c:402-421
        }
    ss->gapOpen = s.gapOpen;
    ss->gapExtend = s.gapExtend;
}
struct axtScoreScheme *axtScoreSchemeDefault(void)
{
    static struct axtScoreScheme *ss = nullptr;     // "Do NOT axtScoreSchemeFree this" (axt.h:96-98)
    if (!ss) {
        ss = static_cast<struct axtScoreScheme *>(malloc(sizeof *ss));
        fillScheme(ss, gathost::ScoreScheme::defaultScheme());
    }
    return ss;
}
struct axtScoreScheme *axtScoreSchemeRead(char *fileName)
{
    try {
        const gathost::ScoreScheme s = gathost::ScoreScheme::read(fileName ? fileName : "");
        struct axtScoreScheme *ss = static_cast<struct axtScoreScheme *>(malloc(sizeof *ss));
        fillScheme(ss, s);
        return ss;
    } catch (const gathost::Error &e) { errAbort("%s", e.message.c_str()); }
}
void axtScoreSchemeFree(struct axtScoreScheme **pObj)
{
    if (!pObj || !*pObj) return;
    free((*pObj)->extra);
    free(*pObj);
    *pObj = nullptr;
}

void gatKentForget(void)
{
    std::lock_guard<std::mutex> hold(gLock);
    gSide[0].valid = gSide[1].valid = false;
}

long gatKentLayout(int which)
{
    switch (which) {
    case 0: return sizeof(struct cBlock);
    case 1: return sizeof(struct chain);
    case 2: return sizeof(struct dnaSeq);
    case 3: return sizeof(struct axtScoreScheme);
    case 10: return offsetof(struct cBlock, tStart);
    case 11: return offsetof(struct cBlock, qStart);
    case 12: return offsetof(struct cBlock, tEnd);
    case 20: return offsetof(struct chain, blockList);
    case 21: return offsetof(struct chain, tStart);
    case 22: return offsetof(struct chain, qStart);
    case 23: return offsetof(struct chain, qStrand);
    case 24: return offsetof(struct chain, id);
    case 30: return offsetof(struct dnaSeq, dna);
    case 31: return offsetof(struct dnaSeq, size);
    case 40: return offsetof(struct axtScoreScheme, matrix);
    case 41: return offsetof(struct axtScoreScheme, gapOpen);
    default: return -1;
    }
}

}  // extern "C"
