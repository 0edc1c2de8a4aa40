// C wrappers of gat_host.hpp (include/gat_host.h) for bindings and tests.
#include <cstring>
#include "gat_host.h"
#include "gat_host.hpp"

using namespace gathost;
static thread_local std::string g_err;
extern "C" const char *gathost_last_error(void) { return g_err.c_str(); }
#define GUARD(body, onError) try { body } catch (const Error &e) { g_err = e.message; return onError; }

struct gathost_gapcalc { GapCalc g; };
extern "C" gathost_gapcalc *gathost_gapcalc_open(const char *name) { GUARD(return new gathost_gapcalc{GapCalc::fromFile(name)};, nullptr) }
extern "C" void gathost_gapcalc_close(gathost_gapcalc *g) { delete g; }
extern "C" int gathost_gapcalc_cost(const gathost_gapcalc *g, int dq, int dt) { return g->g.cost(dq, dt); }
extern "C" int gathost_gapcalc_fill(const gathost_gapcalc *h, gat_scoring *s)
{
    const GapCalc &g = h->g;
    s->smallSize = g.smallSize;
    s->qSmall = g.qSmall.data(); s->tSmall = g.tSmall.data(); s->bSmall = g.bSmall.data();
    s->longCount = (int)g.longPos.size(); s->longPos = g.longPos.data();
    s->qLong = g.qLong.data(); s->tLong = g.tLong.data(); s->bLong = g.bLong.data();
    return 0;
}
extern "C" int gathost_scorescheme(const char *path, int32_t matrix[4][4])
{
    GUARD(ScoreScheme ss = path ? ScoreScheme::read(path) : ScoreScheme::defaultScheme(); memcpy(matrix, ss.matrix, sizeof ss.matrix); return 0;, -1)
}

struct gathost_chains { ChainSet cs; WorkList wl; };
extern "C" gathost_chains *gathost_chains_read(const char *path)
{
    gathost_chains *c = new gathost_chains();
    try { readChains(path, c->cs); buildRecords(c->cs, c->wl); }
    catch (const Error &e) { g_err = e.message; delete c; return nullptr; }
    return c;
}
extern "C" void gathost_chains_close(gathost_chains *c) { delete c; }
extern "C" uint64_t gathost_chains_count(const gathost_chains *c) { return c->cs.chains.size(); }
extern "C" uint64_t gathost_chains_block_count(const gathost_chains *c) { return c->wl.blocks.size(); }
extern "C" const gat_block *gathost_chains_blocks(const gathost_chains *c) { return c->wl.blocks.data(); }
extern "C" int gathost_chains_head(const gathost_chains *c, uint64_t ix, double *score, const char **tName, int *tSize,
                                   int *tStart, int *tEnd, const char **qName, int *qSize, char *qStrand, int *qStart,
                                   int *qEnd, int *id, uint64_t *firstBlock, uint64_t *nBlocks)
{
    if (ix >= c->cs.chains.size()) { g_err = "chain index out of range"; return -1; }
    const ChainHead &h = c->cs.chains[ix];
    *score = h.score; *tName = h.tName.c_str(); *tSize = h.tSize; *tStart = h.tStart; *tEnd = h.tEnd;
    *qName = h.qName.c_str(); *qSize = h.qSize; *qStrand = h.qStrand; *qStart = h.qStart; *qEnd = h.qEnd; *id = h.id;
    *firstBlock = c->wl.chainFirstRecord[ix]; *nBlocks = c->wl.chainFirstRecord[ix + 1] - c->wl.chainFirstRecord[ix];
    return 0;
}
extern "C" int gathost_chains_subset(const gathost_chains *c, uint64_t ix, int subStart, int subEnd, uint64_t *firstBlock,
                                     uint64_t *nBlocks, int32_t *clipStart, int32_t *clipEnd, int64_t *aliBases)
{
    WorkList tmp;
    tmp.chainFirstRecord = c->wl.chainFirstRecord;
    if (!addSubChainJob(c->cs, ix, 0, 0, subStart, subEnd, tmp)) { *firstBlock = *nBlocks = 0; *clipStart = subStart; *clipEnd = subEnd; *aliBases = 0; return 0; }
    *firstBlock = tmp.jobs[0].firstBlock; *nBlocks = tmp.totalJobBlocks;
    *clipStart = tmp.jobs[0].clipStart; *clipEnd = tmp.jobs[0].clipEnd; *aliBases = tmp.aliBases[0];
    return 1;
}

extern "C" int gathost_chains_remove_partial_overlaps(gathost_chains *c, gat_ctx *ctx, const uint32_t *chainT, const uint32_t *chainQ)
{
    const size_t n = c->cs.chains.size();
    GUARD(removePartialOverlaps(ctx, c->cs, std::vector<uint32_t>(chainT, chainT + n), std::vector<uint32_t>(chainQ, chainQ + n));
          buildRecords(c->cs, c->wl); return 0;, -1)
}

struct gathost_compact { CompactWorkList cw; };
extern "C" gathost_compact *gathost_chains_compact(const gathost_chains *c, const uint32_t *chainT, const uint32_t *chainQ)
{
    WorkList wl = c->wl;                 // records as built; one whole-chain job per chain
    wl.jobs.clear(); wl.aliBases.clear(); wl.totalJobBlocks = 0;
    gathost_compact *out = new gathost_compact();
    try {
        for (size_t i = 0; i < c->cs.chains.size(); i++) addChainJob(c->cs, i, chainT[i], chainQ[i], wl);
        if (!packCompact(wl, out->cw)) { g_err = "work-list does not qualify for the compact form"; delete out; return nullptr; }
    } catch (const Error &e) { g_err = e.message; delete out; return nullptr; }
    return out;
}
extern "C" void gathost_compact_free(gathost_compact *p) { delete p; }
extern "C" int gathost_compact_view(const gathost_compact *p, const gat_cjob **jobs, uint64_t *nJobs, const gat_cblock **blocks, uint64_t *nBlocks,
                                    const gat_cabs **abs, uint64_t *nAbs, const gat_cabs **anchors, uint64_t *nAnchors)
{
    *jobs = p->cw.jobs.data(); *nJobs = p->cw.jobs.size(); *blocks = p->cw.blocks.data(); *nBlocks = p->cw.blocks.size();
    *abs = p->cw.abs.data(); *nAbs = p->cw.abs.size(); *anchors = p->cw.anchors.data(); *nAnchors = p->cw.anchors.size();
    return 0;
}

struct gathost_twobit { TwoBitFile tb; explicit gathost_twobit(const char *p) : tb(p) {} };
extern "C" gathost_twobit *gathost_twobit_open(const char *path) { GUARD(return new gathost_twobit(path);, nullptr) }
extern "C" void gathost_twobit_close(gathost_twobit *t) { delete t; }
extern "C" uint32_t gathost_twobit_count(const gathost_twobit *t) { return (uint32_t)t->tb.seqs().size(); }
extern "C" int gathost_twobit_seq(const gathost_twobit *t, uint32_t ix, const char **name, uint32_t *size, const uint8_t **packed,
                                  uint32_t *nRuns, const uint32_t **nStart, const uint32_t **nLen)
{
    if (ix >= t->tb.seqs().size()) { g_err = "sequence index out of range"; return -1; }
    const TwoBitSeq &s = t->tb.seqs()[ix];
    *name = s.name.c_str(); *size = s.size; *packed = s.packed; *nRuns = (uint32_t)s.nStart.size();
    *nStart = s.nStart.data(); *nLen = s.nLen.data();
    return 0;
}

extern "C" int gathost_shard_jobs(const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks, const int64_t *aliBases,
                                  int parts, uint32_t *part)
{
    WorkList wl;
    wl.jobs.assign(jobs, jobs + nJobs);
    wl.aliBases.assign(aliBases, aliBases + nJobs);
    wl.totalJobBlocks = totalJobBlocks;
    const auto shards = shardJobs(wl, parts);
    for (size_t p = 0; p < shards.size(); p++)
        for (uint32_t j : shards[p]) part[j] = (uint32_t)p;
    return 0;
}
