// chainNet -- make alignment nets out of chains.  Drop-in for src/chainNet/chainNet.c of
// hillerlab/GenomeAlignmentTools (same command line, same .net bytes).  Net construction is host
// logic and follows the reference's rules step by step (cited below); what changes is -rescore:
// the reference re-scores every partial target-side fill with one chainSubsetOnT + chainCalcScore
// CPU call while printing (chainNet.c:832-835 -> :230-248).  Under -rescore the decision to print
// a fill never depends on its score (minScore is forced to 0, :1022, and scores are clamped to
// >= 1, :244-245), so here the net is walked once to collect all such fills as clipped jobs, the
// whole batch is scored by the sm_100a kernels behind gat_score(), and a second walk prints.
#include <algorithm>
#include <climits>
#include <cstring>
#include <map>
#include <unistd.h>
#include "gat_host.hpp"
#include <memory>

using namespace gathost;

static const std::vector<OptionSpec> optionSpecs = {
    {"minSpace", OPTION_INT}, {"minFill", OPTION_INT}, {"minScore", OPTION_DOUBLE}, {"inclHap", OPTION_BOOLEAN},
    {"rescore", OPTION_BOOLEAN}, {"tNibDir", OPTION_STRING}, {"qNibDir", OPTION_STRING}, {"scoreScheme", OPTION_STRING},
    {"linearGap", OPTION_STRING}, {"gpus", OPTION_INT},
};

static int minSpace = 25, minFill = 0;
static double minScore = 2000;
static bool inclHap = false, rescore = false;

static void usage()
{   // chainNet.c:70-109
    errAbort(
        "chainNet - Make alignment nets out of chains\n"
        "usage:\n"
        "   chainNet in.chain target.sizes query.sizes target.net query.net\n"
        "where:\n"
        "   in.chain is the chain file sorted by score\n"
        "   target.sizes contains the size of the target sequences\n"
        "   query.sizes contains the size of the query sequences\n"
        "   target.net is the output over the target genome\n"
        "   query.net is the output over the query genome\n"
        "options:\n"
        "   -minSpace=N - minimum gap size to fill, default %d\n"
        "   -minFill=N  - default half of minSpace\n"
        "   -minScore=N - minimum chain score to consider, default %.1lf\n"
        "   -verbose=N - Alter verbosity (default 1)\n"
        "   -inclHap - include query sequences name in the form *_hap*|*_alt*.\n"
        "              Normally these are excluded from nets as being haplotype\n"
        "              pseudochromosomes\n"
        "\n"
        "\n"
        "   -rescore                    compute the real score of the sub-net (instead of approximating it based on the fraction of aligning bases in the subnet)\n"
        "                               The real score will be much more precise especially for imbalanced chains where most aligning blocks are on one side.\n"
        "                               This flag will set minScore=0. Each subnet with a negative score gets score 1. Afterwards, run a non-nested score filter.\n"
        "                               Note: Rescoring is only implemented for the target species net.\n"
        "                               With this flag, you need to give the target and query genome sequence (-tNibDir and -qNibDir) and specify -linearGap\n"
        "   -tNibDir=fileName           target genome file (2bit or nib format)\n"
        "   -qNibDir=fileName           query genome file (2bit or nib format)\n"
        "   -scoreScheme=fileName       Read the scoring matrix from a blastz-format file\n"
        "   -linearGap=<medium|loose|filename> Specify type of linearGap to use.\n"
        "              *Must* specify this argument to one of these choices.\n"
        "              loose is chicken/human linear gap costs.\n"
        "              medium is mouse/human linear gap costs.\n"
        "              Or specify a piecewise linearGap tab delimited file.\n"
        "   -gpus=N                     (B200 build) shard the rescoring jobs over N GPUs, default 1\n"
        "   sample linearGap file (loose)\n"
        "%s",
        minSpace, minScore, GapCalc::sampleFileContents());
}

// ---------------------------------------------------------------- net data (chainNet.c:111-147)
struct Gap { int start, end, oStart, oEnd; std::vector<int> fills; };
struct Fill { int start, end, oStart = 0, oEnd = 0; std::vector<int> gaps; int chain; };
struct Space { int end; int gap; };                          // keyed by start in Chrom::spaces
struct Chrom { std::string name; int size; int root; std::map<int, Space> spaces; };

static std::vector<Gap> gaps;
static std::vector<Fill> fills;

static int newGap(int start, int end, int oStart, int oEnd)
{
    gaps.push_back(Gap{start, end, oStart, oEnd, {}});
    return (int)gaps.size() - 1;
}

// One side's view of a chain's blocks in ascending plus-strand coordinates of that side:
// (s,e) on this side, (os,oe) on the other.  For the query side of a '-' chain this is the
// reversed list with reversed q coordinates (reverseBlocksQ, chainNet.c:546-553).
struct SideBlock { int s, e, os, oe; };

static std::vector<SideBlock> sideBlocks(const ChainSet &cs, const ChainHead &h, bool isQ)
{
    std::vector<SideBlock> v(h.nBlocks);
    const gat_block *b = cs.blocks.data() + h.firstBlock;
    for (uint64_t i = 0; i < h.nBlocks; i++) {
        const int ts = b[i].tStart, te = ts + (int)b[i].size, qs = b[i].qStart, qe = qs + (int)b[i].size;
        if (!isQ) v[i] = SideBlock{ts, te, qs, qe};
        else if (h.qStrand == '-') v[h.nBlocks - 1 - i] = SideBlock{h.qSize - qe, h.qSize - qs, ts, te};
        else v[i] = SideBlock{qs, qe, ts, te};
    }
    return v;
}

// innerBounds, chainNet.c:354-387
static bool innerBounds(const std::vector<SideBlock> &bl, size_t from, int inStart, int inEnd, int &outStart, int &outEnd)
{
    int start = INT_MAX, end = -INT_MAX;
    for (size_t i = from; i < bl.size(); i++) {
        int s = bl[i].s, e = bl[i].e;
        if (e <= inStart) continue;
        if (s >= inEnd) break;
        if (s < inStart) s = inStart;
        if (e > inEnd) e = inEnd;
        if (start > s) start = s;
        if (end < e) end = e;
    }
    if (end < 0 || end - start < minFill) return false;
    outStart = start;
    outEnd = end;
    return true;
}

// addChainT / addChainQ + fillSpace, chainNet.c:487-679
static void addChainSide(Chrom &chrom, const ChainSet &cs, int chainIx, bool isQ)
{
    const ChainHead &h = cs.chains[chainIx];
    const std::vector<SideBlock> bl = sideBlocks(cs, h, isQ);
    int cStart = isQ ? h.qStart : h.tStart, cEnd = isQ ? h.qEnd : h.tEnd;
    if (isQ && h.qStrand == '-') { const int t = cStart; cStart = h.qSize - cEnd; cEnd = h.qSize - t; }
    // snapshot of the spaces that intersect the chain (findSpaces, :533-544); the tree changes below
    struct Hit { int start, end, gap; };
    std::vector<Hit> hits;
    auto it = chrom.spaces.upper_bound(cStart);
    if (it != chrom.spaces.begin()) {
        auto prev = std::prev(it);
        if (prev->second.end > cStart) it = prev;
    }
    for (; it != chrom.spaces.end() && it->first < cEnd; ++it) hits.push_back(Hit{it->first, it->second.end, it->second.gap});
    size_t startBlock = 0;
    for (const Hit &sp : hits) {
        while (startBlock + 1 < bl.size() && bl[startBlock + 1].s <= sp.start) startBlock++;
        int s, e;
        if (!innerBounds(bl, startBlock, sp.start, sp.end, s, e)) continue;
        fills.push_back(Fill());
        const int fillIx = (int)fills.size() - 1;
        fills[fillIx].start = s; fills[fillIx].end = e; fills[fillIx].chain = chainIx;
        chrom.spaces.erase(sp.start);
        if (s - sp.start >= minSpace) chrom.spaces[sp.start] = Space{s, sp.gap};
        if (sp.end - e >= minSpace) chrom.spaces[e] = Space{sp.end, sp.gap};
        gaps[sp.gap].fills.push_back(fillIx);
        for (size_t i = startBlock; i + 1 < bl.size(); i++) {
            const int gapStart = bl[i].e, gapEnd = bl[i + 1].s;
            if (gapStart >= sp.end) break;                  // sorted blocks: nothing further can lie inside
            if (sp.start < gapStart && gapStart + minSpace <= gapEnd && gapEnd < sp.end) {     // strictlyInside, :320-325
                int os, oe;
                if (!isQ) {                                 // other side = query, reported on the plus strand
                    os = bl[i].oe; oe = bl[i + 1].os;
                    if (h.qStrand == '-') { const int t = os; os = h.qSize - oe; oe = h.qSize - t; }
                } else if (h.qStrand == '+') { os = bl[i].oe; oe = bl[i + 1].os; }
                else { os = bl[i + 1].os; oe = bl[i].oe; }  // reversed list, as the reference has it (:655-659)
                const int g = newGap(gapStart, gapEnd, os, oe);
                chrom.spaces[gapStart] = Space{gapEnd, g};
                fills[fillIx].gaps.push_back(g);
            }
        }
    }
}

// sortNet, chainNet.c:697-709
static void sortNet(int gapIx)
{
    std::sort(gaps[gapIx].fills.begin(), gaps[gapIx].fills.end(), [](int a, int b) { return fills[a].start < fills[b].start; });
    for (int f : gaps[gapIx].fills) {
        std::sort(fills[f].gaps.begin(), fills[f].gaps.end(), [](int a, int b) { return gaps[a].start < gaps[b].start; });
        for (int g : fills[f].gaps) sortNet(g);
    }
}

// tFillOtherRange / qFillOtherRange, chainNet.c:389-484: refine the fill to the part of the chain
// actually used and compute the range on the other side.
// per chain and side: do the blocks ascend without overlap?  (-1 not looked at yet)  Then the blocks that meet a
// range are found by bisection instead of the reference's walk over the whole list: a net has one fill per
// aligning stretch of a chain, so walking every chain once per fill is quadratic in the long chains.
static bool ascendingChain(const ChainSet &cs, int chain, bool onQ)
{
    static std::vector<signed char> known[2];
    std::vector<signed char> &v = known[onQ];
    if (v.size() != cs.chains.size()) v.assign(cs.chains.size(), -1);
    if (v[chain] < 0) v[chain] = chainAscends(cs, (size_t)chain, onQ) ? 1 : 0;
    return v[chain] == 1;
}

static void calcOtherRange(Fill &fill, const ChainSet &cs, bool isQ)
{
    const ChainHead &h = cs.chains[fill.chain];
    const bool isRev = h.qStrand == '-';
    int clipStart = fill.start, clipEnd = fill.end;
    if (isQ && isRev) { const int t = clipStart; clipStart = h.qSize - clipEnd; clipEnd = h.qSize - t; }
    int tMin = INT_MAX, tMax = -INT_MAX, qMin = INT_MAX, qMax = -INT_MAX;
    const gat_block *b = cs.blocks.data() + h.firstBlock;
    const uint64_t from = ascendingChain(cs, fill.chain, isQ) ? firstBlockEndingAfter(cs, (size_t)fill.chain, clipStart, isQ) : 0;
    for (uint64_t i = from; i < h.nBlocks; i++) {
        int ts = b[i].tStart, te = ts + (int)b[i].size, qs = b[i].qStart, qe = qs + (int)b[i].size;
        if (isQ) {
            if (qe <= clipStart) continue;
            if (qs >= clipEnd) break;
            if (qs < clipStart) { ts += clipStart - qs; qs = clipStart; }
            if (qe > clipEnd) { te -= qe - clipEnd; qe = clipEnd; }
        } else {
            if (te <= clipStart) continue;
            if (ts >= clipEnd) break;
            if (ts < clipStart) { qs += clipStart - ts; ts = clipStart; }
            if (te > clipEnd) { qe -= te - clipEnd; te = clipEnd; }
        }
        if (qMin > qs) qMin = qs;
        if (qMax < qe) qMax = qe;
        if (tMin > ts) tMin = ts;
        if (tMax < te) tMax = te;
    }
    if (isRev) { const int t = qMin; qMin = h.qSize - qMax; qMax = h.qSize - t; }
    if (isQ) { fill.start = qMin; fill.end = qMax; fill.oStart = tMin; fill.oEnd = tMax; }
    else { fill.start = tMin; fill.end = tMax; fill.oStart = qMin; fill.oEnd = qMax; }
}

static void calcOtherRanges(int gapIx, const ChainSet &cs, bool isQ)
{
    for (int f : gaps[gapIx].fills) {
        calcOtherRange(fills[f], cs, isQ);
        for (int g : fills[f].gaps) calcOtherRanges(g, cs, isQ);
    }
}

// ---------------------------------------------------------------- output (chainNet.c:762-895)
struct NetWriter {
    const ChainSet &cs;
    const std::vector<int> &chainBases;            // chainBaseCount per chain
    bool isQ;
    FILE *f = nullptr;                             // nullptr = collecting pass
    // rescoring
    WorkList *wl = nullptr;
    const std::vector<uint32_t> *chainT = nullptr, *chainQ = nullptr;
    const std::vector<int64_t> *scores = nullptr;  // global score per job, in collection order
    size_t nextJob = 0;
    int depth = 0;

    bool ascending(int chain, bool onQ) { return ascendingChain(cs, chain, onQ); }

    int baseCountSub(int chain, int lo, int hi, bool onQ)
    {   // chainBaseCountSubT / chainBaseCountSubQ, :773-793
        const ChainHead &h = cs.chains[chain];
        int total = 0;
        const gat_block *b = cs.blocks.data() + h.firstBlock;
        if (ascending(chain, onQ)) {
            for (uint64_t i = firstBlockEndingAfter(cs, (size_t)chain, lo, onQ); i < h.nBlocks; i++) {
                const int s = onQ ? b[i].qStart : b[i].tStart, e = s + (int)b[i].size;
                if (s >= hi) break;
                const int x = std::min(e, hi) - std::max(s, lo);
                if (x > 0) total += x;
            }
            return total;
        }
        for (uint64_t i = 0; i < h.nBlocks; i++) {
            const int s = onQ ? b[i].qStart : b[i].tStart, e = s + (int)b[i].size;
            const int x = std::min(e, hi) - std::max(s, lo);
            if (x > 0) total += x;
        }
        return total;
    }

    void subchainInfo(const Fill &fill, int &subSize, double &subScore)
    {   // :795-843
        const ChainHead &h = cs.chains[fill.chain];
        const int fullSize = chainBases[fill.chain];
        int start = fill.start, end = fill.end;
        if (isQ) {
            if (h.qStrand == '-') { const int t = start; start = h.qSize - end; end = h.qSize - t; }
            if (start <= h.qStart && end >= h.qEnd) { subScore = h.score; subSize = fullSize; }
            else { subSize = baseCountSub(fill.chain, start, end, true); subScore = h.score * subSize / fullSize; }
            return;
        }
        if (start <= h.tStart && end >= h.tEnd) { subScore = h.score; subSize = fullSize; return; }
        subSize = baseCountSub(fill.chain, start, end, false);
        if (!rescore) { subScore = h.score * subSize / fullSize; return; }
        if (f == nullptr) {                         // collecting pass: queue the clipped job
            subScore = 1;
            if (subSize >= minFill) {
                const uint64_t hint = ascending(fill.chain, false) ? firstBlockEndingAfter(cs, (size_t)fill.chain, start, false) : 0;
                if (!addSubChainJob(cs, fill.chain, (*chainT)[fill.chain], (*chainQ)[fill.chain], start, end, *wl, hint))
                    errAbort("fill %d-%d of chain %d holds no aligned block", start, end, h.id);
            }
        } else {
            subScore = 1;
            if (subSize >= minFill) {
                const int64_t s = (*scores)[nextJob++];
                subScore = s <= 0 ? 1.0 : (double)s;          // getChainScore, :244-245
            }
        }
    }

    void outFill(int fillIx)
    {   // rOutputFill, :858-878
        const Fill &fill = fills[fillIx];
        const ChainHead &h = cs.chains[fill.chain];
        int subSize;
        double subScore;
        subchainInfo(fill, subSize, subScore);
        if (!(subScore >= minScore && subSize >= minFill)) return;
        ++depth;
        if (f) {
            for (int i = 0; i < depth; i++) fputc(' ', f);
            fprintf(f, "fill %d %d %s %c %d %d id %d score %1.0f ali %d\n", fill.start, fill.end - fill.start,
                    (isQ ? h.tName : h.qName).c_str(), h.qStrand, fill.oStart, fill.oEnd - fill.oStart, h.id, subScore, subSize);
        }
        for (int g : fill.gaps) {               // rOutputGap, :844-856
            ++depth;
            if (f) {
                for (int i = 0; i < depth; i++) fputc(' ', f);
                fprintf(f, "gap %d %d %s %c %d %d\n", gaps[g].start, gaps[g].end - gaps[g].start,
                        (isQ ? h.tName : h.qName).c_str(), h.qStrand, gaps[g].oStart, gaps[g].oEnd - gaps[g].oStart);
            }
            for (int sub : gaps[g].fills) outFill(sub);
            --depth;
        }
        --depth;
    }

    void outSide(const std::vector<Chrom> &chroms)
    {   // outputNetSide, :880-895
        for (const Chrom &c : chroms) {
            depth = 0;
            if (gaps[c.root].fills.empty()) continue;
            if (f) fprintf(f, "net %s %d\n", c.name.c_str(), c.size);
            for (int fi : gaps[c.root].fills) outFill(fi);
        }
    }
};

static void readSizes(const char *path, std::vector<Chrom> &chroms, std::map<std::string, int> &index)
{   // makeChroms, chainNet.c:328-352
    FILE *f = fopen(path, "r");
    if (!f) errAbort("Couldn't open %s , %s", path, strerror(errno));
    char line[4096];
    int lineIx = 0;
    while (fgets(line, sizeof line, f)) {
        lineIx++;
        char name[2048], num[64];
        if (line[0] == '#') continue;
        const int n = sscanf(line, "%2047s %63s", name, num);
        if (n <= 0) continue;
        if (n < 2) errAbort("Expecting 2 words line %d of %s got %d", lineIx, path, n);
        if (index.count(name)) errAbort("Duplicate %s in %s", name, path);
        if (num[0] != '-' && !isdigit((unsigned char)num[0])) errAbort("Expecting number field 2 line %d of %s, got %s", lineIx, path, num);
        Chrom c;
        c.name = name;
        c.size = atoi(num);
        c.root = newGap(0, c.size, 0, 0);
        c.spaces[0] = Space{c.size, c.root};
        index[name] = (int)chroms.size();
        chroms.push_back(std::move(c));
    }
    fclose(f);
}

static int toolMain(int argc, char **argv)
{
    Options opt;
    opt.init(&argc, argv, optionSpecs);
    if (argc != 6) usage();
    minSpace = opt.intVal("minSpace", minSpace);
    minFill = opt.intVal("minFill", minSpace / 2);
    minScore = opt.intVal("minScore", (int)minScore);        // optionInt in the reference (:1016)
    inclHap = opt.exists("inclHap");
    rescore = opt.exists("rescore");
    const char *tNibDir = nullptr, *qNibDir = nullptr, *scoreSchemeName = nullptr, *gapFileName = nullptr;
    ScoreScheme scheme = ScoreScheme::defaultScheme();
    GapCalc gapCalc;
    if (rescore) {
        minScore = 0;
        tNibDir = opt.val("tNibDir", nullptr);
        qNibDir = opt.val("qNibDir", nullptr);
        if (!tNibDir) errAbort("With -rescore you must specify the target genome file (parameter -tNibDir)\n");
        if (!qNibDir) errAbort("With -rescore you must specify the query genome file (parameter -qNibDir)\n");
        gapFileName = opt.val("linearGap", nullptr);
        scoreSchemeName = opt.val("scoreScheme", nullptr);
        if (scoreSchemeName) {
            verbose(1, "Reading scoring matrix from %s\n", scoreSchemeName);
            scheme = ScoreScheme::read(scoreSchemeName);
        }
        if (!gapFileName) errAbort("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults\n");
        gapCalc = GapCalc::fromFile(gapFileName);
        verbose(1, "-rescore is set: read target/query genome from %s and %s. scoreSchemeName %s. gap costs %s.\n", tNibDir, qNibDir,
                scoreSchemeName ? scoreSchemeName : "default", gapFileName);
    }

    const char *chainFile = argv[1], *tSizes = argv[2], *qSizes = argv[3], *tNet = argv[4], *qNet = argv[5];
    ChainSet cs;
    phaseDone("start");
    std::unique_ptr<GpuStarter> gpuStarter;             // -rescore: the CUDA contexts come up while the nets are built
    if (rescore) gpuStarter.reset(new GpuStarter(opt.intVal("gpus", 1)));
    readChains(chainFile, cs);
    FILE *tNetFile = strcmp(tNet, "stdout") == 0 ? stdout : fopen(tNet, "w");
    if (!tNetFile) errAbort("mustOpen: Can't open %s to write: %s", tNet, strerror(errno));
    FILE *qNetFile = strcmp(qNet, "stdout") == 0 ? stdout : fopen(qNet, "w");
    if (!qNetFile) errAbort("mustOpen: Can't open %s to write: %s", qNet, strerror(errno));

    std::vector<Chrom> qChroms, tChroms;
    std::map<std::string, int> qIndex, tIndex;
    readSizes(qSizes, qChroms, qIndex);
    readSizes(tSizes, tChroms, tIndex);
    verbose(1, "Got %d chroms in %s, %d in %s\n", (int)tChroms.size(), tSizes, (int)qChroms.size(), qSizes);

    phaseDone("chains read");
    // build the nets, best chain first (chainNet.c:941-975)
    double lastScore = -1;
    size_t consumed = cs.chains.size();
    std::vector<char> used(cs.chains.size(), 0);
    for (size_t c = 0; c < cs.chains.size(); c++) {
        const ChainHead &h = cs.chains[c];
        if (lastScore >= 0 && h.score > lastScore) errAbort("%s must be sorted in order of score", chainFile);
        lastScore = h.score;
        if (h.score < minScore) { consumed = c + 1; break; }
        verbose(2, "chain %f (%d els) %s %d-%d %c %s %d-%d\n", h.score, (int)h.nBlocks, h.tName.c_str(), h.tStart, h.tEnd, h.qStrand,
                h.qName.c_str(), h.qStart, h.qEnd);
        auto qi = qIndex.find(h.qName);
        if (qi == qIndex.end()) errAbort("%s not found", h.qName.c_str());
        if (qChroms[qi->second].size != h.qSize)
            errAbort("%s is %d in %s but %d in %s", h.qName.c_str(), h.qSize, chainFile, qChroms[qi->second].size, qSizes);
        auto ti = tIndex.find(h.tName);
        if (ti == tIndex.end()) errAbort("%s not found", h.tName.c_str());
        if (tChroms[ti->second].size != h.tSize)
            errAbort("%s is %d in %s but %d in %s", h.tName.c_str(), h.tSize, chainFile, tChroms[ti->second].size, tSizes);
        if (!inclHap && (h.qName.find("_hap") != std::string::npos || h.qName.find("_alt") != std::string::npos)) {
            verbose(2, "skipping chain on query %s\n", h.qName.c_str());
            continue;
        }
        addChainSide(qChroms[qi->second], cs, (int)c, true);
        addChainSide(tChroms[ti->second], cs, (int)c, false);
        used[c] = 1;
    }
    // '#' lines travel to both nets as the reader meets them (lineFileSetMetaDataOutput, :938-939)
    for (size_t i = 0; i < cs.metaLines.size(); i++)
        if (cs.metaLineChain[i] < consumed || consumed == cs.chains.size()) {
            fprintf(tNetFile, "%s\n", cs.metaLines[i].c_str());
            fprintf(qNetFile, "%s\n", cs.metaLines[i].c_str());
        }

    phaseDone("chains added to the nets");
    verbose(1, "Finishing nets\n");
    for (Chrom &c : qChroms)
        if (!gaps[c.root].fills.empty()) { sortNet(c.root); calcOtherRanges(c.root, cs, true); }
    for (Chrom &c : tChroms)
        if (!gaps[c.root].fills.empty()) { sortNet(c.root); calcOtherRanges(c.root, cs, false); }

    phaseDone("nets sorted, other ranges");
    std::vector<int> chainBases(cs.chains.size(), 0);       // chainBaseCount, :762-771
    for (size_t c = 0; c < cs.chains.size(); c++)
        for (uint64_t i = 0; i < cs.chains[c].nBlocks; i++) chainBases[c] += (int)cs.blocks[cs.chains[c].firstBlock + i].size;

    // -rescore: walk the target net once to collect the partial fills, score them in one batch
    WorkList wl;
    std::vector<int64_t> global, local;
    std::vector<uint32_t> chainT(cs.chains.size(), 0), chainQ(cs.chains.size(), 0);
    if (rescore) {
        TwoBitFile tbT(tNibDir), tbQ(qNibDir);
        std::vector<int> useT, useQ, mapT(tbT.seqs().size(), -1), mapQ(tbQ.seqs().size(), -1);
        for (size_t c = 0; c < cs.chains.size(); c++) {
            if (!used[c]) continue;
            const ChainHead &h = cs.chains[c];
            const int ti = tbT.find(h.tName), qi = tbQ.find(h.qName);
            if (ti < 0) errAbort("%s is not in %s", h.tName.c_str(), tNibDir);
            if (qi < 0) errAbort("%s is not in %s", h.qName.c_str(), qNibDir);
            if ((size_t)ti >= mapT.size()) mapT.resize(ti + 1, -1);
        if ((size_t)qi >= mapQ.size()) mapQ.resize(qi + 1, -1);
        if (mapT[ti] < 0) { mapT[ti] = (int)useT.size(); useT.push_back(ti); }
            if (mapQ[qi] < 0) { mapQ[qi] = (int)useQ.size(); useQ.push_back(qi); }
            chainT[c] = (uint32_t)mapT[ti];
            chainQ[c] = (uint32_t)mapQ[qi];
        }
        buildRecords(cs, wl);
        NetWriter collect{cs, chainBases, false};
        collect.wl = &wl; collect.chainT = &chainT; collect.chainQ = &chainQ;
        collect.outSide(tChroms);
        phaseDone("partial fills collected");
        verbose(2, "rescoring %d partial fills (%llu job-blocks) on the GPU\n", (int)wl.jobs.size(), (unsigned long long)wl.totalJobBlocks);
        if (!wl.jobs.empty()) {
            MultiGpu &gpus = gpuStarter->get();
            gpus.prepare(tbT, useT, tbQ, useQ, scheme, gapCalc);
            phaseDone("contexts + genomes");
            gpus.score(wl, global, local);
            phaseDone("fills rescored on the GPU");
        }
    }

    verbose(1, "writing %s\n", tNet);
    NetWriter tw{cs, chainBases, false, tNetFile};
    tw.scores = &global;
    tw.outSide(tChroms);
    verbose(1, "writing %s\n", qNet);
    NetWriter qw{cs, chainBases, true, qNetFile};
    qw.outSide(qChroms);
    if (tNetFile != stdout) fclose(tNetFile);
    if (qNetFile != stdout) fclose(qNetFile);
    phaseDone("nets written");
    return 0;
}

int main(int argc, char **argv) { return runTool(toolMain, argc, argv); }
