// chainCleaner -- remove chain-breaking alignments (suspects) from chains that break nested chains.
// Drop-in for src/chainCleaner/chainCleaner.c of hillerlab/GenomeAlignmentTools: same command line,
// same decisions, same output files.  Break detection and the suspect loop are host logic and
// follow the reference rule by rule (cited below), including the order in which kent's hash
// tables hand out their elements, because that order decides which suspect meets which state of a
// chain.  What changes is the scoring: the reference makes four chainSubsetOnT + chainCalcScore +
// chainCalcScoreLocal CPU calls per tested suspect (chainCleaner.c:1214-1229) and re-scores every
// modified chain at the end (:634-637).  Here the four sub-chains of ALL suspects are scored
// speculatively in one GPU batch before the loop starts; the loop then replays the reference's
// sequential decisions and only goes back to the GPU (one small batch per pass of a breaking
// chain) for suspects whose chains or fill ranges have changed since they were scored.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <tuple>
#include <unistd.h>
#include "gat_host.hpp"

using namespace gathost;

static const std::vector<OptionSpec> optionSpecs = {
    {"net", OPTION_STRING}, {"tSizes", OPTION_STRING}, {"qSizes", OPTION_STRING}, {"scoreScheme", OPTION_STRING},
    {"linearGap", OPTION_STRING}, {"debug", OPTION_BOOLEAN}, {"foldThreshold", OPTION_DOUBLE}, {"LRfoldThreshold", OPTION_DOUBLE},
    {"LRfoldThresholdPairs", OPTION_DOUBLE}, {"maxSuspectBases", OPTION_DOUBLE}, {"maxSuspectScore", OPTION_DOUBLE},
    {"minBrokenChainScore", OPTION_DOUBLE}, {"minLRGapSize", OPTION_INT}, {"doPairs", OPTION_BOOLEAN}, {"maxPairDistance", OPTION_INT},
    {"newChainIDDict", OPTION_STRING}, {"suspectDataFile", OPTION_STRING}, {"onlyThisChr", OPTION_STRING},
    {"onlyThisStart", OPTION_INT}, {"onlyThisEnd", OPTION_INT}, {"gpus", OPTION_INT},
};

// thresholds (chainCleaner.c:95-114)
static double LRfoldThreshold = 2.5, foldThreshold = 0, maxSuspectBases = INT_MAX, maxSuspectScore = 100000,
              minBrokenChainScore = 50000, LRfoldThresholdPairs = 10;
static int minLRGapSize = 0, maxPairDistance = 10000;
static bool doPairs = false;
static const char *onlyThisChr = nullptr;
static int onlyThisStart = -1, onlyThisEnd = -1;

static void usage()
{   // chainCleaner.c:198-245
    errAbort(
        "chainCleaner - Remove chain-breaking alignments from chains that break nested chains.\n"
        "\n"
        "NOTATION: The \"breaking chain\" contains a local alignment block (called \"chain-breaking alignment\" (CBA) or \"suspect\") that breaks a nested chain (\"broken chain\") into two nets.\n"
        "\n"
        "usage:\n"
        "   chainCleaner in.chain tNibDir qNibDir out.chain out.bed -net=in.net \n"
        " OR \n"
        "   chainCleaner in.chain tNibDir qNibDir out.chain out.bed -tSizes=/dir/to/target/chrom.sizes -qSizes=/dir/to/query/chrom.sizes \n"
        " First option:   you have netted the chains and specify the net file via -net=netFile\n"
        " Second option:  you have not netted the chains. Then chainCleaner will net them. In this case, you must specify the chrom.sizes file for the target and query with -tSizes/-qSizes\n"
        " tNibDir/qNibDir are either directories with nib files, or the name of a .2bit file\n\n"
        "\n"
        "output:\n"
        "   out.chain      output file in chain format containing the untouched chains, the original broken chain and the modified breaking chains. NOTE: this file is chainSort-ed.\n"
        "   out.bed        output file in bed format containing the coords and information about the removed chain-breaking alignments.\n"
        "\n"
        "Most important options for deciding which chain-breaking alignments (CBA) to remove:\n"
        "   -LRfoldThreshold=N        threshold for removing local alignment blocks if the score of the left and right fill of brokenChain / CBA score is at least this fold threshold. Default %1.1f\n"
        "   -doPairs                  flag: if set, do test if pairs of CBAs can be removed\n"
        "   -LRfoldThresholdPairs=N   threshold for removing local alignment blocks if the score of the left and right fill of brokenChain / CBA score is at least this fold threshold. Default %1.1f\n"
        "   -maxPairDistance=N        only consider pairs of CBAs where the distance between the end of the upstream CBA and the start of the downstream CBA is at most that many bp (Default %d)\n"
        "\n"
        "   -scoreScheme=fileName       Read the scoring matrix from a blastz-format file\n"
        "   -linearGap=<medium|loose|filename> Specify type of linearGap to use.\n"
        "              *Must* specify this argument to one of these choices.\n"
        "              loose is chicken/human linear gap costs.\n"
        "              medium is mouse/human linear gap costs.\n"
        "              Or specify a piecewise linearGap tab delimited file.\n"
        "   sample linearGap file (loose)\n"
        "%s"
        "\n"
        "\n"
        "Other options for deciding which suspects to remove: \n"
        "   -foldThreshold=N          threshold for removing local alignment blocks if the brokenChain score / suspect score is at least this fold threshold. Default %1.1f\n"
        "   -maxSuspectBases=N        threshold for number of target bases in aligning blocks of the suspect subChain. If higher, do not remove suspect. Default %d\n"
        "   -maxSuspectScore=N        threshold for score of suspect subChain. If higher, do not remove suspect. Default %d\n"
        "   -minBrokenChainScore=N    threshold for minimum score of the entire broken chain. If the broken chain scores lower, it is less likely to be a real alignment and we will not remove the suspect. Default %d\n"
        "   -minLRGapSize=N           threshold for min size of left/right gap (how far the suspect is away from other blocks in the breaking chain). If lower, do not remove suspect (suspect to close to left or right part of breaking chain). Default %d\n"
        "\n"
        "\n"
        "Debug and testing options: \n"
        "   -newChainIDDict=fileName  output 'newChainID{tab}breakingChainID' to this file. Gives a dictionary of the new IDs of chains representing removed suspects and the chain ID of the breaking chain that had the suspect before.\n"
        "   -suspectDataFile=fileName output all the data for suspects to this file in bed format. If set, we do not clean any suspect as this would lead to updating the suspect values (updating the L/R fill region).\n"
        "   -debug                    produces output chain files with the suspect and broken chains, and a bed file with information about all possible suspects. For debugging.\n"
        "   -gpus=N                   (B200 build) shard the scoring batches over N GPUs, default 1\n",
        LRfoldThreshold, LRfoldThresholdPairs, maxPairDistance, GapCalc::sampleFileContents(), foldThreshold, (int)maxSuspectBases,
        (int)maxSuspectScore, (int)minBrokenChainScore, minLRGapSize);
}

// ---------------------------------------------------------------- kent hash order
// The reference keys several hashes by the decimal chain id and walks them with hashTraverseEls /
// hashElListHash (kent/src/lib/hash.c:493-503, 555-569).  Order = buckets ascending, newest first
// inside a bucket; bucket = hashString(name) & (size-1) (hash.c:41-53); tables start at 2^12 and
// double whenever elCount exceeds size (hash.c:136-140, 357-366), which keeps relative order.
class KentOrder {
public:
    bool add(int key)
    {   // hashAdd of a key that is not present yet; returns false if it already was
        if (!present.insert(std::make_pair(key, 1)).second) return false;
        keys.push_back(key);
        return true;
    }
    std::vector<int> traverse() const
    {   // hashTraverseEls order
        size_t power = 12;
        while (keys.size() > ((size_t)1 << power)) power++;      // elCount > size -> double
        const uint32_t mask = (1u << power) - 1;
        std::vector<std::vector<int>> bucket((size_t)1 << power);
        for (int k : keys) {
            char buf[32];
            snprintf(buf, sizeof buf, "%d", k);
            uint32_t h = 0;
            for (const char *s = buf; *s; ++s) h += (h << 3) + (unsigned char)*s;
            bucket[h & mask].push_back(k);
        }
        std::vector<int> out;
        out.reserve(keys.size());
        for (const auto &b : bucket)
            for (auto it = b.rbegin(); it != b.rend(); ++it) out.push_back(*it);
        return out;
    }
    std::vector<int> elList() const
    {   // hashElListHash: slAddHead while traversing = traversal order reversed
        std::vector<int> v = traverse();
        std::reverse(v.begin(), v.end());
        return v;
    }
    bool has(int key) const { return present.count(key) != 0; }
private:
    std::vector<int> keys;
    std::map<int, int> present;
};

// ---------------------------------------------------------------- net file (kent/src/hg/lib/chainNet.c:86-285)
struct NetNode { int tStart, tSize, chainId; std::vector<int> children; };
struct Net { std::string name; int size; std::vector<int> roots; };
static std::vector<NetNode> netNodes;

static void readNets(const char *path, std::vector<Net> &nets)
{
    FILE *f = fopen(path, "r");
    if (!f) errAbort("Couldn't open %s , %s", path, strerror(errno));
    std::vector<std::pair<int, int>> stack;        // (depth = leading spaces, node) of the open ancestors
    char *line = nullptr;
    size_t cap = 0;
    int lineIx = 0;
    while (getline(&line, &cap, f) > 0) {
        lineIx++;
        char *s = line;
        while (*s == ' ' || *s == '\t') s++;
        if (*s == 0 || *s == '\n' || *s == '#') continue;          // lineFileNextReal
        if (strncmp(line, "net ", 4) == 0) {
            char name[1024];
            int size;
            if (sscanf(line + 4, "%1023s %d", name, &size) != 2) errAbort("Expecting at least 3 words line %d of %s", lineIx, path);
            nets.push_back(Net{name, size, {}});
            stack.clear();
            continue;
        }
        if (nets.empty()) errAbort("Expecting 'net' first word of line %d of %s", lineIx, path);
        int depth = 0;
        while (line[depth] == ' ') depth++;
        char type[32], qName[1024], strand[8];
        int tStart, tSize, qStart, qSize, used = 0;
        if (sscanf(line, " %31s %d %d %1023s %7s %d %d%n", type, &tStart, &tSize, qName, strand, &qStart, &qSize, &used) < 7)
            errAbort("Expecting at least 7 words line %d of %s", lineIx, path);
        int chainId = 0;
        char key[64], val[256];
        const char *rest = line + used;
        int adv;
        while (sscanf(rest, " %63s %255s%n", key, val, &adv) == 2) {    // key/value pairs; only "id" matters here
            if (strcmp(key, "id") == 0) chainId = atoi(val);
            rest += adv;
        }
        netNodes.push_back(NetNode{tStart, tSize, chainId, {}});
        const int node = (int)netNodes.size() - 1;
        while (!stack.empty() && stack.back().first >= depth) stack.pop_back();
        if (stack.empty()) nets.back().roots.push_back(node);
        else netNodes[stack.back().second].children.push_back(node);
        stack.push_back(std::make_pair(depth, node));
    }
    free(line);
    fclose(f);
}

// ---------------------------------------------------------------- fills, gaps, breaks
struct FillGap {    // struct fillGapInfo, chainCleaner.c:57-72
    int depth, gapDepth, chainId, parentChainId, chrom, fillStart, fillEnd, gapStart, gapEnd;
};
struct Break {      // struct breakInfo, chainCleaner.c:76-92
    int depth, chainId, parentChainId, chrom;
    int LfillStart, LfillEnd, RfillStart, RfillEnd, LgapStart, LgapEnd, RgapStart, RgapEnd, suspectStart, suspectEnd;
};
typedef std::list<Break> BreakList;

static KentOrder chainId2Count;                               // chains seen at depth > 1, in first-seen order
static std::map<int, std::vector<FillGap>> fillGapsOf;
static KentOrder breakHashOrder, chainsOfInterest;
static std::map<int, BreakList> breaksOf;                     // keyed by breaking (parent) chain id

struct GapState { int chrom, start, end, chainId, depth; };

static void parseFill(const std::vector<int> &list, int depth, int chrom, std::vector<GapState> &depth2gap, std::vector<int> &depth2chain)
{   // chainCleaner.c:786-856
    if (depth + 1 >= (int)depth2gap.size()) { depth2gap.resize(depth + 2); depth2chain.resize(depth + 2); }
    for (int n : list) {
        const NetNode &node = netNodes[n];
        if (node.chainId) {
            depth2chain[depth] = node.chainId;
            if (depth > 1) {
                chainId2Count.add(node.chainId);
                const GapState &g = depth2gap[depth - 1];
                fillGapsOf[node.chainId].push_back(FillGap{depth, g.depth, node.chainId, g.chainId, chrom, node.tStart,
                                                           node.tStart + node.tSize, g.start, g.end});
            }
        } else
            depth2gap[depth] = GapState{chrom, node.tStart, node.tStart + node.tSize, depth2chain[depth - 1], depth};
        if (!node.children.empty()) parseFill(node.children, depth + 1, chrom, depth2gap, depth2chain);
    }
}

// Aligning blocks of the nets (chainCleaner.c:688-762) in a range tree whose overlapping ranges
// merge and pool their chain ids (rangeTreeAddValList, kent/src/lib/rangeTree.c:44-94).  A merged
// range only has to answer "is there an id below X other than P", so two smallest ids suffice.
struct AliRange { int start, end, min1, min2; };
static std::vector<std::vector<AliRange>> aliRanges;          // per net, merged and sorted

static void collectAliBlocks(const std::vector<int> &list, std::vector<AliRange> &out)
{
    for (int n : list) {
        const NetNode &fill = netNodes[n];
        if (fill.chainId) {
            int tStart = fill.tStart;
            for (int c : fill.children) {
                const NetNode &child = netNodes[c];
                if (child.children.empty()) continue;                 // nextGapWithInsert
                out.push_back(AliRange{tStart, child.tStart, fill.chainId, INT_MAX});
                tStart = child.tStart + child.tSize;
            }
            out.push_back(AliRange{tStart, fill.tStart + fill.tSize, fill.chainId, INT_MAX});
        }
        if (!fill.children.empty()) collectAliBlocks(fill.children, out);
    }
}

static void mergeAliRanges(std::vector<AliRange> &v)
{
    v.erase(std::remove_if(v.begin(), v.end(), [](const AliRange &r) { return r.start >= r.end; }), v.end());
    std::sort(v.begin(), v.end(), [](const AliRange &a, const AliRange &b) { return a.start < b.start; });
    std::vector<AliRange> out;
    for (const AliRange &r : v) {
        if (!out.empty() && r.start < out.back().end) {           // strict overlap merges (rangeCmp, rangeTree.c:21-33)
            AliRange &m = out.back();
            m.end = std::max(m.end, r.end);
            for (int id : {r.min1, r.min2}) {
                if (id == m.min1 || id == m.min2) continue;
                if (id < m.min1) { m.min2 = m.min1; m.min1 = id; }
                else if (id < m.min2) m.min2 = id;
            }
        } else out.push_back(r);
    }
    v.swap(out);
}

static bool brokenByAnotherHigherScoringChain(int chrom, int start, int end, int chainId, int parentChainId)
{   // chainCleaner.c:864-883
    const std::vector<AliRange> &v = aliRanges[chrom];
    auto it = std::lower_bound(v.begin(), v.end(), start, [](const AliRange &r, int s) { return r.end <= s; });
    for (; it != v.end() && it->start < end; ++it) {
        if (it->end <= start) continue;
        if ((it->min1 < chainId && it->min1 != parentChainId) || (it->min2 < chainId && it->min2 != parentChainId)) return true;
    }
    return false;
}

static Break newBreak(int depth, int chainId, int parentChainId, int chrom, int LfillStart, int LfillEnd, int RfillStart, int RfillEnd,
                      int LgapStart, int LgapEnd, int RgapStart, int RgapEnd)
{   // chainCleaner.c:910-941
    return Break{depth, chainId, parentChainId, chrom, LfillStart, LfillEnd, RfillStart, RfillEnd, LgapStart, LgapEnd, RgapStart, RgapEnd,
                 LgapEnd, RgapStart};
}

static void getValidBreaks(int chainId, const std::vector<Net> &nets)
{   // chainCleaner.c:969-1086
    std::vector<FillGap> &list = fillGapsOf[chainId];
    if (list.size() <= 1) return;
    for (size_t i = 0; i + 1 < list.size(); i++) {
        const FillGap &a = list[i], &b = list[i + 1];
        if (onlyThisChr && (nets[a.chrom].name != onlyThisChr || onlyThisStart != a.gapEnd || onlyThisEnd != b.gapStart)) continue;
        if (a.depth != b.depth) continue;
        if (a.parentChainId != b.parentChainId) continue;
        if (brokenByAnotherHigherScoringChain(a.chrom, a.fillEnd, b.fillStart, a.chainId, a.parentChainId)) continue;
        if (a.gapStart == b.gapStart && a.gapEnd == b.gapEnd) continue;
        chainsOfInterest.add(a.chainId);
        chainsOfInterest.add(a.parentChainId);
        breakHashOrder.add(a.parentChainId);
        breaksOf[a.parentChainId].push_back(newBreak(a.depth, a.chainId, a.parentChainId, a.chrom, a.fillStart, a.fillEnd, b.fillStart,
                                                     b.fillEnd, a.gapStart, a.gapEnd, b.gapStart, b.gapEnd));
    }
}

// ---------------------------------------------------------------- chains of interest + GPU scoring
struct LiveChain {
    ChainHead head;
    std::vector<gat_block> blocks;      // current blocks (suspects get removed)
    uint32_t tSeq = 0, qSeq = 0;
    std::vector<std::pair<int, int>> removed;   // target ranges whose blocks were removed, in order of removal
    mutable int monotone = -1;          // blocks ascending and disjoint on the target (-1: not checked yet)
};
static std::map<int, LiveChain> live;   // by chain id
static int maxChainId = -1;

struct SubScore { bool isNull = true, whole = false; double global = 0, local = 0; int bases = 0; };

// chainSubsetOnT on the CURRENT blocks of a chain (chain.c:471-558): selection + the header of the sub-chain
struct SubSel { bool isNull, whole; size_t first, count; };
static SubSel selectSub(const LiveChain &c, int subStart, int subEnd)
{
    if (subStart <= c.head.tStart && subEnd >= c.head.tEnd) return SubSel{c.blocks.empty(), true, 0, c.blocks.size()};
    if (c.monotone < 0) {
        c.monotone = 1;
        for (size_t i = 1; i < c.blocks.size(); i++)
            if (c.blocks[i].tStart < c.blocks[i - 1].tStart + (int)c.blocks[i - 1].size) { c.monotone = 0; break; }
    }
    size_t a = 0, e;
    if (c.monotone) {   // the reference walks the list from its head (chain.c:479-510); on sorted blocks a bisection finds the same range
        a = (size_t)(std::partition_point(c.blocks.begin(), c.blocks.end(),
                                          [&](const gat_block &b) { return b.tStart + (int)b.size <= subStart; }) - c.blocks.begin());
        e = (size_t)(std::partition_point(c.blocks.begin() + (long)a, c.blocks.end(),
                                          [&](const gat_block &b) { return b.tStart < subEnd; }) - c.blocks.begin());
    } else {
        while (a < c.blocks.size() && c.blocks[a].tStart + (int)c.blocks[a].size <= subStart) a++;
        e = a;
        while (e < c.blocks.size() && c.blocks[e].tStart < subEnd) e++;
    }
    return SubSel{e == a, false, a, e - a};
}

struct Request { int chainId, subStart, subEnd; };

class Scorer {
public:
    Scorer(GpuStarter &starter, const TwoBitFile &tbT, const TwoBitFile &tbQ, const ScoreScheme &ss, const GapCalc &gc) : gpus(starter.get())
    {
        // like loadTandQSeqs (chainCleaner.c:463-480) only the sequences of chains of interest go to the GPU
        std::vector<int> useT, useQ, mapT(tbT.seqs().size(), -1), mapQ(tbQ.seqs().size(), -1);
        for (auto &kv : live) {
            LiveChain &c = kv.second;
            const int ti = tbT.find(c.head.tName), qi = tbQ.find(c.head.qName);
            if (ti < 0) errAbort("%s is not in %s", c.head.tName.c_str(), tbT.path().c_str());
            if (qi < 0) errAbort("%s is not in %s", c.head.qName.c_str(), tbQ.path().c_str());
            if ((size_t)ti >= mapT.size()) mapT.resize(ti + 1, -1);
        if ((size_t)qi >= mapQ.size()) mapQ.resize(qi + 1, -1);
        if (mapT[ti] < 0) { mapT[ti] = (int)useT.size(); useT.push_back(ti); }
            if (mapQ[qi] < 0) { mapQ[qi] = (int)useQ.size(); useQ.push_back(qi); }
            c.tSeq = (uint32_t)mapT[ti];
            c.qSeq = (uint32_t)mapQ[qi];
        }
        gpus.prepare(tbT, useT, tbQ, useQ, ss, gc);
    }
    // One batch: every request is one chainSubsetOnT + getChainScore of the reference.
    std::vector<SubScore> score(const std::vector<Request> &reqs)
    {
        std::vector<SubScore> out(reqs.size());
        // every request becomes a chain of its own holding just the blocks chainSubsetOnT would keep: what is
        // copied, split into records and uploaded is proportional to the sub-chains, not to the chains they come from
        ChainSet cs;
        std::vector<size_t> chainOf(reqs.size(), (size_t)-1);
        for (size_t i = 0; i < reqs.size(); i++) {
            const LiveChain &c = live.at(reqs[i].chainId);
            const SubSel sel = selectSub(c, reqs[i].subStart, reqs[i].subEnd);
            out[i].whole = sel.whole;
            if (sel.isNull) continue;
            ChainHead h = c.head;
            h.firstBlock = cs.blocks.size();
            h.nBlocks = sel.count;
            cs.blocks.insert(cs.blocks.end(), c.blocks.begin() + (long)sel.first, c.blocks.begin() + (long)(sel.first + sel.count));
            chainOf[i] = cs.chains.size();
            cs.chains.push_back(h);
        }
        WorkList wl;
        buildRecords(cs, wl);
        std::vector<size_t> jobOf(reqs.size(), (size_t)-1);
        for (size_t i = 0; i < reqs.size(); i++) {
            if (chainOf[i] == (size_t)-1) continue;
            const LiveChain &c = live.at(reqs[i].chainId);
            if (!addSubChainJob(cs, chainOf[i], c.tSeq, c.qSeq, reqs[i].subStart, reqs[i].subEnd, wl)) continue;
            jobOf[i] = wl.jobs.size() - 1;
        }
        std::vector<int64_t> global, local;
        gpus.score(wl, global, local);
        gpuCalls++;
        gpuJobs += wl.jobs.size();
        for (size_t i = 0; i < reqs.size(); i++)
            if (jobOf[i] != (size_t)-1) {
                out[i].isNull = false;
                out[i].global = (double)global[jobOf[i]];
                out[i].local = (double)local[jobOf[i]];
                out[i].bases = (int)wl.aliBases[jobOf[i]];
            }
        return out;
    }
    size_t gpuCalls = 0, gpuJobs = 0;
private:
    MultiGpu &gpus;
};

// The four sub-chains of one tested suspect (chainCleaner.c:1214-1217).  Their scores stay valid until blocks are
// removed from the breaking chain inside the suspect range, or from the broken chain inside the fill range: a
// removal elsewhere leaves the blocks chainSubsetOnT selects, and the gaps between them, as they were.
struct TestKey {
    int parentId, brokenId, suspectStart, suspectEnd, LfillStart, RfillEnd;
    bool operator<(const TestKey &o) const
    {
        return std::tie(parentId, brokenId, suspectStart, suspectEnd, LfillStart, RfillEnd) <
               std::tie(o.parentId, o.brokenId, o.suspectStart, o.suspectEnd, o.LfillStart, o.RfillEnd);
    }
};
struct TestScores {
    SubScore suspect, fill, lfill, rfill;
    size_t parentSeen = 0, brokenSeen = 0;      // removals of either chain already reflected in the scores
    bool scored = false;
};
static std::map<TestKey, TestScores> cache;

static TestKey keyOf(const Break &b)
{
    return TestKey{b.parentChainId, b.chainId, b.suspectStart, b.suspectEnd, b.LfillStart, b.RfillEnd};
}

static bool touchedSince(const LiveChain &c, size_t seen, int start, int end)
{
    for (size_t i = seen; i < c.removed.size(); i++)
        if (c.removed[i].first < end && c.removed[i].second > start) return true;
    return false;
}

// is the cache entry of this break still what a fresh scoring would give?
static bool cacheValid(const Break &b)
{
    auto it = cache.find(keyOf(b));
    if (it == cache.end()) return false;
    TestScores &t = it->second;
    if (!t.scored) return true;                 // reserved inside the batch being assembled
    const LiveChain &parent = live.at(b.parentChainId), &broken = live.at(b.chainId);
    if (touchedSince(parent, t.parentSeen, b.suspectStart, b.suspectEnd) || touchedSince(broken, t.brokenSeen, b.LfillStart, b.RfillEnd)) return false;
    t.parentSeen = parent.removed.size();
    t.brokenSeen = broken.removed.size();
    return true;
}

static void requestsFor(const Break &b, std::vector<Request> &reqs)
{
    reqs.push_back(Request{b.parentChainId, b.suspectStart, b.suspectEnd});
    reqs.push_back(Request{b.chainId, b.LfillStart, b.RfillEnd});
    reqs.push_back(Request{b.chainId, b.LfillStart, b.suspectEnd});
    reqs.push_back(Request{b.chainId, b.suspectStart, b.RfillEnd});
}

// Score (speculatively) every break of `todo` that has no valid cache entry, in one GPU batch.
static void ensureScored(Scorer &scorer, const std::vector<const Break *> &todo)
{
    std::vector<Request> reqs;
    std::vector<TestKey> keys;
    for (const Break *b : todo) {
        if (cacheValid(*b)) continue;
        const TestKey k = keyOf(*b);
        cache[k] = TestScores();        // reserve, so duplicates inside this batch are requested once
        keys.push_back(k);
        requestsFor(*b, reqs);
    }
    if (reqs.empty()) return;
    const std::vector<SubScore> res = scorer.score(reqs);
    for (size_t i = 0; i < keys.size(); i++) {
        TestScores t;
        t.suspect = res[4 * i]; t.fill = res[4 * i + 1]; t.lfill = res[4 * i + 2]; t.rfill = res[4 * i + 3];
        t.parentSeen = live.at(keys[i].parentId).removed.size();
        t.brokenSeen = live.at(keys[i].brokenId).removed.size();
        t.scored = true;
        cache[keys[i]] = t;
    }
}

// ---------------------------------------------------------------- the suspect loop
static FILE *finalChainOutFile, *suspectsRemovedOutBedFile, *newChainIDDictFile, *suspectDataFilePointer;
// -debug (chainCleaner.c:1312-1321, 1817-1823): sub-chains and a bed line for every tested suspect
static bool debugMode = false;
static FILE *suspectChainFile, *brokenChainLfillChainFile, *brokenChainRfillChainFile, *brokenChainfillChainFile, *suspectFillBedFile;
static int suspectID = 0;
static std::map<int, int> needsRescoring;
static const std::vector<Net> *netsP;

static void chainRemoveBlocks(LiveChain &c, int tStart, int tEnd)
{   // chainCleaner.c:649-690
    size_t first = 0, cur = 0;
    for (; cur < c.blocks.size(); cur++) {
        if (c.blocks[cur].tStart >= tStart) break;
        first = cur;
    }
    if (first == cur)
        errAbort("ERROR in chainRemoveBlocks: boundaries imply that we remove the first block of chain Id %d (tStart %d - tEnd %d)\n", c.head.id, tStart, tEnd);
    size_t last = first + 1;
    for (; last < c.blocks.size(); last++)
        if (c.blocks[last].tStart >= tEnd) break;
    if (last >= c.blocks.size())
        errAbort("ERROR in chainRemoveBlocks: boundaries imply that we remove the last block of chain Id %d (tStart %d - tEnd %d)\n", c.head.id, tStart, tEnd);
    c.blocks.erase(c.blocks.begin() + first + 1, c.blocks.begin() + last);
    c.removed.emplace_back(tStart, tEnd);
}

// the sub-chain chainFastSubsetOnT would build (chain.c:490-558), for writing a removed suspect
static void writeSubChain(FILE *f, const LiveChain &c, int subStart, int subEnd, double score, int id)
{
    const SubSel sel = selectSub(c, subStart, subEnd);
    std::vector<gat_block> b;
    int qStart = INT_MAX, qEnd = -INT_MAX, tStart = INT_MAX, tEnd = -INT_MAX;
    for (size_t i = sel.first; i < sel.first + sel.count; i++) {
        int ts = c.blocks[i].tStart, te = ts + (int)c.blocks[i].size, qs = c.blocks[i].qStart, qe = qs + (int)c.blocks[i].size;
        if (!sel.whole) {
            if (ts < subStart) { qs += subStart - ts; ts = subStart; }
            if (te > subEnd) { qe -= te - subEnd; te = subEnd; }
        }
        b.push_back(gat_block{ts, qs, (uint32_t)(te - ts)});
        qStart = std::min(qStart, qs); qEnd = std::max(qEnd, qe); tStart = std::min(tStart, ts); tEnd = std::max(tEnd, te);
    }
    ChainHead h = c.head;
    h.score = score; h.id = id; h.firstBlock = 0; h.nBlocks = b.size();
    if (!sel.whole) { h.tStart = tStart; h.tEnd = tEnd; h.qStart = qStart; h.qEnd = qEnd; }
    writeChain(f, h, b.data());
}

// testAndRemoveSuspect, chainCleaner.c:1191-1400.  `up` / `down` may be null.
static bool testAndRemoveSuspect(Scorer &scorer, Break &br, Break *up, Break *down, bool &breaksUpdated, bool isPair, const char *debugInfo)
{
    breaksUpdated = false;
    LiveChain &breakingChain = live.at(br.parentChainId), &brokenChain = live.at(br.chainId);
    const double breakingChainScore = breakingChain.head.score, brokenChainScore = brokenChain.head.score;
    std::vector<const Break *> one(1, &br);
    ensureScored(scorer, one);
    const TestScores &t = cache.at(keyOf(br));
    if (t.suspect.isNull) {
        verbose(3, "\t\tSuspect %d-%d is apparently already deleted as the suspect subChain is NULL\n", br.suspectStart, br.suspectEnd);
        return false;
    }
    if (t.fill.isNull || t.lfill.isNull || t.rfill.isNull)
        errAbort("ERROR: fill sub-chain of broken chain %d is empty (%d-%d)", br.chainId, br.LfillStart, br.RfillEnd);
    // getChainScore stores the global score in chain->score; when chainSubsetOnT handed back the chain
    // itself (chain.c:501-506) that overwrites the broken chain's own score (chainCleaner.c:567)
    if (t.fill.whole || t.lfill.whole || t.rfill.whole) brokenChain.head.score = t.fill.whole ? t.fill.global : (t.lfill.whole ? t.lfill.global : t.rfill.global);
    if (t.suspect.whole) breakingChain.head.score = t.suspect.global;
    const double suspectLocal = t.suspect.local;
    const double ratio = t.fill.global / suspectLocal, ratioL = t.lfill.global / suspectLocal, ratioR = t.rfill.global / suspectLocal;
    const int suspectBases = t.suspect.bases;
    verbose(3, "\t\t\tsuspect subChain            %d - %d   gets score %7d   (local score %d, suspect subChain bases %d, left gap size %d, right gap size %d)\n",
            br.suspectStart, br.suspectEnd, (int)t.suspect.global, (int)suspectLocal, suspectBases, br.LgapEnd - br.LgapStart, br.RgapEnd - br.RgapStart);

    const double lr = isPair ? LRfoldThresholdPairs : LRfoldThreshold;
    bool isRemoved = ratioL >= lr && ratioR >= lr && ratio >= foldThreshold && suspectLocal <= maxSuspectScore && suspectBases <= maxSuspectBases &&
                     brokenChainScore >= minBrokenChainScore && (br.LgapEnd - br.LgapStart) >= minLRGapSize && (br.RgapEnd - br.RgapStart) >= minLRGapSize;
    const char *chrom = (*netsP)[br.chrom].name.c_str();
    if (suspectDataFilePointer) {       // :1282-1309
        isRemoved = false;
        suspectID++;
        fprintf(suspectDataFilePointer, "%s\t%d\t%d\t%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d\n", chrom, br.suspectStart, br.suspectEnd, suspectID,
                br.parentChainId, (int)breakingChainScore, br.chainId, (int)brokenChainScore, (int)suspectLocal, (int)t.fill.global,
                (int)t.lfill.global, (int)t.rfill.global, suspectBases, br.LgapEnd - br.LgapStart, br.RgapEnd - br.RgapStart,
                (int)t.lfill.local, (int)t.rfill.local);
    }
    if (debugMode) {                    // the four sub-chains carry the scores getChainScore left in them (:559-571)
        writeSubChain(suspectChainFile, breakingChain, br.suspectStart, br.suspectEnd, t.suspect.global, breakingChain.head.id);
        writeSubChain(brokenChainLfillChainFile, brokenChain, br.LfillStart, br.suspectEnd, t.lfill.global, brokenChain.head.id);
        writeSubChain(brokenChainRfillChainFile, brokenChain, br.suspectStart, br.RfillEnd, t.rfill.global, brokenChain.head.id);
        writeSubChain(brokenChainfillChainFile, brokenChain, br.LfillStart, br.RfillEnd, t.fill.global, brokenChain.head.id);
        fprintf(suspectFillBedFile, "%s\t%d\t%d\t%s%sSuspect__score_%1.0f__Rleft_%1.2f__Rright_%1.2f\t1000\t+\t%d\t%d\t255,0,0\n", chrom, br.suspectStart,
                br.suspectEnd, isRemoved ? "REMOVED_" : "", debugInfo, suspectLocal, ratioL, ratioR, br.suspectStart, br.suspectEnd);
        fprintf(suspectFillBedFile, "%s\t%d\t%d\t%sFill__score_%1.0f\t1000\t+\t%d\t%d\t0,0,255\n", chrom, br.LfillStart, br.RfillEnd, debugInfo,
                t.fill.global, br.LfillStart, br.RfillEnd);
        fprintf(suspectFillBedFile, "%s\t%d\t%d\t%sLfill__score_%1.0f\t1000\t+\t%d\t%d\t0,125,255\n", chrom, br.LfillStart, br.suspectEnd, debugInfo,
                t.lfill.global, br.LfillStart, br.LfillEnd);
        fprintf(suspectFillBedFile, "%s\t%d\t%d\t%sRfill__score_%1.0f\t1000\t+\t%d\t%d\t0,125,255\n", chrom, br.suspectStart, br.RfillEnd, debugInfo,
                t.rfill.global, br.RfillStart, br.RfillEnd);
    }
    if (!isRemoved) {
        verbose(3, "\t\t\t===> do not remove suspect from breaking chainID %d\n", breakingChain.head.id);
        return false;
    }
    needsRescoring[breakingChain.head.id] = 1;
    verbose(3, "\t\t\t===> REMOVE suspect from breaking chainID %d (this chain will be rescored before writing)\n", breakingChain.head.id);
    fprintf(suspectsRemovedOutBedFile,
            "%s\t%d\t%d\tbreakingChainID_%d_Score_%d_brokenChainID_%d_Score_%d_suspectLocalScore_%d_RatioL_%1.2f_RatioR_%1.2f\t1000\t+\t%d\t%d\t%s\n",
            chrom, br.suspectStart, br.suspectEnd, br.parentChainId, (int)breakingChainScore, br.chainId, (int)brokenChainScore, (int)suspectLocal,
            ratioL, ratioR, br.suspectStart, br.suspectEnd, isPair ? "0,100,255" : "0,0,153");
    // the removed suspect becomes a chain of its own, scored with its global score (:1338-1347)
    maxChainId++;
    writeSubChain(finalChainOutFile, breakingChain, br.suspectStart, br.suspectEnd, t.suspect.global, maxChainId);
    if (newChainIDDictFile) fprintf(newChainIDDictFile, "%d\t%d\n", maxChainId, breakingChain.head.id);
    const double suspectGlobal = t.suspect.global;
    (void)suspectGlobal;
    chainRemoveBlocks(breakingChain, br.suspectStart, br.suspectEnd);
    // neighbouring breaks of the same chain pair inherit the merged fill (:1350-1386)
    if (up && br.chainId == up->chainId && br.parentChainId == up->parentChainId && up->RfillStart == br.LfillStart && up->RfillEnd == br.LfillEnd) {
        breaksUpdated = true;
        up->RfillEnd = br.RfillEnd;
        up->RgapEnd = br.RgapEnd;
    }
    if (down && br.chainId == down->chainId && br.parentChainId == down->parentChainId && down->LfillStart == br.RfillStart && down->LfillEnd == br.RfillEnd) {
        breaksUpdated = true;
        down->LfillStart = br.LfillStart;
        down->LgapStart = br.LgapStart;
    }
    return true;
}

static bool isValidBreakPair(const Break &up, const Break &down)
{   // chainCleaner.c:1409-1444
    if (up.parentChainId != down.parentChainId || up.chainId != down.chainId) return false;
    if (up.depth != down.depth) return false;
    if (down.suspectStart - up.suspectEnd > maxPairDistance) return false;
    return up.RgapStart == down.LgapStart && up.RgapEnd == down.LgapEnd;
}

static void loopOverBreaks(Scorer &scorer)
{   // chainCleaner.c:1452-1632
    const std::vector<int> order = breakHashOrder.elList();
    {   // speculative first pass of every breaking chain in one batch
        std::vector<const Break *> all;
        for (int parent : order)
            for (const Break &b : breaksOf[parent]) all.push_back(&b);
        ensureScored(scorer, all);
    }
    for (int parent : order) {
        BreakList &list = breaksOf[parent];
        int totalNumIteration = 0;
        char debugInfo[64];
        for (;;) {
            for (;;) {      // single breaks until a pass updates nothing
                bool anyUpdated = false;
                totalNumIteration++;
                snprintf(debugInfo, sizeof debugInfo, "SINGLE_%d", totalNumIteration);
                {
                    std::vector<const Break *> pass;
                    for (const Break &b : list) pass.push_back(&b);
                    ensureScored(scorer, pass);
                }
                for (auto it = list.begin(); it != list.end();) {
                    Break *up = it == list.begin() ? nullptr : &*std::prev(it);
                    auto nextIt = std::next(it);
                    Break *down = nextIt == list.end() ? nullptr : &*nextIt;
                    bool updated = false;
                    const bool removed = testAndRemoveSuspect(scorer, *it, up, down, updated, false, debugInfo);
                    if (updated) anyUpdated = true;
                    if (removed) list.erase(it);
                    it = nextIt;
                }
                if (!anyUpdated || list.empty()) break;
            }
            bool anyPairUpdated = false;
            if (doPairs) {
                totalNumIteration++;
                snprintf(debugInfo, sizeof debugInfo, "PAIR_%d", totalNumIteration);
                for (auto it = list.begin(); it != list.end() && std::next(it) != list.end();) {
                    auto downIt = std::next(it);
                    auto afterIt = std::next(downIt);
                    Break *before = it == list.begin() ? nullptr : &*std::prev(it);
                    Break *after = afterIt == list.end() ? nullptr : &*afterIt;
                    if (isValidBreakPair(*it, *downIt)) {
                        Break pair = newBreak(it->depth, it->chainId, it->parentChainId, it->chrom, it->LfillStart, it->LfillEnd,
                                              downIt->RfillStart, downIt->RfillEnd, it->LgapStart, it->LgapEnd, downIt->RgapStart, downIt->RgapEnd);
                        bool updated = false;
                        const bool removed = testAndRemoveSuspect(scorer, pair, before, after, updated, true, debugInfo);
                        if (updated) anyPairUpdated = true;
                        if (removed) { list.erase(it); list.erase(downIt); it = afterIt; }
                        else it = downIt;
                    } else it = downIt;
                }
            }
            if (!anyPairUpdated || list.empty()) break;
        }
    }
}

// ---------------------------------------------------------------- main
static int toolMain(int argc, char **argv)
{
    Options opt;
    opt.init(&argc, argv, optionSpecs);
    if (argc != 6) usage();
    debugMode = opt.exists("debug");     // (the reference tests its flag for the "### DEBUG mode ###" banner before reading the option, :1694/:1710: never printed)
    const char *inChainFile = argv[1], *tNibDir = argv[2], *qNibDir = argv[3], *outChainFile = argv[4], *outRemovedSuspectsFile = argv[5];
    const std::string outChainFileUnsorted = std::string(outChainFile) + ".unsorted";
    const char *inNetFile = opt.val("net", nullptr), *tSizes = opt.val("tSizes", nullptr), *qSizes = opt.val("qSizes", nullptr);
    const char *gapFileName = opt.val("linearGap", nullptr), *scoreSchemeName = opt.val("scoreScheme", nullptr);
    foldThreshold = opt.doubleVal("foldThreshold", foldThreshold);
    LRfoldThreshold = opt.doubleVal("LRfoldThreshold", LRfoldThreshold);
    LRfoldThresholdPairs = opt.doubleVal("LRfoldThresholdPairs", LRfoldThresholdPairs);
    maxSuspectBases = opt.doubleVal("maxSuspectBases", maxSuspectBases);
    maxSuspectScore = opt.doubleVal("maxSuspectScore", maxSuspectScore);
    minBrokenChainScore = opt.doubleVal("minBrokenChainScore", minBrokenChainScore);
    minLRGapSize = opt.intVal("minLRGapSize", minLRGapSize);
    doPairs = opt.exists("doPairs");
    maxPairDistance = opt.intVal("maxPairDistance", maxPairDistance);
    const char *newChainIDDict = opt.val("newChainIDDict", nullptr), *suspectDataFile = opt.val("suspectDataFile", nullptr);
    onlyThisChr = opt.val("onlyThisChr", nullptr);
    onlyThisStart = opt.intVal("onlyThisStart", -1);
    onlyThisEnd = opt.intVal("onlyThisEnd", -1);
    if (onlyThisChr) verbose(1, "ONLY %s %d %d\n", onlyThisChr, onlyThisStart, onlyThisEnd);
    verbose(1, "Verbosity level: %d\n", verboseLevel());
    verbose(1, "foldThreshold: %f    LRfoldThreshold: %f   maxSuspectBases: %d  maxSuspectScore: %d  minBrokenChainScore: %d  minLRGapSize: %d",
            foldThreshold, LRfoldThreshold, (int)maxSuspectBases, (int)maxSuspectScore, (int)minBrokenChainScore, minLRGapSize);
    if (doPairs) verbose(1, " doPairs with LRfoldThreshold: %f   maxPairDistance %d\n", LRfoldThresholdPairs, maxPairDistance);
    else verbose(1, "\n");

    ScoreScheme scheme = ScoreScheme::defaultScheme();
    if (scoreSchemeName) {
        verbose(1, "Reading scoring matrix from %s\n", scoreSchemeName);
        scheme = ScoreScheme::read(scoreSchemeName);
    }
    if (!gapFileName) errAbort("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults\n");
    const GapCalc gapCalc = GapCalc::fromFile(gapFileName);
    if (access(tNibDir, F_OK) != 0) errAbort("ERROR: target 2bit file or nib directory %s does not exist\n", tNibDir);
    if (access(qNibDir, F_OK) != 0) errAbort("ERROR: query 2bit file or nib directory %s does not exist\n", qNibDir);

    GpuStarter gpuStarter(opt.intVal("gpus", 1));      // the CUDA contexts come up while nets and chains are parsed

    // 0. net the chains ourselves if no net was given (chainCleaner.c:1639-1670)
    std::string tmpNet;
    if (!inNetFile) {
        if (!tSizes) errAbort("You must specifiy -tSizes /dir/to/target/chrom.sizes if you do not provide a net file with -net in.net\n");
        if (!qSizes) errAbort("You must specifiy -qSizes /dir/to/query/chrom.sizes if you do not provide a net file with -net in.net\n");
        if (system("which chainNet > /dev/null") != 0)
            errAbort("ERROR: chainNet (kent source code) is not a binary in $PATH. Either install the kent source code or provide the nets as input.\n");
        if (system("which NetFilterNonNested.perl > /dev/null") != 0)
            errAbort("ERROR: NetFilterNonNested.perl (comes with the chainCleaner source code) is not a binary in $PATH. Either install it or provide the nets as input.\n");
        char netFile[] = "tmp.chainCleaner.XXXXXXX.net";
        const int fd = mkstemps(netFile, 4);
        if (fd < 0) errAbort("ERROR: cannot create a tempfile for netting the chain file: %s\n", strerror(errno));
        close(fd);
        verbose(1, "0. need to net the input chains %s (no net file given) ...\n\t\ttempfile for netting: %s\n", inChainFile, netFile);
        const std::string cmd = std::string("bash -c 'set -o pipefail; chainNet -minScore=0 ") + inChainFile + " " + tSizes + " " + qSizes +
                                " stdout /dev/null | NetFilterNonNested.perl /dev/stdin -minScore1 3000 > " + netFile + "'";
        if (system(cmd.c_str()) != 0) errAbort("ERROR: chainNet | NetFilterNonNested.perl failed. Cannot net the chains. Command: %s\n", cmd.c_str());
        tmpNet = netFile;
        inNetFile = tmpNet.c_str();
        verbose(1, "DONE (nets in %s)\n", inNetFile);
    }

    // 1. fills, gaps and valid breaks from the net (chainCleaner.c:1088-1180)
    verbose(1, "1. parsing fills/gaps from %s and getting valid breaks ...\n", inNetFile);
    std::vector<Net> nets;
    readNets(inNetFile, nets);
    netsP = &nets;
    {
        std::vector<GapState> depth2gap(64);
        std::vector<int> depth2chain(64, 0);
        for (size_t n = 0; n < nets.size(); n++) {
            if (onlyThisChr && nets[n].name != onlyThisChr) continue;
            parseFill(nets[n].roots, 1, (int)n, depth2gap, depth2chain);
        }
    }
    aliRanges.resize(nets.size());
    for (size_t n = 0; n < nets.size(); n++) {
        if (onlyThisChr && nets[n].name != onlyThisChr) continue;
        collectAliBlocks(nets[n].roots, aliRanges[n]);
    }
    {   // nets of the same name share one range tree in the reference (genomeRangeTree is keyed by chrom)
        std::map<std::string, size_t> firstOf;
        for (size_t n = 0; n < nets.size(); n++) {
            auto it = firstOf.find(nets[n].name);
            if (it == firstOf.end()) firstOf[nets[n].name] = n;
            else { aliRanges[it->second].insert(aliRanges[it->second].end(), aliRanges[n].begin(), aliRanges[n].end()); aliRanges[n].clear(); }
        }
        for (size_t n = 0; n < nets.size(); n++) mergeAliRanges(aliRanges[n]);
        for (size_t n = 0; n < nets.size(); n++) {
            const size_t first = firstOf[nets[n].name];
            if (first != n) aliRanges[n] = aliRanges[first];
        }
    }
    if (!tmpNet.empty()) unlink(tmpNet.c_str());
    for (int id : chainId2Count.traverse()) getValidBreaks(id, nets);
    verbose(1, "DONE (parsing fills/gaps and getting valid breaks)\n\n");

    // 2. chains: the uninteresting ones go straight to the output (chainCleaner.c:584-618)
    FILE *finalOut = fopen(outChainFileUnsorted.c_str(), "w");
    if (!finalOut) errAbort("mustOpen: Can't open %s to write: %s", outChainFileUnsorted.c_str(), strerror(errno));
    finalChainOutFile = finalOut;
    verbose(1, "2. reading breaking and broken chains from %s and write irrelevant chains to %s ...\n", inChainFile, outChainFileUnsorted.c_str());
    {
        ChainSet cs;
        readChains(inChainFile, cs);
        FILE *interest = nullptr;
        if (debugMode && !(interest = fopen("chainsOfInterest.chain", "w")))
            errAbort("mustOpen: Can't open %s to write: %s", "chainsOfInterest.chain", strerror(errno));
        size_t meta = 0;
        for (size_t c = 0; c < cs.chains.size(); c++) {
            for (; meta < cs.metaLines.size() && cs.metaLineChain[meta] <= c; meta++) fprintf(finalOut, "%s\n", cs.metaLines[meta].c_str());
            const ChainHead &h = cs.chains[c];
            if (maxChainId < h.id) maxChainId = h.id;
            if (onlyThisChr && h.tName != onlyThisChr) continue;
            if (chainsOfInterest.has(h.id)) {
                LiveChain &lc = live[h.id];
                lc.head = h;
                lc.blocks.assign(cs.blocks.begin() + h.firstBlock, cs.blocks.begin() + h.firstBlock + h.nBlocks);
                if (interest) writeChain(interest, h, cs.blocks.data());
            } else writeChain(finalOut, h, cs.blocks.data());
        }
        for (; meta < cs.metaLines.size(); meta++) fprintf(finalOut, "%s\n", cs.metaLines[meta].c_str());
        if (interest) fclose(interest);
    }
    for (int id : chainsOfInterest.traverse())
        if (!live.count(id)) errAbort("ERROR: cannot get chain with Id %d from chainId2chain hash\n", id);
    verbose(1, "DONE\n\n");

    // 3. genomes to the GPU(s)
    verbose(1, "3. reading target and query DNA sequences for breaking and broken chains ...\n");
    TwoBitFile tbT(tNibDir), tbQ(qNibDir);
    std::unique_ptr<Scorer> scorer;
    if (!live.empty()) scorer.reset(new Scorer(gpuStarter, tbT, tbQ, scheme, gapCalc));
    verbose(1, "DONE\n\n");

    // 4. the suspect loop
    verbose(1, "4. loop over all breaks. Remove suspects if they pass our filters and write out deleted suspects to %s ...\n", outRemovedSuspectsFile);
    if (debugMode) {                    // chainCleaner.c:1817-1823
        auto must = [](const char *name) { FILE *f = fopen(name, "w"); if (!f) errAbort("mustOpen: Can't open %s to write: %s", name, strerror(errno)); return f; };
        suspectChainFile = must("suspect.chain");
        brokenChainLfillChainFile = must("brokenChainLfill.chain");
        brokenChainRfillChainFile = must("brokenChainRfill.chain");
        brokenChainfillChainFile = must("brokenChainfill.chain");
        suspectFillBedFile = must("suspectsAndFills.bed");
    }
    suspectsRemovedOutBedFile = fopen(outRemovedSuspectsFile, "w");
    if (!suspectsRemovedOutBedFile) errAbort("mustOpen: Can't open %s to write: %s", outRemovedSuspectsFile, strerror(errno));
    if (newChainIDDict && !(newChainIDDictFile = fopen(newChainIDDict, "w"))) errAbort("mustOpen: Can't open %s to write: %s", newChainIDDict, strerror(errno));
    if (suspectDataFile) {
        if (!(suspectDataFilePointer = fopen(suspectDataFile, "w"))) errAbort("mustOpen: Can't open %s to write: %s", suspectDataFile, strerror(errno));
        doPairs = false;
    }
    if (scorer) loopOverBreaks(*scorer);
    fclose(suspectsRemovedOutBedFile);
    if (newChainIDDictFile) fclose(newChainIDDictFile);
    if (suspectDataFilePointer) fclose(suspectDataFilePointer);
    if (debugMode) {
        fclose(suspectFillBedFile); fclose(suspectChainFile); fclose(brokenChainLfillChainFile); fclose(brokenChainRfillChainFile);
        fclose(brokenChainfillChainFile);
    }
    verbose(1, "DONE\n\n");

    // 5. breaking and broken chains, modified ones re-scored first (chainCleaner.c:625-644)
    verbose(1, "5. write the (new) breaking and the broken chains to %s ...\n", outChainFileUnsorted.c_str());
    {
        const std::vector<int> order = chainsOfInterest.traverse();
        std::vector<Request> reqs;
        for (int id : order)
            if (needsRescoring.count(id)) reqs.push_back(Request{id, INT_MIN, INT_MAX});
        std::map<int, double> newScore;
        if (!reqs.empty()) {
            const std::vector<SubScore> res = scorer->score(reqs);
            for (size_t i = 0; i < reqs.size(); i++) newScore[reqs[i].chainId] = res[i].global;
        }
        for (int id : order) {
            LiveChain &c = live.at(id);
            if (newScore.count(id)) c.head.score = newScore[id];
            ChainHead h = c.head;
            h.firstBlock = 0;
            h.nBlocks = c.blocks.size();
            writeChain(finalOut, h, c.blocks.data());
        }
    }
    fclose(finalOut);
    verbose(1, "DONE\n\n");

    // 6. chainSort (chainCleaner.c:1858-1867)
    verbose(1, "6. chainSort %s %s ...\n", outChainFileUnsorted.c_str(), outChainFile);
    const std::string sortCmd = "chainSort " + outChainFileUnsorted + " " + outChainFile;
    if (system(sortCmd.c_str()) != 0) errAbort("ERROR: chainSort failed. Command: %s\n", sortCmd.c_str());
    unlink(outChainFileUnsorted.c_str());
    verbose(1, "DONE\n\n");
    if (scorer) verbose(2, "GPU scoring: %zu batches, %zu sub-chain jobs\n", scorer->gpuCalls, scorer->gpuJobs);
    verbose(1, "\nALL DONE. New chains are in %s. Deleted suspects in %s\n", outChainFile, outRemovedSuspectsFile);
    if (debugMode)
        verbose(1, "Debug mode created those output files: \n"
                   "\tchainsOfInterest.chain     all breaking and broken chains (chain format)\n"
                   "\tsuspect.chain              subChains of all suspects (chain format)\n"
                   "\tbrokenChainfill.chain      subChains of all left and right parts of broken chains (chain format)\n"
                   "\tbrokenChainLfill.chain     subChains of all left parts of broken chains (chain format)\n"
                   "\tbrokenChainRfill.chain     subChains of all right parts of broken chains (chain format)\n"
                   "\tsuspectsAndFills.bed       coordinates, scores and ratios of the suspects, the left/right parts of broken chains (bed9 format)\n");
    return 0;
}

int main(int argc, char **argv) { return runTool(toolMain, argc, argv); }
