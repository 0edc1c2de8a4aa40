// gat_host.hpp -- host side of the chain tools: everything between the files on disk and the
// batched C ABI of include/gat.h.  It mirrors the pieces of kent the three hillerlab tools use
// around chainCalcScore -- option parsing (lib/options.c), errAbort (lib/errAbort.c), the .2bit
// container (lib/twoBit.c), score schemes (lib/axt.c), gap tables (lib/gapCalc.c), .chain I/O
// and chainSubsetOnT (lib/chain.c) -- with the same file formats, messages and exit codes, but
// its output is a CSR work-list for the GPU instead of per-chain CPU calls.  No scoring here.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "gat.h"

namespace gathost {

// ---- errAbort / verbose (kent/src/lib/errAbort.c:182-222, verbose.c:19-31)
[[noreturn]] void errAbort(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void verbose(int level, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void verboseSetLevel(int level);
// GAT_TOOL_TIMING=1 in the environment: wall clock of each phase of a tool on stderr
void phaseDone(const char *what);
int verboseLevel();
// Library code throws; tools call runTool() which turns an Error into errAbort.
struct Error { std::string message; };
[[noreturn]] void fail(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
int runTool(int (*toolMain)(int, char **), int argc, char **argv);

// ---- options (kent/src/lib/options.c:41-203, 306)
enum OptionType { OPTION_BOOLEAN, OPTION_STRING, OPTION_INT, OPTION_FLOAT, OPTION_DOUBLE, OPTION_LONG_LONG };
struct OptionSpec { const char *name; OptionType type; };
class Options {
public:
    // Removes the option words from argv (like optionInit) and validates them against specs
    // (+ the common -verbose=N).  "--" stops option parsing.
    void init(int *argc, char **argv, const std::vector<OptionSpec> &specs);
    bool exists(const char *name) const { return values.count(name) != 0; }
    const char *val(const char *name, const char *dflt) const;
    int intVal(const char *name, int dflt) const;
    double doubleVal(const char *name, double dflt) const;
private:
    std::map<std::string, std::string> values;
};

// ---- .2bit container (kent/src/lib/twoBit.c:422-513, 574-650); the payload stays packed
struct TwoBitSeq {
    std::string name;
    uint32_t size = 0;
    const uint8_t *packed = nullptr;              // (size+3)/4 bytes inside the file image
    std::vector<uint32_t> nStart, nLen, maskStart, maskLen;
};
class TwoBitFile {
public:
    static bool isTwoBit(const std::string &path);           // twoBitIsFile: name ends in .2bit
    // A .2bit file, or -- what chainNet / chainCleaner also accept as tNibDir / qNibDir
    // (chainNet.c:166-170) -- a directory of <name>.nib files (kent/src/lib/nib.c:83-235: 4 bits
    // per base, T=0 C=1 A=2 G=3 N=4, bit 3 = soft-masked).  A nib sequence is converted to the
    // .2bit payload + N runs when find() first asks for it, so everything downstream is unchanged.
    explicit TwoBitFile(const std::string &path);
    ~TwoBitFile();
    TwoBitFile(const TwoBitFile &) = delete;
    TwoBitFile &operator=(const TwoBitFile &) = delete;
    int find(const std::string &name) const;                 // -1 if absent (nib directory: loads <name>.nib on demand)
    const std::vector<TwoBitSeq> &seqs() const { return seqs_; }
    const std::string &path() const { return path_; }
private:
    std::string path_;
    uint8_t *image_ = nullptr;
    size_t imageSize_ = 0;
    bool mapped_ = false;
    bool nibDir_ = false;
    mutable std::vector<TwoBitSeq> seqs_;
    mutable std::unordered_map<std::string, int> index_;
    mutable std::vector<std::vector<uint8_t>> nibPayload_;   // converted nib sequences (owned)
    int loadNib(const std::string &name) const;
};

// Upload the sequences `use` (indices into tb.seqs(), in that order) with gat_load_genome.
void uploadGenome(gat_ctx *ctx, int side, const TwoBitFile &tb, const std::vector<int> &use);

// ---- scoring parameters
struct ScoreScheme {                               // struct axtScoreScheme, kent/src/inc/axt.h:83-91
    int32_t matrix[4][4];                          // [q][t], kent codes T=0 C=1 A=2 G=3
    int gapOpen = 400, gapExtend = 30;
    static ScoreScheme defaultScheme();            // axtScoreSchemeDefault, axt.c:423-458
    static ScoreScheme read(const std::string &path);   // axtScoreSchemeRead, axt.c:692-834
};
struct GapCalc {                                   // struct gapCalc, kent/src/lib/gapCalc.c:12-37
    int smallSize = 0;
    std::vector<int32_t> qSmall, tSmall, bSmall, longPos;
    std::vector<double> qLong, tLong, bLong;
    static GapCalc fromFile(const char *name);     // gapCalcFromFile: "loose", "medium" or a file
    static GapCalc fromString(const std::string &text);
    static const char *sampleFileContents();       // gapCalcSampleFileContents, gapCalc.c:75-79
    int cost(int dq, int dt) const;                // gapCalcCost (host restatement for -verbose paths / tests)
};
void setScoring(gat_ctx *ctx, const ScoreScheme &ss, const GapCalc &gc);

// ---- chains (kent/src/inc/chain.h:48-63, lib/chain.c:200-346, 471-558)
struct ChainHead {
    double score = 0;
    std::string tName, qName;
    int tSize = 0, tStart = 0, tEnd = 0, qSize = 0, qStart = 0, qEnd = 0, id = 0;
    char qStrand = '+';
    uint64_t firstBlock = 0, nBlocks = 0;          // into ChainSet::blocks
};
struct ChainSet {
    std::vector<ChainHead> chains;
    std::vector<gat_block> blocks;                 // exactly the blocks of the file (never split)
    std::vector<std::string> metaLines;            // '#' lines (chainNet / chainCleaner pass them through)
    std::vector<size_t> metaLineChain;             // how many chains had been read when the line was met
};
// Reads every chain of a file (plain, or .gz through `gzip -dc` like linefile.c:40-53).
void readChains(const std::string &path, ChainSet &out);
void writeChain(FILE *f, const ChainHead &c, const gat_block *blocks);     // chainWrite, chain.c:211-227
void writeChains(FILE *f, const std::vector<ChainHead> &chains, const gat_block *blocks);   // all of them, formatted in parallel

// ---- work-list
struct WorkList {
    std::vector<gat_job> jobs;
    std::vector<gat_block> blocks;                 // device records: long blocks split into JOINED pieces
    uint64_t totalJobBlocks = 0;
    std::vector<uint64_t> chainFirstRecord;        // first device record of every chain (+ sentinel)
    std::vector<uint64_t> blockFirstRecord;        // first device record of every file block (+ sentinel); empty when
                                                   // no block was split (record index == block index)
    std::vector<int64_t> aliBases;                 // per job: sum of clipped sizes (scoreChain.c:182)
};
// Device records for all chains of a set; tSeq/qSeq come from the name->index maps.
void buildRecords(const ChainSet &cs, WorkList &wl);
// One whole-chain job (what scoreChain scores).
void addChainJob(const ChainSet &cs, size_t chainIx, uint32_t tSeq, uint32_t qSeq, WorkList &wl);
// chainSubsetOnT (chain.c:471-558) as a job; returns false for kent's NULL sub-chain (no job added).
// firstKeptHint: callers that know the chain's blocks ascend on the target may pass the index (within the chain) of the
// first block with tEnd > subStart, found by bisection; the default walks the chain from its head like the reference.
bool addSubChainJob(const ChainSet &cs, size_t chainIx, uint32_t tSeq, uint32_t qSeq, int subStart, int subEnd, WorkList &wl,
                    uint64_t firstKeptHint = 0);
// Index (within the chain) of the first block whose end on the target (onQ: query) lies beyond `pos`, by bisection;
// only for chains whose blocks ascend and do not overlap on that side (chainAscends).
bool chainAscends(const ChainSet &cs, size_t chainIx, bool onQ);
uint64_t firstBlockEndingAfter(const ChainSet &cs, size_t chainIx, int pos, bool onQ);
struct CompactWorkList {
    std::vector<gat_cjob> jobs;
    std::vector<gat_cblock> blocks;
    std::vector<gat_cabs> abs, anchors;
};
bool packCompact(const WorkList &wl, CompactWorkList &out);
// Greedy longest-processing-time split of jobs over `parts` GPUs by aligned bases.
std::vector<std::vector<uint32_t>> shardJobs(const WorkList &wl, int parts);
// Sub-work-list holding the given jobs (block records are shared, so only jobs are re-indexed).
void extractShard(const WorkList &wl, const std::vector<uint32_t> &jobIx, std::vector<gat_job> &jobs, uint64_t &totalJobBlocks);

// Score a work-list on `nGpus` devices, one context and one host thread per device; every device holds a full copy of
// both genomes, shards are independent (no collective), results land in global/local indexed like wl.jobs.
//   * a whole-chain job above 1/(4 nGpus) of the aligned bases is cut at block boundaries into pieces; a piece comes back
//     as a tuple (gat_request_tuples) and the pieces are joined on the host with the gap cost between them, bit-identical
//     to the unsplit chain;
//   * jobs (and pieces) are balanced by greedy aligned-base load; every GPU receives only the records its jobs reference,
//     from a pinned staging buffer, as a compact work-list (gat_score_compact) when the shard holds whole chains only.
// Devices are 0..nGpus-1 unless GAT_DEVICES names them ("0,0" puts two contexts on device 0: how a one-GPU box tests this).
struct MultiGpu {
    std::vector<gat_ctx *> ctx;
    explicit MultiGpu(int nGpus);
    ~MultiGpu();
    // genomes and scoring parameters to every context (in parallel); the gap tables are kept for joining pieces
    void prepare(const TwoBitFile &tbT, const std::vector<int> &useT, const TwoBitFile &tbQ, const std::vector<int> &useQ,
                 const ScoreScheme &ss, const GapCalc &gc);
    void score(const WorkList &wl, std::vector<int64_t> &global, std::vector<int64_t> &local);
    struct ShardStats { uint64_t jobs = 0, records = 0, h2dBytes = 0, pieces = 0; bool compact = false;
                        double buildS = 0, stageS = 0, scoreS = 0; };   // host seconds: shard's work-list, pinned staging, the scoring call
    std::vector<ShardStats> lastShards;     // what the last score() sent to each GPU (tests, -verbose)
    // GAT_DUMP_WORKLIST=<prefix> in the environment: every score() call also writes its work-list to <prefix>.<k>.jobs / .blocks
    // (raw gat_job / gat_block arrays) and the sequence names in upload order to <prefix>.tseqs / .qseqs, so that bench.py can
    // time exactly the batches a tool sends (configs 2 and 3 of BASELINE.json)
private:
    std::vector<std::string> tNames_, qNames_;
    int dumpCount_ = 0;
    GapCalc gap_;
    bool haveGap_ = false;
    struct Staging { void *p = nullptr; size_t cap = 0; };
    std::vector<Staging> staging_;          // pinned, one per GPU, grown on demand
    void *pinned(size_t g, size_t bytes);
};

// chainRemovePartialOverlaps (kent/src/lib/chainConnect.c:255-344) for every chain of a set: where adjacent blocks
// overlap, the crossover point comes from the device (gat_crossover, one batch per sweep), the trimming and the rare
// removal of a dried-up block follow the reference step by step.  chainT / chainQ: sequence index per chain (as
// uploaded).  Blocks, block counts and bounds of `cs` are rewritten in place.
void removePartialOverlaps(gat_ctx *ctx, ChainSet &cs, const std::vector<uint32_t> &chainT, const std::vector<uint32_t> &chainQ);

// Starts creating the CUDA contexts on a background thread; get() waits for them.
struct GpuStarter {
    struct Impl;
    Impl *impl;
    explicit GpuStarter(int nGpus);
    ~GpuStarter();
    MultiGpu &get();
};

}  // namespace gathost
