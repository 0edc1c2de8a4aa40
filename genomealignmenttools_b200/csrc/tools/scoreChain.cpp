// scoreChain -- (re)score existing chains.  Drop-in for src/scoreChain/scoreChain.c of
// hillerlab/GenomeAlignmentTools: same command line, same output bytes.  The reference scores
// chain by chain on the CPU (scoreChain.c:301-331 -> getChainScore :207-220 -> chainCalcScore +
// chainCalcScoreLocal); here all chains of the file become one CSR work-list that is scored by
// the sm_100a kernels behind gat_score() (include/gat.h), optionally sharded over several GPUs.
#include <cstring>
#include <unistd.h>
#include "gat_host.hpp"

using namespace gathost;

static const std::vector<OptionSpec> optionSpecs = {
    {"scoreScheme", OPTION_STRING},
    {"linearGap", OPTION_STRING},
    {"doLocalScore", OPTION_BOOLEAN},
    {"forceLocalScore", OPTION_BOOLEAN},
    {"returnOnlyScore", OPTION_BOOLEAN},
    {"returnOnlyScoreAndCoords", OPTION_BOOLEAN},
    {"gpus", OPTION_INT},               // extension: shard the chains over this many GPUs (default 1)
};

static void usage()
{   // scoreChain.c:52-79
    errAbort(
        "scoreChain - (re)score existing chains\n"
        "usage:\n"
        "   scoreChain in.chainFile reference.2bit query.2bit out.chain  -linearGap=loose|medium|filename\n"
        "Where reference.2bit and query.2bit are the names of a .2bit files for the reference and query\n"
        "options:\n"
        " Local score = we set score = 0 if score < 0 and return the max of the score that we reach for a chain\n"
        "   -returnOnlyScore             default=FALSE. Just return chain ID{tab}globalScore{tab}localScore{tab}totalAligningBases, not the entire chain\n"
        "   -returnOnlyScoreAndCoords    default=FALSE. Just return chain ID{tab}chainStartInRef{tab}chainEndInRef{tab}localScore{tab}totalAligningBases, not the entire chain\n"
        "   -doLocalScore                default=FALSE. Only if the global score of a chain is negative, compute and output the local score in the chain file.\n"
        "   -forceLocalScore             default=FALSE. Always output the local score in the chain file.\n"
        "   -scoreScheme=fileName        Read the scoring matrix from a blastz-format file\n"
        "   -linearGap=<medium|loose|filename>    Specify type of linearGap to use.\n"
        "              *Must* specify this argument to one of these choices.\n"
        "              loose is chicken/human linear gap costs.\n"
        "              medium is mouse/human linear gap costs.\n"
        "              Or specify a piecewise linearGap tab delimited file.\n"
        "   -gpus=N                      (B200 build) shard the chains over N GPUs, default 1\n"
        "   sample linearGap file (loose)\n"
        "%s",
        GapCalc::sampleFileContents());
}

static int toolMain(int argc, char **argv)
{
    Options opt;
    opt.init(&argc, argv, optionSpecs);
    const char *gapFileName = opt.val("linearGap", nullptr);
    const char *scoreSchemeName = opt.val("scoreScheme", nullptr);
    const bool doLocalScore = opt.exists("doLocalScore"), forceLocalScore = opt.exists("forceLocalScore");
    const bool returnOnlyScore = opt.exists("returnOnlyScore"), returnOnlyScoreAndCoords = opt.exists("returnOnlyScoreAndCoords");
    if (argc != 5) usage();
    if (returnOnlyScore && returnOnlyScoreAndCoords)
        errAbort("ERROR: You cannot specify both returnOnlyScore and returnOnlyScoreAndCoords\n");

    ScoreScheme scheme;
    if (scoreSchemeName) {
        verbose(2, "Reading scoring matrix from %s\n", scoreSchemeName);
        scheme = ScoreScheme::read(scoreSchemeName);
    } else
        scheme = ScoreScheme::defaultScheme();
    if (gapFileName == nullptr) errAbort("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults\n");
    const GapCalc gapCalc = GapCalc::fromFile(gapFileName);

    const char *t2bit = argv[2], *q2bit = argv[3];
    if (access(t2bit, F_OK) != 0) errAbort("ERROR: target 2bit file or nib directory %s does not exist\n", t2bit);
    if (access(q2bit, F_OK) != 0) errAbort("ERROR: query 2bit file or nib directory %s does not exist\n", q2bit);
    if (!TwoBitFile::isTwoBit(t2bit)) errAbort("ERROR: only 2bit files are supported, not %s\n", t2bit);
    if (!TwoBitFile::isTwoBit(q2bit)) errAbort("ERROR: only 2bit files are supported, not %s\n", q2bit);
    phaseDone("start");
    GpuStarter gpuStarter(opt.intVal("gpus", 1));      // contexts come up while the inputs are parsed
    TwoBitFile tbT(t2bit), tbQ(q2bit);
    phaseDone("2bit indexes");

    FILE *f = strcmp(argv[4], "stdout") == 0 ? stdout : fopen(argv[4], "w");
    if (!f) errAbort("mustOpen: Can't open %s to write: %s", argv[4], strerror(errno));

    ChainSet cs;
    readChains(argv[1], cs);
    phaseDone("chain file parsed");

    // like the reference (loadSeq, scoreChain.c:100-116) only sequences that chains name are loaded
    std::vector<int> useT, useQ, mapT(tbT.seqs().size(), -1), mapQ(tbQ.seqs().size(), -1);
    std::vector<uint32_t> chainT(cs.chains.size()), chainQ(cs.chains.size());
    for (size_t c = 0; c < cs.chains.size(); c++) {
        const ChainHead &h = cs.chains[c];
        const int ti = tbT.find(h.tName), qi = tbQ.find(h.qName);
        if (ti < 0) errAbort("%s is not in %s", h.tName.c_str(), t2bit);
        if (qi < 0) errAbort("%s is not in %s", h.qName.c_str(), q2bit);
        if ((size_t)ti >= mapT.size()) mapT.resize(ti + 1, -1);
        if ((size_t)qi >= mapQ.size()) mapQ.resize(qi + 1, -1);
        if (mapT[ti] < 0) { mapT[ti] = (int)useT.size(); useT.push_back(ti); verbose(3, "\t\tLoaded %d bases of %s from %s\n", (int)tbT.seqs()[ti].size, h.tName.c_str(), t2bit); }
        if (mapQ[qi] < 0) { mapQ[qi] = (int)useQ.size(); useQ.push_back(qi); verbose(3, "\t\tLoaded %d bases of %s from %s\n", (int)tbQ.seqs()[qi].size, h.qName.c_str(), q2bit); }
        chainT[c] = (uint32_t)mapT[ti];
        chainQ[c] = (uint32_t)mapQ[qi];
    }

    std::vector<int64_t> global, local;
    WorkList wl;
    if (!cs.chains.empty()) {
        MultiGpu &gpus = gpuStarter.get();
        phaseDone("CUDA contexts");
        gpus.prepare(tbT, useT, tbQ, useQ, scheme, gapCalc);
        phaseDone("genomes uploaded");
        buildRecords(cs, wl);
        for (size_t c = 0; c < cs.chains.size(); c++) addChainJob(cs, c, chainT[c], chainQ[c], wl);
        phaseDone("work-list built");
        gpus.score(wl, global, local);
        phaseDone("scored on the GPU");
    }

    for (size_t c = 0; c < cs.chains.size(); c++) {
        ChainHead &h = cs.chains[c];
        const double globalScore = (double)global[c], localScore = (double)local[c];
        verbose(2, "adjust score for chain %d (t %s %d-%d  q %c %s %d-%d) from %f to ", h.id, h.tName.c_str(), h.tStart, h.tEnd,
                h.qStrand, h.qName.c_str(), h.qStart, h.qEnd, h.score);
        if (forceLocalScore) h.score = localScore;                      // scoreChain.c:313-321
        else {
            h.score = globalScore;
            if (h.score <= 0 && doLocalScore) {
                verbose(2, "\tSCORE IS NEGATIVE --> doLocal is set --> set global score %f to local score %f\n ", h.score, localScore);
                h.score = localScore;
            }
        }
        verbose(2, "Final score for chainID %d: %f\n ", h.id, h.score);
        if (returnOnlyScore) fprintf(f, "%d\t%1.0f\t%1.0f\t%d\n", h.id, globalScore, localScore, (int)wl.aliBases[c]);
        else if (returnOnlyScoreAndCoords)
            fprintf(f, "%d\t%d\t%d\t%1.0f\t%1.0f\t%d\n", h.id, h.tStart, h.tEnd, globalScore, localScore, (int)wl.aliBases[c]);
    }
    if (!returnOnlyScore && !returnOnlyScoreAndCoords) writeChains(f, cs.chains, cs.blocks.data());
    if (f != stdout && fclose(f) != 0) errAbort("Error closing %s", argv[4]);
    phaseDone("output written");
    return 0;
}

int main(int argc, char **argv) { return runTool(toolMain, argc, argv); }
