// chainSort -- sort chains by score (or target / query position).  chainCleaner finishes by
// running `chainSort out.unsorted out.chain` (src/chainCleaner/chainCleaner.c:1863); this is the
// same tool as kent/src/hg/mouseStuff/chainSort/chainSort.c: chains are read into a list that
// ends up reversed (slAddHead), then sorted with a stable merge sort, so ties keep reversed
// input order.  Pure host code.
#include <algorithm>
#include <cstring>
#include "gat_host.hpp"

using namespace gathost;

static const std::vector<OptionSpec> optionSpecs = {{"target", OPTION_BOOLEAN}, {"query", OPTION_BOOLEAN}, {"index", OPTION_STRING}};

static void usage()
{
    errAbort(
        "chainSort - Sort chains.  By default sorts by score.\n"
        "Note this loads all chains into memory, so it is not\n"
        "suitable for large sets.  Instead, run chainSort on\n"
        "multiple small files, followed by chainMergeSort.\n"
        "usage:\n"
        "   chainSort inFile outFile\n"
        "Note that inFile and outFile can be the same\n"
        "options:\n"
        "   -target sort on target start rather than score\n"
        "   -query sort on query start rather than score\n"
        "   -index=out.tab build simple two column index file\n"
        "                    <out file position>  <value>\n"
        "                  where <value> is score, target, or query \n"
        "                  depending on the sort.\n");
}

static int toolMain(int argc, char **argv)
{
    Options opt;
    opt.init(&argc, argv, optionSpecs);
    if (argc != 3) usage();
    const bool isQuery = opt.exists("query"), isTarget = opt.exists("target");
    ChainSet cs;
    readChains(argv[1], cs);
    FILE *f = strcmp(argv[2], "stdout") == 0 ? stdout : fopen(argv[2], "w");
    if (!f) errAbort("mustOpen: Can't open %s to write: %s", argv[2], strerror(errno));
    FILE *index = nullptr;
    if (const char *indexName = opt.val("index", nullptr)) {
        index = fopen(indexName, "w");
        if (!index) errAbort("mustOpen: Can't open %s to write: %s", indexName, strerror(errno));
    }
    for (const std::string &m : cs.metaLines) fprintf(f, "%s\n", m.c_str());      // lineFileSetMetaDataOutput
    std::vector<size_t> order(cs.chains.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = order.size() - 1 - i;     // slAddHead: last chain first
    // Tie order (chains with equal keys): kent's slSort is libc qsort over the list that slAddHead reversed.  glibc's qsort is a
    // stable merge sort whenever it can allocate its buffer (every version the reference's fixtures were made with, up to 2.36),
    // which std::stable_sort over the reversed order reproduces; glibc 2.37+ may use introsort, whose tie order is unspecified
    // -- against a reference built there, equal-score chains can come out in another order.
    const auto &ch = cs.chains;
    if (isTarget)
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            const int d = strcmp(ch[a].tName.c_str(), ch[b].tName.c_str());
            return d != 0 ? d < 0 : ch[a].tStart < ch[b].tStart;
        });
    else if (isQuery)
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            const int d = strcmp(ch[a].qName.c_str(), ch[b].qName.c_str());
            return d != 0 ? d < 0 : ch[a].qStart < ch[b].qStart;
        });
    else
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ch[a].score > ch[b].score; });
    double lastScore = -1;
    std::string lastName;
    for (size_t i : order) {
        const ChainHead &h = ch[i];
        if (index) {
            if (isTarget || isQuery) {
                const std::string &name = isTarget ? h.tName : h.qName;
                if (name != lastName) { lastName = name; fprintf(index, "%lx\t%s\n", ftell(f), name.c_str()); }
            } else if (h.score != lastScore) { lastScore = h.score; fprintf(index, "%lx\t%1.0f\n", ftell(f), h.score); }
        }
        writeChain(f, h, cs.blocks.data());
    }
    if (index) fclose(index);
    if (f != stdout) fclose(f);
    return 0;
}

int main(int argc, char **argv) { return runTool(toolMain, argc, argv); }
