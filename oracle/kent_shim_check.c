/* oracle/kent_shim_check.c -- TEST INFRASTRUCTURE.  A C program written against kent's OWN headers (struct chain, cBlock,
 * dnaSeq, axtScoreScheme and the prototypes of chainConnect.h / gapCalc.h / axt.h come from $(REF)/kent/src/inc) that is
 * linked against libgatkent.so instead of kent's library: proves that the shims of include/gat_kent.h are a drop-in for
 * those symbols.  Reads a small case description, calls the kent-named entry points, prints what they return; the test
 * (tests/test_gpu_kent_shims.py) compares with the oracle.
 *
 *   kent_shim_check case.txt
 * case.txt:  gap <loose|medium|file>\n scheme <file|->\n target <dna>\n query <dna>\n chain <id> <nBlocks>\n
 *            <tStart> <tEnd> <qStart> <qEnd> per block\n pairs <n>\n <dq> <dt> per pair */
#include "common.h"
#include "dnaseq.h"
#include "chain.h"
#include "axt.h"
#include "gapCalc.h"
#include "chainConnect.h"
#include <stddef.h>

long gatKentLayout(int which);
double chainCalcScoreLocal(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query,
                           struct dnaSeq *target, int *retAliBases);

static char *readWord(FILE *f, size_t cap)
{
    char *buf = malloc(cap + 2);
    if (fscanf(f, "%s", buf) != 1) { fprintf(stderr, "case file ended early\n"); exit(2); }
    return buf;
}

int main(int argc, char **argv)
{
    if (argc != 2) { fprintf(stderr, "usage: kent_shim_check case.txt\n"); return 2; }
    /* the layouts the shim assumes are kent's */
    long want[][2] = {{0, sizeof(struct cBlock)}, {1, sizeof(struct chain)}, {2, sizeof(struct dnaSeq)}, {3, sizeof(struct axtScoreScheme)},
                      {10, offsetof(struct cBlock, tStart)}, {11, offsetof(struct cBlock, qStart)}, {12, offsetof(struct cBlock, tEnd)},
                      {20, offsetof(struct chain, blockList)}, {21, offsetof(struct chain, tStart)}, {22, offsetof(struct chain, qStart)},
                      {23, offsetof(struct chain, qStrand)}, {24, offsetof(struct chain, id)}, {30, offsetof(struct dnaSeq, dna)},
                      {31, offsetof(struct dnaSeq, size)}, {40, offsetof(struct axtScoreScheme, matrix)}, {41, offsetof(struct axtScoreScheme, gapOpen)}};
    for (size_t i = 0; i < sizeof want / sizeof want[0]; i++)
        if (gatKentLayout((int)want[i][0]) != want[i][1]) { printf("layout MISMATCH %ld: shim %ld kent %ld\n", want[i][0], gatKentLayout((int)want[i][0]), want[i][1]); return 1; }
    printf("layout ok\n");

    FILE *f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    char key[64];
    struct gapCalc *gc = NULL;
    struct axtScoreScheme *ss = NULL;
    struct dnaSeq target, query;
    memset(&target, 0, sizeof target); memset(&query, 0, sizeof query);
    struct chain ch;
    memset(&ch, 0, sizeof ch);
    size_t cap = 1 << 26;
    while (fscanf(f, "%63s", key) == 1) {
        if (!strcmp(key, "gap")) { char *w = readWord(f, 4096); gc = gapCalcFromFile(w); }
        else if (!strcmp(key, "scheme")) { char *w = readWord(f, 4096); ss = strcmp(w, "-") ? axtScoreSchemeRead(w) : axtScoreSchemeDefault(); }
        else if (!strcmp(key, "target")) { target.dna = readWord(f, cap); target.size = (int)strlen(target.dna); target.name = "t"; }
        else if (!strcmp(key, "query")) { query.dna = readWord(f, cap); query.size = (int)strlen(query.dna); query.name = "q"; }
        else if (!strcmp(key, "chain")) {
            int n;
            if (fscanf(f, "%d %d", &ch.id, &n) != 2) return 2;
            struct cBlock *tail = NULL;
            for (int i = 0; i < n; i++) {
                struct cBlock *b = calloc(1, sizeof *b);
                if (fscanf(f, "%d %d %d %d", &b->tStart, &b->tEnd, &b->qStart, &b->qEnd) != 4) return 2;
                if (tail) tail->next = b; else ch.blockList = b;
                tail = b;
                if (i == 0) { ch.tStart = b->tStart; ch.qStart = b->qStart; }
                ch.tEnd = b->tEnd; ch.qEnd = b->qEnd;
            }
            ch.tName = "t"; ch.qName = "q"; ch.tSize = target.size; ch.qSize = query.size; ch.qStrand = '+';
            printf("global %.0f\n", chainCalcScore(&ch, ss, gc, &query, &target));
            int ali = 0;
            double local = chainCalcScoreLocal(&ch, ss, gc, &query, &target, &ali);
            printf("local %.0f ali %d\n", local, ali);
            /* chainCalcScoreSubChain: sequences that hold the chain's span only (chainConnect.c:42-59) */
            struct dnaSeq st = target, sq = query;
            st.dna = cloneStringZ(target.dna + ch.tStart, ch.tEnd - ch.tStart); st.size = ch.tEnd - ch.tStart;
            sq.dna = cloneStringZ(query.dna + ch.qStart, ch.qEnd - ch.qStart); sq.size = ch.qEnd - ch.qStart;
            printf("sub %.0f\n", chainCalcScoreSubChain(&ch, ss, gc, &sq, &st));
            struct cBlock *b = ch.blockList;
            printf("block0 %.0f\n", chainScoreBlock(query.dna + b->qStart, target.dna + b->tStart, b->tEnd - b->tStart, ss->matrix));
        } else if (!strcmp(key, "pairs")) {
            int n;
            if (fscanf(f, "%d", &n) != 1) return 2;
            for (int i = 0; i < n; i++) {
                int dq, dt;
                if (fscanf(f, "%d %d", &dq, &dt) != 2) return 2;
                printf("gap %d %d %d\n", dq, dt, gapCalcCost(gc, dq, dt));
            }
        } else { fprintf(stderr, "unknown key %s\n", key); return 2; }
    }
    return 0;
}

/* the only kent library function used above (kent/src/lib/common.c cloneStringZ), so that nothing of kent's library is linked */
char *cloneStringZ(const char *s, int size)
{
    char *d = calloc((size_t)size + 1, 1);
    memcpy(d, s, (size_t)size);
    return d;
}
