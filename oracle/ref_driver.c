/* oracle/ref_driver.c -- TEST/BENCH INFRASTRUCTURE.  Times the UNMODIFIED reference scoring loop
 * (src/scoreChain/scoreChain.c:301-311: getChainScore = chainCalcScore + chainCalcScoreLocal)
 * over an in-memory chain set, so the CPU number excludes text parsing and .2bit unpacking
 * (reported separately as load_s).  One process = one core, as the reference is single-threaded;
 * bench.py runs one of these per host core on chain shards split by target chromosome.
 *
 * usage: ref_driver in.chain t.2bit q.2bit linearGap scoreScheme|- reps [scores.out]
 * prints one JSON line. */
#include "common.h"
#include "linefile.h"
#include "hash.h"
#include "dnaseq.h"
#include "twoBit.h"
#include "axt.h"
#include "gapCalc.h"
#include "chain.h"
#include "chainConnect.h"
#include <time.h>

extern struct gapCalc *gapCalc;
extern struct axtScoreScheme *scoreScheme;
extern char *t2bit, *q2bit;
extern struct hash *tSeqHash, *qSeqHash, *qSeqMinusStrandHash;
extern struct twoBitFile *ttbf, *qtbf;
void loadSeq(char *seqPath, boolean isTarget, char *newName, struct hash *hash);
struct dnaSeq *getSeqFromHash(char *chrom, char strand, struct hash *hash);
double getChainScore(struct chain *chain, double *globalScore, double *localScore, int *aliBases);

static double now(void)
{
struct timespec ts;
clock_gettime(CLOCK_MONOTONIC, &ts);
return ts.tv_sec + 1e-9*ts.tv_nsec;
}

int main(int argc, char *argv[])
{
if (argc < 7)
    errAbort("usage: ref_driver in.chain t.2bit q.2bit linearGap scoreScheme|- reps [scores.out]");
int reps = atoi(argv[6]);
double t0 = now();
scoreScheme = sameString(argv[5], "-") ? axtScoreSchemeDefault() : axtScoreSchemeRead(argv[5]);
gapCalc = gapCalcFromFile(argv[4]);
dnaUtilOpen();
t2bit = argv[2]; q2bit = argv[3];
ttbf = twoBitOpen(t2bit); qtbf = twoBitOpen(q2bit);
tSeqHash = newHash(0); qSeqHash = newHash(0); qSeqMinusStrandHash = newHash(0);
struct lineFile *lf = lineFileOpen(argv[1], TRUE);
struct chain *chain, *chainList = NULL;
long long nChains = 0, nBlocks = 0, aligned = 0;
while ((chain = chainRead(lf)) != NULL)
    {
    loadSeq(t2bit, TRUE, chain->tName, tSeqHash);
    loadSeq(q2bit, FALSE, chain->qName, qSeqHash);
    if (chain->qStrand == '-')
	getSeqFromHash(chain->qName, '-', qSeqHash);	/* pay the lazy whole-chromosome rev-comp up front */
    slAddHead(&chainList, chain);
    ++nChains;
    struct cBlock *b;
    for (b = chain->blockList; b != NULL; b = b->next)
	{ ++nBlocks; aligned += b->tEnd - b->tStart; }
    }
slReverse(&chainList);
lineFileClose(&lf);
double loadS = now() - t0;

FILE *out = (argc > 7) ? mustOpen(argv[7], "w") : NULL;
double best = 1e30, total = 0, checksum = 0;
int r;
for (r = 0; r < reps; ++r)
    {
    double g, l; int ali;
    double t1 = now();
    checksum = 0;
    for (chain = chainList; chain != NULL; chain = chain->next)
	{
	getChainScore(chain, &g, &l, &ali);
	checksum += g + 3*l + 7*ali;
	if (out != NULL && r == 0)
	    fprintf(out, "%d\t%1.0f\t%1.0f\t%d\n", chain->id, g, l, ali);
	}
    double dt = now() - t1;
    total += dt;
    if (dt < best) best = dt;
    }
if (out != NULL) carefulClose(&out);
printf("{\"chains\": %lld, \"blocks\": %lld, \"aligned_bp\": %lld, \"load_s\": %.6f, "
       "\"score_s_best\": %.6f, \"score_s_mean\": %.6f, \"reps\": %d, \"checksum\": %.0f}\n",
       nChains, nBlocks, aligned, loadS, best, reps ? total/reps : 0.0, reps, checksum);
return 0;
}
