/* oracle/ref_shim.c -- TEST INFRASTRUCTURE.  Thin ctypes-friendly entry points in front of the
 * UNMODIFIED reference objects (kent jkweb sources + src/scoreChain/scoreChain.c compiled with
 * main renamed).  Every score below is produced by the reference's own chainCalcScore
 * (kent/src/lib/chainConnect.c:24), chainCalcScoreLocal (src/scoreChain/scoreChain.c:176),
 * chainSubsetOnT (kent/src/lib/chain.c:471) and gapCalcCost (kent/src/lib/gapCalc.c:298);
 * this file only moves arguments in and results out. */
#include "common.h"
#include "linefile.h"
#include "hash.h"
#include "dnaseq.h"
#include "twoBit.h"
#include "axt.h"
#include "gapCalc.h"
#include "chain.h"
#include "chainConnect.h"

/* globals and functions defined (non-static) in src/scoreChain/scoreChain.c:18-38,100-220 */
extern struct gapCalc *gapCalc;
extern struct axtScoreScheme *scoreScheme;
extern char *t2bit, *q2bit;
extern struct hash *tSeqHash, *qSeqHash, *qSeqMinusStrandHash;
extern struct twoBitFile *ttbf, *qtbf;
void loadSeq(char *seqPath, boolean isTarget, char *newName, struct hash *hash);
struct dnaSeq *getSeqFromHash(char *chrom, char strand, struct hash *hash);
double chainCalcScoreLocal(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc,
	struct dnaSeq *query, struct dnaSeq *target, int *retAliBases);
double getChainScore(struct chain *chain, double *globalScore, double *localScore, int *aliBases);

static struct chain **chains = NULL;
static int chainCount = 0, chainAlloc = 0;

int ref_set_scoring(const char *scoreSchemeFile, const char *linearGap)
/* Same calls as scoreChain.c:253-263. */
{
if (scoreSchemeFile != NULL && scoreSchemeFile[0] != 0)
    scoreScheme = axtScoreSchemeRead((char *)scoreSchemeFile);
else
    scoreScheme = axtScoreSchemeDefault();
gapCalc = gapCalcFromFile((char *)linearGap);
return 0;
}

int ref_open_genomes(const char *tPath, const char *qPath)
/* Same calls as scoreChain.c:265-291. */
{
dnaUtilOpen();
t2bit = cloneString((char *)tPath);
q2bit = cloneString((char *)qPath);
ttbf = twoBitOpen(t2bit);
qtbf = twoBitOpen(q2bit);
tSeqHash = newHash(0);
qSeqHash = newHash(0);
qSeqMinusStrandHash = newHash(0);
return 0;
}

int ref_load_chains(const char *chainFile)
/* chainRead every chain of the file (kent/src/lib/chain.c:337) and load its sequences. */
{
struct lineFile *lf = lineFileOpen((char *)chainFile, TRUE);
struct chain *chain;
chainCount = 0;
while ((chain = chainRead(lf)) != NULL)
    {
    if (chainCount == chainAlloc)
	{
	int newAlloc = chainAlloc ? 2*chainAlloc : 1024;
	chains = needMoreMem(chains, chainAlloc*sizeof(chains[0]), newAlloc*sizeof(chains[0]));
	chainAlloc = newAlloc;
	}
    loadSeq(t2bit, TRUE, chain->tName, tSeqHash);
    loadSeq(q2bit, FALSE, chain->qName, qSeqHash);
    chains[chainCount++] = chain;
    }
lineFileClose(&lf);
return chainCount;
}

int ref_chain_count(void) { return chainCount; }

void ref_chain_info(int ix, int *id, int *tStart, int *tEnd, int *nBlocks)
{
struct chain *c = chains[ix];
*id = c->id; *tStart = c->tStart; *tEnd = c->tEnd; *nBlocks = slCount(c->blockList);
}

void ref_score_all(double *global, double *local, int *aliBases)
/* scoreChain's own getChainScore for every loaded chain. */
{
int i;
for (i=0; i<chainCount; ++i)
    getChainScore(chains[i], &global[i], &local[i], &aliBases[i]);
}

void ref_score_sub(int n, const int *chainIx, const int *subStart, const int *subEnd,
	double *global, double *local, int *aliBases, int *isNull)
/* chainSubsetOnT then the reference scorers, exactly as chainNet.c:832-835 /
 * chainCleaner.c:1214-1229 do.  Scores go to the out arrays; chain->score is left alone. */
{
int i;
for (i=0; i<n; ++i)
    {
    struct chain *chain = chains[chainIx[i]], *sub = NULL, *toFree = NULL;
    chainSubsetOnT(chain, subStart[i], subEnd[i], &sub, &toFree);
    if (sub == NULL)
	{
	isNull[i] = 1; global[i] = local[i] = 0; aliBases[i] = 0;
	continue;
	}
    isNull[i] = 0;
    struct dnaSeq *qSeq = getSeqFromHash(sub->qName, sub->qStrand, qSeqHash);
    struct dnaSeq *tSeq = getSeqFromHash(sub->tName, '+', tSeqHash);
    global[i] = chainCalcScore(sub, scoreScheme, gapCalc, qSeq, tSeq);
    local[i] = chainCalcScoreLocal(sub, scoreScheme, gapCalc, qSeq, tSeq, &aliBases[i]);
    chainFree(&toFree);
    }
}

int ref_gap_cost(int dq, int dt) { return gapCalcCost(gapCalc, dq, dt); }
int ref_matrix(int q, int t) { return scoreScheme->matrix[q][t]; }

double ref_score_block(const char *q, const char *t, int size)
{ return chainScoreBlock((char *)q, (char *)t, size, scoreScheme->matrix); }

int ref_seq_dna(int isTarget, const char *name, char strand, char **retDna)
/* Hand out the unpacked (and, on '-', reverse-complemented) sequence the reference scores on. */
{
struct dnaSeq *seq = getSeqFromHash((char *)name, strand, isTarget ? tSeqHash : qSeqHash);
*retDna = seq->dna;
return seq->size;
}

void ref_crossover(const char *tName, const char *qName, char qStrand, int n,
	const int *leftTEnd, const int *leftQEnd, const int *rightTStart, const int *rightQStart,
	const int *overlap, int *pos, int *adjust)
/* cBlockFindCrossover (kent/src/lib/chainConnect.c:61-105) on blocks that carry just the fields it reads. */
{
int i;
struct dnaSeq *qSeq, *tSeq;
loadSeq(t2bit, TRUE, (char *)tName, tSeqHash);
loadSeq(q2bit, FALSE, (char *)qName, qSeqHash);
qSeq = getSeqFromHash((char *)qName, qStrand, qSeqHash);
tSeq = getSeqFromHash((char *)tName, '+', tSeqHash);
for (i=0; i<n; ++i)
    {
    struct cBlock left, right;
    ZeroVar(&left); ZeroVar(&right);
    left.tEnd = leftTEnd[i]; left.qEnd = leftQEnd[i];
    left.tStart = left.tEnd - overlap[i]; left.qStart = left.qEnd - overlap[i];
    right.tStart = rightTStart[i]; right.qStart = rightQStart[i];
    right.tEnd = right.tStart + overlap[i]; right.qEnd = right.qStart + overlap[i];
    cBlockFindCrossover(&left, &right, qSeq, tSeq, overlap[i], scoreScheme->matrix, &pos[i], &adjust[i]);
    }
}

int ref_remove_partial_overlaps(int chainIx)
/* chainRemovePartialOverlaps (kent/src/lib/chainConnect.c:255-344) on a loaded chain, in place; returns its block count. */
{
struct chain *chain = chains[chainIx];
struct dnaSeq *qSeq = getSeqFromHash(chain->qName, chain->qStrand, qSeqHash);
struct dnaSeq *tSeq = getSeqFromHash(chain->tName, '+', tSeqHash);
chainRemovePartialOverlaps(chain, qSeq, tSeq, scoreScheme->matrix);
return slCount(chain->blockList);
}

void ref_chain_blocks(int chainIx, int *tStart, int *qStart, int *size, int *bounds)
/* Blocks and bounds (tStart tEnd qStart qEnd) of a loaded chain. */
{
struct chain *chain = chains[chainIx];
struct cBlock *b;
int i = 0;
for (b = chain->blockList; b != NULL; b = b->next, ++i)
    { tStart[i] = b->tStart; qStart[i] = b->qStart; size[i] = b->tEnd - b->tStart; }
bounds[0] = chain->tStart; bounds[1] = chain->tEnd; bounds[2] = chain->qStart; bounds[3] = chain->qEnd;
}
