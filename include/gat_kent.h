/* include/gat_kent.h -- kent's own C entry points of the chain-rescoring path, kept by name and signature, on top of
 * the batched GPU ABI of include/gat.h (libgatkent.so; SURVEY 8b "C entry points that keep their names").
 *
 * A program written against kent's headers
 *     kent/src/inc/chainConnect.h:34-44   chainScoreBlock, chainCalcScore, chainCalcScoreSubChain
 *     kent/src/inc/gapCalc.h:11-35        gapCalcDefault / Original / FromFile / FromString / Free, gapCalcCost
 *     kent/src/inc/axt.h:93-121           axtScoreSchemeDefault / Read / Free
 * links this library instead of jkweb.a for those symbols and keeps calling them one chain at a time: synchronous, the
 * caller owns chain / dnaSeq / score scheme, results by value, fatal errors through errAbort (message on stderr, exit(-1),
 * kent/src/lib/errAbort.c:182-197).  Every call becomes a one-job batch of gat_score() on the GPU -- there is no CPU
 * scoring loop in here; gapCalcCost evaluates on the device as well (gat_gap_cost).  The tools of this repository do
 * NOT go through these shims: they batch.
 *
 * What the shims assume about kent's types (checked against kent's own headers by oracle/kent_shim_check.c through
 * gatKentLayout): struct cBlock, struct chain (kent/src/inc/chain.h:17-25, 48-63), struct dnaSeq
 * (kent/src/inc/dnaseq.h:18-26), struct axtScoreScheme (kent/src/inc/axt.h:83-91).  struct gapCalc is private to kent
 * (gapCalc.h:8-9) and private here.
 *
 * Limits, all reported through errAbort: a score scheme may only have entries for a/c/g/t in either case (what
 * axtScoreSchemeRead produces: axt.c:431-454 writes no other cell, and propagateCase copies them to the other case);
 * any other character of a sequence scores 0 like kent's N.  Sequences are uploaded when first seen and cached by
 * (pointer, size, sampled content): call gatKentForget() after changing a sequence in place.
 */
#ifndef GAT_KENT_H
#define GAT_KENT_H
#ifdef __cplusplus
extern "C" {
#endif

struct chain;
struct cBlock;
struct dnaSeq;
struct axtScoreScheme;
struct gapCalc;

/* kent/src/inc/chainConnect.h:34-44 */
double chainScoreBlock(char *q, char *t, int size, int matrix[256][256]);
double chainCalcScore(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query,
                      struct dnaSeq *target);
double chainCalcScoreSubChain(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc,
                              struct dnaSeq *query, struct dnaSeq *target);
/* hillerlab's addition with the same arguments (static in src/scoreChain/scoreChain.c:176-198 and
 * src/chainCleaner/chainCleaner.c:531-551): the local score; *retAliBases (may be NULL) = sum of block sizes */
double chainCalcScoreLocal(struct chain *chain, struct axtScoreScheme *ss, struct gapCalc *gapCalc, struct dnaSeq *query,
                           struct dnaSeq *target, int *retAliBases);

/* kent/src/inc/gapCalc.h:11-35 */
struct gapCalc *gapCalcDefault(void);
struct gapCalc *gapCalcOriginal(void);
struct gapCalc *gapCalcFromFile(char *fileName);
struct gapCalc *gapCalcFromString(char *s);
void gapCalcFree(struct gapCalc **pGapCalc);
int gapCalcCost(struct gapCalc *gapCalc, int dq, int dt);
char *gapCalcSampleFileContents(void);

/* kent/src/inc/axt.h:93-121 */
struct axtScoreScheme *axtScoreSchemeDefault(void);      /* static singleton: do NOT free (axt.c:423-430) */
struct axtScoreScheme *axtScoreSchemeRead(char *fileName);
void axtScoreSchemeFree(struct axtScoreScheme **pObj);

/* not kent's: drop the cached copies of sequences (after editing a dnaSeq in place), and the layout the shims assume
 * (which = 0..3: sizeof cBlock / chain / dnaSeq / axtScoreScheme; 10..: offsets, see gat_kent.cpp) */
void gatKentForget(void);
long gatKentLayout(int which);

#ifdef __cplusplus
}
#endif
#endif
