/* oracle/chain_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the chain-rescoring path of hillerlab/GenomeAlignmentTools
 * (kent chainCalcScore / chainScoreBlock / gapCalcCost + hillerlab chainCalcScoreLocal).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the
 * product (genomealignmenttools_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's own golden chains (kent axtChain/tests/expected/{newStyleLastz,oldStyleBlastz}.chain: 671823, 671644),
 *   - the gapCalcCost / chrM known-answer vectors of SURVEY.md section 4,
 *   - the unmodified reference compiled into oracle/_ref (libkentref.so) on seeded random inputs,
 *   - fixtures under tests/golden/ that were produced by that compiled reference.
 */
#ifndef CHAIN_ORACLE_H
#define CHAIN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_genome orc_genome;     /* a parsed .2bit, sequences unpacked on demand */
typedef struct orc_scoring orc_scoring;   /* 256x256 matrix + gap tables */
typedef struct orc_chainset orc_chainset; /* a parsed .chain file */

/* Work-list records.  Layout is deliberately identical to include/gat.h (gat_block / gat_job)
 * so a test can hand the very same arrays to both sides; the definitions are independent. */
typedef struct { int32_t tStart, qStart, size; } orc_block;
typedef struct {
    uint32_t tSeq;        /* index of target sequence in the target .2bit */
    uint32_t qSeq;        /* index of query sequence; bit 31 set = '-' strand */
    uint32_t firstBlock;  /* first block of the job in blocks[] */
    uint32_t blockPtr;    /* CSR row pointer: number of job-blocks of all earlier jobs */
    int32_t clipStart, clipEnd; /* target-side clip range (chainSubsetOnT), INT32_MIN/MAX = none */
} orc_job;

const char *orc_last_error(void);

/* ---- .2bit (kent/src/lib/twoBit.c:422-513, 574-650, 725-878) */
orc_genome *orc_genome_open(const char *path);
void orc_genome_close(orc_genome *g);
int orc_genome_count(const orc_genome *g);
const char *orc_genome_name(const orc_genome *g, int ix);
int64_t orc_genome_size(const orc_genome *g, int ix);
int orc_genome_find(const orc_genome *g, const char *name);
/* 1 char/base, mixed case, N for N-blocks; strand '-' gives the reverse complement
 * (dnautil.c:466-470).  Cached; owned by the genome. */
const char *orc_genome_dna(orc_genome *g, int ix, char strand);

/* ---- scoring scheme (axt.c:402-458, 692-834) + gap costs (gapCalc.c) */
orc_scoring *orc_scoring_new(const char *matrixFile /* NULL = default */, const char *linearGap);
void orc_scoring_free(orc_scoring *s);
int orc_matrix_at(const orc_scoring *s, int qChar, int tChar);
int orc_gap_cost(const orc_scoring *s, int dq, int dt);
int orc_gap_small_size(const orc_scoring *s);

/* ---- scoring (chainConnect.c:14-40, scoreChain.c:176-198, chain.c:471-558) */
double orc_score_block(const orc_scoring *s, const char *q, const char *t, int size);
/* Score every job of a CSR work-list.  Returns 0, or -1 with orc_last_error() set. */
int orc_score_jobs(const orc_scoring *s, orc_genome *tg, orc_genome *qg,
                   const orc_job *jobs, int64_t nJobs, int64_t totalJobBlocks,
                   const orc_block *blocks, int64_t nBlocks,
                   int64_t *global, int64_t *local, int64_t *aliBases);

/* Crossover point of two overlapping blocks (cBlockFindCrossover, chainConnect.c:61-105): left ends at
 * (leftTEnd, leftQEnd), right starts at (rightTStart, rightQStart), the last / first `overlap` bases of the two
 * cover the same stretch.  pos = offset from the start of the overlap at which to switch to the right block,
 * adjust = what the pair loses against the sum of both overlaps.  q/t: 1 char/base sequences (query in chain strand). */
void orc_find_crossover(const orc_scoring *s, const char *q, const char *t, int leftTEnd, int leftQEnd,
                        int rightTStart, int rightQStart, int overlap, int *pos, int *adjust);

/* ---- .chain (chain.c:256-346) */
orc_chainset *orc_chains_read(const char *path);
void orc_chains_free(orc_chainset *cs);
int64_t orc_chains_count(const orc_chainset *cs);
int64_t orc_chains_total_blocks(const orc_chainset *cs);
/* header fields of chain ix; names are owned by the chain set */
void orc_chains_header(const orc_chainset *cs, int64_t ix, double *score, const char **tName,
                       int *tSize, int *tStart, int *tEnd, const char **qName, int *qSize,
                       char *qStrand, int *qStart, int *qEnd, int *id,
                       int64_t *firstBlock, int64_t *nBlocks);
const orc_block *orc_chains_blocks(const orc_chainset *cs);
/* chainSubsetOnT's block selection (chain.c:479-510): which blocks of chain ix survive the
 * target range and what clip applies.  Returns 0 if the sub-chain is empty (NULL in kent). */
int orc_chains_subset(const orc_chainset *cs, int64_t ix, int subStart, int subEnd,
                      int64_t *firstBlock, int64_t *nBlocks, int32_t *clipStart, int32_t *clipEnd);

#ifdef __cplusplus
}
#endif
#endif
