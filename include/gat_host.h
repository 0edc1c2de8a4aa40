/* include/gat_host.h -- C view of the host-side library (libgathost.so): the parsers and the
 * work-list builder the tools use in front of include/gat.h.  Exists so that bindings and tests
 * can exercise exactly the code the command-line tools run.  Every function returns 0 or -1 with
 * gathost_last_error() set; nothing in here computes a score. */
#ifndef GAT_HOST_H
#define GAT_HOST_H
#include <stdint.h>
#include "gat.h"
#ifdef __cplusplus
extern "C" {
#endif

const char *gathost_last_error(void);

/* gapCalcFromFile (kent/src/lib/gapCalc.c:233-256): tables for "loose", "medium" or a file.
 * Arrays are owned by the handle. */
typedef struct gathost_gapcalc gathost_gapcalc;
gathost_gapcalc *gathost_gapcalc_open(const char *name);
void gathost_gapcalc_close(gathost_gapcalc *g);
int gathost_gapcalc_cost(const gathost_gapcalc *g, int dq, int dt);      /* gapCalcCost, :298-331 */
int gathost_gapcalc_fill(const gathost_gapcalc *g, gat_scoring *out);    /* gap part of gat_scoring */

/* axtScoreSchemeRead / axtScoreSchemeDefault (kent/src/lib/axt.c:423-458, 692-834); path NULL = default. */
int gathost_scorescheme(const char *path, int32_t matrix[4][4]);

/* chainRead over a whole file (kent/src/lib/chain.c:256-346) into CSR arrays owned by the handle. */
typedef struct gathost_chains gathost_chains;
gathost_chains *gathost_chains_read(const char *path);
void gathost_chains_close(gathost_chains *c);
uint64_t gathost_chains_count(const gathost_chains *c);
uint64_t gathost_chains_block_count(const gathost_chains *c);
const gat_block *gathost_chains_blocks(const gathost_chains *c);
int gathost_chains_head(const gathost_chains *c, uint64_t ix, double *score, const char **tName, int *tSize,
                        int *tStart, int *tEnd, const char **qName, int *qSize, char *qStrand, int *qStart,
                        int *qEnd, int *id, uint64_t *firstBlock, uint64_t *nBlocks);
/* chainSubsetOnT's selection (chain.c:471-558): returns 1 and the record range + clip, 0 for NULL. */
int gathost_chains_subset(const gathost_chains *c, uint64_t ix, int subStart, int subEnd, uint64_t *firstBlock,
                          uint64_t *nBlocks, int32_t *clipStart, int32_t *clipEnd, int64_t *aliBases);

/* chainRemovePartialOverlaps (kent/src/lib/chainConnect.c:255-344) on every chain of the set, crossover points from
 * gat_crossover() on `ctx` (genomes and scoring loaded); chainT / chainQ = sequence index per chain as uploaded.
 * Blocks, counts and bounds are rewritten in place (read them back with gathost_chains_head / _blocks). */
int gathost_chains_remove_partial_overlaps(gathost_chains *c, gat_ctx *ctx, const uint32_t *chainT, const uint32_t *chainQ);

/* The chains of the set as the compact work-list of gat_score_compact (gathost::packCompact): one whole-chain job per
 * chain.  NULL (gathost_last_error) if the set does not qualify (a record beyond GAT_CBLOCK_MAX_SIZE cannot happen:
 * records are cut at GAT_SPLIT_BASES; more than 32768 sequences can).  Arrays are owned by the handle. */
typedef struct gathost_compact gathost_compact;
gathost_compact *gathost_chains_compact(const gathost_chains *c, const uint32_t *chainT, const uint32_t *chainQ);
void gathost_compact_free(gathost_compact *p);
int gathost_compact_view(const gathost_compact *p, const gat_cjob **jobs, uint64_t *nJobs, const gat_cblock **blocks, uint64_t *nBlocks,
                         const gat_cabs **abs, uint64_t *nAbs, const gat_cabs **anchors, uint64_t *nAnchors);

/* .2bit container (kent/src/lib/twoBit.c:422-650). */
typedef struct gathost_twobit gathost_twobit;
gathost_twobit *gathost_twobit_open(const char *path);
void gathost_twobit_close(gathost_twobit *t);
uint32_t gathost_twobit_count(const gathost_twobit *t);
int gathost_twobit_seq(const gathost_twobit *t, uint32_t ix, const char **name, uint32_t *size, const uint8_t **packed,
                       uint32_t *nRuns, const uint32_t **nStart, const uint32_t **nLen);

/* Greedy aligned-base sharding of jobs over `parts` GPUs; part[j] receives the GPU of job j. */
int gathost_shard_jobs(const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks, const int64_t *aliBases,
                       int parts, uint32_t *part);

#ifdef __cplusplus
}
#endif
#endif
