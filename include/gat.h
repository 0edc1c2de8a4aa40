/* include/gat.h -- the drop-in boundary: a C ABI for batched chain rescoring on one B200.
 *
 * What it replaces.  In hillerlab/GenomeAlignmentTools every chain / net fill / suspect
 * sub-chain is scored by one synchronous CPU call
 *     double chainCalcScore(struct chain*, struct axtScoreScheme*, struct gapCalc*,
 *                           struct dnaSeq *query, struct dnaSeq *target)
 *                                                   kent/src/inc/chainConnect.h:34-38
 * (implementation kent/src/lib/chainConnect.c:14-40, gap costs kent/src/lib/gapCalc.c:298-331)
 * plus the hillerlab addition chainCalcScoreLocal (src/scoreChain/scoreChain.c:176-198,
 * src/chainCleaner/chainCleaner.c:531-551), reached from the tool wrappers getChainScore()
 * at src/scoreChain/scoreChain.c:207-220, src/chainNet/chainNet.c:230-248 and
 * src/chainCleaner/chainCleaner.c:559-571.  The host side of the tools keeps parsing
 * .chain/.2bit/.net; instead of calling chainCalcScore per chain it builds a CSR work-list
 * and makes ONE gat_score() call.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions.  Plain pointers and sizes only.  Every function returns 0 on success or a
 * negative GAT_E* code; gat_last_error() describes the failure (tool wrappers turn that into
 * kent's errAbort: message on stderr, exit(-1), kent/src/lib/errAbort.c:182-197).  There is
 * no CPU fallback: without a CUDA device gat_create fails.  One context drives one GPU; calls on
 * a context are serialised by the caller (the reference is single-threaded).  Scores are int64
 * and equal, as integers, the doubles kent computes (it sums ints into a double; exact < 2^53).
 */
#ifndef GAT_H
#define GAT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAT_OK 0
#define GAT_ECUDA (-1)    /* CUDA runtime failure */
#define GAT_EINVAL (-2)   /* bad argument */
#define GAT_ESTATE (-3)   /* genome / scoring not loaded yet */
#define GAT_EWORKLIST (-4)/* the device found an out-of-range record in the work-list */
#define GAT_ENOMEM (-5)

#define GAT_TARGET 0
#define GAT_QUERY 1

#define GAT_QSEQ_MINUS 0x80000000u  /* gat_job.qSeq bit: query on '-' strand (chain.h:58 qStrand) */
#define GAT_BLOCK_JOINED 0x80000000u/* gat_block.size bit: this record continues the previous record's
                                     * gapless block (a long block split by the host for load
                                     * balance): no gap cost, no local-score clamp between them */
#define GAT_MAX_BLOCK_BASES ((1u << 20) - 1) /* largest gat_block.size the kernels accept; with matrix entries beyond
                                             * +-2047 the limit drops to (2^31-1)/max|M| (32-bit record sums), never
                                             * below GAT_SPLIT_BASES: see gat_max_record_bases() */
#define GAT_SPLIT_BASES 4096u                /* recommended record size: hosts cut longer gapless blocks into
                                             * JOINED records of this many bases so that a few very long blocks
                                             * spread over many warps (gathost::buildRecords does) */
#define GAT_NO_CLIP_START INT32_MIN
#define GAT_NO_CLIP_END INT32_MAX

/* One gapless block (struct cBlock, kent/src/inc/chain.h:17-25, minus the list pointer):
 * target [tStart, tStart+size), query [qStart, qStart+size), 0-based half-open, per sequence;
 * query coordinates of a '-' chain are in reverse-complement space (chainFormat.doc). */
typedef struct gat_block {
    int32_t tStart;
    int32_t qStart;
    uint32_t size;
} gat_block;

/* One scoring job = one chainCalcScore call of the reference: a chain, or the part of a chain
 * inside a target range (chainSubsetOnT, kent/src/lib/chain.c:471-558).  Jobs form a CSR over
 * "job-blocks": job j owns job-blocks [blockPtr[j], blockPtr[j+1]) which are the records
 * blocks[firstBlock .. firstBlock + n).  Several jobs may point into the same records (net
 * fills / sub-chains of one chain).  Every kept record is clipped to [clipStart, clipEnd) on
 * the target with the query moved by the same delta (chain.c:513-522). */
typedef struct gat_job {
    uint32_t tSeq;       /* target sequence index (order of gat_load_genome) */
    uint32_t qSeq;       /* query sequence index | GAT_QSEQ_MINUS */
    uint32_t firstBlock; /* index into blocks[] */
    uint32_t blockPtr;   /* CSR row pointer: sum of block counts of jobs 0..j-1 */
    int32_t clipStart;   /* GAT_NO_CLIP_START when the whole chain is scored */
    int32_t clipEnd;     /* GAT_NO_CLIP_END   "                              */
} gat_job;

/* One run of N in a sequence (nStarts/nSizes of the .2bit record, kent/src/inc/twoBit.h:9-22). */
typedef struct gat_nrun {
    uint32_t seq;
    uint32_t start;
    uint32_t len;
} gat_nrun;

/* Scoring parameters = struct axtScoreScheme's live 4x4 (kent/src/inc/axt.h:83-91; every other
 * cell of its 256x256 matrix is 0, which is how N scores 0) + the tables struct gapCalc holds
 * (kent/src/lib/gapCalc.c:12-37). */
typedef struct gat_scoring {
    int32_t matrix[4][4];   /* [query base][target base], base codes T=0 C=1 A=2 G=3 (dnautil.h:23-27) */
    int32_t smallSize;      /* gaps < smallSize are table look-ups */
    const int32_t *qSmall;  /* [smallSize], entry 0 is 0 */
    const int32_t *tSmall;
    const int32_t *bSmall;
    int32_t longCount;      /* knots for gaps >= smallSize; longPos[0] == smallSize */
    const int32_t *longPos; /* [longCount] */
    const double *qLong;    /* [longCount] */
    const double *tLong;
    const double *bLong;
} gat_scoring;

typedef struct gat_ctx gat_ctx;
typedef struct gat_worklist gat_worklist;

/* Timing / traffic of the most recent scoring call, for bench.py and the tools' -verbose line. */
typedef struct gat_stats {
    float score_kernel_ms;  /* the block-scoring kernel alone (CUDA events on the ctx stream) */
    float all_kernels_ms;   /* chunk index + scoring + cross-chunk fix-up */
    float h2d_ms, d2h_ms;   /* 0 for resident work-lists */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t kernel_launches;
    uint32_t chunks;        /* CTAs of the scoring kernel */
    uint32_t long_streamed; /* 1: the instantiation that streams blocks of more than 1056 bases ran (picked from a sample of
                               the list's block sizes; both instantiations score every list exactly) */
    uint32_t reserved;
} gat_stats;

const char *gat_last_error(void);
int gat_device_count(void);

/* Context on CUDA device `device`.  `stream` is a cudaStream_t to launch on (e.g. the caller's
 * current stream, so that its CUDA events bracket our kernels); NULL = a stream the ctx owns. */
int gat_create(gat_ctx **out, int device, void *stream);
void gat_destroy(gat_ctx *ctx);

/* Upload one genome and keep it resident (replaces twoBitReadSeqFrag's unpack-to-chars,
 * kent/src/lib/twoBit.c:725-878, and the whole-chromosome reverseComplement copies of
 * scoreChain.c:123-149).  `packed` holds each sequence's .2bit payload (4 bases/byte, first
 * base in bits 7..6) starting at byte seqByteOffset[i]; nRuns must be grouped by sequence in
 * file order. */
int gat_load_genome(gat_ctx *ctx, int side, const uint8_t *packed, uint64_t packedBytes,
                    const uint64_t *seqByteOffset, const uint32_t *seqSize, uint32_t nSeq,
                    const gat_nrun *nRuns, uint64_t nNRuns);

int gat_set_scoring(gat_ctx *ctx, const gat_scoring *scoring);
/* Longest record (gat_block.size) gat_score accepts under the current scoring parameters. */
uint32_t gat_max_record_bases(const gat_ctx *ctx);

/* The hot call.  Host arrays in, host arrays out, blocking.  `totalJobBlocks` closes the CSR
 * (= blockPtr of a virtual job nJobs).  global[j] = chainCalcScore, local[j] =
 * chainCalcScoreLocal of job j.  (aliBases is the plain sum of clipped sizes: host arithmetic.) */
int gat_score(gat_ctx *ctx, const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks,
              const gat_block *blocks, uint64_t nBlocks, int64_t *global, int64_t *local);

/* Same work split into upload / run / download, for callers that keep a work-list resident
 * (repeated scoring with different matrices, benchmarking the kernels without PCIe). */
int gat_worklist_create(gat_ctx *ctx, const gat_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks,
                        const gat_block *blocks, uint64_t nBlocks, gat_worklist **out);
int gat_worklist_run(gat_ctx *ctx, gat_worklist *wl);           /* asynchronous on the ctx stream */
int gat_worklist_results(gat_ctx *ctx, gat_worklist *wl, int64_t *global, int64_t *local);
void gat_worklist_destroy(gat_ctx *ctx, gat_worklist *wl);

/* The same scoring call fed with a compact work-list: what crosses PCIe is about half of what gat_score() needs
 * (6 bytes per block instead of 12, 8 per chain instead of 24), and a .chain file stores exactly these numbers
 * ("size dt dq" lines, kent/src/lib/chain.c:211-227).  The device expands it into gat_job / gat_block records and then
 * runs the kernels of gat_score().  Whole chains only (job j owns records [blockPtr[j], blockPtr[j+1]), no clip).
 *   gat_cblock: size (<= GAT_CBLOCK_MAX_SIZE; cut longer blocks into GAT_CBLOCK_JOINED pieces) and the gap in front of
 *               the block on both sequences.  GAT_CBLOCK_ABS: the block does not start relative to its predecessor (first
 *               block of a chain, a gap of 65536 or more, a negative gap): dt | dq << 16 is an index into abs[].
 *   anchors[g]: start of record 1024 * g, so that groups of 1024 records expand independently. */
#define GAT_CBLOCK_ABS 0x8000u
#define GAT_CBLOCK_JOINED 0x4000u
#define GAT_CBLOCK_MAX_SIZE 0x3fffu
#define GAT_CJOB_MINUS 0x8000u
#define GAT_CGROUP 1024u
typedef struct gat_cblock { uint16_t size, dt, dq; } gat_cblock;
typedef struct gat_cjob { uint32_t blockPtr; uint16_t tSeq, qSeq; /* qSeq | GAT_CJOB_MINUS */ } gat_cjob;
typedef struct gat_cabs { int32_t tStart, qStart; } gat_cabs;
int gat_score_compact(gat_ctx *ctx, const gat_cjob *jobs, uint64_t nJobs, const gat_cblock *blocks, uint64_t nBlocks,
                      const gat_cabs *abs, uint64_t nAbs, const gat_cabs *anchors, int64_t *global, int64_t *local);

/* The same list in 4 bytes per block (gat_score_packed): what most records of a .chain file need.  gat_pblock is one word,
 *   bits 0..11 size (<= GAT_PBLOCK_MAX_SIZE; longer blocks are cut into GAT_PBLOCK_JOINED pieces), bit 12 GAT_PBLOCK_JOINED,
 *   bit 13 GAT_PBLOCK_ABS, bits 14..22 dt, bits 23..31 dq (the gap in front of the block, each <= 511).
 * A block that opens a chain, or whose gap is negative or above 511 on either side, is absolute: it takes the NEXT entry of abs[] in
 * list order (no index is stored: entry k belongs to the k-th absolute record of the list), and absBase[g] = number of absolute
 * records in front of record 1024 * g lets the groups expand independently.  jobs, abs and anchors as in gat_score_compact.
 * About 59 MB instead of 77 MB per 10 M blocks of a typical chain file; the scores are those of gat_score on the expanded list. */
#define GAT_PBLOCK_MAX_SIZE 0xfffu
#define GAT_PBLOCK_JOINED 0x1000u
#define GAT_PBLOCK_ABS 0x2000u
#define GAT_PBLOCK_MAX_GAP 511u
typedef uint32_t gat_pblock;
int gat_score_packed(gat_ctx *ctx, const gat_cjob *jobs, uint64_t nJobs, const gat_pblock *blocks, uint64_t nBlocks,
                     const gat_cabs *abs, uint64_t nAbs, const gat_cabs *anchors, const uint32_t *absBase, int64_t *global, int64_t *local);

/* Crossover points of overlapping adjacent blocks, in one batch: what kent's chainRemovePartialOverlaps asks of
 *     void cBlockFindCrossover(struct cBlock *left, struct cBlock *right, struct dnaSeq *qSeq, struct dnaSeq *tSeq,
 *                              int overlap, int matrix[256][256], int *retPos, int *retScoreAdjustment)
 *                                                   kent/src/inc/chainConnect.h, kent/src/lib/chainConnect.c:61-105
 * once per overlapping pair (kent/src/lib/chainConnect.c:284-296; axtChain's last step before it rescoring).
 * The left block ends at (leftTEnd, leftQEnd), the right block starts at (rightTStart, rightQStart); their last / first
 * `overlap` bases are compared base by base.  pos[i] = offset from the start of the overlap at which the right block
 * takes over (0..overlap, the first best one), adjust[i] = retScoreAdjustment.  Query coordinates are the chain's own
 * (reverse-complement space on '-').  Uses the matrix of gat_set_scoring; gap tables play no part. */
typedef struct gat_xpair {
    uint32_t tSeq;        /* target sequence index */
    uint32_t qSeq;        /* query sequence index | GAT_QSEQ_MINUS */
    int32_t leftTEnd, leftQEnd;
    int32_t rightTStart, rightQStart;
    int32_t overlap;      /* >= 0 */
} gat_xpair;
int gat_crossover(gat_ctx *ctx, const gat_xpair *pairs, uint64_t nPairs, int32_t *pos, int32_t *adjust);

/* Chains split over several GPUs (a job with more than ~1/(4 nGPU) of the aligned bases is cut at block
 * boundaries).  A part is scored as a job of its own; its global score alone cannot be joined into the chain's local
 * score (chainCalcScoreLocal clamps at 0 and keeps a running maximum, src/scoreChain/scoreChain.c:181-195), so a part
 * comes back as the 4-number tuple (d, c, e, f) meaning "entered with running score s and best M, the part leaves
 * s' = max(c, s + d) and M' = max(M, s + e, f)"; d alone is the part's global score.  Parts join on the host:
 *     acc = tuple(part 0); for every further part: gat_tuple_join(&acc, gapCalcCost between the parts, &tuple(part));
 *     gat_tuple_scores(&acc, &global, &local);
 * which is bit-identical to scoring the unsplit chain.  gat_request_tuples() names the jobs of the NEXT scoring call on the
 * context (gat_score, gat_score_compact or gat_worklist_run + gat_worklist_results) whose tuples are wanted; out[k] is
 * filled when that call returns its results.  Every such job must own at least GAT_TUPLE_MIN_BLOCKS job-blocks (parts
 * are large by construction) and must not start with a GAT_BLOCK_JOINED record. */
#define GAT_TUPLE_MIN_BLOCKS 1024u
typedef struct gat_tuple { int64_t d, c, e, f; } gat_tuple;
int gat_request_tuples(gat_ctx *ctx, const uint32_t *jobIx, uint64_t n, gat_tuple *out);
void gat_tuple_join(gat_tuple *acc, int64_t gapCost, const gat_tuple *next);
void gat_tuple_scores(const gat_tuple *t, int64_t *global, int64_t *local);

/* gapCalcCost (kent/src/lib/gapCalc.c:298-331) for a batch of (dq, dt) pairs, evaluated on the device by the routines the
 * scoring kernel uses, under the tables of gat_set_scoring.  (kent's own gapCalcCost in include/gat_kent.h is this with n = 1.) */
int gat_gap_cost(gat_ctx *ctx, const int32_t *dq, const int32_t *dt, uint64_t n, int32_t *out);

int gat_synchronize(gat_ctx *ctx);
int gat_get_stats(gat_ctx *ctx, gat_stats *out);
/* When on, gat_worklist_run/gat_score bracket each kernel with CUDA events (filled into gat_stats). */
int gat_set_profiling(gat_ctx *ctx, int on);

/* Pinned host memory for work-lists and results (cudaHostAlloc): PCIe at full rate. */
void *gat_host_alloc(size_t bytes);
void gat_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
