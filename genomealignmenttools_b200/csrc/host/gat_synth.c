/* gat_synth.c -- host-side generator for the synthetic benchmark genomes.
 * The reference ships no genome data (hg38/mm10 .2bit must be downloaded, README.md:136-140) and
 * its example chain blob is missing, so throughput is reported on synthetic .2bit payloads at the
 * real chrom.sizes.  Random genomes give every chain a hopeless score, so the query is made
 * homologous: along every block the target bases are copied into the query with substitutions
 * (transitions twice as likely as each transversion), reverse-complemented on '-' chains.
 * Pure host code, no scoring in here. */
#include <stdint.h>
#include <stddef.h>

typedef struct { int32_t tStart, qStart; uint32_t size; } synth_block;
typedef struct { uint32_t tSeq, qSeq, firstBlock, blockPtr; int32_t clipStart, clipEnd; } synth_job;

static inline uint64_t mix64(uint64_t x)
{   /* splitmix64 finaliser: a counter-based generator, so results do not depend on visiting order */
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline unsigned getBase(const uint8_t *p, uint64_t pos)
{
    return (p[pos >> 2] >> (6 - 2 * (pos & 3))) & 3u;
}
static inline void setBase(uint8_t *p, uint64_t pos, unsigned code)
{
    unsigned sh = 6 - 2 * (pos & 3);
    p[pos >> 2] = (uint8_t)((p[pos >> 2] & ~(3u << sh)) | (code << sh));
}

/* Fill n bytes with reproducible pseudo-random bases. */
void gat_synth_fill(uint8_t *dst, uint64_t n, uint64_t seed)
{
    uint64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t r = mix64(seed ^ (i * 0x2545F4914F6CDD1Dull));
        for (int k = 0; k < 8; k++) dst[i + k] = (uint8_t)(r >> (8 * k));
    }
    if (i < n) {
        uint64_t r = mix64(seed ^ (i * 0x2545F4914F6CDD1Dull));
        for (int k = 0; i < n; i++, k++) dst[i] = (uint8_t)(r >> (8 * k));
    }
}

/* Overwrite the query along every job-block with a mutated copy of the target.
 * substPer64k: substitution probability in 1/65536 units.  Jobs are visited in order; where
 * chains overlap on the query the later one wins (deterministic). */
void gat_synth_plant(const uint8_t *tPacked, const uint64_t *tByteOff,
                     uint8_t *qPacked, const uint64_t *qByteOff, const uint32_t *qSize,
                     const synth_job *jobs, uint64_t nJobs, uint64_t totalJobBlocks,
                     const synth_block *blocks, uint32_t substPer64k, uint64_t seed)
{
    for (uint64_t j = 0; j < nJobs; j++) {
        const synth_job *job = &jobs[j];
        uint64_t nb = (j + 1 < nJobs ? jobs[j + 1].blockPtr : totalJobBlocks) - job->blockPtr;
        uint32_t qs = job->qSeq & 0x7fffffffu;
        int minus = job->qSeq >> 31;
        const uint8_t *t = tPacked + tByteOff[job->tSeq];
        uint8_t *q = qPacked + qByteOff[qs];
        uint64_t qLen = qSize[qs];
        for (uint64_t k = 0; k < nb; k++) {
            uint64_t bi = (uint64_t)job->firstBlock + k;
            const synth_block *b = &blocks[bi];
            uint32_t n = b->size & 0x7fffffffu;
            uint64_t key = mix64(seed ^ (bi << 20));
            for (uint32_t i = 0; i < n; i++) {
                unsigned code = getBase(t, (uint64_t)b->tStart + i);
                uint64_t r = mix64(key + i);
                if ((r & 0xffff) < substPer64k) {
                    unsigned kind = (r >> 16) & 3;          /* 0,1 transition; 2,3 the two transversions */
                    code ^= (kind <= 1) ? 1u : (kind == 2 ? 2u : 3u);
                }
                uint64_t p = (uint64_t)b->qStart + i;
                if (minus) setBase(q, qLen - 1 - p, code ^ 2u);  /* complement: T<->A, C<->G */
                else setBase(q, p, code);
            }
        }
    }
}
