/* oracle/chain_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see chain_oracle.h).
 *
 * A from-scratch CPU restatement of what the reference computes on the chain-rescoring path.
 * It deliberately keeps the reference's *numerical shape* -- one char per base, a 256x256 int
 * matrix indexed [query char][target char], block sums accumulated in a double, gap costs from
 * double interpolation truncated to int -- so that every intermediate has the same value as in
 * kent's code.  Each function names the reference lines it restates.  Compile with
 * -ffp-contract=off (the reference binaries are plain x86-64 SSE2: no fused multiply-add).
 */
#include "chain_oracle.h"
#include <ctype.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char errBuf[512];
const char *orc_last_error(void) { return errBuf; }
static void setErr(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errBuf, sizeof errBuf, fmt, ap);
    va_end(ap);
}

/* ====================================================================== .2bit */

typedef struct {
    char *name;
    uint64_t offset;       /* file offset of the record */
    uint32_t size;         /* bases */
    uint32_t nCount, mCount;
    uint32_t *nStart, *nLen, *mStart, *mLen;
    unsigned char *packed; /* (size+3)/4 bytes, first base in the two high bits */
    char *fwd, *rev;       /* unpacked caches */
} orcSeq;

struct orc_genome {
    int count;
    orcSeq *seq;
};

static uint32_t rd32(const unsigned char *p, int swapped)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return swapped ? __builtin_bswap32(v) : v;
}
static uint64_t rd64(const unsigned char *p, int swapped)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return swapped ? __builtin_bswap64(v) : v;
}

orc_genome *orc_genome_open(const char *path)
/* File layout: twoBit.c:312-397 (writer) and :442-513, :574-650 (reader).
 * Signature 0x1A412743 (sig.h:58-62), byte-swapped files accepted, version 0 or 1. */
{
    FILE *f = fopen(path, "rb");
    if (!f) { setErr("cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long fileSize = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *buf = malloc(fileSize > 0 ? fileSize : 1);
    if (fread(buf, 1, fileSize, f) != (size_t)fileSize) { fclose(f); free(buf); setErr("short read %s", path); return NULL; }
    fclose(f);
    if (fileSize < 16) { free(buf); setErr("%s too short", path); return NULL; }
    uint32_t sig;
    memcpy(&sig, buf, 4);
    int swapped;
    if (sig == 0x1A412743u) swapped = 0;
    else if (sig == 0x4327411Au) swapped = 1;
    else { free(buf); setErr("%s doesn't have a valid twoBitSig", path); return NULL; }
    uint32_t version = rd32(buf + 4, swapped);
    if (version > 1) { free(buf); setErr("2bit version %u unsupported", version); return NULL; }
    orc_genome *g = calloc(1, sizeof *g);
    g->count = (int)rd32(buf + 8, swapped);
    g->seq = calloc(g->count ? g->count : 1, sizeof(orcSeq));
    size_t pos = 16;
    for (int i = 0; i < g->count; i++) {
        int nameLen = buf[pos++];
        g->seq[i].name = malloc(nameLen + 1);
        memcpy(g->seq[i].name, buf + pos, nameLen);
        g->seq[i].name[nameLen] = 0;
        pos += nameLen;
        if (version == 1) { g->seq[i].offset = rd64(buf + pos, swapped); pos += 8; }
        else { g->seq[i].offset = rd32(buf + pos, swapped); pos += 4; }
    }
    for (int i = 0; i < g->count; i++) {
        orcSeq *s = &g->seq[i];
        const unsigned char *p = buf + s->offset;
        s->size = rd32(p, swapped); p += 4;
        s->nCount = rd32(p, swapped); p += 4;
        s->nStart = malloc(sizeof(uint32_t) * (s->nCount + 1));
        s->nLen = malloc(sizeof(uint32_t) * (s->nCount + 1));
        for (uint32_t k = 0; k < s->nCount; k++) s->nStart[k] = rd32(p + 4 * k, swapped);
        p += 4 * (size_t)s->nCount;
        for (uint32_t k = 0; k < s->nCount; k++) s->nLen[k] = rd32(p + 4 * k, swapped);
        p += 4 * (size_t)s->nCount;
        s->mCount = rd32(p, swapped); p += 4;
        s->mStart = malloc(sizeof(uint32_t) * (s->mCount + 1));
        s->mLen = malloc(sizeof(uint32_t) * (s->mCount + 1));
        for (uint32_t k = 0; k < s->mCount; k++) s->mStart[k] = rd32(p + 4 * k, swapped);
        p += 4 * (size_t)s->mCount;
        for (uint32_t k = 0; k < s->mCount; k++) s->mLen[k] = rd32(p + 4 * k, swapped);
        p += 4 * (size_t)s->mCount;
        p += 4; /* reserved */
        size_t packedBytes = ((size_t)s->size + 3) / 4;
        s->packed = malloc(packedBytes ? packedBytes : 1);
        memcpy(s->packed, p, packedBytes);
    }
    free(buf);
    return g;
}

void orc_genome_close(orc_genome *g)
{
    if (!g) return;
    for (int i = 0; i < g->count; i++) {
        orcSeq *s = &g->seq[i];
        free(s->name); free(s->nStart); free(s->nLen); free(s->mStart); free(s->mLen);
        free(s->packed); free(s->fwd); free(s->rev);
    }
    free(g->seq);
    free(g);
}

int orc_genome_count(const orc_genome *g) { return g->count; }
const char *orc_genome_name(const orc_genome *g, int ix) { return g->seq[ix].name; }
int64_t orc_genome_size(const orc_genome *g, int ix) { return g->seq[ix].size; }
int orc_genome_find(const orc_genome *g, const char *name)
{
    for (int i = 0; i < g->count; i++)
        if (strcmp(g->seq[i].name, name) == 0) return i;
    return -1;
}

static void unpackSeq(orcSeq *s)
/* twoBitReadSeqFragExt for the whole sequence with doMask=TRUE (twoBit.c:725-878):
 * 2 bits/base, T=0 C=1 A=2 G=3 (dnautil.h:23-27), first base in bits 7..6 (:811-818);
 * N blocks overlay 'n' (loop stops at the first block starting at/after the end, :838-850),
 * everything upper-cased (:855), then mask blocks lower-cased (:856-873). */
{
    static const char code[4] = { 'T', 'C', 'A', 'G' };
    int64_t n = s->size;
    char *d = malloc(n + 1);
    for (int64_t i = 0; i < n; i++)
        d[i] = code[(s->packed[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    d[n] = 0;
    for (uint32_t k = 0; k < s->nCount; k++) {
        int64_t a = (int)s->nStart[k], b = a + (int)s->nLen[k];
        if (a >= n) break;
        if (a < 0) a = 0;
        if (b > n) b = n;
        for (int64_t i = a; i < b; i++) d[i] = 'N';
    }
    for (uint32_t k = 0; k < s->mCount; k++) {
        int64_t a = (int)s->mStart[k], b = a + (int)s->mLen[k];
        if (a >= n) break;
        if (a < 0) a = 0;
        if (b > n) b = n;
        for (int64_t i = a; i < b; i++) d[i] = (char)tolower((unsigned char)d[i]);
    }
    s->fwd = d;
}

static char complementChar(char c)
/* ntCompTable for the characters a .2bit can yield (dnautil.c:403-425): a<->t, c<->g, n->n. */
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;
    }
}

const char *orc_genome_dna(orc_genome *g, int ix, char strand)
{
    orcSeq *s = &g->seq[ix];
    if (!s->fwd) unpackSeq(s);
    if (strand != '-') return s->fwd;
    if (!s->rev) {
        /* reverseComplement of the whole chromosome (scoreChain.c:141-142, dnautil.c:466-470) */
        int64_t n = s->size;
        char *r = malloc(n + 1);
        for (int64_t i = 0; i < n; i++) r[i] = complementChar(s->fwd[n - 1 - i]);
        r[n] = 0;
        s->rev = r;
    }
    return s->rev;
}

/* ====================================================================== scoring scheme */

struct orc_scoring {
    int matrix[256][256];
    int smallSize;
    int *qSmall, *tSmall, *bSmall;
    int longCount;
    int *longPos;
    double *qLong, *tLong, *bLong;
    int lastPos;
    double qLastVal, tLastVal, bLastVal, qLastSlope, tLastSlope, bLastSlope;
};

static const char *mediumSpec = /* "original" costs, gapCalc.c:40-46 */
    "tableSize 11\nsmallSize 111\n"
    "position 1 2 3 11 111 2111 12111 32111 72111 152111 252111\n"
    "qGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900\n"
    "tGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900\n"
    "bothGap 750 825 850 1000 1300 3300 23300 58300 118300 218300 318300\n";
static const char *looseSpec = /* "default" costs, gapCalc.c:50-56 */
    "tableSize 11\nsmallSize 111\n"
    "position 1 2 3 11 111 2111 12111 32111 72111 152111 252111\n"
    "qGap 325 360 400 450 600 1100 3600 7600 15600 31600 56600\n"
    "tGap 325 360 400 450 600 1100 3600 7600 15600 31600 56600\n"
    "bothGap 625 660 700 750 900 1400 4000 8000 16000 32000 57000\n";

static int truncToInt(double d)
/* C's (int)double on x86-64 is cvttsd2si: truncation toward zero, 0x80000000 when out of range. */
{
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
    return (int)d;
}

static int interp(int x, const int *pos, const double *val, int n)
/* gapCalc.c:82-104.  Exact knot -> its value; inside -> v[i-1] + dv*(x-s[i-1])/ds evaluated
 * left to right in double; beyond the last knot -> extrapolate the last segment. */
{
    for (int i = 0; i < n; i++) {
        if (x == pos[i]) return truncToInt(val[i]);
        if (x < pos[i]) {
            int ds = pos[i] - pos[i - 1];
            double dv = val[i] - val[i - 1];
            double prod = dv * (double)(x - pos[i - 1]);
            double quot = prod / (double)ds;
            return truncToInt(val[i - 1] + quot);
        }
    }
    int ds = pos[n - 1] - pos[n - 2];
    double dv = val[n - 1] - val[n - 2];
    double prod = dv * (double)(x - pos[n - 2]);
    double quot = prod / (double)ds;
    return truncToInt(val[n - 2] + quot);
}

static char *nextRealLine(char **cursor)
/* lineFileNextReal (linefile.c:878-891): skip blank lines and lines whose first non-space is '#'. */
{
    while (**cursor) {
        char *line = *cursor, *nl = strchr(line, '\n');
        if (nl) { *nl = 0; *cursor = nl + 1; } else *cursor = line + strlen(line);
        char *s = line;
        while (*s && isspace((unsigned char)*s)) s++;
        if (*s && *s != '#') return line;
    }
    return NULL;
}

static int taggedNumbers(char **cursor, const char *tag, int count, int *iOut, double *dOut)
/* readTaggedNumLine (gapCalc.c:112-144): tag is compared case-insensitively (sameWord). */
{
    char *line = nextRealLine(cursor);
    if (!line) { setErr("gap spec ends before %s", tag); return -1; }
    char *save = NULL, *w = strtok_r(line, " \t\r", &save);
    if (!w || strcasecmp(w, tag) != 0) { setErr("Expecting %s got %s", tag, w ? w : "(nothing)"); return -1; }
    for (int i = 0; i < count; i++) {
        w = strtok_r(NULL, " \t\r", &save);
        if (!w) { setErr("Not enough numbers on %s line", tag); return -1; }
        if (!isdigit((unsigned char)w[0])) { setErr("Expecting number got %s", w); return -1; }
        if (iOut) iOut[i] = atoi(w);
        if (dOut) dOut[i] = atof(w);
    }
    if (strtok_r(NULL, " \t\r", &save)) { setErr("Too many numbers on %s line", tag); return -1; }
    return 0;
}

static int buildGapTables(orc_scoring *s, const char *specText)
/* gapCalcRead, gapCalc.c:146-222. */
{
    char *text = strdup(specText), *cur = text;
    int tableSize = 0, rc = -1;
    int *pos = NULL;
    double *qv = NULL, *tv = NULL, *bv = NULL;
    if (taggedNumbers(&cur, "tableSize", 1, &tableSize, NULL)) goto done;
    if (taggedNumbers(&cur, "smallSize", 1, &s->smallSize, NULL)) goto done;
    if (tableSize < 2 || s->smallSize < 1) { setErr("bad gap table sizes"); goto done; }
    pos = calloc(tableSize, sizeof *pos);
    qv = calloc(tableSize, sizeof *qv); tv = calloc(tableSize, sizeof *tv); bv = calloc(tableSize, sizeof *bv);
    if (taggedNumbers(&cur, "position", tableSize, pos, NULL)) goto done;
    if (taggedNumbers(&cur, "qGap", tableSize, NULL, qv)) goto done;
    if (taggedNumbers(&cur, "tGap", tableSize, NULL, tv)) goto done;
    if (taggedNumbers(&cur, "bothGap", tableSize, NULL, bv)) goto done;
    if (pos[0] > 1) { setErr("gap table must start at position 1 (reference reads out of bounds otherwise)"); goto done; }
    s->qSmall = calloc(s->smallSize, sizeof(int));
    s->tSmall = calloc(s->smallSize, sizeof(int));
    s->bSmall = calloc(s->smallSize, sizeof(int));
    for (int i = 1; i < s->smallSize; i++) { /* entry 0 stays 0: gapCalc.c:171-182 starts at 1 */
        s->qSmall[i] = interp(i, pos, qv, tableSize);
        s->tSmall[i] = interp(i, pos, tv, tableSize);
        s->bSmall[i] = interp(i, pos, bv, tableSize);
    }
    int startLong = -1;
    for (int i = 0; i < tableSize; i++)
        if (pos[i] == s->smallSize) { startLong = i; break; }
    if (startLong < 0) { setErr("No position %d in gapCalcRead()", s->smallSize); goto done; }
    s->longCount = tableSize - startLong;
    if (s->longCount < 2) { setErr("need two long positions"); goto done; }
    s->longPos = malloc(sizeof(int) * s->longCount);
    s->qLong = malloc(sizeof(double) * s->longCount);
    s->tLong = malloc(sizeof(double) * s->longCount);
    s->bLong = malloc(sizeof(double) * s->longCount);
    for (int i = 0; i < s->longCount; i++) {
        s->longPos[i] = pos[startLong + i];
        s->qLong[i] = qv[startLong + i]; s->tLong[i] = tv[startLong + i]; s->bLong[i] = bv[startLong + i];
    }
    int L = s->longCount;
    s->lastPos = s->longPos[L - 1];
    s->qLastVal = s->qLong[L - 1]; s->tLastVal = s->tLong[L - 1]; s->bLastVal = s->bLong[L - 1];
    double dx = (double)s->lastPos - (double)s->longPos[L - 2]; /* calcSlope, gapCalc.c:106-110 */
    s->qLastSlope = (s->qLastVal - s->qLong[L - 2]) / dx;
    s->tLastSlope = (s->tLastVal - s->tLong[L - 2]) / dx;
    s->bLastSlope = (s->bLastVal - s->bLong[L - 2]) / dx;
    rc = 0;
done:
    free(text); free(pos); free(qv); free(tv); free(bv);
    return rc;
}

static char *slurp(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *t = malloc(n + 1);
    if (fread(t, 1, n, f) != (size_t)n) { fclose(f); free(t); return NULL; }
    t[n] = 0;
    fclose(f);
    return t;
}

static void spreadCase(orc_scoring *s)
/* propagateCase, axt.c:402-421: upper/lower/mixed all score like lower/lower. */
{
    static const char lo[4] = { 'a', 'c', 'g', 't' }, up[4] = { 'A', 'C', 'G', 'T' };
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            int v = s->matrix[(int)lo[a]][(int)lo[b]];
            s->matrix[(int)lo[a]][(int)up[b]] = v;
            s->matrix[(int)up[a]][(int)lo[b]] = v;
            s->matrix[(int)up[a]][(int)up[b]] = v;
        }
}

static int chopWords(char *line, char **words, int maxWords)
{
    int n = 0;
    char *save = NULL;
    for (char *w = strtok_r(line, " \t\r", &save); w && n < maxWords; w = strtok_r(NULL, " \t\r", &save))
        words[n++] = w;
    return n;
}

static char *nextChopLine(char **cursor)
/* lineFileChopNext (linefile.c:907-922): skip lines that START with '#' and lines with no words. */
{
    while (**cursor) {
        char *line = *cursor, *nl = strchr(line, '\n');
        if (nl) { *nl = 0; *cursor = nl + 1; } else *cursor = line + strlen(line);
        if (line[0] == '#') continue;
        char *s = line;
        while (*s && isspace((unsigned char)*s)) s++;
        if (*s) return line;
    }
    return NULL;
}

static int readMatrixFile(orc_scoring *s, const char *path)
/* axtScoreSchemeReadLf, axt.c:692-819: optional key=value lines, then "A C G T", then four rows
 * of four numbers (five words = row label first).  File order is A,C,G,T; rows are the first
 * matrix index.  A following line, if present, must hold O= and E= (both > 0). */
{
    static const char order[4] = { 'a', 'c', 'g', 't' };
    char *text = slurp(path), *cur = text, *w[6];
    if (!text) { setErr("cannot open %s", path); return -1; }
    int rc = -1;
    for (;;) {
        char *line = nextChopLine(&cur);
        if (!line) { setErr("Scoring matrix file %s too short", path); goto done; }
        char *copy = strdup(line);
        int n = chopWords(copy, w, 6);
        int isSetting = strchr(w[0], '=') != NULL || (n > 1 && strchr(w[1], '=') != NULL);
        if (isSetting) { free(copy); continue; }
        int ok = n >= 4 && w[0][0] == 'A' && w[1][0] == 'C' && w[2][0] == 'G' && w[3][0] == 'T';
        free(copy);
        if (!ok) { setErr("%s doesn't seem to be a score matrix file", path); goto done; }
        break;
    }
    for (int i = 0; i < 4; i++) {
        char *line = nextChopLine(&cur);
        if (!line) { setErr("Scoring matrix file %s too short", path); goto done; }
        int n = chopWords(line, w, 6);
        int first = (n == 5) ? 1 : 0;
        if (n < first + 4) { setErr("matrix row %d of %s too short", i + 1, path); goto done; }
        for (int j = 0; j < 4; j++) {
            const char *a = w[first + j];
            if (a[0] != '-' && !isdigit((unsigned char)a[0])) { setErr("Expecting number got %s in %s", a, path); goto done; }
            s->matrix[(int)order[i]][(int)order[j]] = atoi(a);
        }
    }
    if (*cur) { /* lineFileNext: the very next raw line, blank or not */
        char *line = cur, *nl = strchr(line, '\n');
        if (nl) *nl = 0;
        int gotO = 0, gotE = 0, o = 0, e = 0, n = 0;
        char *parts[32], *save = NULL;
        for (char *p = strtok_r(line, " =,\t", &save); p && n < 32; p = strtok_r(NULL, " =,\t", &save)) parts[n++] = p;
        for (int i = 0; i + 1 < n; i += 2) {
            if (strcmp(parts[i], "O") == 0) { gotO = 1; o = atoi(parts[i + 1]); }
            if (strcmp(parts[i], "E") == 0) { gotE = 1; e = atoi(parts[i + 1]); }
        }
        if (!gotO || !gotE) { setErr("Expecting O = and E = in last line of %s", path); goto done; }
        if (o <= 0 || e <= 0) { setErr("Must have positive gap scores"); goto done; }
    }
    rc = 0;
done:
    free(text);
    return rc;
}

orc_scoring *orc_scoring_new(const char *matrixFile, const char *linearGap)
{
    orc_scoring *s = calloc(1, sizeof *s);
    if (matrixFile && matrixFile[0]) {
        if (readMatrixFile(s, matrixFile)) { free(s); return NULL; }
    } else {
        /* axtScoreSchemeDefault, axt.c:423-458 (blastz default), rows = first index */
        static const char o[4] = { 'a', 'c', 'g', 't' };
        static const int def[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 },
                                       { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) s->matrix[(int)o[i]][(int)o[j]] = def[i][j];
    }
    spreadCase(s);
    int rc;
    if (strcmp(linearGap, "loose") == 0) rc = buildGapTables(s, looseSpec);        /* gapCalc.c:238-242 */
    else if (strcmp(linearGap, "medium") == 0) rc = buildGapTables(s, mediumSpec); /* gapCalc.c:243-247 */
    else {
        char *t = slurp(linearGap);
        if (!t) { setErr("cannot open gap file %s", linearGap); free(s); return NULL; }
        rc = buildGapTables(s, t);
        free(t);
    }
    if (rc) { orc_scoring_free(s); return NULL; }
    return s;
}

void orc_scoring_free(orc_scoring *s)
{
    if (!s) return;
    free(s->qSmall); free(s->tSmall); free(s->bSmall);
    free(s->longPos); free(s->qLong); free(s->tLong); free(s->bLong);
    free(s);
}

int orc_matrix_at(const orc_scoring *s, int q, int t) { return s->matrix[q & 255][t & 255]; }
int orc_gap_small_size(const orc_scoring *s) { return s->smallSize; }

static int oneSidedCost(const orc_scoring *s, int v, const int *small, const double *lng, double lastVal, double lastSlope)
{
    if (v < s->smallSize) return small[v];
    if (v >= s->lastPos) {
        double ext = lastSlope * (double)(v - s->lastPos);
        return truncToInt(lastVal + ext);
    }
    return interp(v, s->longPos, lng, s->longCount);
}

int orc_gap_cost(const orc_scoring *s, int dq, int dt)
/* gapCalcCost, gapCalc.c:298-331. */
{
    if (dt < 0) dt = 0;
    if (dq < 0) dq = 0;
    if (dt == 0) return oneSidedCost(s, dq, s->qSmall, s->qLong, s->qLastVal, s->qLastSlope);
    if (dq == 0) return oneSidedCost(s, dt, s->tSmall, s->tLong, s->tLastVal, s->tLastSlope);
    int both = (int)((unsigned)dq + (unsigned)dt);
    return oneSidedCost(s, both, s->bSmall, s->bLong, s->bLastVal, s->bLastSlope);
}

/* ====================================================================== scoring */

double orc_score_block(const orc_scoring *s, const char *q, const char *t, int size)
/* chainScoreBlock, chainConnect.c:14-22: sum of matrix[q[i]][t[i]] held in a double. */
{
    double total = 0;
    for (int i = 0; i < size; i++)
        total += s->matrix[(unsigned char)q[i]][(unsigned char)t[i]];
    return total;
}

void orc_find_crossover(const orc_scoring *s, const char *q, const char *t, int leftTEnd, int leftQEnd,
                        int rightTStart, int rightQStart, int overlap, int *pos, int *adjust)
/* chainConnect.c:80-104: start from "all of the overlap goes to the right block", move the switch point base by
 * base (the left block gains a base, the right one loses it) and keep the first best position. */
{
    const char *rq = q + rightQStart, *lq = q + leftQEnd - overlap;
    const char *rt = t + rightTStart, *lt = t + leftTEnd - overlap;
    double rScore = orc_score_block(s, rq, rt, overlap), lScore = orc_score_block(s, lq, lt, overlap);
    double score = rScore, best = rScore;
    int bestPos = 0;
    for (int i = 0; i < overlap; i++) {
        score += s->matrix[(unsigned char)lq[i]][(unsigned char)lt[i]];
        score -= s->matrix[(unsigned char)rq[i]][(unsigned char)rt[i]];
        if (score > best) { best = score; bestPos = i + 1; }
    }
    *pos = bestPos;
    *adjust = (int)(rScore + lScore - best);
}

int orc_score_jobs(const orc_scoring *s, orc_genome *tg, orc_genome *qg,
                   const orc_job *jobs, int64_t nJobs, int64_t totalJobBlocks,
                   const orc_block *blocks, int64_t nBlocks,
                   int64_t *global, int64_t *local, int64_t *aliBases)
/* Per job: clip like chainFastSubsetOnT (chain.c:510-522), then run the global recurrence of
 * chainCalcScore (chainConnect.c:30-38) and the local one of chainCalcScoreLocal
 * (scoreChain.c:181-195) side by side. */
{
    for (int64_t j = 0; j < nJobs; j++) {
        const orc_job *job = &jobs[j];
        int64_t nb = (j + 1 < nJobs ? jobs[j + 1].blockPtr : totalJobBlocks) - job->blockPtr;
        int tIx = (int)job->tSeq, qIx = (int)(job->qSeq & 0x7fffffffu);
        char strand = (job->qSeq >> 31) ? '-' : '+';
        if (tIx >= tg->count || qIx >= qg->count) { setErr("job %lld: sequence index out of range", (long long)j); return -1; }
        if (nb < 0 || (int64_t)job->firstBlock + nb > nBlocks) { setErr("job %lld: block range out of bounds", (long long)j); return -1; }
        const char *tDna = orc_genome_dna(tg, tIx, '+');
        const char *qDna = orc_genome_dna(qg, qIx, strand);
        int64_t tSize = tg->seq[tIx].size, qSize = qg->seq[qIx].size;
        double score = 0, run = 0, best = 0;
        int64_t ali = 0;
        int prevTe = 0, prevQe = 0;
        for (int64_t k = 0; k < nb; k++) {
            const orc_block *b = &blocks[job->firstBlock + k];
            int ts = b->tStart, te = b->tStart + b->size, qs = b->qStart, qe = b->qStart + b->size;
            if (ts < job->clipStart) { qs += job->clipStart - ts; ts = job->clipStart; }
            if (te > job->clipEnd) { qe -= te - job->clipEnd; te = job->clipEnd; }
            if (k > 0) {
                int cost = orc_gap_cost(s, qs - prevQe, ts - prevTe);
                score -= cost;
                run -= cost;
                if (run < 0) run = 0;
            }
            int n = te - ts;
            if (n > 0) {
                if (ts < 0 || qs < 0 || (int64_t)ts + n > tSize || (int64_t)qs + n > qSize) {
                    setErr("job %lld block %lld runs past its sequence", (long long)j, (long long)k);
                    return -1;
                }
                double bs = orc_score_block(s, qDna + qs, tDna + ts, n);
                score += bs;
                run += bs;
            }
            if (run > best) best = run;
            ali += n;
            prevTe = te;
            prevQe = qe;
        }
        global[j] = (int64_t)score;
        local[j] = (int64_t)best;
        aliBases[j] = ali;
    }
    return 0;
}

/* ====================================================================== .chain */

typedef struct {
    double score;
    char *tName, *qName;
    int tSize, tStart, tEnd, qSize, qStart, qEnd, id;
    char qStrand;
    int64_t firstBlock, nBlocks;
} orcChain;

struct orc_chainset {
    orcChain *chain;
    int64_t count, alloc;
    orc_block *block;
    int64_t blockCount, blockAlloc;
};

static int needNum(const char *w, int *out)
/* lineFileNeedNum, linefile.c:1015-1025 */
{
    if (w[0] != '-' && !isdigit((unsigned char)w[0])) { setErr("Expecting number, got %s", w); return -1; }
    *out = atoi(w);
    return 0;
}

orc_chainset *orc_chains_read(const char *path)
/* chainReadChainLine + chainReadBlocks, chain.c:256-346.  Lines starting with '#' and blank
 * lines are skipped (lineFileChop).  Chains without an id get 1,2,3,... (chain.c:189-198, 276-279). */
{
    char *text = slurp(path), *cur = text, *w[16];
    if (!text) { setErr("cannot open %s", path); return NULL; }
    orc_chainset *cs = calloc(1, sizeof *cs);
    int nextId = 1;
    for (;;) {
        char *line = nextChopLine(&cur);
        if (!line) break;
        int n = chopWords(line, w, 13);
        if (n < 12) { setErr("Expecting at least 12 words in chain line"); goto fail; }
        if (strcmp(w[0], "chain") != 0) { setErr("Expecting 'chain' got %s", w[0]); goto fail; }
        if (cs->count == cs->alloc) {
            cs->alloc = cs->alloc ? 2 * cs->alloc : 256;
            cs->chain = realloc(cs->chain, cs->alloc * sizeof(orcChain));
        }
        orcChain *c = &cs->chain[cs->count];
        memset(c, 0, sizeof *c);
        c->score = atof(w[1]);
        c->tName = strdup(w[2]);
        c->qName = strdup(w[7]);
        c->qStrand = w[9][0];
        if (needNum(w[3], &c->tSize) || needNum(w[5], &c->tStart) || needNum(w[6], &c->tEnd) ||
            needNum(w[8], &c->qSize) || needNum(w[10], &c->qStart) || needNum(w[11], &c->qEnd)) goto fail;
        if (n >= 13) { if (needNum(w[12], &c->id)) goto fail; }
        else c->id = nextId++;
        cs->count++;
        if (c->qStart >= c->qEnd || c->tStart >= c->tEnd) { setErr("End before start in chain %d", c->id); goto fail; }
        if (c->qStart < 0 || c->tStart < 0) { setErr("Start before zero in chain %d", c->id); goto fail; }
        if (c->qEnd > c->qSize || c->tEnd > c->tSize) { setErr("Past end of sequence in chain %d", c->id); goto fail; }
        c->firstBlock = cs->blockCount;
        int q = c->qStart, t = c->tStart;
        for (;;) {
            line = nextChopLine(&cur);
            if (!line) { setErr("chain %d ends early", c->id); goto fail; }
            n = chopWords(line, w, 3);
            int size, dt, dq;
            if (needNum(w[0], &size)) goto fail;
            if (cs->blockCount == cs->blockAlloc) {
                cs->blockAlloc = cs->blockAlloc ? 2 * cs->blockAlloc : 4096;
                cs->block = realloc(cs->block, cs->blockAlloc * sizeof(orc_block));
            }
            orc_block *b = &cs->block[cs->blockCount++];
            b->tStart = t; b->qStart = q; b->size = size;
            t += size; q += size;
            if (n == 1) break;
            if (n < 3) { setErr("Expecting 1 or 3 words in block line of chain %d", c->id); goto fail; }
            if (needNum(w[1], &dt) || needNum(w[2], &dq)) goto fail;
            t += dt; q += dq;
        }
        c->nBlocks = cs->blockCount - c->firstBlock;
        if (q != c->qEnd) { setErr("q end mismatch %d vs %d in chain %d", q, c->qEnd, c->id); goto fail; }
        if (t != c->tEnd) { setErr("t end mismatch %d vs %d in chain %d", t, c->tEnd, c->id); goto fail; }
    }
    free(text);
    return cs;
fail:
    free(text);
    orc_chains_free(cs);
    return NULL;
}

void orc_chains_free(orc_chainset *cs)
{
    if (!cs) return;
    for (int64_t i = 0; i < cs->count; i++) { free(cs->chain[i].tName); free(cs->chain[i].qName); }
    free(cs->chain); free(cs->block); free(cs);
}

int64_t orc_chains_count(const orc_chainset *cs) { return cs->count; }
int64_t orc_chains_total_blocks(const orc_chainset *cs) { return cs->blockCount; }
const orc_block *orc_chains_blocks(const orc_chainset *cs) { return cs->block; }

void orc_chains_header(const orc_chainset *cs, int64_t ix, double *score, const char **tName,
                       int *tSize, int *tStart, int *tEnd, const char **qName, int *qSize,
                       char *qStrand, int *qStart, int *qEnd, int *id,
                       int64_t *firstBlock, int64_t *nBlocks)
{
    const orcChain *c = &cs->chain[ix];
    *score = c->score; *tName = c->tName; *tSize = c->tSize; *tStart = c->tStart; *tEnd = c->tEnd;
    *qName = c->qName; *qSize = c->qSize; *qStrand = c->qStrand; *qStart = c->qStart; *qEnd = c->qEnd;
    *id = c->id; *firstBlock = c->firstBlock; *nBlocks = c->nBlocks;
}

int orc_chains_subset(const orc_chainset *cs, int64_t ix, int subStart, int subEnd,
                      int64_t *firstBlock, int64_t *nBlocks, int32_t *clipStart, int32_t *clipEnd)
/* chainSubsetOnT + chainFastSubsetOnT, chain.c:471-558:
 *  - range covering the chain -> the chain itself, unclipped (:501-506);
 *  - else first block with tEnd > subStart (:479-484), stop at the first with tStart >= subEnd (:510);
 *  - every kept block is clipped to [subStart, subEnd) with q moved by the same delta (:513-522). */
{
    const orcChain *c = &cs->chain[ix];
    if (subStart <= c->tStart && subEnd >= c->tEnd) {
        *firstBlock = c->firstBlock; *nBlocks = c->nBlocks;
        *clipStart = INT32_MIN; *clipEnd = INT32_MAX;
        return 1;
    }
    int64_t a = 0;
    while (a < c->nBlocks) {
        const orc_block *b = &cs->block[c->firstBlock + a];
        if (b->tStart + b->size > subStart) break;
        a++;
    }
    int64_t e = a;
    while (e < c->nBlocks && cs->block[c->firstBlock + e].tStart < subEnd) e++;
    *firstBlock = c->firstBlock + a; *nBlocks = e - a;
    *clipStart = subStart; *clipEnd = subEnd;
    return e > a;
}
