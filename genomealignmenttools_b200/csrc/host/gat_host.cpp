// gat_host.cpp -- see gat_host.hpp.  Compiled with -ffp-contract=off: the small gap tables are
// built with the same double arithmetic, in the same order, as kent/src/lib/gapCalc.c:82-104.
#include "gat_host.hpp"
#include <time.h>
#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <memory>
#include <queue>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <chrono>
#include <thread>
#include <tuple>
#include <unistd.h>

namespace gathost {

// ------------------------------------------------------------------ errAbort / verbose
static int g_verbose = 1;
void phaseDone(const char *what)
{
    static const bool on = getenv("GAT_TOOL_TIMING") != nullptr;
    static struct timespec last = {0, 0};
    if (!on) return;
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if (last.tv_sec || last.tv_nsec)
        fprintf(stderr, "[timing] %-28s %8.3f s\n", what, (double)(now.tv_sec - last.tv_sec) + 1e-9 * (double)(now.tv_nsec - last.tv_nsec));
    last = now;
}

void verboseSetLevel(int level) { g_verbose = level; }
int verboseLevel() { return g_verbose; }

void verbose(int level, const char *fmt, ...)
{
    if (level > g_verbose) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fflush(stderr);
}

void errAbort(const char *fmt, ...)
{   // message + newline on stderr, exit(-1): errAbort.c:182-222
    va_list ap;
    va_start(ap, fmt);
    fflush(stdout);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
    exit(-1);
}

void fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error{buf};
}

int runTool(int (*toolMain)(int, char **), int argc, char **argv)
{
    try {
        const int rc = toolMain(argc, argv);
        // like the kent tools nothing is torn down at the end: flush and leave (skips the CUDA context teardown)
        fflush(nullptr);
        _exit(rc);
    } catch (const Error &e) {
        errAbort("%s", e.message.c_str());
    }
}

// ------------------------------------------------------------------ options
void Options::init(int *argc, char **argv, const std::vector<OptionSpec> &specs)
{
    auto findSpec = [&](const std::string &name, OptionType &type) {
        for (const auto &s : specs)
            if (name == s.name) { type = s.type; return true; }
        if (name == "verbose") { type = OPTION_INT; return true; }      // commonOptions, options.c:23-26
        return false;
    };
    int out = 1, i = 1;
    for (; i < *argc; i++) {
        char *arg = argv[i];
        if (strcmp(arg, "--") == 0) { i++; break; }
        char *eq = strchr(arg, '=');
        bool isOption = (eq != nullptr) || arg[0] == '-';
        if (isOption && arg[0] == '-' && (arg[1] == 0 || isspace((unsigned char)arg[1]))) isOption = false;
        if (isOption && eq != nullptr)      // this=that counts only if everything before '=' is a word (options.c:158-170)
            for (char *s = arg; s < eq; ++s)
                if (*s != '_' && *s != '-' && !isalnum((unsigned char)*s)) { isOption = false; break; }
        if (!isOption) { argv[out++] = arg; continue; }
        const char *nameStart = arg[0] == '-' ? arg + 1 : arg;
        std::string name = eq ? std::string(nameStart, (size_t)(eq - nameStart)) : std::string(nameStart);
        const char *val = eq ? eq + 1 : nullptr;
        OptionType type;
        if (!findSpec(name, type)) fail("-%s is not a valid option", name.c_str());
        char *end = nullptr;
        switch (type) {
        case OPTION_BOOLEAN:
            if (val) fail("boolean option -%s must not have value", name.c_str());
            break;
        case OPTION_STRING:
            if (!val) fail("string option -%s must have a value", name.c_str());
            break;
        case OPTION_INT:
        case OPTION_LONG_LONG:
            if (!val) fail("int option -%s must have a value", name.c_str());
            (void)strtoll(val, &end, 10);
            if (*val == 0 || *end != 0)
                fail("value of -%s is not a valid %s: \"%s\"", name.c_str(), type == OPTION_INT ? "integer" : "long long", val);
            break;
        case OPTION_FLOAT:
        case OPTION_DOUBLE:
            if (!val) fail("%s option -%s must have a value", type == OPTION_FLOAT ? "float" : "double", name.c_str());
            (void)strtod(val, &end);
            if (*val == 0 || *end != 0)
                fail("value of -%s is not a valid %s: \"%s\"", name.c_str(), type == OPTION_FLOAT ? "float" : "double", val);
            break;
        }
        values[name] = val ? val : "on";
    }
    for (; i < *argc; i++) argv[out++] = argv[i];
    *argc = out;
    argv[out] = nullptr;
    if (exists("verbose")) verboseSetLevel(intVal("verbose", 0));
}

const char *Options::val(const char *name, const char *dflt) const
{
    auto it = values.find(name);
    return it == values.end() ? dflt : it->second.c_str();
}
int Options::intVal(const char *name, int dflt) const
{
    auto it = values.find(name);
    return it == values.end() ? dflt : atoi(it->second.c_str());
}
double Options::doubleVal(const char *name, double dflt) const
{
    auto it = values.find(name);
    return it == values.end() ? dflt : atof(it->second.c_str());
}

// ------------------------------------------------------------------ .2bit
static uint32_t rd32(const uint8_t *p, bool swap)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return swap ? __builtin_bswap32(v) : v;
}
static uint64_t rd64(const uint8_t *p, bool swap)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return swap ? __builtin_bswap64(v) : v;
}

bool TwoBitFile::isTwoBit(const std::string &path)
{   // twoBitIsFile, twoBit.c: endsWith(fileName, ".2bit")
    return path.size() >= 5 && path.compare(path.size() - 5, 5, ".2bit") == 0;
}

TwoBitFile::TwoBitFile(const std::string &path) : path_(path)
{
    struct stat pathStat;
    if (!isTwoBit(path) && stat(path.c_str(), &pathStat) == 0 && S_ISDIR(pathStat.st_mode)) {
        nibDir_ = true;             // a directory of .nib files; sequences are loaded when asked for
        return;
    }
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) fail("Can't open %s to read: %s", path.c_str(), strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); fail("Can't stat %s", path.c_str()); }
    imageSize_ = (size_t)st.st_size;
    void *m = imageSize_ ? mmap(nullptr, imageSize_, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
    if (m != MAP_FAILED) { image_ = static_cast<uint8_t *>(m); mapped_ = true; }
    else {
        image_ = static_cast<uint8_t *>(malloc(imageSize_ ? imageSize_ : 1));
        size_t got = 0;
        while (got < imageSize_) {
            ssize_t r = read(fd, image_ + got, imageSize_ - got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (got != imageSize_) { close(fd); fail("%s is truncated", path.c_str()); }
    }
    close(fd);
    if (imageSize_ < 16) fail("%s doesn't have a valid twoBitSig", path.c_str());
    uint32_t sig;
    memcpy(&sig, image_, 4);
    bool swap;
    if (sig == 0x1A412743u) swap = false;                 // twoBitSig, sig.h:58-62
    else if (sig == 0x4327411Au) swap = true;
    else fail("%s doesn't have a valid twoBitSig", path.c_str());
    const uint32_t version = rd32(image_ + 4, swap), count = rd32(image_ + 8, swap);
    if (version > 1) fail("Can only handle version 0 or version 1 of this file. This is version %d", (int)version);
    size_t pos = 16;
    std::vector<uint64_t> offsets(count);
    seqs_.resize(count);
    for (uint32_t i = 0; i < count; i++) {
        if (pos + 1 > imageSize_) fail("%s is truncated", path.c_str());
        const size_t len = image_[pos++];
        if (pos + len + (version ? 8 : 4) > imageSize_) fail("%s is truncated", path.c_str());
        seqs_[i].name.assign(reinterpret_cast<const char *>(image_ + pos), len);
        pos += len;
        if (version == 1) { offsets[i] = rd64(image_ + pos, swap); pos += 8; }
        else { offsets[i] = rd32(image_ + pos, swap); pos += 4; }
        index_[seqs_[i].name] = (int)i;
    }
    for (uint32_t i = 0; i < count; i++) {
        TwoBitSeq &s = seqs_[i];
        size_t p = (size_t)offsets[i];
        auto need = [&](size_t n) { if (p + n > imageSize_) fail("%s is truncated", path.c_str()); };
        need(8);
        s.size = rd32(image_ + p, swap); p += 4;
        const uint32_t nc = rd32(image_ + p, swap); p += 4;
        need((size_t)nc * 8 + 4);
        s.nStart.resize(nc); s.nLen.resize(nc);
        for (uint32_t k = 0; k < nc; k++) s.nStart[k] = rd32(image_ + p + 4 * (size_t)k, swap);
        p += 4 * (size_t)nc;
        for (uint32_t k = 0; k < nc; k++) s.nLen[k] = rd32(image_ + p + 4 * (size_t)k, swap);
        p += 4 * (size_t)nc;
        const uint32_t mc = rd32(image_ + p, swap); p += 4;
        need((size_t)mc * 8 + 4);
        s.maskStart.resize(mc); s.maskLen.resize(mc);
        for (uint32_t k = 0; k < mc; k++) s.maskStart[k] = rd32(image_ + p + 4 * (size_t)k, swap);
        p += 4 * (size_t)mc;
        for (uint32_t k = 0; k < mc; k++) s.maskLen[k] = rd32(image_ + p + 4 * (size_t)k, swap);
        p += 4 * (size_t)mc;
        p += 4;     // reserved
        need(((size_t)s.size + 3) / 4);
        s.packed = image_ + p;
    }
}

TwoBitFile::~TwoBitFile()
{
    if (mapped_) munmap(image_, imageSize_);
    else free(image_);
}

int TwoBitFile::find(const std::string &name) const
{
    auto it = index_.find(name);
    if (it != index_.end()) return it->second;
    return nibDir_ ? loadNib(name) : -1;
}

int TwoBitFile::loadNib(const std::string &name) const
{   // nibOpenVerify + nibInput, kent/src/lib/nib.c:83-235
    const std::string fileName = path_ + "/" + name + ".nib";
    FILE *f = fopen(fileName.c_str(), "rb");
    if (!f) fail("Can't open %s to read: %s", fileName.c_str(), strerror(errno));
    uint32_t head[2];
    if (fread(head, 4, 2, f) != 2) { fclose(f); fail("%s is not a good .nib file.", fileName.c_str()); }
    uint32_t sig = head[0], size = head[1];
    if (sig != 0x6BE93D3Au) {       // nibSig, sig.h:49; byte-swapped files are accepted
        sig = __builtin_bswap32(sig); size = __builtin_bswap32(size);
        if (sig != 0x6BE93D3Au) { fclose(f); fail("%s is not a good .nib file.", fileName.c_str()); }
    }
    std::vector<uint8_t> nib(((size_t)size + 1) / 2);
    if (fread(nib.data(), 1, nib.size(), f) != nib.size()) { fclose(f); fail("Read error 2 in %s", fileName.c_str()); }
    fclose(f);
    TwoBitSeq s;
    s.name = name;
    s.size = size;
    std::vector<uint8_t> packed(((size_t)size + 3) / 4, 0);
    bool inN = false;
    for (uint32_t i = 0; i < size; i++) {
        const unsigned v = (i & 1) ? (nib[i >> 1] & 0xf) : (nib[i >> 1] >> 4);      // first base in the high nibble
        const unsigned code = v & 7;                // bit 3 is the soft-mask flag: irrelevant to scores (axt.c:402-421)
        const bool isN = code >= 4;                 // N (and the undefined codes 5..7) have all-zero matrix rows
        if (!isN) packed[i >> 2] |= (uint8_t)(code << (6 - 2 * (i & 3)));
        if (isN && !inN) { s.nStart.push_back(i); s.nLen.push_back(0); }
        if (isN) s.nLen.back()++;
        inN = isN;
    }
    nibPayload_.push_back(std::move(packed));
    s.packed = nibPayload_.back().data();
    seqs_.push_back(std::move(s));
    index_[name] = (int)seqs_.size() - 1;
    return (int)seqs_.size() - 1;
}

void uploadGenome(gat_ctx *ctx, int side, const TwoBitFile &tb, const std::vector<int> &use)
{
    std::vector<uint64_t> offs(use.size());
    std::vector<uint32_t> sizes(use.size());
    std::vector<gat_nrun> runs;
    uint64_t total = 0;
    for (size_t i = 0; i < use.size(); i++) {
        const TwoBitSeq &s = tb.seqs()[use[i]];
        offs[i] = total;
        sizes[i] = s.size;
        total += ((uint64_t)s.size + 3) / 4;
        for (size_t k = 0; k < s.nStart.size(); k++) runs.push_back(gat_nrun{(uint32_t)i, s.nStart[k], s.nLen[k]});
    }
    // the C ABI wants one buffer: gather the used payloads, a slice per host thread.  Plain memory: pinning
    // hundreds of MB for one copy costs more than the staged pageable copy it would save.
    std::unique_ptr<uint8_t[]> staging(new uint8_t[total ? total : 1]);
    {
        unsigned T = std::thread::hardware_concurrency();
        T = std::max(1u, std::min(T ? T : 1u, std::min(16u, (unsigned)(total >> 24) + 1u)));
        auto body = [&](unsigned k) {       // bytes [lo, hi) of the gathered buffer
            const uint64_t lo = total / T * k, hi = k + 1 == T ? total : total / T * (k + 1);
            for (size_t i = 0; i < use.size(); i++) {
                const uint64_t b = offs[i], e = offs[i] + ((uint64_t)sizes[i] + 3) / 4;
                const uint64_t from = std::max(b, lo), to = std::min(e, hi);
                if (from < to) memcpy(staging.get() + from, tb.seqs()[use[i]].packed + (from - b), (size_t)(to - from));
            }
        };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < T; k++) th.emplace_back(body, k);
        body(0);
        for (auto &t : th) t.join();
    }
    const int rc = gat_load_genome(ctx, side, staging.get(), total, offs.data(), sizes.data(), (uint32_t)use.size(), runs.data(), runs.size());
    if (rc != GAT_OK) fail("%s", gat_last_error());
}

// ------------------------------------------------------------------ text helpers
static std::string slurp(const std::string &path)
{
    const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
    std::string text;
    char buf[1 << 16];
    if (gz) {
        // linefile.c:40-53: compressed files are read through a decompressor child.  Spawned without a shell (a file name is
        // data, not a command line) and its exit status is checked: a corrupt or truncated .gz is an error, not a short input.
        if (access(path.c_str(), R_OK) != 0) fail("Couldn't open %s , %s", path.c_str(), strerror(errno));
        int fds[2];
        if (pipe(fds) != 0) fail("pipe failed: %s", strerror(errno));
        const pid_t child = fork();
        if (child < 0) fail("fork failed: %s", strerror(errno));
        if (child == 0) {
            dup2(fds[1], STDOUT_FILENO);
            close(fds[0]); close(fds[1]);
            execlp("gzip", "gzip", "-dc", "--", path.c_str(), (char *)nullptr);
            _exit(127);
        }
        close(fds[1]);
        ssize_t n;
        while ((n = read(fds[0], buf, sizeof buf)) > 0 || (n < 0 && errno == EINTR))
            if (n > 0) text.append(buf, (size_t)n);
        close(fds[0]);
        int status = 0;
        while (waitpid(child, &status, 0) < 0 && errno == EINTR) {}
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0)
            fail("gzip -dc %s failed (%s %d)", path.c_str(), WIFEXITED(status) ? "exit status" : "signal",
                 WIFEXITED(status) ? WEXITSTATUS(status) : WTERMSIG(status));
        return text;
    }
    FILE *f = (path == "stdin") ? stdin : fopen(path.c_str(), "rb");
    if (!f) fail("Couldn't open %s , %s", path.c_str(), strerror(errno));
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    if (f != stdin) fclose(f);
    return text;
}

// splits text into lines in place (replaces '\n' by 0); no copy
struct LineCursor {
    std::string &text;
    size_t pos = 0;
    int lineIx = 0;
    explicit LineCursor(std::string &t) : text(t) {}
    char *next()
    {
        if (pos >= text.size()) return nullptr;
        char *line = &text[pos];
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) pos = text.size();
        else { text[nl] = 0; pos = nl + 1; }
        lineIx++;
        return line;
    }
};

static int chop(char *line, char **words, int maxWords)
{   // chopByWhite
    int n = 0;
    char *s = line;
    for (;;) {
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s || n == maxWords) break;
        words[n++] = s;
        while (*s && !isspace((unsigned char)*s)) s++;
        if (!*s) break;
        *s++ = 0;
    }
    return n;
}

// ------------------------------------------------------------------ score scheme
ScoreScheme ScoreScheme::defaultScheme()
{
    static const int acgt[4][4] = {{91, -114, -31, -123}, {-114, 100, -125, -31}, {-31, -125, 100, -114}, {-123, -31, -114, 91}};
    static const int code[4] = {2, 1, 3, 0};    // A C G T (file order) -> kent codes
    ScoreScheme ss;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) ss.matrix[code[i]][code[j]] = acgt[i][j];
    return ss;
}

ScoreScheme ScoreScheme::read(const std::string &path)
{
    static const int code[4] = {2, 1, 3, 0};
    std::string text = slurp(path);
    LineCursor lc(text);
    char *w[6];
    auto chopNext = [&](int &n) -> bool {      // lineFileChopNext, linefile.c:907-922
        for (char *line; (line = lc.next()) != nullptr;) {
            if (line[0] == '#') continue;
            n = chop(line, w, 6);
            if (n) return true;
        }
        return false;
    };
    int n;
    for (;;) {
        if (!chopNext(n)) fail("Scoring matrix file %s too short\n", path.c_str());
        if (strchr(w[0], '=') || (n > 1 && strchr(w[1], '='))) continue;        // lastz settings lines
        if (n < 4 || w[0][0] != 'A' || w[1][0] != 'C' || w[2][0] != 'G' || w[3][0] != 'T')
            fail("%s doesn't seem to be a score matrix file", path.c_str());
        break;
    }
    ScoreScheme ss;
    for (int i = 0; i < 4; i++) {
        if (!chopNext(n)) fail("Scoring matrix file %s too short\n", path.c_str());
        const int first = (n == 5) ? 1 : 0;
        if (n < first + 4) fail("Expecting 4 numbers line %d of %s", lc.lineIx, path.c_str());
        for (int j = 0; j < 4; j++) {
            const char *a = w[first + j];
            if (a[0] != '-' && !isdigit((unsigned char)a[0]))
                fail("Expecting number field %d line %d of %s, got %s", first + j + 1, lc.lineIx, path.c_str(), a);
            ss.matrix[code[i]][code[j]] = atoi(a);
        }
    }
    if (char *line = lc.next()) {       // the very next raw line must carry O= and E= (axt.c:785-805)
        bool gotO = false, gotE = false;
        std::vector<char *> parts;
        for (char *p = strtok(line, " =,\t"); p; p = strtok(nullptr, " =,\t")) parts.push_back(p);
        for (size_t i = 0; i + 1 < parts.size(); i += 2) {
            if (strcmp(parts[i], "O") == 0) { gotO = true; ss.gapOpen = atoi(parts[i + 1]); }
            if (strcmp(parts[i], "E") == 0) { gotE = true; ss.gapExtend = atoi(parts[i + 1]); }
        }
        if (!gotO || !gotE) fail("Expecting O = and E = in last line of %s", path.c_str());
        if (ss.gapOpen <= 0 || ss.gapExtend <= 0) fail("Must have positive gap scores");
    }
    return ss;
}

// ------------------------------------------------------------------ gap costs
static const char *looseCosts =
    "tablesize       11\n"
    "smallSize       111\n"
    "position        1       2       3       11      111     2111    12111   32111   72111   152111  252111\n"
    "qGap    325     360     400     450     600     1100    3600    7600    15600   31600   56600\n"
    "tGap    325     360     400     450     600     1100    3600    7600    15600   31600   56600\n"
    "bothGap 625     660     700     750     900     1400    4000    8000    16000   32000   57000\n";
static const char *mediumCosts =
    "tableSize 11\n"
    "smallSize 111\n"
    "position 1 2 3 11 111 2111 12111 32111 72111 152111 252111\n"
    "qGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900\n"
    "tGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900\n"
    "bothGap 750 825 850 1000 1300 3300 23300 58300 118300 218300 318300\n";

const char *GapCalc::sampleFileContents() { return looseCosts; }

static int truncToInt(double d)
{   // (int)double on x86-64: toward zero, INT_MIN when out of range
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
    return (int)d;
}

static int interpolate(int x, const std::vector<int> &s, const std::vector<double> &v)
{   // gapCalc.c:82-104
    const int n = (int)s.size();
    for (int i = 0; i < n; i++) {
        if (x == s[i]) return truncToInt(v[i]);
        if (x < s[i]) {
            const int ds = s[i] - s[i - 1];
            const double dv = v[i] - v[i - 1];
            const double prod = dv * (double)(x - s[i - 1]);
            return truncToInt(v[i - 1] + prod / (double)ds);
        }
    }
    const int ds = s[n - 1] - s[n - 2];
    const double dv = v[n - 1] - v[n - 2];
    const double prod = dv * (double)(x - s[n - 2]);
    return truncToInt(v[n - 2] + prod / (double)ds);
}

GapCalc GapCalc::fromString(const std::string &spec)
{   // gapCalcRead, gapCalc.c:146-222
    std::string text = spec;
    LineCursor lc(text);
    auto tagged = [&](const char *tag, int count, std::vector<int> *iOut, std::vector<double> *dOut) {
        char *line;
        for (;;) {      // lineFileNextReal: skip blank lines and lines whose first non-space is '#'
            line = lc.next();
            if (!line) fail("Unexpected end of file in gap cost specification");
            char *s = line;
            while (*s && isspace((unsigned char)*s)) s++;
            if (*s && *s != '#') break;
        }
        std::vector<char *> words(count + 3);
        const int n = chop(line, words.data(), count + 2);
        if (n == 0 || strcasecmp(words[0], tag) != 0) fail("Expecting %s got %s line %d of gap costs", tag, n ? words[0] : "", lc.lineIx);
        if (n - 1 < count) fail("Not enough numbers line %d of gap costs", lc.lineIx);
        if (n - 1 > count) fail("Too many numbers line %d of gap costs", lc.lineIx);
        for (int i = 0; i < count; i++) {
            if (!isdigit((unsigned char)words[i + 1][0])) fail("Expecting number got %s line %d of gap costs", words[i + 1], lc.lineIx);
            if (iOut) iOut->push_back(atoi(words[i + 1]));
            if (dOut) dOut->push_back(atof(words[i + 1]));
        }
    };
    GapCalc g;
    std::vector<int> one, pos;
    std::vector<double> qv, tv, bv;
    tagged("tableSize", 1, &one, nullptr);
    const int tableSize = one[0];
    one.clear();
    tagged("smallSize", 1, &one, nullptr);
    g.smallSize = one[0];
    if (tableSize < 2 || g.smallSize < 1) fail("bad tableSize/smallSize in gap costs");
    tagged("position", tableSize, &pos, nullptr);
    tagged("qGap", tableSize, nullptr, &qv);
    tagged("tGap", tableSize, nullptr, &tv);
    tagged("bothGap", tableSize, nullptr, &bv);
    if (pos[0] > 1) fail("gap cost positions must start at 1");
    g.qSmall.assign(g.smallSize, 0); g.tSmall.assign(g.smallSize, 0); g.bSmall.assign(g.smallSize, 0);
    for (int i = 1; i < g.smallSize; i++) {
        g.qSmall[i] = interpolate(i, pos, qv);
        g.tSmall[i] = interpolate(i, pos, tv);
        g.bSmall[i] = interpolate(i, pos, bv);
    }
    int startLong = -1;
    for (int i = 0; i < tableSize; i++)
        if (pos[i] == g.smallSize) { startLong = i; break; }
    if (startLong < 0) fail("No position %d in gapCalcRead()\n", g.smallSize);
    g.longPos.assign(pos.begin() + startLong, pos.end());
    g.qLong.assign(qv.begin() + startLong, qv.end());
    g.tLong.assign(tv.begin() + startLong, tv.end());
    g.bLong.assign(bv.begin() + startLong, bv.end());
    if (g.longPos.size() < 2) fail("gap costs need at least two positions from smallSize on");
    return g;
}

GapCalc GapCalc::fromFile(const char *name)
{
    if (name == nullptr) fail("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults\n");
    if (strcmp(name, "loose") == 0) { verbose(2, "using loose linear gap costs (chicken/human)\n"); return fromString(looseCosts); }
    if (strcmp(name, "medium") == 0) { verbose(2, "using medium (original) linear gap costs (mouse/human)\n"); return fromString(mediumCosts); }
    return fromString(slurp(name));
}

int GapCalc::cost(int dq, int dt) const
{   // gapCalcCost, gapCalc.c:298-331
    if (dt < 0) dt = 0;
    if (dq < 0) dq = 0;
    const std::vector<int32_t> *small;
    const std::vector<double> *lng;
    int v;
    if (dt == 0) { small = &qSmall; lng = &qLong; v = dq; }
    else if (dq == 0) { small = &tSmall; lng = &tLong; v = dt; }
    else { small = &bSmall; lng = &bLong; v = (int)((unsigned)dq + (unsigned)dt); }
    if (v < smallSize) return (*small)[v];
    const int L = (int)longPos.size(), last = longPos[L - 1];
    if (v >= last) {
        const double slope = ((*lng)[L - 1] - (*lng)[L - 2]) / ((double)last - (double)longPos[L - 2]);
        const double ext = slope * (double)(v - last);
        return truncToInt((*lng)[L - 1] + ext);
    }
    return interpolate(v, std::vector<int>(longPos.begin(), longPos.end()), *lng);
}

void setScoring(gat_ctx *ctx, const ScoreScheme &ss, const GapCalc &gc)
{
    gat_scoring s;
    memcpy(s.matrix, ss.matrix, sizeof s.matrix);
    s.smallSize = gc.smallSize;
    s.qSmall = gc.qSmall.data(); s.tSmall = gc.tSmall.data(); s.bSmall = gc.bSmall.data();
    s.longCount = (int)gc.longPos.size();
    s.longPos = gc.longPos.data();
    s.qLong = gc.qLong.data(); s.tLong = gc.tLong.data(); s.bLong = gc.bLong.data();
    if (gat_set_scoring(ctx, &s) != GAT_OK) fail("%s", gat_last_error());
}

// ------------------------------------------------------------------ chains
static inline bool parseInt(const char *w, int &out)
{   // lineFileNeedNum: first char '-' or digit, then atoi
    if (w[0] != '-' && !(w[0] >= '0' && w[0] <= '9')) return false;
    out = atoi(w);
    return true;
}

// chainRead (chain.c:256-346) over text[begin, end), which starts at a chain header (or at file start).
// Chains without an id column get id INT_MIN here; readChains numbers them in file order afterwards.
static void parseChainText(std::string &text, size_t begin, size_t end, int lineBase, bool moreFollows, const char *fn, ChainSet &out)
{
    struct Cursor {     // splits [pos, end) into lines in place (replaces '\n' by 0); no copy
        std::string &text; size_t pos, end; int lineIx;
        char *next()
        {
            if (pos >= end) return nullptr;
            char *line = &text[pos];
            const void *nl = memchr(line, '\n', end - pos);
            if (!nl) { pos = end; if (end < text.size()) text[end] = 0; }
            else { const size_t at = (size_t)((const char *)nl - text.data()); text[at] = 0; pos = at + 1; }
            lineIx++;
            return line;
        }
    } lc{text, begin, end, lineBase};
    char *w[16];
    auto chopNext = [&](int maxWords) -> int {
        for (char *line; (line = lc.next()) != nullptr;) {
            if (line[0] == '#') { out.metaLines.emplace_back(line); out.metaLineChain.push_back(out.chains.size()); continue; }
            const int n = chop(line, w, maxWords);
            if (n) return n;
        }
        return 0;
    };
    for (;;) {
        int n = chopNext(13);
        if (n == 0) break;
        if (n < 12) fail("Expecting at least 12 words line %d of %s", lc.lineIx, fn);
        if (strcmp(w[0], "chain") != 0) fail("Expecting 'chain' line %d of %s", lc.lineIx, fn);
        ChainHead c;
        c.score = atof(w[1]);
        c.tName = w[2];
        c.qName = w[7];
        c.qStrand = w[9][0];
        auto need = [&](int ix, int &v) {
            if (!parseInt(w[ix], v)) fail("Expecting number field %d line %d of %s, got %s", ix + 1, lc.lineIx, fn, w[ix]);
        };
        need(3, c.tSize);
        if (n >= 13) need(12, c.id);
        else c.id = INT_MIN;
        need(5, c.tStart); need(6, c.tEnd); need(8, c.qSize); need(10, c.qStart); need(11, c.qEnd);
        if (c.qStart >= c.qEnd || c.tStart >= c.tEnd) fail("End before start line %d of %s", lc.lineIx, fn);
        if (c.qStart < 0 || c.tStart < 0) fail("Start before zero line %d of %s", lc.lineIx, fn);
        if (c.qEnd > c.qSize || c.tEnd > c.tSize) fail("Past end of sequence line %d of %s", lc.lineIx, fn);
        c.firstBlock = out.blocks.size();
        int q = c.qStart, t = c.tStart;
        for (;;) {
            n = chopNext(3);
            if (n == 0) {
                // a piece ends inside a chain only if the next piece's header interrupts it: say what a sequential read says
                if (moreFollows) fail("Expecting number field 1 line %d of %s, got %s", lc.lineIx + 1, fn, "chain");
                fail("Unexpected end of file in %s", fn);
            }
            int size, dt, dq;
            if (!parseInt(w[0], size)) fail("Expecting number field 1 line %d of %s, got %s", lc.lineIx, fn, w[0]);
            out.blocks.push_back(gat_block{t, q, (uint32_t)size});
            t += size; q += size;
            if (n == 1) break;
            if (n < 3) fail("Expecting 1 or 3 words line %d of %s\n", lc.lineIx, fn);
            if (!parseInt(w[1], dt)) fail("Expecting number field 2 line %d of %s, got %s", lc.lineIx, fn, w[1]);
            if (!parseInt(w[2], dq)) fail("Expecting number field 3 line %d of %s, got %s", lc.lineIx, fn, w[2]);
            t += dt; q += dq;
        }
        c.nBlocks = out.blocks.size() - c.firstBlock;
        if (q != c.qEnd) fail("q end mismatch %d vs %d line %d of %s\n", q, c.qEnd, lc.lineIx, fn);
        if (t != c.tEnd) fail("t end mismatch %d vs %d line %d of %s\n", t, c.tEnd, lc.lineIx, fn);
        out.chains.push_back(std::move(c));
    }
}

static unsigned hostThreads(size_t work, size_t perThread)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    if (hw > 32) hw = 32;
    const size_t want = work / perThread;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(hw, want));
}

void readChains(const std::string &path, ChainSet &out)
{   // The file is cut at chain headers ("\nchain " can only be one: block lines are numbers, meta lines start
    // with '#') into one piece per host thread; the pieces are parsed concurrently and joined in file order,
    // so chains, blocks, meta lines, sequential ids and the first error are those of a sequential read.
    std::string text = slurp(path);
    const char *fn = path.c_str();
    // GAT_PARSE_PIECE_BYTES: smallest piece worth a thread (tests set it to a few bytes to cut small files)
    const char *pieceEnv = getenv("GAT_PARSE_PIECE_BYTES");
    const unsigned T = hostThreads(text.size(), pieceEnv && atol(pieceEnv) > 0 ? (size_t)atol(pieceEnv) : (size_t)4 << 20);
    std::vector<size_t> cut(T + 1, text.size());
    cut[0] = 0;
    for (unsigned i = 1; i < T; i++) {
        const size_t from = std::max(cut[i - 1], text.size() / T * i);
        const size_t at = text.find("\nchain ", from);
        cut[i] = at == std::string::npos ? text.size() : at + 1;
    }
    std::vector<int> lineBase(T + 1, 0);
    std::vector<ChainSet> part(T);
    std::vector<std::string> error(T);
    std::vector<char> failed(T, 0);
    auto run = [&](auto &&body) {
        std::vector<std::thread> th;
        for (unsigned i = 1; i < T; i++) th.emplace_back(body, i);
        body(0u);
        for (auto &t : th) t.join();
    };
    if (T > 1) {
        std::vector<int> lines(T, 0);
        run([&](unsigned i) { lines[i] = (int)std::count(text.begin() + (long)cut[i], text.begin() + (long)cut[i + 1], '\n'); });
        for (unsigned i = 0; i < T; i++) lineBase[i + 1] = lineBase[i] + lines[i];
    }
    run([&](unsigned i) {
        try { parseChainText(text, cut[i], cut[i + 1], lineBase[i], cut[i + 1] < text.size(), fn, i == 0 ? out : part[i]); }
        catch (const Error &e) { failed[i] = 1; error[i] = e.message; }
    });
    for (unsigned i = 0; i < T; i++)
        if (failed[i]) fail("%s", error[i].c_str());
    size_t nChains = out.chains.size(), nBlocks = out.blocks.size();
    std::vector<size_t> chainAt(T, 0), blockAt(T, 0);
    for (unsigned i = 1; i < T; i++) { chainAt[i] = nChains; blockAt[i] = nBlocks; nChains += part[i].chains.size(); nBlocks += part[i].blocks.size(); }
    out.chains.resize(nChains);
    out.blocks.resize(nBlocks);
    run([&](unsigned i) {
        if (i == 0) return;
        std::copy(part[i].blocks.begin(), part[i].blocks.end(), out.blocks.begin() + (long)blockAt[i]);
        for (size_t c = 0; c < part[i].chains.size(); c++) {
            part[i].chains[c].firstBlock += blockAt[i];
            out.chains[chainAt[i] + c] = std::move(part[i].chains[c]);
        }
    });
    for (unsigned i = 1; i < T; i++)
        for (size_t m = 0; m < part[i].metaLines.size(); m++) {
            out.metaLines.push_back(std::move(part[i].metaLines[m]));
            out.metaLineChain.push_back(part[i].metaLineChain[m] + chainAt[i]);
        }
    int nextId = 1;                             // chain.c:276-279
    for (ChainHead &c : out.chains)
        if (c.id == INT_MIN) c.id = nextId++;
}

static inline char *putInt(char *p, int v)
{   // "%d"
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

void writeChain(FILE *f, const ChainHead &c, const gat_block *blocks)
{   // chainWriteHead + chainWrite, chain.c:200-227.  Block lines are formatted by hand into a buffer:
    // after the GPU has scored a file, fprintf per block would be the longest phase of scoreChain.
    fprintf(f, "chain %1.0f %s %d + %d %d %s %d %c %d %d %d\n", c.score, c.tName.c_str(), c.tSize, c.tStart, c.tEnd,
            c.qName.c_str(), c.qSize, c.qStrand, c.qStart, c.qEnd, c.id);
    const gat_block *b = blocks + c.firstBlock;
    char buf[1 << 16];
    char *p = buf;
    for (uint64_t i = 0; i < c.nBlocks; i++) {
        if (p > buf + sizeof buf - 64) { fwrite(buf, 1, (size_t)(p - buf), f); p = buf; }
        p = putInt(p, (int)b[i].size);
        if (i + 1 < c.nBlocks) {
            *p++ = '\t';
            p = putInt(p, b[i + 1].tStart - (b[i].tStart + (int)b[i].size));
            *p++ = '\t';
            p = putInt(p, b[i + 1].qStart - (b[i].qStart + (int)b[i].size));
        }
        *p++ = '\n';
    }
    *p++ = '\n';
    fwrite(buf, 1, (size_t)(p - buf), f);
}

static void formatChain(std::string &o, const ChainHead &c, const gat_block *blocks)
{
    char head[512];
    const int hn = snprintf(head, sizeof head, "chain %1.0f %s %d + %d %d %s %d %c %d %d %d\n", c.score, c.tName.c_str(), c.tSize,
                            c.tStart, c.tEnd, c.qName.c_str(), c.qSize, c.qStrand, c.qStart, c.qEnd, c.id);
    if (hn < 0 || (size_t)hn >= sizeof head) {      // very long sequence names: let the C library size it
        std::string big((size_t)hn + 1, 0);
        snprintf(&big[0], big.size(), "chain %1.0f %s %d + %d %d %s %d %c %d %d %d\n", c.score, c.tName.c_str(), c.tSize,
                 c.tStart, c.tEnd, c.qName.c_str(), c.qSize, c.qStrand, c.qStart, c.qEnd, c.id);
        o.append(big.c_str());
    } else o.append(head, (size_t)hn);
    const gat_block *b = blocks + c.firstBlock;
    char buf[4096];
    char *p = buf;
    for (uint64_t i = 0; i < c.nBlocks; i++) {
        if (p > buf + sizeof buf - 64) { o.append(buf, (size_t)(p - buf)); p = buf; }
        p = putInt(p, (int)b[i].size);
        if (i + 1 < c.nBlocks) {
            *p++ = '\t';
            p = putInt(p, b[i + 1].tStart - (b[i].tStart + (int)b[i].size));
            *p++ = '\t';
            p = putInt(p, b[i + 1].qStart - (b[i].qStart + (int)b[i].size));
        }
        *p++ = '\n';
    }
    *p++ = '\n';
    o.append(buf, (size_t)(p - buf));
}

void writeChains(FILE *f, const std::vector<ChainHead> &chains, const gat_block *blocks)
{   // chainWrite for a whole set: pieces of about equal size are formatted on the host threads, written in order
    uint64_t total = 0;
    for (const ChainHead &c : chains) total += c.nBlocks + 4;
    const unsigned T = hostThreads((size_t)total, (size_t)1 << 18);
    std::vector<size_t> cut(T + 1, chains.size());
    cut[0] = 0;
    {
        uint64_t acc = 0; unsigned k = 1;
        for (size_t c = 0; c < chains.size() && k < T; c++) {
            acc += chains[c].nBlocks + 4;
            if (acc >= total / T * k) cut[k++] = c + 1;
        }
    }
    std::vector<std::string> piece(T);
    std::vector<std::thread> th;
    auto body = [&](unsigned i) {
        for (size_t c = cut[i]; c < cut[i + 1]; c++) formatChain(piece[i], chains[c], blocks);
    };
    for (unsigned i = 1; i < T; i++) th.emplace_back(body, i);
    body(0);
    for (auto &t : th) t.join();
    for (unsigned i = 0; i < T; i++) fwrite(piece[i].data(), 1, piece[i].size(), f);
}

// ------------------------------------------------------------------ work-list
void buildRecords(const ChainSet &cs, WorkList &wl)
{
    wl.blocks.clear();
    wl.blocks.reserve(cs.blocks.size());
    wl.chainFirstRecord.assign(cs.chains.size() + 1, 0);
    wl.blockFirstRecord.clear();
    const uint32_t maxPiece = GAT_SPLIT_BASES;
    bool anySplit = false;
    for (const gat_block &b : cs.blocks)
        if (b.size > maxPiece) { anySplit = true; break; }
    if (anySplit) wl.blockFirstRecord.assign(cs.blocks.size() + 1, 0);
    for (size_t c = 0; c < cs.chains.size(); c++) {
        wl.chainFirstRecord[c] = wl.blocks.size();
        const ChainHead &h = cs.chains[c];
        for (uint64_t i = 0; i < h.nBlocks; i++) {
            const gat_block &b = cs.blocks[h.firstBlock + i];
            if (anySplit) wl.blockFirstRecord[h.firstBlock + i] = wl.blocks.size();
            if (b.size <= maxPiece) { wl.blocks.push_back(b); continue; }
            // a long gapless block is cut into JOINED pieces (they score exactly like the block) so that its
            // bases spread over many warps
            for (uint32_t off = 0; off < b.size; off += maxPiece) {
                const uint32_t n = std::min(maxPiece, b.size - off);
                wl.blocks.push_back(gat_block{b.tStart + (int)off, b.qStart + (int)off, n | (off ? GAT_BLOCK_JOINED : 0u)});
            }
        }
    }
    wl.chainFirstRecord[cs.chains.size()] = wl.blocks.size();
    if (anySplit) wl.blockFirstRecord[cs.blocks.size()] = wl.blocks.size();
    if (wl.blocks.size() > 0xffffffffull) fail("more than 2^32 block records in one batch");
}

static void pushJob(WorkList &wl, uint32_t tSeq, uint32_t qSeq, char strand, uint64_t firstRecord, uint64_t nRecords,
                    int clipStart, int clipEnd, int64_t ali)
{
    if (wl.totalJobBlocks + nRecords > 0xffffffffull) fail("more than 2^32 job-blocks in one batch");
    gat_job j;
    j.tSeq = tSeq;
    j.qSeq = qSeq | (strand == '-' ? GAT_QSEQ_MINUS : 0u);
    j.firstBlock = (uint32_t)firstRecord;
    j.blockPtr = (uint32_t)wl.totalJobBlocks;
    j.clipStart = clipStart;
    j.clipEnd = clipEnd;
    wl.jobs.push_back(j);
    wl.aliBases.push_back(ali);
    wl.totalJobBlocks += nRecords;
}

void addChainJob(const ChainSet &cs, size_t c, uint32_t tSeq, uint32_t qSeq, WorkList &wl)
{
    const ChainHead &h = cs.chains[c];
    int64_t ali = 0;
    for (uint64_t i = 0; i < h.nBlocks; i++) ali += (int)cs.blocks[h.firstBlock + i].size;
    pushJob(wl, tSeq, qSeq, h.qStrand, wl.chainFirstRecord[c], wl.chainFirstRecord[c + 1] - wl.chainFirstRecord[c],
            GAT_NO_CLIP_START, GAT_NO_CLIP_END, ali);
}

bool chainAscends(const ChainSet &cs, size_t c, bool onQ)
{
    const ChainHead &h = cs.chains[c];
    const gat_block *b = cs.blocks.data() + h.firstBlock;
    for (uint64_t i = 1; i < h.nBlocks; i++) {
        const int prevEnd = (onQ ? b[i - 1].qStart : b[i - 1].tStart) + (int)b[i - 1].size;
        if ((onQ ? b[i].qStart : b[i].tStart) < prevEnd) return false;
    }
    return true;
}

uint64_t firstBlockEndingAfter(const ChainSet &cs, size_t c, int pos, bool onQ)
{
    const ChainHead &h = cs.chains[c];
    const gat_block *b = cs.blocks.data() + h.firstBlock;
    return (uint64_t)(std::partition_point(b, b + h.nBlocks, [&](const gat_block &x) {
                          return (onQ ? x.qStart : x.tStart) + (int)x.size <= pos;
                      }) - b);
}

bool addSubChainJob(const ChainSet &cs, size_t c, uint32_t tSeq, uint32_t qSeq, int subStart, int subEnd, WorkList &wl,
                    uint64_t firstKeptHint)
{   // chainSubsetOnT + chainFastSubsetOnT, chain.c:471-558
    const ChainHead &h = cs.chains[c];
    if (subStart <= h.tStart && subEnd >= h.tEnd) {     // the chain itself, unclipped (:501-506)
        addChainJob(cs, c, tSeq, qSeq, wl);
        return true;
    }
    const gat_block *b = cs.blocks.data() + h.firstBlock;
    uint64_t a = firstKeptHint;
    while (a < h.nBlocks && b[a].tStart + (int)b[a].size <= subStart) a++;       // first block with tEnd > subStart (:479-484)
    uint64_t e = a;
    int64_t ali = 0;
    while (e < h.nBlocks && b[e].tStart < subEnd) {                               // stop at tStart >= subEnd (:510)
        const int ts = std::max(b[e].tStart, subStart), te = std::min(b[e].tStart + (int)b[e].size, subEnd);
        ali += te - ts;
        e++;
    }
    if (e == a) return false;
    // file blocks [a, e) as device records: the same indices unless a block of the batch was split
    uint64_t ra, re;
    if (wl.blockFirstRecord.empty()) { ra = wl.chainFirstRecord[c] + a; re = wl.chainFirstRecord[c] + e; }
    else { ra = wl.blockFirstRecord[h.firstBlock + a]; re = wl.blockFirstRecord[h.firstBlock + e]; }
    pushJob(wl, tSeq, qSeq, h.qStrand, ra, re - ra, subStart, subEnd, ali);
    return true;
}

std::vector<std::vector<uint32_t>> shardJobs(const WorkList &wl, int parts)
{
    std::vector<std::vector<uint32_t>> out(parts);
    if (parts <= 1) {
        out[0].resize(wl.jobs.size());
        for (size_t i = 0; i < wl.jobs.size(); i++) out[0][i] = (uint32_t)i;
        return out;
    }
    // weight = aligned bases + a per-record term (short blocks cost more per base than long ones)
    std::vector<uint32_t> order(wl.jobs.size());
    std::vector<int64_t> weight(wl.jobs.size());
    for (size_t i = 0; i < wl.jobs.size(); i++) {
        order[i] = (uint32_t)i;
        const uint64_t next = i + 1 < wl.jobs.size() ? wl.jobs[i + 1].blockPtr : wl.totalJobBlocks;
        weight[i] = std::max<int64_t>(wl.aliBases[i], 0) + 64 * (int64_t)(next - wl.jobs[i].blockPtr) + 64;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
    typedef std::pair<int64_t, int> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int p = 0; p < parts; p++) heap.push(Load(0, p));
    for (uint32_t j : order) {
        Load l = heap.top();
        heap.pop();
        out[l.second].push_back(j);
        heap.push(Load(l.first + weight[j], l.second));
    }
    for (auto &v : out) std::sort(v.begin(), v.end());      // keep file order inside a shard: neighbours share sectors
    return out;
}

void extractShard(const WorkList &wl, const std::vector<uint32_t> &jobIx, std::vector<gat_job> &jobs, uint64_t &total)
{
    jobs.resize(jobIx.size());
    total = 0;
    for (size_t k = 0; k < jobIx.size(); k++) {
        const uint32_t i = jobIx[k];
        const uint64_t next = (size_t)i + 1 < wl.jobs.size() ? wl.jobs[i + 1].blockPtr : wl.totalJobBlocks;
        jobs[k] = wl.jobs[i];
        jobs[k].blockPtr = (uint32_t)total;
        total += next - wl.jobs[i].blockPtr;
    }
}

MultiGpu::MultiGpu(int nGpus)
{
    const int have = gat_device_count();
    if (have == 0) fail("no CUDA device: chain scoring has no CPU path (%s)", gat_last_error());
    std::vector<int> devices;
    if (const char *list = getenv("GAT_DEVICES")) {         // e.g. "0,0": two contexts on one device
        for (const char *p = list; *p;) {
            devices.push_back(atoi(p));
            while (*p && *p != ',') p++;
            if (*p == ',') p++;
        }
        if ((int)devices.size() < nGpus) fail("-gpus=%d but GAT_DEVICES names %d device(s)", nGpus, (int)devices.size());
        devices.resize(nGpus);
    } else {
        if (nGpus > have) fail("-gpus=%d but only %d CUDA device(s) visible", nGpus, have);
        for (int d = 0; d < nGpus; d++) devices.push_back(d);
    }
    for (int d : devices) {
        gat_ctx *c = nullptr;
        if (gat_create(&c, d, nullptr) != GAT_OK) fail("%s", gat_last_error());
        ctx.push_back(c);
    }
    staging_.resize(ctx.size());
}

MultiGpu::~MultiGpu()
{
    for (auto &st : staging_) gat_host_free(st.p);
    for (gat_ctx *c : ctx) gat_destroy(c);
}

void *MultiGpu::pinned(size_t g, size_t bytes)
{
    Staging &st = staging_[g];
    if (bytes > st.cap) {
        gat_host_free(st.p);
        st.cap = bytes + bytes / 4 + 4096;
        st.p = gat_host_alloc(st.cap);
        if (!st.p) { st.cap = 0; fail("%s", gat_last_error()); }
    }
    return st.p;
}

void MultiGpu::prepare(const TwoBitFile &tbT, const std::vector<int> &useT, const TwoBitFile &tbQ, const std::vector<int> &useQ,
                       const ScoreScheme &ss, const GapCalc &gc)
{
    gap_ = gc;
    haveGap_ = true;
    tNames_.clear(); qNames_.clear();
    for (int i : useT) tNames_.push_back(tbT.seqs()[i].name);
    for (int i : useQ) qNames_.push_back(tbQ.seqs()[i].name);
    std::vector<std::string> errors(ctx.size());
    std::vector<std::thread> threads;
    for (size_t g = 0; g < ctx.size(); g++)
        threads.emplace_back([&, g]() {
            try {
                uploadGenome(ctx[g], GAT_TARGET, tbT, useT);
                uploadGenome(ctx[g], GAT_QUERY, tbQ, useQ);
                setScoring(ctx[g], ss, gc);
            } catch (const Error &e) { errors[g] = e.message; }
        });
    for (auto &t : threads) t.join();
    for (const auto &e : errors)
        if (!e.empty()) fail("%s", e.c_str());
}

// ------------------------------------------------------------------ chainRemovePartialOverlaps
namespace {
struct XBlock { int tStart, tEnd, qStart, qEnd; };
struct XKey {
    uint32_t tSeq, qSeq; int lt, lq, rt, rq, ov;
    bool operator<(const XKey &o) const
    {
        return std::tie(tSeq, qSeq, lt, lq, rt, rq, ov) < std::tie(o.tSeq, o.qSeq, o.lt, o.lq, o.rt, o.rq, o.ov);
    }
};
typedef std::map<XKey, std::pair<int, int>> XCache;     // -> (crossover, scoreAdjustment); crossover -1 = asked, not answered yet

// One chain, chainConnect.c:255-344.  Returns false if a crossover was missing from the cache (its key is in `want`).
bool trimChain(const ChainHead &h, uint32_t tSeq, uint32_t qSeq, std::vector<XBlock> &bl, XCache &cache, std::vector<XKey> &want)
{
    for (size_t i = 1; i < bl.size(); i++)      // checkChainIncreases, :182-198
        if (bl[i - 1].qStart >= bl[i].qStart || bl[i - 1].tStart >= bl[i].tStart)
            fail("a (%d %d) not before b (%d %d) %s", bl[i - 1].qStart, bl[i - 1].tStart, bl[i].qStart, bl[i].tStart, "before removePartialOverlaps");
    for (;;) {
        bool totalTrimA = false;
        size_t a = 0, b = 1;
        while (b < bl.size()) {
            bool totalTrimB = false;
            const int dq = bl[b].qStart - bl[a].qEnd, dt = bl[b].tStart - bl[a].tEnd;
            if (dq < 0 || dt < 0) {
                const int overlap = -std::min(dq, dt);
                const int aSize = bl[a].qEnd - bl[a].qStart, bSize = bl[b].qEnd - bl[b].qStart;
                if (overlap >= aSize || overlap >= bSize) totalTrimB = true;
                else {
                    const XKey k{tSeq, qSeq | (h.qStrand == '-' ? GAT_QSEQ_MINUS : 0u), bl[a].tEnd, bl[a].qEnd, bl[b].tStart, bl[b].qStart, overlap};
                    auto it = cache.find(k);
                    if (it == cache.end() || it->second.first < 0) {
                        if (it == cache.end()) { cache[k] = std::make_pair(-1, 0); want.push_back(k); }
                        return false;
                    }
                    const int crossover = it->second.first, invCross = overlap - crossover;
                    bl[b].qStart += crossover; bl[b].tStart += crossover;
                    bl[a].qEnd -= invCross; bl[a].tEnd -= invCross;
                    if (bl[b].qEnd <= bl[b].qStart) totalTrimB = true;
                    else if (bl[a].qEnd <= bl[a].qStart) totalTrimA = true;
                }
            }
            if (totalTrimA) {           // removeNegativeBlocks, :214-240, then start over
                bl.erase(std::remove_if(bl.begin(), bl.end(), [](const XBlock &x) { return x.qStart >= x.qEnd || x.tStart >= x.tEnd; }), bl.end());
                break;
            } else if (totalTrimB) bl.erase(bl.begin() + (long)b);
            else { a = b; b++; }
        }
        if (!totalTrimA) break;
        if (bl.size() < 2) break;
    }
    for (size_t i = 1; i < bl.size(); i++)      // checkChainGaps, :164-180
        if (bl[i - 1].qEnd > bl[i].qStart || bl[i - 1].tEnd > bl[i].tStart)
            fail("Negative gap between (%d %d - %d %d) and (%d %d - %d %d) %s", bl[i - 1].qStart, bl[i - 1].tStart, bl[i - 1].qEnd, bl[i - 1].tEnd,
                 bl[i].qStart, bl[i].tStart, bl[i].qEnd, bl[i].tEnd, "after removePartialOverlaps");
    for (const XBlock &x : bl)                  // checkStartBeforeEnd, :150-162
        if (x.qStart >= x.qEnd || x.tStart >= x.tEnd)
            fail("Start after end in (%d %d) to (%d %d) %s", x.qStart, x.tStart, x.qEnd, x.tEnd, "after removePartialOverlaps");
    return true;
}
}  // namespace

void removePartialOverlaps(gat_ctx *ctx, ChainSet &cs, const std::vector<uint32_t> &chainT, const std::vector<uint32_t> &chainQ)
{
    const size_t n = cs.chains.size();
    auto original = [&](size_t c) {
        const ChainHead &h = cs.chains[c];
        std::vector<XBlock> bl(h.nBlocks);
        for (uint64_t i = 0; i < h.nBlocks; i++) {
            const gat_block &b = cs.blocks[h.firstBlock + i];
            bl[i] = XBlock{b.tStart, b.tStart + (int)b.size, b.qStart, b.qStart + (int)b.size};
        }
        return bl;
    };
    XCache cache;
    std::vector<XKey> want;
    // first sweep, speculative: every overlapping neighbour pair as the file has it (later trims only change a pair's
    // inputs after a block dried up, which is rare)
    for (size_t c = 0; c < n; c++) {
        const ChainHead &h = cs.chains[c];
        const std::vector<XBlock> bl = original(c);
        for (size_t i = 1; i < bl.size(); i++) {
            const int dq = bl[i].qStart - bl[i - 1].qEnd, dt = bl[i].tStart - bl[i - 1].tEnd;
            if (dq >= 0 && dt >= 0) continue;
            const int overlap = -std::min(dq, dt);
            if (overlap >= bl[i - 1].qEnd - bl[i - 1].qStart || overlap >= bl[i].qEnd - bl[i].qStart) continue;
            const XKey k{chainT[c], chainQ[c] | (h.qStrand == '-' ? GAT_QSEQ_MINUS : 0u), bl[i - 1].tEnd, bl[i - 1].qEnd, bl[i].tStart, bl[i].qStart, overlap};
            if (cache.emplace(k, std::make_pair(-1, 0)).second) want.push_back(k);
        }
    }
    std::vector<std::vector<XBlock>> done(n);
    std::vector<char> finished(n, 0);
    for (;;) {
        if (!want.empty()) {
            std::vector<gat_xpair> pairs(want.size());
            std::vector<int32_t> pos(want.size()), adj(want.size());
            for (size_t i = 0; i < want.size(); i++)
                pairs[i] = gat_xpair{want[i].tSeq, want[i].qSeq, want[i].lt, want[i].lq, want[i].rt, want[i].rq, want[i].ov};
            if (gat_crossover(ctx, pairs.data(), pairs.size(), pos.data(), adj.data()) != GAT_OK) fail("%s", gat_last_error());
            for (size_t i = 0; i < want.size(); i++) cache[want[i]] = std::make_pair((int)pos[i], (int)adj[i]);
            want.clear();
        }
        bool pending = false;
        for (size_t c = 0; c < n; c++) {
            if (finished[c]) continue;
            std::vector<XBlock> bl = original(c);
            if (trimChain(cs.chains[c], chainT[c], chainQ[c], bl, cache, want)) { done[c] = std::move(bl); finished[c] = 1; }
            else pending = true;
        }
        if (!pending) break;
    }
    std::vector<gat_block> blocks;
    blocks.reserve(cs.blocks.size());
    for (size_t c = 0; c < n; c++) {
        ChainHead &h = cs.chains[c];
        h.firstBlock = blocks.size();
        for (const XBlock &x : done[c]) blocks.push_back(gat_block{x.tStart, x.qStart, (uint32_t)(x.tEnd - x.tStart)});
        h.nBlocks = done[c].size();
        if (!done[c].empty()) {                 // setChainBounds, :242-253
            h.tStart = done[c].front().tStart; h.qStart = done[c].front().qStart;
            h.tEnd = done[c].back().tEnd; h.qEnd = done[c].back().qEnd;
        }
    }
    cs.blocks.swap(blocks);
}

// CUDA context creation takes most of a second: tools start it first and parse their inputs meanwhile.
struct GpuStarter::Impl {
    std::thread th;
    MultiGpu *gpus = nullptr;
    std::string error;
};
GpuStarter::GpuStarter(int nGpus) : impl(new Impl)
{
    Impl *im = impl;
    // load the kernels with the context, on this background thread, instead of at the first launch
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    im->th = std::thread([im, nGpus]() {
        try { im->gpus = new MultiGpu(nGpus); } catch (const Error &e) { im->error = e.message; }
    });
}
MultiGpu &GpuStarter::get()
{
    if (impl->th.joinable()) impl->th.join();
    if (!impl->gpus) fail("%s", impl->error.c_str());
    return *impl->gpus;
}
GpuStarter::~GpuStarter()
{
    if (impl->th.joinable()) impl->th.join();
    // the contexts are left to process exit on purpose (tools leave through _exit, see runTool)
    delete impl;
}

// The work-list of whole chains as gat_score_compact takes it (include/gat.h): per block its size and the gap in front of
// it, absolute starts only where a chain begins or a gap does not fit 16 bits.  False if the list does not qualify
// (clipped or shared-record jobs, records beyond GAT_CBLOCK_MAX_SIZE, more than 32768 sequences).
bool packCompact(const WorkList &wl, CompactWorkList &out)
{
    const size_t nJobs = wl.jobs.size(), n = wl.blocks.size();
    if (wl.totalJobBlocks != n) return false;
    out.jobs.resize(nJobs);
    for (size_t j = 0; j < nJobs; j++) {
        const gat_job &job = wl.jobs[j];
        if (job.firstBlock != job.blockPtr || job.clipStart != GAT_NO_CLIP_START || job.clipEnd != GAT_NO_CLIP_END) return false;
        if (job.tSeq > 0xffffu || (job.qSeq & 0x7fffffffu) > 0x7fffu) return false;
        out.jobs[j] = gat_cjob{job.blockPtr, (uint16_t)job.tSeq, (uint16_t)((job.qSeq & 0x7fffu) | ((job.qSeq >> 31) ? GAT_CJOB_MINUS : 0u))};
    }
    out.blocks.resize(n);
    out.abs.clear();
    out.anchors.resize((n + GAT_CGROUP - 1) / GAT_CGROUP);
    size_t nextJob = 0;
    for (size_t i = 0; i < n; i++) {
        const gat_block &b = wl.blocks[i];
        const uint32_t size = b.size & 0x7fffffffu;
        if (size > GAT_CBLOCK_MAX_SIZE) return false;
        bool isAbs = false;
        while (nextJob < nJobs && wl.jobs[nextJob].blockPtr <= i) {        // a chain starts here (empty jobs start nothing)
            const uint64_t end = nextJob + 1 < nJobs ? wl.jobs[nextJob + 1].blockPtr : n;
            if (wl.jobs[nextJob].blockPtr == i && end > i) isAbs = true;
            nextJob++;
        }
        long long dt = 0, dq = 0;
        if (i > 0) {
            const gat_block &p = wl.blocks[i - 1];
            const long long ps = p.size & 0x7fffffffu;
            dt = (long long)b.tStart - ((long long)p.tStart + ps);
            dq = (long long)b.qStart - ((long long)p.qStart + ps);
        } else isAbs = true;
        if (dt < 0 || dt > 0xffff || dq < 0 || dq > 0xffff) isAbs = true;
        uint16_t s16 = (uint16_t)(size | ((b.size & GAT_BLOCK_JOINED) ? GAT_CBLOCK_JOINED : 0u));
        if (isAbs) {
            const uint64_t ix = out.abs.size();
            if (ix >> 32) return false;
            out.abs.push_back(gat_cabs{b.tStart, b.qStart});
            out.blocks[i] = gat_cblock{(uint16_t)(s16 | GAT_CBLOCK_ABS), (uint16_t)(ix & 0xffff), (uint16_t)(ix >> 16)};
        } else out.blocks[i] = gat_cblock{s16, (uint16_t)dt, (uint16_t)dq};
        if (i % GAT_CGROUP == 0) out.anchors[i / GAT_CGROUP] = gat_cabs{b.tStart, b.qStart};
    }
    return true;
}

void MultiGpu::score(const WorkList &wl, std::vector<int64_t> &global, std::vector<int64_t> &local)
{
    const size_t nJobs = wl.jobs.size(), nGpus = ctx.size();
    global.assign(nJobs, 0);
    local.assign(nJobs, 0);
    lastShards.assign(nGpus, ShardStats());
    if (wl.jobs.empty()) return;
    if (const char *prefix = getenv("GAT_DUMP_WORKLIST")) {
        auto put = [&](const std::string &path, const void *p, size_t bytes) {
            FILE *f = fopen(path.c_str(), "wb");
            if (!f || (bytes && fwrite(p, 1, bytes, f) != bytes)) fail("cannot write %s", path.c_str());
            fclose(f);
        };
        const std::string base = std::string(prefix) + "." + std::to_string(dumpCount_++);
        put(base + ".jobs", wl.jobs.data(), wl.jobs.size() * sizeof(gat_job));
        put(base + ".blocks", wl.blocks.data(), wl.blocks.size() * sizeof(gat_block));
        const std::string meta = "{\"jobs\": " + std::to_string(wl.jobs.size()) + ", \"blocks\": " + std::to_string(wl.blocks.size()) +
                                 ", \"totalJobBlocks\": " + std::to_string(wl.totalJobBlocks) + "}\n";
        put(base + ".meta", meta.data(), meta.size());
        std::string names;
        for (const auto &n : tNames_) names += n + "\n";
        put(std::string(prefix) + ".tseqs", names.data(), names.size());
        names.clear();
        for (const auto &n : qNames_) names += n + "\n";
        put(std::string(prefix) + ".qseqs", names.data(), names.size());
    }
    if (nGpus == 1) {
        CompactWorkList cw;
        lastShards[0].jobs = nJobs; lastShards[0].records = wl.blocks.size();
        if (packCompact(wl, cw)) {      // whole chains (scoreChain): half the bytes over PCIe
            lastShards[0].compact = true;
            lastShards[0].h2dBytes = cw.jobs.size() * sizeof(gat_cjob) + cw.blocks.size() * sizeof(gat_cblock) + (cw.abs.size() + cw.anchors.size()) * sizeof(gat_cabs);
            if (gat_score_compact(ctx[0], cw.jobs.data(), cw.jobs.size(), cw.blocks.data(), cw.blocks.size(), cw.abs.data(), cw.abs.size(),
                                  cw.anchors.data(), global.data(), local.data()) != GAT_OK)
                fail("%s", gat_last_error());
            if (getenv("GAT_TOOL_TIMING"))
                fprintf(stderr, "gpu shard 0: %llu jobs, %llu records, %llu bytes host->device (compact), 0 pieces of cut chains\n",
                        (unsigned long long)nJobs, (unsigned long long)wl.blocks.size(), (unsigned long long)lastShards[0].h2dBytes);
            return;
        }
        lastShards[0].h2dBytes = nJobs * sizeof(gat_job) + wl.blocks.size() * sizeof(gat_block);
        if (gat_score(ctx[0], wl.jobs.data(), nJobs, wl.totalJobBlocks, wl.blocks.data(), wl.blocks.size(),
                      global.data(), local.data()) != GAT_OK)
            fail("%s", gat_last_error());
        if (getenv("GAT_TOOL_TIMING"))
            fprintf(stderr, "gpu shard 0: %llu jobs, %llu records, %llu bytes host->device (plain), 0 pieces of cut chains\n",
                    (unsigned long long)nJobs, (unsigned long long)wl.blocks.size(), (unsigned long long)lastShards[0].h2dBytes);
        return;
    }

    // ---- pieces: every job is one piece, except whole-chain jobs above 1/(4 nGpus) of the aligned bases
    struct Piece { uint32_t job; uint64_t first, count; int64_t weight; };
    std::vector<Piece> pieces;
    std::vector<uint32_t> firstPiece(nJobs + 1);
    int64_t allBases = 0;
    for (size_t j = 0; j < nJobs; j++) allBases += std::max<int64_t>(wl.aliBases[j], 0);
    const int64_t limit = std::max<int64_t>(1, allBases / (4 * (int64_t)nGpus));
    pieces.reserve(nJobs + 8 * nGpus);
    for (size_t j = 0; j < nJobs; j++) {
        const gat_job &job = wl.jobs[j];
        const uint64_t count = (j + 1 < nJobs ? wl.jobs[j + 1].blockPtr : wl.totalJobBlocks) - job.blockPtr;
        firstPiece[j] = (uint32_t)pieces.size();
        const bool whole = job.clipStart == GAT_NO_CLIP_START && job.clipEnd == GAT_NO_CLIP_END;
        if (haveGap_ && whole && wl.aliBases[j] > limit && count >= 2 * (uint64_t)GAT_TUPLE_MIN_BLOCKS) {
            const int k = (int)((wl.aliBases[j] + limit - 1) / limit);
            std::vector<uint64_t> cuts{0};          // records (within the job) at which a piece starts
            uint64_t i = 0;
            int64_t run = 0;
            for (int part = 1; part < k; part++) {
                const int64_t until = wl.aliBases[j] * part / k;
                while (i < count && run < until) { run += wl.blocks[job.firstBlock + i].size & 0x7fffffffu; i++; }
                // never start a piece with a continuation record; a piece (and what is left behind it) must be long enough for
                // the fix-up kernel to finish it
                while (i < count && (wl.blocks[job.firstBlock + i].size & GAT_BLOCK_JOINED)) { run += wl.blocks[job.firstBlock + i].size & 0x7fffffffu; i++; }
                if (i - cuts.back() >= GAT_TUPLE_MIN_BLOCKS && count - i >= GAT_TUPLE_MIN_BLOCKS) cuts.push_back(i);
            }
            for (size_t c = 0; c < cuts.size(); c++)
                pieces.push_back(Piece{(uint32_t)j, job.firstBlock + cuts[c], (c + 1 < cuts.size() ? cuts[c + 1] : count) - cuts[c], 0});
        } else pieces.push_back(Piece{(uint32_t)j, job.firstBlock, count, 0});
    }
    firstPiece[nJobs] = (uint32_t)pieces.size();
    for (Piece &p : pieces) {
        int64_t bases = 0;
        if (firstPiece[p.job + 1] - firstPiece[p.job] == 1) bases = std::max<int64_t>(wl.aliBases[p.job], 0);
        else for (uint64_t r = 0; r < p.count; r++) bases += wl.blocks[p.first + r].size & 0x7fffffffu;
        p.weight = bases + 64 * (int64_t)p.count + 64;      // short blocks cost more per base than long ones
    }
    // ---- greedy longest-processing-time balance, over units: a heavy piece is a unit of its own, light pieces go in runs of
    // neighbours in file order (neighbours share genome sectors, and a few hundred units sort and balance in no time where
    // a million pieces do not)
    int64_t allWeight = 0;
    for (const Piece &p : pieces) allWeight += p.weight;
    const int64_t heavy = std::max<int64_t>(1, allWeight / (64 * (int64_t)nGpus));
    struct Unit { uint32_t first, count; int64_t weight; };
    std::vector<Unit> units;
    for (uint32_t i = 0; i < pieces.size();) {
        Unit u{i, 1, pieces[i].weight};
        i++;
        if (u.weight < heavy)
            while (i < pieces.size() && pieces[i].weight < heavy && u.weight + pieces[i].weight <= heavy) { u.weight += pieces[i].weight; u.count++; i++; }
        units.push_back(u);
    }
    std::vector<uint32_t> order(units.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return units[a].weight > units[b].weight; });
    typedef std::pair<int64_t, int> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (size_t g = 0; g < nGpus; g++) heap.push(Load(0, (int)g));
    std::vector<std::vector<uint32_t>> shardUnits(nGpus);
    for (uint32_t ui : order) {
        Load l = heap.top();
        heap.pop();
        shardUnits[l.second].push_back(ui);
        heap.push(Load(l.first + units[ui].weight, l.second));
    }
    std::vector<std::vector<uint32_t>> shard(nGpus);
    for (size_t g = 0; g < nGpus; g++) {
        std::sort(shardUnits[g].begin(), shardUnits[g].end());  // file order inside a shard
        for (uint32_t ui : shardUnits[g])
            for (uint32_t k = 0; k < units[ui].count; k++) shard[g].push_back(units[ui].first + k);
    }

    // ---- one host thread per GPU: compact the shard's records, stage them in pinned memory, score
    std::vector<int64_t> pGlobal(pieces.size()), pLocal(pieces.size());
    std::vector<gat_tuple> pTuple(pieces.size());
    std::vector<std::string> errors(nGpus);
    std::vector<std::thread> threads;
    for (size_t g = 0; g < nGpus; g++)
        threads.emplace_back([&, g]() {
            try {
                const std::vector<uint32_t> &mine = shard[g];
                if (mine.empty()) return;
                const auto t0 = std::chrono::steady_clock::now();
                auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
                WorkList sw;            // the shard as a work-list of its own: records copied range by range
                sw.jobs.resize(mine.size());
                uint64_t total = 0;
                for (uint32_t pi : mine) total += pieces[pi].count;
                sw.blocks.resize(total);
                std::vector<uint32_t> wantTuple;
                uint64_t at = 0;
                for (size_t k = 0; k < mine.size(); k++) {
                    const Piece &p = pieces[mine[k]];
                    gat_job job = wl.jobs[p.job];
                    job.firstBlock = job.blockPtr = (uint32_t)at;
                    sw.jobs[k] = job;
                    memcpy(sw.blocks.data() + at, wl.blocks.data() + p.first, p.count * sizeof(gat_block));
                    at += p.count;
                    if (firstPiece[p.job + 1] - firstPiece[p.job] > 1) wantTuple.push_back((uint32_t)k);
                }
                sw.totalJobBlocks = total;
                ShardStats &st = lastShards[g];
                st.jobs = mine.size(); st.records = total; st.pieces = wantTuple.size();
                std::vector<int64_t> gl(mine.size()), lo(mine.size());
                std::vector<gat_tuple> tup(wantTuple.size());
                if (!wantTuple.empty() && gat_request_tuples(ctx[g], wantTuple.data(), wantTuple.size(), tup.data()) != GAT_OK)
                    fail("%s", gat_last_error());
                CompactWorkList cw;
                int rc;
                const bool packed = packCompact(sw, cw);
                st.buildS = since(t0);
                const auto t1 = std::chrono::steady_clock::now();
                if (packed) {
                    const size_t bj = cw.jobs.size() * sizeof(gat_cjob), bb = cw.blocks.size() * sizeof(gat_cblock),
                                 ba = cw.abs.size() * sizeof(gat_cabs), bn = cw.anchors.size() * sizeof(gat_cabs);
                    auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
                    char *base = static_cast<char *>(pinned(g, up(bj) + up(bb) + up(ba) + up(bn) + 64));
                    char *pj = base, *pb = pj + up(bj), *pa = pb + up(bb), *pn = pa + up(ba);
                    memcpy(pj, cw.jobs.data(), bj); memcpy(pb, cw.blocks.data(), bb); memcpy(pa, cw.abs.data(), ba); memcpy(pn, cw.anchors.data(), bn);
                    st.compact = true; st.h2dBytes = bj + bb + ba + bn;
                    st.stageS = since(t1);
                    rc = gat_score_compact(ctx[g], reinterpret_cast<gat_cjob *>(pj), cw.jobs.size(), reinterpret_cast<gat_cblock *>(pb), cw.blocks.size(),
                                           reinterpret_cast<gat_cabs *>(pa), cw.abs.size(), reinterpret_cast<gat_cabs *>(pn), gl.data(), lo.data());
                } else {
                    const size_t bj = sw.jobs.size() * sizeof(gat_job), bb = sw.blocks.size() * sizeof(gat_block);
                    char *base = static_cast<char *>(pinned(g, bj + bb + 128));
                    char *pj = base, *pb = base + ((bj + 63) & ~(size_t)63);
                    memcpy(pj, sw.jobs.data(), bj); memcpy(pb, sw.blocks.data(), bb);
                    st.h2dBytes = bj + bb;
                    st.stageS = since(t1);
                    rc = gat_score(ctx[g], reinterpret_cast<gat_job *>(pj), sw.jobs.size(), total, reinterpret_cast<gat_block *>(pb), sw.blocks.size(),
                                   gl.data(), lo.data());
                }
                if (rc != GAT_OK) fail("%s", gat_last_error());
                st.scoreS = since(t1) - st.stageS;
                for (size_t k = 0; k < mine.size(); k++) { pGlobal[mine[k]] = gl[k]; pLocal[mine[k]] = lo[k]; }
                for (size_t k = 0; k < wantTuple.size(); k++) pTuple[mine[wantTuple[k]]] = tup[k];
            } catch (const Error &e) { errors[g] = e.message; }
        });
    for (auto &t : threads) t.join();
    for (const auto &e : errors)
        if (!e.empty()) fail("%s", e.c_str());

    if (getenv("GAT_TOOL_TIMING"))
        for (size_t g = 0; g < nGpus; g++)
            fprintf(stderr, "gpu shard %zu: %llu jobs, %llu records, %llu bytes host->device (%s), %llu pieces of cut chains; host: build %.3f s, stage %.3f s, "
                    "scoring call %.3f s\n", g,
                    (unsigned long long)lastShards[g].jobs, (unsigned long long)lastShards[g].records, (unsigned long long)lastShards[g].h2dBytes,
                    lastShards[g].compact ? "compact" : "plain", (unsigned long long)lastShards[g].pieces, lastShards[g].buildS, lastShards[g].stageS,
                    lastShards[g].scoreS);
    // ---- back to jobs; the pieces of a cut chain are joined in order with the gap cost between them
    for (size_t j = 0; j < nJobs; j++) {
        const uint32_t p0 = firstPiece[j], p1 = firstPiece[j + 1];
        if (p1 - p0 == 1) { global[j] = pGlobal[p0]; local[j] = pLocal[p0]; continue; }
        gat_tuple acc = pTuple[p0];
        for (uint32_t p = p0 + 1; p < p1; p++) {
            const gat_block &last = wl.blocks[pieces[p - 1].first + pieces[p - 1].count - 1], &first = wl.blocks[pieces[p].first];
            const int ls = (int)(last.size & 0x7fffffffu);
            gat_tuple_join(&acc, gap_.cost(first.qStart - (last.qStart + ls), first.tStart - (last.tStart + ls)), &pTuple[p]);
        }
        gat_tuple_scores(&acc, &global[j], &local[j]);
    }
}

}  // namespace gathost
