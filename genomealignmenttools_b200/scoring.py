"""Host side of the scoring parameters: the substitution matrix (struct axtScoreScheme,
kent/src/inc/axt.h:83-91) and the piecewise-linear gap costs (struct gapCalc,
kent/src/lib/gapCalc.c:12-37).  Same names, same file formats, same errors as the reference;
the output is the flat gat_scoring the C ABI takes.  gapCalcCost() below exists for host-side
consumers (tests, tools' -verbose output); scoring itself happens on the GPU.
"""
import numpy as np

# kent base codes (dnautil.h:23-27)
T, C, A, G = 0, 1, 2, 3
_FILE_ORDER = (A, C, G, T)  # matrix files are written A C G T (axt.c:697-701)

_MEDIUM = """tableSize 11
smallSize 111
position 1 2 3 11 111 2111 12111 32111 72111 152111 252111
qGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900
tGap 350 425 450 600 900 2900 22900 57900 117900 217900 317900
bothGap 750 825 850 1000 1300 3300 23300 58300 118300 218300 318300
"""  # gapCalc.c:40-46 ("original", -linearGap=medium)
_LOOSE = """tablesize 11
smallSize 111
position 1 2 3 11 111 2111 12111 32111 72111 152111 252111
qGap 325 360 400 450 600 1100 3600 7600 15600 31600 56600
tGap 325 360 400 450 600 1100 3600 7600 15600 31600 56600
bothGap 625 660 700 750 900 1400 4000 8000 16000 32000 57000
"""  # gapCalc.c:50-56 ("default", -linearGap=loose)


class ScoreSchemeError(ValueError):
    pass


def _atoi(word):
    """C atoi: optional sign, leading digits, stops at the first non-digit."""
    s = word.strip()
    i, sign = 0, 1
    if s[:1] in "+-":
        sign = -1 if s[0] == "-" else 1
        i = 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    return sign * int(s[i:j]) if j > i else 0


class ScoreScheme:
    """4x4 substitution scores, matrix[q][t] in kent base codes.  Every other character pair
    (N, IUPAC) scores 0 in the reference because its 256x256 table is zero-initialised and only
    the a/c/g/t cells are ever written (axt.c:431-454, 773-783)."""

    def __init__(self, matrix, gap_open=400, gap_extend=30):
        self.matrix = np.asarray(matrix, dtype=np.int32).reshape(4, 4)
        self.gap_open, self.gap_extend = gap_open, gap_extend

    @classmethod
    def default(cls):
        """axtScoreSchemeDefault (axt.c:423-458): the blastz matrix."""
        acgt = [[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]]
        return cls._from_file_order(acgt)

    @classmethod
    def _from_file_order(cls, rows, **kw):
        m = np.zeros((4, 4), dtype=np.int32)
        for i, qi in enumerate(_FILE_ORDER):
            for j, tj in enumerate(_FILE_ORDER):
                m[qi, tj] = rows[i][j]
        return cls(m, **kw)

    @classmethod
    def read(cls, path):
        """axtScoreSchemeRead (axt.c:692-834): blastz and lastz-settings formats."""
        with open(path, "r") as f:
            lines = f.read().split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        pos = 0

        def chop_next():
            nonlocal pos
            while pos < len(lines):
                line = lines[pos]; pos += 1
                if line.startswith("#"):
                    continue
                w = line.split()
                if w:
                    return w[:6]
            return None

        while True:
            w = chop_next()
            if w is None:
                raise ScoreSchemeError("Scoring matrix file %s too short" % path)
            if "=" in w[0] or (len(w) > 1 and "=" in w[1]):
                continue
            if len(w) < 4 or w[0][0] != "A" or w[1][0] != "C" or w[2][0] != "G" or w[3][0] != "T":
                raise ScoreSchemeError("%s doesn't seem to be a score matrix file" % path)
            break
        rows = []
        for _ in range(4):
            w = chop_next()
            if w is None:
                raise ScoreSchemeError("Scoring matrix file %s too short" % path)
            first = 1 if len(w) == 5 else 0
            if len(w) < first + 4:
                raise ScoreSchemeError("matrix row of %s too short" % path)
            vals = []
            for a in w[first:first + 4]:
                if a[0] != "-" and not a[0].isdigit():
                    raise ScoreSchemeError("Expecting number, got %s in %s" % (a, path))
                vals.append(_atoi(a))
            rows.append(vals)
        gap_open, gap_extend = 400, 30
        if pos < len(lines):  # lineFileNext: the very next raw line, blank or not (axt.c:785-805)
            parts = [p for p in lines[pos].replace("=", " ").replace(",", " ").replace("\t", " ").split(" ") if p]
            got = {}
            for i in range(0, len(parts) - 1, 2):
                if parts[i] in ("O", "E"):
                    got[parts[i]] = _atoi(parts[i + 1])
            if "O" not in got or "E" not in got:
                raise ScoreSchemeError("Expecting O = and E = in last line of %s" % path)
            if got["O"] <= 0 or got["E"] <= 0:
                raise ScoreSchemeError("Must have positive gap scores")
            gap_open, gap_extend = got["O"], got["E"]
        return cls._from_file_order(rows, gap_open=gap_open, gap_extend=gap_extend)


class GapCalcError(ValueError):
    pass


def _trunc_int(d):
    """(int)double as x86-64 does it: toward zero, INT_MIN when out of range."""
    if not (-2147483649.0 < d < 2147483648.0):
        return -(2 ** 31)
    return int(d)


def _interpolate(x, pos, val):
    """gapCalc.c:82-104, doubles evaluated in the reference's order."""
    n = len(pos)
    for i in range(n):
        if x == pos[i]:
            return _trunc_int(val[i])
        if x < pos[i]:
            ds = pos[i] - pos[i - 1]
            dv = val[i] - val[i - 1]
            return _trunc_int(val[i - 1] + dv * float(x - pos[i - 1]) / float(ds))
    ds = pos[n - 1] - pos[n - 2]
    dv = val[n - 1] - val[n - 2]
    return _trunc_int(val[n - 2] + dv * float(x - pos[n - 2]) / float(ds))


class GapCalc:
    """gapCalcFromFile / gapCalcRead (gapCalc.c:146-256)."""

    def __init__(self, text):
        lines = [l for l in text.split("\n") if l.strip() and not l.lstrip().startswith("#")]
        it = iter(lines)

        def tagged(tag, count, as_float):
            try:
                words = next(it).split()
            except StopIteration:
                raise GapCalcError("gap spec ends before %s" % tag)
            if words[0].lower() != tag.lower():
                raise GapCalcError("Expecting %s got %s" % (tag, words[0]))
            nums = words[1:]
            if len(nums) < count:
                raise GapCalcError("Not enough numbers on %s line" % tag)
            if len(nums) > count:
                raise GapCalcError("Too many numbers on %s line" % tag)
            out = []
            for w in nums:
                if not w[0].isdigit():
                    raise GapCalcError("Expecting number got %s" % w)
                out.append(float(w) if as_float else _atoi(w))
            return out

        table_size = tagged("tableSize", 1, False)[0]
        self.small_size = tagged("smallSize", 1, False)[0]
        pos = tagged("position", table_size, False)
        qv = tagged("qGap", table_size, True)
        tv = tagged("tGap", table_size, True)
        bv = tagged("bothGap", table_size, True)
        if pos[0] > 1:
            raise GapCalcError("gap table must start at position 1")
        s = self.small_size
        self.q_small = np.zeros(s, dtype=np.int32)
        self.t_small = np.zeros(s, dtype=np.int32)
        self.b_small = np.zeros(s, dtype=np.int32)
        for i in range(1, s):  # entry 0 stays 0 (gapCalc.c:171-182)
            self.q_small[i] = _interpolate(i, pos, qv)
            self.t_small[i] = _interpolate(i, pos, tv)
            self.b_small[i] = _interpolate(i, pos, bv)
        if s not in pos:
            raise GapCalcError("No position %d in gapCalcRead()" % s)
        k = pos.index(s)
        self.long_pos = np.asarray(pos[k:], dtype=np.int32)
        self.q_long = np.asarray(qv[k:], dtype=np.float64)
        self.t_long = np.asarray(tv[k:], dtype=np.float64)
        self.b_long = np.asarray(bv[k:], dtype=np.float64)
        if len(self.long_pos) < 2:
            raise GapCalcError("need two long positions")

    @classmethod
    def from_file(cls, name):
        """gapCalcFromFile (gapCalc.c:233-256): 'loose', 'medium' or a file name."""
        if name is None:
            raise GapCalcError("Must specify linear gap costs.  Use 'loose' or 'medium' for defaults")
        if name == "loose":
            return cls(_LOOSE)
        if name == "medium":
            return cls(_MEDIUM)
        with open(name, "r") as f:
            return cls(f.read())

    def cost(self, dq, dt):
        """gapCalcCost (gapCalc.c:298-331) -- host restatement for tools/tests, not the hot path."""
        dt = max(dt, 0); dq = max(dq, 0)
        if dt == 0:
            small, lng, v = self.q_small, self.q_long, dq
        elif dq == 0:
            small, lng, v = self.t_small, self.t_long, dt
        else:
            small, lng, v = self.b_small, self.b_long, dq + dt
        if v < self.small_size:
            return int(small[v])
        last = int(self.long_pos[-1])
        if v >= last:
            slope = (lng[-1] - lng[-2]) / (float(last) - float(self.long_pos[-2]))
            return _trunc_int(float(lng[-1]) + float(slope) * float(v - last))
        return _interpolate(v, [int(p) for p in self.long_pos], [float(x) for x in lng])


class Scoring:
    """What gat_set_scoring takes: a ScoreScheme and a GapCalc."""

    def __init__(self, scheme=None, gap="loose"):
        self.scheme = scheme if scheme is not None else ScoreScheme.default()
        self.gap = gap if isinstance(gap, GapCalc) else GapCalc.from_file(gap)

    @classmethod
    def from_options(cls, score_scheme=None, linear_gap=None):
        """-scoreScheme= / -linearGap= as the three tools parse them (scoreChain.c:253-263)."""
        scheme = ScoreScheme.read(score_scheme) if score_scheme else ScoreScheme.default()
        return cls(scheme, GapCalc.from_file(linear_gap))
