""".chain text I/O (kent/src/lib/chain.c:200-346) into the CSR work-list layout.

Python edition for tests and bench plumbing; the CLI tools use the C++ reader in csrc/host."""
import numpy as np
from .records import BLOCK_DTYPE, JOB_DTYPE, NO_CLIP_START, NO_CLIP_END, QSEQ_MINUS


class ChainFormatError(ValueError):
    pass


class ChainSet:
    """All chains of a file: per-chain header columns + one flat BLOCK_DTYPE array."""

    def __init__(self):
        self.score, self.tName, self.tSize, self.tStart, self.tEnd = [], [], [], [], []
        self.qName, self.qSize, self.qStrand, self.qStart, self.qEnd, self.id = [], [], [], [], [], []
        self.firstBlock, self.nBlocks = [], []
        self.blocks = np.zeros(0, dtype=BLOCK_DTYPE)
        self.meta = []  # '#' lines, passed through by chainNet / chainCleaner (chainNet.c:938-939)

    def __len__(self):
        return len(self.id)

    @classmethod
    def read(cls, path):
        """chainRead (chain.c:337-346): header (:256-296) then 'size [dt dq]' lines (:298-335)."""
        cs = cls()
        bt, bq, bs = [], [], []
        next_id = 1
        opener = open
        if str(path).endswith(".gz"):
            import gzip
            opener = gzip.open
        with opener(path, "rt") as f:
            lines = f.read().split("\n")
        i, n = 0, len(lines)

        def next_words():
            nonlocal i
            while i < n:
                line = lines[i]; i += 1
                if line.startswith("#"):
                    cs.meta.append(line)
                    continue
                w = line.split()
                if w:
                    return w
            return None

        def need_num(w):
            if w[0] != "-" and not w[0].isdigit():
                raise ChainFormatError("Expecting number, got %s line %d of %s" % (w, i, path))
            j = 1
            while j < len(w) and w[j].isdigit():
                j += 1
            return int(w[:j]) if j > 1 or w[0].isdigit() else 0

        while True:
            w = next_words()
            if w is None:
                break
            if len(w) < 12:
                raise ChainFormatError("Expecting at least 12 words line %d of %s" % (i, path))
            if w[0] != "chain":
                raise ChainFormatError("Expecting 'chain' line %d of %s" % (i, path))
            tSize, tStart, tEnd = need_num(w[3]), need_num(w[5]), need_num(w[6])
            qSize, qStart, qEnd = need_num(w[8]), need_num(w[10]), need_num(w[11])
            if len(w) >= 13:
                cid = need_num(w[12])
            else:
                cid = next_id; next_id += 1
            if qStart >= qEnd or tStart >= tEnd:
                raise ChainFormatError("End before start line %d of %s" % (i, path))
            if qStart < 0 or tStart < 0:
                raise ChainFormatError("Start before zero line %d of %s" % (i, path))
            if qEnd > qSize or tEnd > tSize:
                raise ChainFormatError("Past end of sequence line %d of %s" % (i, path))
            cs.score.append(float(w[1])); cs.tName.append(w[2]); cs.tSize.append(tSize)
            cs.tStart.append(tStart); cs.tEnd.append(tEnd); cs.qName.append(w[7]); cs.qSize.append(qSize)
            cs.qStrand.append(w[9][0]); cs.qStart.append(qStart); cs.qEnd.append(qEnd); cs.id.append(cid)
            cs.firstBlock.append(len(bs))
            q, t = qStart, tStart
            while True:
                w = next_words()
                if w is None:
                    raise ChainFormatError("chain %d ends early in %s" % (cid, path))
                size = need_num(w[0])
                bt.append(t); bq.append(q); bs.append(size)
                t += size; q += size
                if len(w) == 1:
                    break
                if len(w) < 3:
                    raise ChainFormatError("Expecting 1 or 3 words line %d of %s" % (i, path))
                t += need_num(w[1]); q += need_num(w[2])
            cs.nBlocks.append(len(bs) - cs.firstBlock[-1])
            if q != qEnd:
                raise ChainFormatError("q end mismatch %d vs %d line %d of %s" % (q, qEnd, i, path))
            if t != tEnd:
                raise ChainFormatError("t end mismatch %d vs %d line %d of %s" % (t, tEnd, i, path))
        cs.blocks = np.zeros(len(bs), dtype=BLOCK_DTYPE)
        cs.blocks["tStart"] = bt; cs.blocks["qStart"] = bq; cs.blocks["size"] = bs
        return cs

    def jobs(self, t_genome, q_genome):
        """One whole-chain job per chain (what scoreChain scores, scoreChain.c:301-311)."""
        jobs = np.zeros(len(self), dtype=JOB_DTYPE)
        jobs["tSeq"] = [t_genome.index(n) for n in self.tName]
        jobs["qSeq"] = np.array([q_genome.index(n) for n in self.qName], dtype=np.uint32) | \
            np.where(np.array(self.qStrand) == "-", np.uint32(QSEQ_MINUS), np.uint32(0))
        jobs["firstBlock"] = self.firstBlock
        jobs["blockPtr"] = self.firstBlock
        jobs["clipStart"] = NO_CLIP_START
        jobs["clipEnd"] = NO_CLIP_END
        return jobs, len(self.blocks)

    def subset_job(self, ix, sub_start, sub_end):
        """chainSubsetOnT's block selection (chain.c:471-558) as (firstBlock, nBlocks, clipStart,
        clipEnd); nBlocks == 0 is kent's NULL sub-chain."""
        fb, nb = self.firstBlock[ix], self.nBlocks[ix]
        if sub_start <= self.tStart[ix] and sub_end >= self.tEnd[ix]:
            return fb, nb, NO_CLIP_START, NO_CLIP_END
        b = self.blocks[fb:fb + nb]
        t_end = b["tStart"].astype(np.int64) + b["size"].astype(np.int64)
        keep = np.nonzero(t_end > sub_start)[0]
        a = int(keep[0]) if len(keep) else nb
        stop = np.nonzero(b["tStart"][a:] >= sub_end)[0]
        e = a + int(stop[0]) if len(stop) else nb
        return fb + a, e - a, sub_start, sub_end


def write_chains(path, headers, blocks, first_block, n_blocks, scores=None):
    """chainWrite (chain.c:211-227).  headers: iterable of
    (score, tName, tSize, tStart, tEnd, qName, qSize, qStrand, qStart, qEnd, id)."""
    out = []
    ts, qs, sz = blocks["tStart"], blocks["qStart"], blocks["size"]
    for c, h in enumerate(headers):
        score = h[0] if scores is None else scores[c]
        out.append("chain %.0f %s %d + %d %d %s %d %s %d %d %d\n" % ((score,) + tuple(h[1:])))
        fb, nb = int(first_block[c]), int(n_blocks[c])
        t = ts[fb:fb + nb].astype(np.int64); q = qs[fb:fb + nb].astype(np.int64); s = sz[fb:fb + nb].astype(np.int64)
        dt = t[1:] - (t[:-1] + s[:-1]); dq = q[1:] - (q[:-1] + s[:-1])
        lines = ["%d\t%d\t%d" % (s[k], dt[k], dq[k]) for k in range(nb - 1)]
        lines.append("%d" % s[nb - 1])
        out.append("\n".join(lines) + "\n\n")
    with open(path, "w") as f:
        f.write("".join(out))
