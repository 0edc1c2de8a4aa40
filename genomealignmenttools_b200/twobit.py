""".2bit container (kent/src/lib/twoBit.c): header + index + per-sequence records.

Only the container is handled here -- the payload stays packed (4 bases/byte, T=0 C=1 A=2 G=3,
first base in bits 7..6) and goes to the GPU as is; nothing is unpacked to characters.
File layout: twoBit.c:312-397 (writer), :442-513 / :574-650 (reader); signature sig.h:58-62.
"""
import struct
import numpy as np
from .records import NRUN_DTYPE

SIG = 0x1A412743
SIG_SWAPPED = 0x4327411A


class PackedGenome:
    def __init__(self, names, sizes, packed, byte_offsets, n_runs=None, mask_runs=None):
        self.names = list(names)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        self.packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self.byte_offsets = np.ascontiguousarray(byte_offsets, dtype=np.uint64)
        self.n_runs = np.zeros(0, dtype=NRUN_DTYPE) if n_runs is None else np.ascontiguousarray(n_runs, dtype=NRUN_DTYPE)
        self.mask_runs = np.zeros(0, dtype=NRUN_DTYPE) if mask_runs is None else np.ascontiguousarray(mask_runs, dtype=NRUN_DTYPE)
        self._index = {n: i for i, n in enumerate(self.names)}

    def index(self, name):
        return self._index[name]

    @property
    def total_bases(self):
        return int(self.sizes.astype(np.int64).sum())

    # ------------------------------------------------------------------ read
    @classmethod
    def read_2bit(cls, path):
        data = np.fromfile(path, dtype=np.uint8)
        buf = data.tobytes() if data.size < (1 << 26) else memoryview(data)
        sig = struct.unpack_from("<I", buf, 0)[0]
        if sig == SIG:
            e = "<"
        elif sig == SIG_SWAPPED:
            e = ">"
        else:
            raise ValueError("%s doesn't have a valid twoBitSig" % path)
        version, count, _ = struct.unpack_from(e + "III", buf, 4)
        if version not in (0, 1):
            raise ValueError("Can only handle version 0 or version 1 of this file. This is version %d" % version)
        pos = 16
        names, offsets = [], []
        for _ in range(count):
            ln = buf[pos]
            names.append(bytes(buf[pos + 1:pos + 1 + ln]).decode())
            pos += 1 + ln
            if version == 1:
                offsets.append(struct.unpack_from(e + "Q", buf, pos)[0]); pos += 8
            else:
                offsets.append(struct.unpack_from(e + "I", buf, pos)[0]); pos += 4
        sizes = np.zeros(count, dtype=np.uint32)
        byte_offsets = np.zeros(count, dtype=np.uint64)
        nr, mr = [], []
        u4 = np.dtype(e + "u4")
        for i, off in enumerate(offsets):
            p = off
            size, ncount = struct.unpack_from(e + "II", buf, p); p += 8
            ns = np.frombuffer(buf, dtype=u4, count=ncount, offset=p); p += 4 * ncount
            nl = np.frombuffer(buf, dtype=u4, count=ncount, offset=p); p += 4 * ncount
            mcount = struct.unpack_from(e + "I", buf, p)[0]; p += 4
            ms = np.frombuffer(buf, dtype=u4, count=mcount, offset=p); p += 4 * mcount
            ml = np.frombuffer(buf, dtype=u4, count=mcount, offset=p); p += 4 * mcount
            p += 4  # reserved
            sizes[i] = size
            byte_offsets[i] = p
            if ncount:
                r = np.zeros(ncount, dtype=NRUN_DTYPE); r["seq"] = i; r["start"] = ns; r["len"] = nl; nr.append(r)
            if mcount:
                r = np.zeros(mcount, dtype=NRUN_DTYPE); r["seq"] = i; r["start"] = ms; r["len"] = ml; mr.append(r)
        n_runs = np.concatenate(nr) if nr else None
        mask_runs = np.concatenate(mr) if mr else None
        # the payloads are used in place: `packed` is the whole file, byte_offsets point into it
        return cls(names, sizes, data, byte_offsets, n_runs, mask_runs)

    # ------------------------------------------------------------------ write
    def write_2bit(self, path, version=0, swapped=False):
        e = ">" if swapped else "<"
        count = len(self.names)
        index_size = sum(1 + len(n.encode()) + (8 if version == 1 else 4) for n in self.names)
        pos = 16 + index_size
        recs, offsets = [], []
        for i in range(count):
            nr = self.n_runs[self.n_runs["seq"] == i]
            mr = self.mask_runs[self.mask_runs["seq"] == i]
            nbytes = (int(self.sizes[i]) + 3) // 4
            start = int(self.byte_offsets[i])
            head = struct.pack(e + "II", int(self.sizes[i]), len(nr))
            head += nr["start"].astype(e + "u4").tobytes() + nr["len"].astype(e + "u4").tobytes()
            head += struct.pack(e + "I", len(mr))
            head += mr["start"].astype(e + "u4").tobytes() + mr["len"].astype(e + "u4").tobytes()
            head += struct.pack(e + "I", 0)
            offsets.append(pos)
            recs.append((head, self.packed[start:start + nbytes]))
            pos += len(head) + nbytes
        with open(path, "wb") as f:
            f.write(struct.pack(e + "IIII", SIG, version, count, 0))
            for n, off in zip(self.names, offsets):
                b = n.encode()
                f.write(struct.pack("B", len(b)) + b + struct.pack(e + ("Q" if version == 1 else "I"), off))
            for head, payload in recs:
                f.write(head)
                payload.tofile(f)

    # ------------------------------------------------------------------ helpers for tests
    @classmethod
    def from_codes(cls, names, code_arrays, n_runs=None, mask_runs=None):
        """Pack per-sequence arrays of base codes (T=0 C=1 A=2 G=3) the way twoBitFromDnaSeq does."""
        sizes, chunks, offs, cur = [], [], [], 0
        for codes in code_arrays:
            codes = np.asarray(codes, dtype=np.uint8)
            n = len(codes)
            pad = (-n) % 4
            c = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
            by = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
            sizes.append(n); offs.append(cur); chunks.append(by.astype(np.uint8)); cur += len(by)
        packed = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
        return cls(names, sizes, packed, offs, n_runs, mask_runs)

    def codes(self, i):
        """Unpack sequence i to base codes (test helper only; the product never unpacks)."""
        n = int(self.sizes[i]); start = int(self.byte_offsets[i])
        by = self.packed[start:start + (n + 3) // 4]
        out = np.empty((len(by), 4), dtype=np.uint8)
        out[:, 0] = by >> 6; out[:, 1] = (by >> 4) & 3; out[:, 2] = (by >> 2) & 3; out[:, 3] = by & 3
        return out.reshape(-1)[:n]
