#!/usr/bin/env python
"""Static SASS instruction counts per source-line range of one kernel (no GPU needed).

  tools/sass_lines.py <lib.so> <kernel-name-substring> [file:lo-hi=label ...]

Extracts the cubin, disassembles with line info and counts instructions (and their pipe class) per label.
Used to budget the instruction count of a loop body before spending GPU time on it."""
import collections
import re
import subprocess
import sys
import tempfile
import os

ALU = {"LOP3", "IADD3", "SHF", "SEL", "ISETP", "PRMT", "LEA", "VIADD", "VIMNMX", "VIADDMNMX", "VIMNMX3", "PLOP3", "MOV", "IABS", "FSETP", "BMSK", "SGXT", "FLO", "IADD", "LOP", "P2R", "R2P", "CS2R", "FMNMX"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
XU = {"POPC", "MUFU", "I2F", "F2I", "BREV", "FLO"}
LSU = {"LDG", "STG", "LDS", "STS", "SHFL", "ATOMS", "ATOMG", "RED", "LDL", "STL", "LD", "ST", "LDC", "ATOM", "LDSM", "UBLKCP", "SYNCS", "CCTL", "MATCH", "REDUX", "VOTE"}


def main():
    lib, kern = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        m = re.match(r"(.+):(\d+)-(\d+)=(.+)", a)
        ranges.append((m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)))
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
        out = ""
        for c in cubins:
            out += subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
    cur_file, cur_line, in_kernel = None, 0, False
    per = collections.defaultdict(lambda: collections.Counter())
    lines = collections.Counter()
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            in_kernel = kern in m.group(1)
            continue
        if ln.startswith(".section") or re.match(r"\s*\.section", ln):
            in_kernel = False
        if not in_kernel:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
        if not m:
            continue
        op = m.group(1)
        cls = "alu" if op in ALU else "fma" if op in FMA else "xu" if op in XU else "lsu" if op in LSU else "uni" if op.startswith("U") or op in ("R2UR", "S2UR") else "oth"
        label = "other"
        for f, lo, hi, lab in ranges:
            if cur_file == f and lo <= cur_line <= hi:
                label = lab
                break
        per[label][cls] += 1
        per[label]["all"] += 1
        lines[(cur_file, cur_line)] += 1
    for lab, c in per.items():
        print("%-22s all %5d  alu %4d fma %4d xu %3d lsu %4d uni %4d oth %4d" % (lab, c["all"], c["alu"], c["fma"], c["xu"], c["lsu"], c["uni"], c["oth"]))
    if "--lines" in sys.argv or True:
        top = sorted(lines.items(), key=lambda kv: -kv[1])[:int(os.environ.get("TOP", "0"))]
        for (f, l), n in top:
            print("  %s:%d  %d" % (f, l, n))


if __name__ == "__main__":
    main()
