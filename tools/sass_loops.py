#!/usr/bin/env python
"""Static loop sizes of one kernel from its SASS (no GPU needed): every backward branch closes a loop
[target, branch]; prints the instruction count and opcode-class mix of each loop body.

  tools/sass_loops.py <lib.so> <mangled-kernel-substring>

The scoring kernel's dynamic instruction count is dominated by two loops (the sub-tile loop, run 4x per tile, and
the item-round loop), so their static sizes are the budget to watch while editing."""
import collections, re, subprocess, sys

ALU = {"LOP3", "IADD3", "SHF", "SEL", "ISETP", "PRMT", "LEA", "VIADD", "VIMNMX", "VIADDMNMX", "VIMNMX3", "PLOP3", "MOV", "IABS", "BMSK", "SGXT", "FLO", "P2R", "R2P", "CS2R"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
XU = {"POPC", "MUFU", "I2F", "F2I", "BREV"}
MEM = {"LDG", "STG", "LDS", "STS", "SHFL", "ATOMS", "ATOMG", "RED", "LDL", "STL", "LD", "ST", "LDC", "LDCU", "UBLKCP", "SYNCS", "CCTL", "REDUX", "VOTE", "S2R", "S2UR"}
CTL = {"BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "NOP", "YIELD", "BAR"}


def main():
    lib, kern = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ins, on = [], False
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            on = kern in m.group(1)
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            addr = int(m.group(1), 16)
            text = m.group(2)
            mm = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", text)
            op = mm.group(2) if mm else "?"
            ins.append((addr, op, text))
    print("kernel instructions:", len(ins))
    loops = []
    for i, (addr, op, text) in enumerate(ins):
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)\s*$", text)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
    for lo, hi in sorted(set(loops)):
        body = [x for x in ins if lo <= x[0] <= hi]
        c = collections.Counter()
        for _, op, _ in body:
            c["alu" if op in ALU else "fma" if op in FMA else "xu" if op in XU else "mem" if op in MEM else "ctl" if op in CTL else "uni" if op[0] == "U" or op == "R2UR" else "oth"] += 1
        ops = collections.Counter(op for _, op, _ in body)
        print("loop %05x-%05x: %4d instr  alu %3d fma %3d xu %2d mem %3d ctl %3d uni %3d oth %3d | %s" % (
            lo, hi, len(body), c["alu"], c["fma"], c["xu"], c["mem"], c["ctl"], c["uni"], c["oth"],
            " ".join("%s:%d" % kv for kv in ops.most_common(9))))


if __name__ == "__main__":
    main()
