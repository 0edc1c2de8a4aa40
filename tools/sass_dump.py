#!/usr/bin/env python
"""Print the SASS of one kernel between two addresses: tools/sass_dump.py lib.so kernel-substring lo hi (hex)"""
import re, subprocess, sys
lib, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        on = kern in m.group(1)
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m and lo <= int(m.group(1), 16) <= hi:
        print("%05x  %s" % (int(m.group(1), 16), m.group(2)))
