#!/usr/bin/env python
"""Summarise an ncu report of scoreChunksKernel: headline metrics + instructions / stall samples per
source-line range.  usage: tools/ncu_phase.py report.ncu-rep [marker=lo-hi ...]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print('%-66s %-10s %s' % (w, rows[1][i], rows[2][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--print-kernel-base", "function"],
                     capture_output=True, text=True).stdout
def split_row(line, ncols=None):
    # ncu does not escape quotes inside the source column: split it off by position
    line = line.rstrip('\n')
    if not line.startswith('"'):
        return next(csv.reader([line]), [])
    body = line[1:-1] if line.endswith('"') else line[1:]
    if ncols is None:
        return body.split('","')
    first, rest = body.split('","', 1)
    tail = rest.rsplit('","', ncols - 2)
    return [first] + tail
raw_lines = src.split('\n')
hdr_cols = None
rows = []
for ln in raw_lines:
    if ln.startswith('"Line No"'):
        hdr_cols = len(split_row(ln))
        rows.append(split_row(ln))
    elif hdr_cols and ln.startswith('"'):
        rows.append(split_row(ln, hdr_cols))
    else:
        rows.append(next(csv.reader([ln]), []) if ln else [])
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
ci, si = h.index('Instructions Executed'), h.index('# Samples')
end = hi[1] - 3 if len(hi) > 1 else len(rows)
lines = {}
for r in rows[hi[0] + 1:end]:
    if len(r) <= ci or not r[0].strip().isdigit():
        continue
    lines[int(r[0])] = (int(r[ci]), int(r[si]), r[1])
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print('total warp instructions (sampled kernel):', tot, ' samples:', ts)
text = open('/root/repo/genomealignmenttools_b200/csrc/gat_kernels.cuh').read().split('\n')
def find(s):
    for i, l in enumerate(text):
        if s in l:
            return i + 1
    raise KeyError(s)
marks = [('tuple helpers', 'struct Tup {', 'struct GapView'), ('gap cost', '// ------------------------------------------------------------------ gap cost', '// ------------------------------------------------------------------ base windows'),
         ('window loads', '// ------------------------------------------------------------------ base windows', '// does [g0, g0+len) touch'),
         ('mayTouchN', '// does [g0, g0+len) touch', '// ------------------------------------------------------------------ 32 base pairs'),
         ('scoreWindow', '// ------------------------------------------------------------------ 32 base pairs', '// ------------------------------------------------------------------ chunk index'),
         ('phase3 reduce fn', '// Phase 3 of scoreChunksKernel for one warp', 'scoreChunksKernel(const __grid_constant__'),
         ('staging+phase0', 'scoreChunksKernel(const __grid_constant__', '// ---- phase 1'),
         ('phase1 descriptors', '// ---- phase 1', '// ---- phase 2'), ('phase2 items', '// ---- phase 2', '// ---- phase 3'),
         ('phase3', '// ---- phase 3', '// ------------------------------------------------------------------ cross-chunk fix-up')]
for name, a, b in marks:
    try:
        la, lb = find(a.split('\n')[0]), find(b.split('\n')[0])
    except KeyError:
        continue
    n = sum(v[0] for l, v in lines.items() if la <= l < lb); s = sum(v[1] for l, v in lines.items() if la <= l < lb)
    print('%-22s L%3d-%3d inst %6.2f%% (%6.1fM)  samples %6.2f%%' % (name, la, lb, 100 * n / tot, n / 1e6, 100 * s / ts))
names = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
agg = {n: 0 for n in names}
for r in rows[hi[0] + 1:end]:
    if len(r) <= ci or not r[0].strip().isdigit():
        continue
    for n in names:
        try:
            agg[n] += int(r[h.index(n)])
        except ValueError:
            pass
t = sum(agg.values())
print({k: round(100 * v / t, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]})
if '--top' in sys.argv:
    for n, s, l, srcl in sorted(((v[0], v[1], l, v[2]) for l, v in lines.items()), reverse=True)[:40]:
        print('%6.2f%% %6d smp  L%-4d %s' % (100 * n / tot, s, l, srcl[:110]))
