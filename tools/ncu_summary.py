#!/usr/bin/env python
"""Summarise an ncu report of scoreTilesKernel (taken with --set full --import-source on): headline metrics, stall reasons,
opcode mix, shared-memory wavefronts per opcode, and the dynamic instruction / stall-sample share of contiguous code regions
(regions = runs of SASS instructions with about the same execution count: per tile, per sub-tile, per item round ...).

usage: tools/ncu_summary.py report.ncu-rep > profiles/<name>_summary.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum', 'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed_op_tma_ld.sum']
print("== headline metrics (%s)" % rows[2][hdr.index('Kernel Name')] if 'Kernel Name' in hdr else "== headline metrics")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print('%-70s %-14s %s' % (w, units[i], vals[i]))
print("\n== warps stalled per issue slot, by reason (> 0.05)")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and float(vals[i] or 0) > 0.05:
        print('%-28s %s' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vals[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
recs = []
byop, samp, wav = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(h):
        continue
    n = int(r[ix['Instructions Executed']] or 0)
    sass = r[ix['Source']]
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?', sass)
    op = m.group(2) if m else '?'
    byop[op] += n
    samp[op] += int(r[ix['# Samples']] or 0)
    wav[op + (m.group(3) or '') if m else '?'] += int(r[ix['L1 Wavefronts Shared']] or 0)
    recs.append((int(r[ix['Address']], 16), n, int(r[ix['# Samples']] or 0), sass, int(r[ix['stall_long_sb']] or 0),
                 int(r[ix['stall_short_sb']] or 0), int(r[ix['stall_wait']] or 0)))
tot = sum(byop.values()); ts = sum(samp.values())
print("\n== opcode mix (warp instructions executed; share; stall samples)")
for op, n in byop.most_common(24):
    print('%-12s %11d %5.1f%%  %d' % (op, n, 100.0 * n / tot, samp[op]))
print("\n== shared-memory wavefronts by opcode")
for op, n in wav.most_common(8):
    if n:
        print('%-14s %d' % (op, n))
recs.sort()
base = recs[0][0]
regs, cur = [], [recs[0]]
for x in recs[1:]:
    a, b = cur[-1][1], x[1]
    if max(a, b) > 1.3 * min(a, b) + 1000:
        regs.append(cur); cur = [x]
    else:
        cur.append(x)
regs.append(cur)
print("\n== code regions (SASS offset range, instructions, mean executions, share of dynamic instructions, share of stall samples, long / short scoreboard / wait samples)")
for g in regs:
    dyn = sum(x[1] for x in g)
    if dyn > 0.006 * tot or sum(x[2] for x in g) > 0.015 * ts:
        print('%05x-%05x n=%4d avg=%8d dyn=%5.1f%% smp=%5.1f%% long=%5d short=%5d wait=%5d   %s' % (
            g[0][0] - base, g[-1][0] - base, len(g), dyn // len(g), 100.0 * dyn / tot, 100.0 * sum(x[2] for x in g) / ts,
            sum(x[4] for x in g), sum(x[5] for x in g), sum(x[6] for x in g), g[0][3].strip()[:40]))
print("\ntotal warp instructions %d, stall samples %d" % (tot, ts))
