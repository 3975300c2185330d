#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + the hottest SASS lines.
usage: tools/ncu_summary.py report.ncu-rep [n_top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg.per_second']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for wname in want:
    if wname in hdr:
        i = hdr.index(wname)
        vals = [r[i] for r in data]
        try:
            if all(float(v) < 0.05 for v in vals) and 'stalled' in wname: continue
        except ValueError: pass
        print("%-88s %-10s %s" % (wname, units[i], vals))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
start = 1
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 1:]:
    if len(r) < len(hdr): break
    data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "sass lines", len(data))
keys = [h for h in hdr if h.startswith('stall_') and '(' not in h]
top = sorted(enumerate(data), key=lambda t: -int(t[1][ix['# Samples']]))[:ntop]
for n, r in sorted(top):
    st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]]) > 0.1 * int(r[ix['# Samples']])}
    print(n, r[ix['Source']].strip()[:58].ljust(58), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(10), st)

# where the executed instructions are: SASS lines grouped by how often they ran (a loop body's
# lines share one count), heaviest groups first
from collections import defaultdict
grp = defaultdict(lambda: [0, 0])
for r in data:
    try:
        e = int(r[ix['Instructions Executed']])
    except ValueError:
        continue
    if e <= 0: continue
    import math
    key = round(math.log(e, 1.08))          # counts within 8% of each other
    grp[key][0] += 1; grp[key][1] += e
total_exec = sum(v[1] for v in grp.values())
print("instructions executed (warp level) by execution-count group: total %.4g" % total_exec)
for key, (nl, ex) in sorted(grp.items(), key=lambda t: -t[1][1])[:14]:
    print("  ~%.4g executions x %4d SASS lines = %.4g (%.1f%%)" % (ex / nl, nl, ex, 100.0 * ex / total_exec))

# shared-memory / L1 wavefronts per instruction (bank conflicts show up as "excessive")
wcols = [h for h in hdr if 'avefronts' in h]
def num(r, h):
    try: return float(r[ix[h]])
    except (ValueError, KeyError): return 0.0
for h in wcols:
    print("total %-52s %.4g" % (h, sum(num(r, h) for r in data)))
exc = [h for h in wcols if 'Shared' in h and 'Excessive' in h]
if exc:
    h = exc[0]
    top = [t for t in sorted(enumerate(data), key=lambda t: -num(t[1], h))[:12] if num(t[1], h) > 0]
    print("instructions with excessive shared-memory wavefronts:", len(top))
    for n, r in sorted(top):
        print(n, r[ix['Source']].strip()[:58].ljust(58), "executed", r[ix['Instructions Executed']].rjust(10),
              " ".join("%s=%.4g" % (c.replace('L1 Wavefronts ', ''), num(r, c)) for c in wcols if 'Shared' in c))
