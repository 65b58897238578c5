#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page): key metrics, instruction mix, top stall lines."""
import csv, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_elapsed.max']
want += [k for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
for v in rows[2:]:
    print("-----")
    for i, k in enumerate(h):
        if k in want: print(f"{k} [{u[i]}] = {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# first kernel only
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
byop = collections.Counter(); tot = 0; recs = []
for n, r in enumerate(rows[2:]):
    if len(r) < 10 or r[0] == 'Kernel Name' or r[0] == 'Address': 
        if r and r[0] == 'Kernel Name' and n > 0: break
        continue
    s = r[ix['Source']].strip(); ex = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']])
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    byop[op] += ex; tot += ex
    st = {k: int(r[ix[k]]) for k in hdr if k.startswith('stall_') and '(' not in k}
    recs.append((sm, n, s, ex, st))
print("===== instruction mix (warp-level executed), total", tot)
for op, c in byop.most_common(36): print(f"{op:30s} {c:13d} {100*c/tot:5.1f}%")
ts = sum(x[0] for x in recs)
print("===== top stall lines, total samples", ts)
for sm, n, s, ex, st in sorted(recs, reverse=True)[:30]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{sm:7d} {100*sm/ts:4.1f}% line {n:5d} ex={ex:11d} {s[:52]:52s} {top}")
