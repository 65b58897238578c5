#!/usr/bin/env python
"""Executed-instruction and DRAM figures of the dominant kernel from one `ncu --set full` capture -> profiles/kernel_counts.json
(read by bench.py for roofline.traffic and roofline.executed).

    tools/kernel_counts.py <file.ncu-rep> <workload> <combinations in the captured launch> <source note> [--no-traffic]

popc_warp_inst_per_comb : POPC warp instructions (source page) / combinations  (x 32 = POPC32 per combination)
alu_warp_inst_per_comb  : ALU-pipe warp instructions = sm__pipe_alu_cycles_active (pct of elapsed) x cycles x SMs x 2 / combinations
                          (the ALU pipe takes one warp instruction every two cycles per SM sub-partition: 16 lanes)
dram_bytes_per_launch   : dram__bytes_read.sum + dram__bytes_write.sum (only meaningful when the capture ran the full workload)
"""
import csv, json, os, subprocess, sys
rep, workload, combs, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
m = {k: (x, un) for k, un, x in zip(h, u, v)}
def num(k):
    x, un = m[k]
    f = float(x.replace(",", ""))
    return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(un, 1)
cycles = num("sm__cycles_elapsed.max")
alu = num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100 * cycles * 148 * 2
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
popc = 0
for r in rows[2:]:
    if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
        continue
    s = r[ix["Source"]].strip()
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    if op.startswith("POPC"):
        popc += int(r[ix["Instructions Executed"]])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counts.json")
db = json.load(open(path)) if os.path.exists(path) else {}
e = {"kernel": m["Kernel Name"][0], "captured_combinations": combs, "kernel_ms_under_ncu": num("gpu__time_duration.sum") / (1e6 if m["gpu__time_duration.sum"][1] == "ns" else (1e3 if m["gpu__time_duration.sum"][1] in ("us", "usecond") else 1)),
     "popc_warp_inst_per_comb": popc / combs, "alu_warp_inst_per_comb": alu / combs,
     "xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), "alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "issue_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"), "source": note}
if "--no-traffic" not in sys.argv:
    e["dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
db[workload] = e
json.dump(db, open(path, "w"), indent=1)
print(json.dumps(e, indent=1))
