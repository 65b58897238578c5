#!/usr/bin/env python
"""Static SASS instruction mix of one kernel of a .so / cubin: python tools/sass_mix.py <lib> <substring of the mangled name> [--dump]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur: funcs[cur].append(m.group(2).strip())
for name, ins in funcs.items():
    if pat not in name: continue
    mix = collections.Counter()
    for i in ins:
        t = i.split()
        op = t[1] if t[0].startswith('@') else t[0]
        mix[op.split('.')[0] if '--full' not in sys.argv else op] += 1
    print(name, len(ins), "instructions")
    print("  " + ", ".join(f"{k} {v}" for k, v in mix.most_common(24)))
    if "--dump" in sys.argv:
        for n, i in enumerate(ins): print(f"{n:5d}  {i}")
