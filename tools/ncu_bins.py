import csv, subprocess, sys, collections
rep = sys.argv[1]; binw = int(sys.argv[2]) if len(sys.argv) > 2 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
bins = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot = 0
for n, r in enumerate(rows[2:]):
    if len(r) < 10 or r[0] in ('Kernel Name', 'Address'):
        continue
    s = r[ix['Source']].strip(); ex = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']])
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    b = bins[n // binw]; b[0] += sm; b[1] += ex; b[2][op.split('.')[0]] += ex
    for k in hdr:
        if k.startswith('stall_') and '(' not in k: b[3][k] += int(r[ix[k]])
    tot += sm
for k in sorted(bins):
    b = bins[k]
    print(f"lines {k*binw:5d}-{(k+1)*binw-1:5d} samples {b[0]:7d} {100*b[0]/tot:5.1f}% ex {b[1]:11d}  ops {dict(b[2].most_common(4))} stalls {dict(b[3].most_common(3))}")
