#!/usr/bin/env python
"""One-line digest of a bench.py JSON line read from stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(f"n={d.get('n_gpus')} value={d.get('value', 0) / 1e9:.2f}G ms/step={d.get('ms_per_step', 0):.3f} e2e={d.get('e2e', {}).get('value', 0) / 1e9:.2f}G "
          f"kernel_ms={r.get('kernel_ms', 0):.3f} frac={r.get('frac', 0):.3f} share={r.get('kernel_share_of_step', 0):.3f} clocks={d.get('clocks', {}).get('sm_mhz')}")
