#!/bin/bash
# array lists on/off on c2, alternating (same box)
for V in 1 0 1 0; do
  echo "LIST_SCAN=$V"; HPGV_LIST_SCAN=$V timeout 60 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_short.py
done
