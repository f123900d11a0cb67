#!/bin/bash
# GPU-side A/B of compile-time flags: tools/flag_ab.sh "<EXTRA flags>" workload [workload ...]
FLAGS=$1; shift
for f in "" "$FLAGS"; do
  make -B -s -C opengjk_b200/csrc EXTRA="$f" 2>&1 | grep -i error
  for w in "$@"; do
    for rep in 1 2; do
    python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('[$f]', '$w', '%.1f Mpairs/s' % (j['value']/1e6), 'frac', round(j['roofline']['frac'],3))"
    done
  done
done
