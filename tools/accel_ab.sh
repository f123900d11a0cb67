#!/bin/bash
# GPU-side A/B: resident-store throughput of the large-hull workloads without cluster tables
# and with them at 8 / 16 / 32 lanes per pair.
one() {
  python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', '$2', '%.1f Mpairs/s' % (j['value']/1e6), 'frac', round(j['roofline']['frac'],3), j['roofline']['kernel'])"
}
for w in ${WORKLOADS:-hulls1k_f32 hulls1k10k_f32}; do
  OGJK_ACCEL_MIN=0 one $w "no tables"
  for l in 8 16 32; do OGJK_TABLE_LANES=$l one $w "tables lanes=$l"; done
done
