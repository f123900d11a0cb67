#!/bin/bash
# GPU-side: multi-pass scheduling sweep (OGJK_PASSES=passes:iters1:iters2) on the small-hull workloads.
for w in ${WORKLOADS:-hulls64_f64 hulls64_f32}; do
  for cfg in 1:4:3 2:3:3 2:4:3 2:5:3 3:3:2 3:3:3 3:4:2 3:4:3; do
    OGJK_PASSES=$cfg python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$w passes=$cfg', '%.1f Mpairs/s' % (j['value']/1e6), 'launches', j['gpu_launches'])"
  done
done
