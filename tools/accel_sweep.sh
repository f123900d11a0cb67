#!/bin/bash
# GPU-side: sweep the pruned warp-per-pair scan's clusters-per-trip (OGJK_ACCEL_K) and launch bounds.
# usage: tools/accel_sweep.sh "K LB" ["K LB" ...]
run() {
  for w in hulls1k_f32 hulls1k10k_f32; do
    python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', '$w', '%.1f Mpairs/s' % (j['value']/1e6), 'frac', round(j['roofline']['frac'],3))"
  done
}
for cfg in "$@"; do
  set -- $cfg
  make -B -s -C opengjk_b200/csrc EXTRA="-DOGJK_ACCEL_K=$1 -DOGJK_LB_TABLE_F32=$2 $3" 2>&1 | grep -i error
  grep -A2 "gjk_table_kernelIfLi32" opengjk_b200/csrc/ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo
  run "K=$1 lb=$2 $3"
done
