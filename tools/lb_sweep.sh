#!/bin/bash
# GPU-side: rebuild with different __launch_bounds__ minimum-blocks settings and sweep.
for v in "5 4 4" "6 4 4" "4 5 4" "4 3 4" "4 4 5" "4 4 3"; do
  set -- $v
  rm -f opengjk_b200/libopengjk_b200.so
  make -s -C opengjk_b200/csrc EXTRA="-DOGJK_LB_GROUP_F32=$1 -DOGJK_LB_GROUP_F64=$2 -DOGJK_LB_SMALL_F32=$3" > /dev/null 2>&1
  echo "== group_f32=$1 group_f64=$2 small_f32=$3"
  grep -E "Compiling|registers|spill" opengjk_b200/csrc/ptxas.log | paste - - - | grep "_group_kernelIfLi2E\|_group_kernelIdLi2E\|gjk_small_kernelIf\|_group_kernelIfLi32" | grep -v persist | sed -E 's/.*ogjk[0-9]+([a-z_]+)I([fd])(Li([0-9]+))?.* ([0-9]+) bytes spill stores.*Used ([0-9]+) registers.*/\1 \2 G=\4 spill=\5 regs=\6/' | tr '\n' ';'; echo
  python tools/tune_classes.py hulls64_f64 hulls64_f32 boxes_f32 hulls1k_f32 2>&1 | grep "hulls64_f64    lanes=  2\|hulls64_f32    lanes=  2\|boxes_f32      lanes=  0\|hulls1k_f32    lanes= 32"
done
