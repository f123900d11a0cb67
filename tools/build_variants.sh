#!/bin/bash
# Builds A/B variants of the product library HERE (nvcc cross-compiles without a GPU) so that
# a single gpurun call can sweep them: opengjk_b200/variants/libogjk_<name>.so
#   tools/build_variants.sh name1:"-DFLAG=1 -DOTHER=2" name2:"..." ...
# Built artefacts are git-ignored (*.so) but travel with the gpurun snapshot.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/opengjk_b200/variants
mkdir -p $OUT
NVFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-fvisibility=hidden,-pthread -Xptxas -v"
pids=()
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc $NVFLAGS $flags -shared $ROOT/opengjk_b200/csrc/abi.cu -o $OUT/libogjk_$name.so 2> $OUT/ptxas_$name.log \
      && echo "built $name [$flags]" || { echo "FAILED $name"; tail -5 $OUT/ptxas_$name.log; } ) &
  pids+=($!)
  if (( ${#pids[@]} >= 4 )); then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
done
wait
