#!/bin/bash
# Copies the reference's OWN Python test suites and ctypes wrappers -- unmodified -- from the
# read-only reference tree into the git-ignored baseline/_ref/ (it travels to the GPU box with
# gpurun; /root/reference does not exist there).  tests/test_reference_suites.py runs them
# against this repo's reference-named flavour libraries.
#   baseline/fetch_ref.sh            (called by __graft_entry__.build() when the reference is mounted)
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF/scalar/examples/python_ctypes" ] || { echo "fetch_ref: $REF not mounted, nothing to do"; exit 0; }
mkdir -p $OUT/scalar/pyopengjk $OUT/scalar/test $OUT/gpu/pyopengjk_gpu
cp $REF/scalar/examples/python_ctypes/src/pyopengjk/*.py $OUT/scalar/pyopengjk/
cp $REF/scalar/examples/python_ctypes/test/test_min_distance.py $OUT/scalar/test/
cp $REF/scalar/examples/c/userP.dat $REF/scalar/examples/c/userQ.dat $OUT/scalar/
cp $REF/gpu/examples/python/src/pyopengjk_gpu/*.py $OUT/gpu/pyopengjk_gpu/
cp $REF/gpu/examples/python/test_examples.py $OUT/gpu/
( cd $OUT && find . -type f \( -name '*.py' -o -name '*.dat' \) | sort | xargs sha256sum ) > $OUT/MANIFEST.sha256
echo "fetch_ref: $(wc -l < $OUT/MANIFEST.sha256) reference files under baseline/_ref"
