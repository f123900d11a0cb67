for L in 8 4 2; do
  make -B -s -C opengjk_b200/csrc EXTRA=-DOGJK_AABB_LANES=$L 2>&1 | grep -i error
  echo "lanes per polytope = $L"; python tools/broadphase_bench.py 2>&1 | grep aabb_kernel | cut -c1-170
done
