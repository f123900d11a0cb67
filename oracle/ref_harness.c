/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin batch driver around the UNMODIFIED reference scalar implementation.
 * oracle/Makefile compiles this file together with
 * /root/reference/scalar/openGJK.c (read in place, never copied) into
 * oracle/_ref/libopengjk_ref_{f64,f32}.so.  It adapts the flat
 * [x y z x y z ...] pool layout used by the tests/bench to the reference's
 * gkFloat** row-pointer layout (scalar/include/openGJK/openGJK.h:68-78) and
 * calls the reference's public compute_minimum_distance() (openGJK.h:110)
 * once per pair.  The reference has no batched or threaded CPU entry point;
 * the OpenMP split over pairs below is the "all host cores" CPU baseline.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "openGJK/openGJK.h"

/* Row pointers for a whole pool; built outside any timed region. */
gkFloat** ref_rows_create(gkFloat* coords, int64_t nverts) {
  gkFloat** rows = (gkFloat**)malloc(sizeof(gkFloat*) * (size_t)(nverts > 0 ? nverts : 1));
  if (!rows) return NULL;
  for (int64_t i = 0; i < nverts; i++) rows[i] = coords + 3 * i;
  return rows;
}

void ref_rows_free(gkFloat** rows) { free(rows); }

gkFloat ref_gjk_single(int n1, gkFloat* c1, int n2, gkFloat* c2, gkSimplex* s) {
  gkFloat** r1 = ref_rows_create(c1, n1);
  gkFloat** r2 = ref_rows_create(c2, n2);
  gkPolytope a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  a.numpoints = n1;
  a.coord = r1;
  b.numpoints = n2;
  b.coord = r2;
  gkFloat d = compute_minimum_distance(a, b, s);
  free(r1);
  free(r2);
  return d;
}

void ref_gjk_batch(int64_t npairs, gkFloat** rows, const int64_t* off,
                   const int* cnt, const int* pa, const int* pb,
                   gkSimplex* simplices, gkFloat* distances, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t p = 0; p < npairs; p++) {
    gkPolytope a, b;
    memset(&a, 0, sizeof(a));
    memset(&b, 0, sizeof(b));
    a.numpoints = cnt[pa[p]];
    a.coord = rows + off[pa[p]];
    b.numpoints = cnt[pb[p]];
    b.coord = rows + off[pb[p]];
    gkSimplex* s = &simplices[p];
    memset(s, 0, sizeof(*s));
    distances[p] = compute_minimum_distance(a, b, s);
  }
}

void ref_sub(gkSimplex* s, gkFloat* v) {
  switch (s->nvrtx) {
    case 4: opengjk_test_S3D(s, v); break;
    case 3: opengjk_test_S2D(s, v); break;
    case 2: opengjk_test_S1D(s, v); break;
    default: break;
  }
}

int ref_sizeof_simplex(void) { return (int)sizeof(gkSimplex); }
int ref_sizeof_polytope(void) { return (int)sizeof(gkPolytope); }
