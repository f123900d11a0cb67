/* compat_scalar.c -- reference-named scalar flavour (include/openGJK/openGJK.h).
 * Plain C host code over the C ABI; one lazily created context guarded by a mutex
 * (the reference function is re-entrant, so concurrent callers are serialised here). */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "openGJK/openGJK.h"
#include "opengjk_b200.h"

#ifdef USE_32BITS
#define SINGLE ogjk_single_rows_f32
#define SUBALG ogjk_test_subalgorithm_f32
typedef ogjk_simplex_f32 simplex_t;
#else
#define SINGLE ogjk_single_rows_f64
#define SUBALG ogjk_test_subalgorithm_f64
typedef ogjk_simplex_f64 simplex_t;
#endif

static ogjk_ctx* g_ctx;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

/* Called with g_lock held (every entry point takes it first), so creation happens once. */
static ogjk_ctx* context(void) {
  static int tried;
  if (!tried) {
    tried = 1;
    if (ogjk_ctx_create(0, &g_ctx) != OGJK_OK) {
      g_ctx = 0;
      fprintf(stderr, "opengjk_ce (opengjk_b200): %s\n", ogjk_last_error());
    }
  }
  return g_ctx;
}

gkFloat compute_minimum_distance(gkPolytope bd1, gkPolytope bd2, gkSimplex* s) {
  gkFloat d = (gkFloat)NAN;
  pthread_mutex_lock(&g_lock);
  ogjk_ctx* c = context();
  if (c && s) SINGLE(c, bd1.numpoints, bd1.coord, bd2.numpoints, bd2.coord, (simplex_t*)s, &d);
  pthread_mutex_unlock(&g_lock);
  return d;
}

static void sub(gkSimplex* s, gkFloat* v) {
  pthread_mutex_lock(&g_lock);
  ogjk_ctx* c = context();
  if (c && s && v) SUBALG(c, (simplex_t*)s, v);
  pthread_mutex_unlock(&g_lock);
}
void opengjk_test_S1D(gkSimplex* s, gkFloat* v) { sub(s, v); }
void opengjk_test_S2D(gkSimplex* s, gkFloat* v) { sub(s, v); }
void opengjk_test_S3D(gkSimplex* s, gkFloat* v) { sub(s, v); }

/* The reference's entry point for C# callers (scalar/openGJK.c:1154-1220, built there under
 * CS_MONO_BUILD; P/Invoke in scalar/examples/cs/main.cs): coordinates arrive component-major,
 * inCoords[i * n + j] = component i of vertex j; returns the distance. */
OPENGJK_EXPORT gkFloat csFunction(int nCoordsA, gkFloat* inCoordsA, int nCoordsB, gkFloat* inCoordsB) {
  if (nCoordsA < 1 || nCoordsB < 1 || !inCoordsA || !inCoordsB) return (gkFloat)NAN;
  const int n = nCoordsA + nCoordsB;
  gkFloat* rows = (gkFloat*)malloc((size_t)n * 3 * sizeof(gkFloat));
  gkFloat** ptr = (gkFloat**)malloc((size_t)n * sizeof(gkFloat*));
  gkFloat d = (gkFloat)NAN;
  if (rows && ptr) {
    for (int j = 0; j < nCoordsA; ++j)
      for (int i = 0; i < 3; ++i) rows[3 * j + i] = inCoordsA[i * nCoordsA + j];
    for (int j = 0; j < nCoordsB; ++j)
      for (int i = 0; i < 3; ++i) rows[3 * (nCoordsA + j) + i] = inCoordsB[i * nCoordsB + j];
    for (int j = 0; j < n; ++j) ptr[j] = rows + 3 * j;
    gkPolytope a, b;
    gkSimplex s;
    a.numpoints = nCoordsA; a.coord = ptr; a.s_idx = 0; a.s[0] = a.s[1] = a.s[2] = 0;
    b.numpoints = nCoordsB; b.coord = ptr + nCoordsA; b.s_idx = 0; b.s[0] = b.s[1] = b.s[2] = 0;
    s.nvrtx = 0;
    d = compute_minimum_distance(a, b, &s);
  }
  free(rows);
  free(ptr);
  return d;
}
