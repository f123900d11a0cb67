/*
 * openGJK/openGJK.h -- drop-in header of the reference-named flavour libraries
 * (libopengjk_ce.so = fp64, libopengjk_ce_f32.so = fp32 with -DUSE_32BITS).
 *
 * Declares the same public surface as the reference's scalar header
 * (/root/reference/scalar/include/openGJK/openGJK.h:54-122): the gkFloat switch, the
 * gkPolytope / gkSimplex structs (same field names, order and layout: fp64 48 B /
 * 184 B, fp32 32 B / 108 B) and compute_minimum_distance() taking both polytopes BY
 * VALUE with `coord` an array of row pointers.  The implementation runs the query as
 * a batch of one on the GPU through include/opengjk_b200.h (ogjk_single_rows_*).
 */
#ifndef OPENGJK_H__
#define OPENGJK_H__

#include <float.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef USE_32BITS
#define gkFloat float
#define gkEpsilon FLT_EPSILON
#else
#define gkFloat double
#define gkEpsilon DBL_EPSILON
#endif

#ifndef OPENGJK_EXPORT
#define OPENGJK_EXPORT __attribute__((visibility("default")))
#endif

typedef struct gkPolytope_ {
  int numpoints;   /* vertices of the body */
  gkFloat s[3];    /* scratch: last support point (ignored on input) */
  int s_idx;       /* scratch: its index */
  gkFloat** coord; /* numpoints row pointers, 3 coordinates each; owned by the caller */
} gkPolytope;

typedef struct gkSimplex_ {
  int nvrtx;               /* vertices of the final simplex, 1..4 */
  gkFloat vrtx[4][3];      /* Minkowski-difference points */
  int vrtx_idx[4][2];      /* their source vertex indices [vertex][body] */
  gkFloat witnesses[2][3]; /* closest points on body 1 / body 2 */
} gkSimplex;

/* Minimum distance between two convex point clouds (0 when they touch or intersect).
 * Returns NaN if no CUDA device is available: there is no CPU fallback. */
OPENGJK_EXPORT gkFloat compute_minimum_distance(gkPolytope bd1, gkPolytope bd2, gkSimplex* s);

/* Sub-solver test hooks (reference: openGJK.h:120-122): s->nvrtx = 2 / 3 / 4 with the
 * newest vertex in the highest slot; v receives the point of the simplex closest to the origin. */
OPENGJK_EXPORT void opengjk_test_S1D(gkSimplex* s, gkFloat* v);
OPENGJK_EXPORT void opengjk_test_S2D(gkSimplex* s, gkFloat* v);
OPENGJK_EXPORT void opengjk_test_S3D(gkSimplex* s, gkFloat* v);

/* Entry point for C# callers (reference: scalar/openGJK.c:1154-1220 under CS_MONO_BUILD):
 * component-major coordinates, inCoords[i * n + j] = component i of vertex j. */
OPENGJK_EXPORT gkFloat csFunction(int nCoordsA, gkFloat* inCoordsA, int nCoordsB, gkFloat* inCoordsB);

#ifdef __cplusplus
}
#endif
#endif /* OPENGJK_H__ */
