/*
 * epa_oracle.c -- TEST INFRASTRUCTURE ONLY (not part of the product path).
 *
 * Sequential CPU restatement of the reference's EPA (Expanding Polytope
 * Algorithm) for intersecting pairs.  The reference has no CPU EPA: its only
 * implementation is the warp-per-pair CUDA kernel compute_epa_kernel
 * (/root/reference/gpu/opengjk_gpu.cu:2022-2971 with helpers :1641-2018).
 * This file states that kernel's algorithm as one thread would run it: lane-0
 * blocks stay as written, and the two lane-partitioned reductions (support
 * argmax, closest-face argmin) are emulated lane for lane, because their
 * __shfl_down tree does not resolve exact ties to the lowest index.
 *
 * Parity status: PINNED against outputs of the reference itself.  No reference test pins
 * EPA numerically (gpu/examples/python/test_examples.py only checks loose ranges) and there is
 * no CPU EPA to link against, so the pin is tests/golden/epa_ref_cuda_{f32,f64}.npz: results of
 * the reference's own CUDA compute_gjk_epa / compute_epa (gpu/opengjk_gpu.cu compiled unmodified
 * on a B200 by tests/golden/make_epa_golden.py) on 7 000 seeded pairs per precision.  Against the
 * -fmad=false build this restatement reproduces distance and both witnesses bit for bit (one
 * fp64 pair of 8 000 differs in the last bits); against the default, FMA-contracting build the
 * depth agrees within 1e-5 / 1e-10 on every pair (tests/test_epa_golden.py).  The reference
 * fp32 kernel faults (CUDA error 700) on axis-aligned overlapping cubes; that workload is
 * pinned in fp64 only.  Our CUDA kernel is compared with this file bit for bit
 * (tests/test_gpu_epa.py) and with the same golden files directly.
 *
 * Choices where the reference leaves behaviour open (all documented in DESIGN.md):
 *   - a degenerate face (|n|^2 <= eps^2) becomes invalid for everybody (:1712-1716);
 *   - exits that write no output in the reference ("no closest face" :2605-2607,
 *     vertex cap :2533) keep the pre-filled defaults: the GJK witnesses, the
 *     GJK distance and the normal rule of the non-colliding path (:2061-2080).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef OGJK_ORACLE_F32
typedef float real;
#define REAL_EPS FLT_EPSILON
#define SUFFIX(name) name##_f32
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FMAX fmaxf
#define R_FMIN fminf
#else
typedef double real;
#define REAL_EPS DBL_EPSILON
#define SUFFIX(name) name##_f64
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMAX fmax
#define R_FMIN fmin
#endif

#define MAX_FACES 128              /* opengjk_gpu.cu:49 */
#define MAX_VERTS (MAX_FACES + 4)  /* :50 */
#define MAX_EDGES (MAX_FACES * 3)  /* :2765 */
#define EPA_MAX_ITER 64            /* :2528 */
#define BIG ((real)1e10f)          /* :1715, :1747 */

typedef struct {
  int nvrtx;
  real vrtx[4][3];
  int vrtx_idx[4][2];
  real witnesses[2][3];
} osimplex;

typedef struct {
  int v[3];
  int v_idx[3][2];
  real normal[3];
  real distance;
  int valid;
} face_t;

typedef struct {
  real vertices[MAX_VERTS][3];
  int vertex_indices[MAX_VERTS][2];
  int num_vertices;
  face_t faces[MAX_FACES];
  int max_face_index;
} poly_t;

typedef struct {
  int v1, v2;
  int i1[2], i2[2];
  int valid;
} edge_t;

static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline real nrm2(const real* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
static inline void cross3(const real* a, const real* b, real* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* :1659-1717 */
static void face_normal_distance(poly_t* p, int fi, const real* centroid) {
  face_t* f = &p->faces[fi];
  if (!f->valid) return;
  const real* v0 = p->vertices[f->v[0]];
  const real* v1 = p->vertices[f->v[1]];
  const real* v2 = p->vertices[f->v[2]];
  real e0[3], e1[3];
  for (int i = 0; i < 3; i++) {
    e0[i] = v1[i] - v0[i];
    e1[i] = v2[i] - v0[i];
  }
  cross3(e0, e1, f->normal);
  const real nsq = nrm2(f->normal);
  if (nsq > REAL_EPS * REAL_EPS) {
    const real nrm = R_SQRT(nsq);
    for (int i = 0; i < 3; i++) f->normal[i] /= nrm;
    real tc[3];
    for (int i = 0; i < 3; i++) tc[i] = centroid[i] - v0[i];
    if (dot3(f->normal, tc) > 0)
      for (int i = 0; i < 3; i++) f->normal[i] = -f->normal[i];
    f->distance = dot3(f->normal, v0);
    if (f->distance < 0) {
      for (int i = 0; i < 3; i++) f->normal[i] = -f->normal[i];
      f->distance = -f->distance;
    }
  } else {
    f->valid = 0;
    f->distance = BIG;
  }
}

/* The reference reduces per-lane partial results with a __shfl_down tree (offsets 16, 8, 4,
 * 2, 1; an out-of-range source returns the caller's own value) and strict comparisons
 * (:1780-1801, :2582-2593).  Lanes own ascending index chunks, but on an exact tie between
 * lanes the survivor is whatever that tree leaves in lane 0, which is NOT always the lowest
 * index (lane 0 may already hold lane 16's candidate when it meets lane 4's).  Overlapping
 * boxes tie all the time, so the tree is emulated literally: every lane updates from a
 * snapshot of the previous step. */
#define LANES 32
static int tree_argmax(real* val, int* idx) {
  for (int off = LANES / 2; off > 0; off >>= 1) {
    real nv[LANES];
    int ni[LANES];
    for (int l = 0; l < LANES; l++) {
      const int src = l + off < LANES ? l + off : l;
      nv[l] = val[l]; ni[l] = idx[l];
      if (val[src] > val[l]) { nv[l] = val[src]; ni[l] = idx[src]; }
    }
    memcpy(val, nv, sizeof(nv));
    memcpy(idx, ni, sizeof(ni));
  }
  return idx[0];
}
static int tree_closest_face(real* dist, int* face) {
  for (int off = LANES / 2; off > 0; off >>= 1) {
    real nd[LANES];
    int nf[LANES];
    for (int l = 0; l < LANES; l++) {
      const int src = l + off < LANES ? l + off : l;
      nd[l] = dist[l]; nf[l] = face[l];
      if (face[src] >= 0 && (face[l] < 0 || dist[src] < dist[l])) { nd[l] = dist[src]; nf[l] = face[src]; }
    }
    memcpy(dist, nd, sizeof(nd));
    memcpy(face, nf, sizeof(nf));
  }
  return face[0];
}

/* :1735-1825: lane l scans indices [l*ppt, (l+1)*ppt) of both bodies, ppt = ceil(max(n1,n2)/32),
 * first-index argmax from -1e10 inside the lane, then the tree above. */
static void support_epa(int n1, const real* c1, int n2, const real* c2, const real* dir, real* out, int* out_idx) {
  const int max_points = n1 > n2 ? n1 : n2;
  const int ppt = (max_points + LANES - 1) / LANES;
  const real nd[3] = {-dir[0], -dir[1], -dir[2]};
  real m1[LANES], m2[LANES];
  int i1[LANES], i2[LANES];
  for (int l = 0; l < LANES; l++) {
    const int start = l * ppt;
    const int end = start + ppt < max_points ? start + ppt : max_points;
    m1[l] = m2[l] = -BIG;
    i1[l] = i2[l] = -1;
    for (int i = start; i < n1 && i < end; i++) {
      const real s = dot3(c1 + 3 * (size_t)i, dir);
      if (s > m1[l]) { m1[l] = s; i1[l] = i; }
    }
    for (int i = start; i < n2 && i < end; i++) {
      const real s = dot3(c2 + 3 * (size_t)i, nd);
      if (s > m2[l]) { m2[l] = s; i2[l] = i; }
    }
  }
  int b1 = tree_argmax(m1, i1), b2 = tree_argmax(m2, i2);
  if (b1 < 0) b1 = 0; /* unreachable for finite inputs below 1e10 */
  if (b2 < 0) b2 = 0;
  for (int i = 0; i < 3; i++) out[i] = c1[3 * (size_t)b1 + i] - c2[3 * (size_t)b2 + i];
  out_idx[0] = b1;
  out_idx[1] = b2;
}

/* :2568-2597: lane l takes faces [l*fpt, (l+1)*fpt), fpt = ceil(max_face_index/32), first-index
 * argmin of the non-negative distances inside the lane, then the tree above. */
static int closest_face(const poly_t* p, real* closest_distance) {
  const int fpt = (p->max_face_index + LANES - 1) / LANES;
  real d[LANES];
  int f[LANES];
  for (int l = 0; l < LANES; l++) {
    const int start = l * fpt;
    const int end = start + fpt < p->max_face_index ? start + fpt : p->max_face_index;
    d[l] = BIG;
    f[l] = -1;
    for (int i = start; i < end; ++i) {
      if (!p->faces[i].valid) continue;
      if (p->faces[i].distance >= 0.0f && p->faces[i].distance < d[l]) { d[l] = p->faces[i].distance; f[l] = i; }
    }
  }
  const int best = tree_closest_face(d, f);
  *closest_distance = d[0];
  return best;
}

/* test probe: one support query (tests/test_epa_golden.py pins the tie rule with it) */
void SUFFIX(oracle_epa_support)(int n1, const real* c1, int n2, const real* c2, const real* dir, real* out, int* out_idx) {
  support_epa(n1, c1, n2, c2, dir, out, out_idx);
}

/* contact normal from witness1 to witness2 (:2061-2080 and the copies at :2146-2164 ...) */
static void normal_from_witnesses(const real* w1, const real* w2, real* nrm_out) {
  real d[3];
  real norm = 0.0f;
  for (int i = 0; i < 3; i++) {
    d[i] = w2[i] - w1[i];
    norm += d[i] * d[i];
  }
  norm = R_SQRT(norm);
  if (norm > REAL_EPS) {
    for (int i = 0; i < 3; i++) nrm_out[i] = d[i] / norm;
  } else {
    nrm_out[0] = 1.0f;
    nrm_out[1] = 0.0f;
    nrm_out[2] = 0.0f;
  }
}

/* "no new support point": depth 0, support vertices as witnesses (:2136-2166 and copies) */
static void finish_no_progress(const real* c1, const real* c2, const int* idx, real* distance, real* w1, real* w2,
                               real* normal) {
  *distance = 0.0f;
  for (int c = 0; c < 3; c++) {
    w1[c] = c1[3 * (size_t)idx[0] + c];
    w2[c] = c2[3 * (size_t)idx[1] + c];
  }
  normal_from_witnesses(w1, w2, normal);
}

static int is_new_point(const osimplex* s, const real* nv) {
  const real eps_sq = REAL_EPS * REAL_EPS;
  for (int k = 0; k < s->nvrtx; ++k) {
    const real dx = nv[0] - s->vrtx[k][0];
    const real dy = nv[1] - s->vrtx[k][1];
    const real dz = nv[2] - s->vrtx[k][2];
    const real d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < eps_sq) return 0;
  }
  return 1;
}

static void push_vertex(osimplex* s, const real* nv, const int* idx) {
  const int k = s->nvrtx;
  for (int c = 0; c < 3; ++c) s->vrtx[k][c] = nv[c];
  s->vrtx_idx[k][0] = idx[0];
  s->vrtx_idx[k][1] = idx[1];
  s->nvrtx = k + 1;
}

/* :1828-1937 */
static void init_polytope(poly_t* p, const osimplex* s, real* centroid) {
  for (int i = 0; i < MAX_FACES; ++i) {
    p->faces[i].valid = 0;
    p->faces[i].distance = BIG;
  }
  p->num_vertices = 4;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 3; j++) p->vertices[i][j] = s->vrtx[i][j];
    p->vertex_indices[i][0] = s->vrtx_idx[i][0];
    p->vertex_indices[i][1] = s->vrtx_idx[i][1];
  }
  centroid[0] = centroid[1] = centroid[2] = 0.0f;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 3; j++) centroid[j] += p->vertices[i][j];
  for (int j = 0; j < 3; j++) centroid[j] /= 4.0f;

  static const int tet[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
  for (int f = 0; f < 4; f++) {
    for (int v = 0; v < 3; v++) {
      const int vi = tet[f][v];
      p->faces[f].v[v] = vi;
      p->faces[f].v_idx[v][0] = s->vrtx_idx[vi][0];
      p->faces[f].v_idx[v][1] = s->vrtx_idx[vi][1];
    }
    p->faces[f].valid = 1;
  }
  for (int f = 0; f < 4; f++) {
    face_t* fc = &p->faces[f];
    const real* v0 = p->vertices[fc->v[0]];
    const real* v1 = p->vertices[fc->v[1]];
    const real* v2 = p->vertices[fc->v[2]];
    real e0[3], e1[3], n[3], tc[3];
    for (int i = 0; i < 3; i++) {
      e0[i] = v1[i] - v0[i];
      e1[i] = v2[i] - v0[i];
    }
    cross3(e0, e1, n);
    for (int i = 0; i < 3; i++) tc[i] = centroid[i] - v0[i];
    if (dot3(n, tc) > 0) {
      int t = fc->v[1]; fc->v[1] = fc->v[2]; fc->v[2] = t;
      int t0 = fc->v_idx[1][0], t1 = fc->v_idx[1][1];
      fc->v_idx[1][0] = fc->v_idx[2][0]; fc->v_idx[1][1] = fc->v_idx[2][1];
      fc->v_idx[2][0] = t0; fc->v_idx[2][1] = t1;
    }
  }
  p->max_face_index = 4;
}

/* :1940-2011 */
static void barycentric_origin(const real* v0, const real* v1, const real* v2, real* a0, real* a1, real* a2) {
  real e0[3], e1[3], v0n[3];
  for (int i = 0; i < 3; i++) {
    e0[i] = v1[i] - v0[i];
    e1[i] = v2[i] - v0[i];
    v0n[i] = -v0[i];
  }
  const real d00 = dot3(e0, e0), d01 = dot3(e0, e1), d11 = dot3(e1, e1);
  const real d20 = dot3(v0n, e0), d21 = dot3(v0n, e1);
  const real denom = d00 * d11 - d01 * d01;
  if (R_FABS(denom) < REAL_EPS) {
    *a0 = *a1 = *a2 = (real)1.0 / (real)3.0;
    return;
  }
  const real inv = (real)1.0 / denom;
  const real u = (d11 * d20 - d01 * d21) * inv;
  const real v = (d00 * d21 - d01 * d20) * inv;
  const real w = (real)1.0 - u - v;
  if (w < 0) {
    real e12[3], v1n[3];
    for (int i = 0; i < 3; i++) {
      e12[i] = v2[i] - v1[i];
      v1n[i] = -v1[i];
    }
    real t = dot3(v1n, e12) / dot3(e12, e12);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    *a0 = 0; *a1 = (real)1.0 - t; *a2 = t;
  } else if (u < 0) {
    real t = dot3(v0n, e1) / dot3(e1, e1);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    *a0 = (real)1.0 - t; *a1 = 0; *a2 = t;
  } else if (v < 0) {
    real t = dot3(v0n, e0) / dot3(e0, e0);
    t = R_FMAX((real)0.0, R_FMIN((real)1.0, t));
    *a0 = (real)1.0 - t; *a1 = t; *a2 = 0;
  } else {
    *a0 = w; *a1 = u; *a2 = v;
  }
}

/* result from the closest face (:2626-2660, repeated at :2683-2714 and :2936-2969) */
static void finish_on_face(const poly_t* p, int fi, real closest_distance, const real* c1, const real* c2,
                           real* distance, real* w1, real* w2, real* normal) {
  const face_t* f = &p->faces[fi];
  real a0, a1, a2;
  barycentric_origin(p->vertices[f->v[0]], p->vertices[f->v[1]], p->vertices[f->v[2]], &a0, &a1, &a2);
  for (int i = 0; i < 3; i++) {
    w1[i] = c1[3 * (size_t)f->v_idx[0][0] + i] * a0 + c1[3 * (size_t)f->v_idx[1][0] + i] * a1 +
            c1[3 * (size_t)f->v_idx[2][0] + i] * a2;
    w2[i] = c2[3 * (size_t)f->v_idx[0][1] + i] * a0 + c2[3 * (size_t)f->v_idx[1][1] + i] * a1 +
            c2[3 * (size_t)f->v_idx[2][1] + i] * a2;
  }
  *distance = -closest_distance;
  for (int i = 0; i < 3; i++) normal[i] = f->normal[i];
}

/* compute_epa_kernel for one pair.  `s_in` is the post-witness GJK simplex (not
 * modified: the reference grows a private copy, :2049), `*distance` the GJK
 * distance on entry and the signed result on exit.  Returns the number of EPA
 * iterations (0 when the main loop never ran). */
int SUFFIX(oracle_epa)(int n1, const real* c1, int n2, const real* c2, const osimplex* s_in, real* distance, real* w1,
                       real* w2, real* normal) {
  osimplex s = *s_in;
  /* defaults for exits that write nothing in the reference */
  for (int i = 0; i < 3; i++) {
    w1[i] = s.witnesses[0][i];
    w2[i] = s.witnesses[1][i];
  }
  normal_from_witnesses(w1, w2, normal);
  if (*distance > REAL_EPS) return 0; /* :2053-2083 separated: GJK result stands */

  real dir[3], nv[3];
  int nvi[2];
  if (s.nvrtx != 4) {
    if (s.nvrtx == 1) { /* :2088-2185 */
      for (int c = 0; c < 3; c++) dir[c] = s.vrtx[0][c];
      support_epa(n1, c1, n2, c2, dir, nv, nvi);
      if (is_new_point(&s, nv)) {
        push_vertex(&s, nv, nvi);
      } else {
        finish_no_progress(c1, c2, nvi, distance, w1, w2, normal);
        return 0;
      }
    }
    if (s.nvrtx == 2) { /* :2186-2306 */
      real edge[3];
      for (int c = 0; c < 3; ++c) edge[c] = s.vrtx[1][c] - s.vrtx[0][c];
      real axis[3] = {1.0f, 0.0f, 0.0f};
      const real en = R_SQRT(edge[0] * edge[0] + edge[1] * edge[1] + edge[2] * edge[2]);
      if (en > REAL_EPS && R_FABS(edge[0]) > 0.9f * en) {
        axis[0] = 0.0f; axis[1] = 1.0f; axis[2] = 0.0f;
      }
      cross3(edge, axis, dir);
      if (nrm2(dir) < REAL_EPS) {
        axis[0] = 0.0f; axis[1] = 0.0f; axis[2] = 1.0f;
        cross3(edge, axis, dir);
      }
      support_epa(n1, c1, n2, c2, dir, nv, nvi);
      if (is_new_point(&s, nv)) {
        push_vertex(&s, nv, nvi);
      } else {
        finish_no_progress(c1, c2, nvi, distance, w1, w2, normal);
        return 0;
      }
    }
    if (s.nvrtx == 3) { /* :2307-2459 */
      real e0[3], e1[3];
      for (int c = 0; c < 3; ++c) {
        e0[c] = s.vrtx[1][c] - s.vrtx[0][c];
        e1[c] = s.vrtx[2][c] - s.vrtx[0][c];
      }
      cross3(e0, e1, dir);
      support_epa(n1, c1, n2, c2, dir, nv, nvi);
      if (is_new_point(&s, nv)) {
        push_vertex(&s, nv, nvi);
      } else {
        for (int c = 0; c < 3; c++) dir[c] = -dir[c];
        support_epa(n1, c1, n2, c2, dir, nv, nvi);
        if (is_new_point(&s, nv)) {
          push_vertex(&s, nv, nvi);
        } else {
          finish_no_progress(c1, c2, nvi, distance, w1, w2, normal);
          return 0;
        }
      }
    }
    if (s.nvrtx != 4) { /* :2461-2486 (nvrtx outside 1..4) */
      *distance = 0.0f;
      return 0;
    }
  }

  poly_t poly;
  real centroid[3];
  init_polytope(&poly, &s, centroid);

  const real tolerance = (real)REAL_EPS * 1e2f; /* eps_tot22 :42, :2529 */
  int iteration = 0;
  while (iteration < EPA_MAX_ITER && poly.num_vertices < MAX_VERTS - 1) {
    iteration++;
    for (int i = 0; i < poly.max_face_index; ++i)
      if (poly.faces[i].valid) face_normal_distance(&poly, i, centroid);

    real closest_distance = BIG;
    const int closest = closest_face(&poly, &closest_distance);
    if (closest < 0) break; /* :2605-2607 nothing written */

    const real d[3] = {poly.faces[closest].normal[0], poly.faces[closest].normal[1], poly.faces[closest].normal[2]};
    support_epa(n1, c1, n2, c2, d, nv, nvi);

    const real dist_to_new = d[0] * nv[0] + d[1] * nv[1] + d[2] * nv[2];
    const real improvement = dist_to_new - closest_distance;
    if (improvement < tolerance) { /* :2624 converged */
      finish_on_face(&poly, closest, closest_distance, c1, c2, distance, w1, w2, normal);
      break;
    }

    int dup = 0; /* :2665-2677 */
    {
      const real eps_sq = REAL_EPS * REAL_EPS;
      for (int i = 0; i < poly.num_vertices; i++) {
        const real dx = nv[0] - poly.vertices[i][0];
        const real dy = nv[1] - poly.vertices[i][1];
        const real dz = nv[2] - poly.vertices[i][2];
        if (dx * dx + dy * dy + dz * dz < eps_sq) { dup = 1; break; }
      }
    }
    if (dup) {
      finish_on_face(&poly, closest, closest_distance, c1, c2, distance, w1, w2, normal);
      break;
    }

    /* add vertex, incremental centroid (:2719-2736) */
    const int new_id = poly.num_vertices;
    for (int i = 0; i < 3; i++) poly.vertices[new_id][i] = nv[i];
    poly.vertex_indices[new_id][0] = nvi[0];
    poly.vertex_indices[new_id][1] = nvi[1];
    poly.num_vertices++;
    {
      const real n = (real)poly.num_vertices;
      for (int i = 0; i < 3; i++) centroid[i] = centroid[i] * (n - 1.0f) / n + nv[i] / n;
    }

    /* visible faces (:1720-1732, :2758-2762) */
    for (int i = 0; i < poly.max_face_index; i++) {
      face_t* f = &poly.faces[i];
      if (!f->valid) continue;
      const real* v0 = poly.vertices[f->v[0]];
      const real df[3] = {nv[0] - v0[0], nv[1] - v0[1], nv[2] - v0[2]};
      if (dot3(f->normal, df) > REAL_EPS) f->valid = 0;
    }

    /* edges of every invalid face below max_face_index (:2765-2810) */
    edge_t edges[MAX_EDGES];
    int ne = 0;
    for (int f = 0; f < poly.max_face_index; f++) {
      const face_t* fc = &poly.faces[f];
      if (fc->valid) continue;
      for (int k = 0; k < 3; k++) {
        if (ne < MAX_EDGES) {
          const int a = k, b = (k + 1) % 3;
          edges[ne].v1 = fc->v[a];
          edges[ne].v2 = fc->v[b];
          edges[ne].i1[0] = fc->v_idx[a][0]; edges[ne].i1[1] = fc->v_idx[a][1];
          edges[ne].i2[0] = fc->v_idx[b][0]; edges[ne].i2[1] = fc->v_idx[b][1];
          edges[ne].valid = 1;
          ne++;
        }
      }
    }
    /* cancel shared edges (:2813-2826); the inner loop keeps going after a hit */
    for (int i = 0; i < ne; i++) {
      if (!edges[i].valid) continue;
      for (int j = i + 1; j < ne; j++) {
        if (!edges[j].valid) continue;
        if ((edges[i].v1 == edges[j].v1 && edges[i].v2 == edges[j].v2) ||
            (edges[i].v1 == edges[j].v2 && edges[i].v2 == edges[j].v1)) {
          edges[i].valid = 0;
          edges[j].valid = 0;
        }
      }
    }
    /* one new face per surviving edge, first free slot (:2829-2897) */
    for (int i = 0; i < ne; i++) {
      if (!edges[i].valid) continue;
      int slot = -1;
      for (int j = 0; j < MAX_FACES; j++)
        if (!poly.faces[j].valid) { slot = j; break; }
      if (slot < 0) break;
      face_t* nf = &poly.faces[slot];
      nf->v[0] = edges[i].v1; nf->v[1] = edges[i].v2; nf->v[2] = new_id;
      nf->v_idx[0][0] = edges[i].i1[0]; nf->v_idx[0][1] = edges[i].i1[1];
      nf->v_idx[1][0] = edges[i].i2[0]; nf->v_idx[1][1] = edges[i].i2[1];
      nf->v_idx[2][0] = nvi[0]; nf->v_idx[2][1] = nvi[1];
      nf->valid = 1;
      const real* f0 = poly.vertices[nf->v[0]];
      const real* f1 = poly.vertices[nf->v[1]];
      const real* f2 = poly.vertices[nf->v[2]];
      real fe0[3], fe1[3], fn[3], tc[3];
      for (int c = 0; c < 3; c++) {
        fe0[c] = f1[c] - f0[c];
        fe1[c] = f2[c] - f0[c];
      }
      cross3(fe0, fe1, fn);
      for (int c = 0; c < 3; c++) tc[c] = centroid[c] - f0[c];
      if (dot3(fn, tc) > 0) {
        int t = nf->v[1]; nf->v[1] = nf->v[2]; nf->v[2] = t;
        int t0 = nf->v_idx[1][0], t1 = nf->v_idx[1][1];
        nf->v_idx[1][0] = nf->v_idx[2][0]; nf->v_idx[1][1] = nf->v_idx[2][1];
        nf->v_idx[2][0] = t0; nf->v_idx[2][1] = t1;
      }
      if (slot >= poly.max_face_index) poly.max_face_index = slot + 1;
    }
  }

  if (iteration >= EPA_MAX_ITER) { /* :2919-2970 */
    for (int i = 0; i < poly.max_face_index; ++i)
      if (poly.faces[i].valid) face_normal_distance(&poly, i, centroid);
    int closest = -1;
    real closest_distance = BIG;
    for (int i = 0; i < poly.max_face_index; ++i) {
      if (!poly.faces[i].valid) continue;
      if (poly.faces[i].distance >= 0.0f && poly.faces[i].distance < closest_distance) {
        closest_distance = poly.faces[i].distance;
        closest = i;
      }
    }
    if (closest >= 0 && poly.faces[closest].valid)
      finish_on_face(&poly, closest, closest_distance, c1, c2, distance, w1, w2, normal);
  }
  return iteration;
}

real SUFFIX(oracle_gjk)(int n1, const real* c1, int n2, const real* c2, osimplex* s, int* iters);

/* GJK followed by EPA on every pair (compute_gjk_epa, opengjk_gpu.cu:3055-3099);
 * with run_gjk == 0 it is compute_epa (:3007-3053): simplices/distances are inputs. */
void SUFFIX(oracle_gjk_epa_batch)(int64_t npairs, const real* coords, const int64_t* off, const int* cnt, const int* pa,
                                  const int* pb, osimplex* simplices, real* distances, real* w1, real* w2, real* normals,
                                  int* epa_iters, int run_gjk, int nthreads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t p = 0; p < npairs; p++) {
    const int a = pa[p], b = pb[p];
    const real* c1 = coords + 3 * off[a];
    const real* c2 = coords + 3 * off[b];
    if (run_gjk) {
      memset(&simplices[p], 0, sizeof(osimplex));
      distances[p] = SUFFIX(oracle_gjk)(cnt[a], c1, cnt[b], c2, &simplices[p], 0);
    }
    const int it = SUFFIX(oracle_epa)(cnt[a], c1, cnt[b], c2, &simplices[p], &distances[p], w1 + 3 * p, w2 + 3 * p,
                                      normals + 3 * p);
    if (epa_iters) epa_iters[p] = it;
  }
}
