/*
 * gjk_oracle.c -- TEST INFRASTRUCTURE ONLY (not part of the product path).
 *
 * Plain-C CPU restatement of the reference GJK distance query
 * (/root/reference/scalar/openGJK.c), written against a flat
 * [x0 y0 z0 x1 y1 z1 ...] vertex array instead of the reference's gkFloat**.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The CUDA product path never
 * links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the reference's own sub-solver known-answer vectors
 *     (scalar/tests/subalg_{1,2,3}simplex_test.c, HFFtest_test.c) via
 *     tests/golden/subsolver_kat.json,
 *   - the reference's end-to-end analytic cases
 *     (scalar/examples/python_ctypes/test/test_min_distance.py),
 *   - bit-for-bit agreement with the unmodified reference compiled into
 *     oracle/_ref/libopengjk_ref_{f64,f32}.so on seeded random batches, and
 *   - tests/golden/gjk_golden_*.npz produced by that reference build.
 *
 * Every arithmetic expression keeps the reference's operand order and is
 * compiled with -ffp-contract=off so results are bit-identical to the
 * reference built for baseline x86-64 (no FMA).
 *
 * Build: see oracle/Makefile.  Compiled twice: default fp64 (suffix _f64) and
 * -DOGJK_ORACLE_F32 (suffix _f32), mirroring the reference's USE_32BITS switch
 * (scalar/include/openGJK/openGJK.h:54-60).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef OGJK_ORACLE_F32
typedef float real;
#define REAL_EPS FLT_EPSILON
#define SUFFIX(name) name##_f32
#else
typedef double real;
#define REAL_EPS DBL_EPSILON
#define SUFFIX(name) name##_f64
#endif

/* openGJK.c:44-56 */
#define MAX_ITER 25
#define EPS_REL ((real)(REAL_EPS * 1e4))
#define EPS_ABS ((real)(REAL_EPS * 1e2))

/* Same memory layout as the reference gkSimplex (openGJK.h:84-94). */
typedef struct {
  int nvrtx;
  real vrtx[4][3];
  int vrtx_idx[4][2];
  real witnesses[2][3];
} osimplex;

/* One Minkowski-difference point with the indices of its source vertices. */
typedef struct {
  real p[3];
  int id[2];
} mvert;

/* openGJK.c:58-59 -- left-to-right evaluation. */
static inline real dot3(const real* a, const real* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline real nrm2(const real* a) {
  return a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
}

/* openGJK.c:181-187 */
static inline real det3(const real* p, const real* q, const real* r) {
  return p[0] * ((q[1] * r[2]) - (r[1] * q[2])) -
         p[1] * (q[0] * r[2] - r[0] * q[2]) +
         p[2] * (q[0] * r[1] - r[0] * q[1]);
}

/* openGJK.c:189-195 */
static inline void cross3(const real* a, const real* b, real* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* openGJK.c:197-210 */
static inline void proj_line(const real* p, const real* q, real* v) {
  real pq[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
  const real t = dot3(p, pq) / dot3(pq, pq);
  for (int i = 0; i < 3; i++) v[i] = p[i] - pq[i] * t;
}

/* openGJK.c:212-231 */
static inline void proj_plane(const real* p, const real* q, const real* r,
                              real* v) {
  real pq[3], pr[3], n[3];
  for (int i = 0; i < 3; i++) pq[i] = p[i] - q[i];
  for (int i = 0; i < 3; i++) pr[i] = p[i] - r[i];
  cross3(pq, pr, n);
  const real t = dot3(n, p) / dot3(n, n);
  for (int i = 0; i < 3; i++) v[i] = n[i] * t;
}

/* openGJK.c:233-244 */
static inline int hf1(const real* p, const real* q) {
  real acc = 0;
  for (int i = 0; i < 3; i++) acc += (p[i] * p[i] - p[i] * q[i]);
  return acc > 0;
}

/* openGJK.c:246-262 */
static inline int hf2(const real* p, const real* q, const real* r) {
  real pq[3], pr[3], m[3], n[3];
  for (int i = 0; i < 3; i++) pq[i] = q[i] - p[i];
  for (int i = 0; i < 3; i++) pr[i] = r[i] - p[i];
  cross3(pq, pr, m);
  cross3(pq, m, n);
  return dot3(p, n) < 0;
}

/* openGJK.c:264-277 */
static inline int hf3(const real* p, const real* q, const real* r) {
  real pq[3], pr[3], n[3];
  for (int i = 0; i < 3; i++) pq[i] = q[i] - p[i];
  for (int i = 0; i < 3; i++) pr[i] = r[i] - p[i];
  cross3(pq, pr, n);
  return dot3(p, n) <= 0;
}

/* ---- simplex slot helpers (replace the reference's macros :61-179) ------ */
static inline void slot_get(const osimplex* s, int k, mvert* m) {
  for (int t = 0; t < 3; t++) m->p[t] = s->vrtx[k][t];
  m->id[0] = s->vrtx_idx[k][0];
  m->id[1] = s->vrtx_idx[k][1];
}
static inline void slot_put(osimplex* s, int k, const mvert* m) {
  for (int t = 0; t < 3; t++) s->vrtx[k][t] = m->p[t];
  s->vrtx_idx[k][0] = m->id[0];
  s->vrtx_idx[k][1] = m->id[1];
}
static inline void slot_copy(osimplex* s, int dst, int src) {
  for (int t = 0; t < 3; t++) s->vrtx[dst][t] = s->vrtx[src][t];
  s->vrtx_idx[dst][0] = s->vrtx_idx[src][0];
  s->vrtx_idx[dst][1] = s->vrtx_idx[src][1];
}
/* select_1xy (:67-92): keep newest (slot 3 -> 2), then a in slot 1, b in 0. */
static inline void keep_tri(osimplex* s, const mvert* a, const mvert* b) {
  s->nvrtx = 3;
  slot_copy(s, 2, 3);
  slot_put(s, 1, a);
  slot_put(s, 0, b);
}
/* select_1x (:94-113): newest (slot 3 -> 1), a in slot 0. */
static inline void keep_seg(osimplex* s, const mvert* a) {
  s->nvrtx = 2;
  slot_copy(s, 1, 3);
  slot_put(s, 0, a);
}

/* openGJK.c:279-290 (+ S1Dregion1 :132-141) */
static void sub_1d(osimplex* s, real* v) {
  const real* a = s->vrtx[1];
  const real* b = s->vrtx[0];
  if (hf1(a, b)) {
    proj_line(a, b, v);
  } else {
    for (int t = 0; t < 3; t++) v[t] = s->vrtx[1][t];
    s->nvrtx = 1;
    slot_copy(s, 0, 1);
  }
}

/* openGJK.c:292-335 (+ region macros :143-168) */
static void sub_2d(osimplex* s, real* v) {
  const real* a = s->vrtx[2];
  const real* b = s->vrtx[1];
  const real* c = s->vrtx[0];
  const int ab = hf1(a, b);
  const int ac = hf1(a, c);

  if (ab) {
    if (!hf2(a, b, c)) {
      if (ac) {
        if (!hf2(a, c, b)) {
          proj_plane(a, b, c, v);
        } else {
          proj_line(a, c, v);
          s->nvrtx = 2; /* S2Dregion13 */
          slot_copy(s, 1, 2);
        }
      } else {
        proj_plane(a, b, c, v);
      }
    } else {
      proj_line(a, b, v);
      s->nvrtx = 2; /* S2Dregion12 */
      slot_copy(s, 0, 2);
    }
  } else if (ac) {
    if (!hf2(a, c, b)) {
      proj_plane(a, b, c, v);
    } else {
      proj_line(a, c, v);
      s->nvrtx = 2; /* S2Dregion13 */
      slot_copy(s, 1, 2);
    }
  } else {
    for (int t = 0; t < 3; t++) v[t] = s->vrtx[2][t]; /* S2Dregion1 */
    s->nvrtx = 1;
    slot_copy(s, 0, 2);
  }
}

/* openGJK.c:337-605.  The reference's untouched-state branches are kept:
 * a plane-sum of 0 with all three hf1 tests positive leaves s and v alone
 * (:534-601), as does any fall-through in the plane-sum 1 / dotTotal 2 arm. */
static void sub_3d(osimplex* s, real* v) {
  mvert s1, s2, s3, si, sj, sk;
  real s4[3], e12[3], e13[3], e14[3];
  int i, j, k;

  slot_get(s, 3, &s1);
  slot_get(s, 2, &s2);
  slot_get(s, 1, &s3);
  for (int t = 0; t < 3; t++) s4[t] = s->vrtx[0][t];
  for (int t = 0; t < 3; t++) {
    e12[t] = s2.p[t] - s->vrtx[3][t];
    e13[t] = s3.p[t] - s->vrtx[3][t];
    e14[t] = s4[t] - s->vrtx[3][t];
  }

  int h1[3];
  h1[2] = hf1(s1.p, s2.p);
  h1[1] = hf1(s1.p, s3.p);
  h1[0] = hf1(s1.p, s4);
  const int line3 = h1[1];
  const int line4 = h1[0];
  const int ndot = h1[2] + line3 + line4;

  if (ndot == 0) { /* S3Dregion1 :170-179 */
    for (int t = 0; t < 3; t++) v[t] = s1.p[t];
    s->nvrtx = 1;
    slot_put(s, 0, &s1);
    return;
  }

  const real d134 = det3(e13, e14, e12);
  const int sss = (d134 <= 0);

  int pl2 = hf3(s1.p, s3.p, s4) - sss;
  pl2 = pl2 * pl2;
  int pl3 = hf3(s1.p, s4, s2.p) - sss;
  pl3 = pl3 * pl3;
  int pl4 = hf3(s1.p, s2.p, s3.p) - sss;
  pl4 = pl4 * pl4;

  switch (pl2 + pl3 + pl4) {
    case 3: /* origin inside: S3Dregion1234 :61-65 */
      v[0] = 0;
      v[1] = 0;
      v[2] = 0;
      s->nvrtx = 4;
      break;

    case 2: /* :382-413 drop the vertex opposite the one failing face */
      s->nvrtx = 3;
      if (!pl2) {
        slot_copy(s, 2, 3);
      } else if (!pl3) {
        slot_put(s, 1, &s2);
        slot_copy(s, 2, 3);
      } else if (!pl4) {
        slot_put(s, 0, &s3);
        slot_put(s, 1, &s2);
        slot_copy(s, 2, 3);
      }
      sub_2d(s, v);
      break;

    case 1: /* :414-532 */
      s->nvrtx = 3;
      if (pl2) {
        k = 2; i = 1; j = 0;
      } else if (pl3) {
        k = 1; i = 0; j = 2;
      } else {
        k = 0; i = 2; j = 1;
      }
      slot_get(s, i, &si);
      slot_get(s, j, &sj);
      slot_get(s, k, &sk);

      if (ndot == 1) {
        if (h1[k]) {
          if (!hf2(s1.p, sk.p, si.p)) {
            keep_tri(s, &si, &sk);
            proj_plane(s1.p, si.p, sk.p, v);
          } else if (!hf2(s1.p, sk.p, sj.p)) {
            keep_tri(s, &sj, &sk);
            proj_plane(s1.p, sj.p, sk.p, v);
          } else {
            keep_seg(s, &sk);
            proj_line(s1.p, sk.p, v);
          }
        } else if (h1[i]) {
          if (!hf2(s1.p, si.p, sk.p)) {
            keep_tri(s, &si, &sk);
            proj_plane(s1.p, si.p, sk.p, v);
          } else {
            keep_seg(s, &si);
            proj_line(s1.p, si.p, v);
          }
        } else {
          if (!hf2(s1.p, sj.p, sk.p)) {
            keep_tri(s, &sj, &sk);
            proj_plane(s1.p, sj.p, sk.p, v);
          } else {
            keep_seg(s, &sj);
            proj_line(s1.p, sj.p, v);
          }
        }
      } else if (ndot == 2) {
        if (h1[i]) {
          if (!hf2(s1.p, sk.p, si.p)) {
            if (!hf2(s1.p, si.p, sk.p)) {
              keep_tri(s, &si, &sk);
              proj_plane(s1.p, si.p, sk.p, v);
            } else {
              keep_seg(s, &sk);
              proj_line(s1.p, sk.p, v);
            }
          } else {
            if (!hf2(s1.p, sk.p, sj.p)) {
              keep_tri(s, &sj, &sk);
              proj_plane(s1.p, sj.p, sk.p, v);
            } else {
              keep_seg(s, &sk);
              proj_line(s1.p, sk.p, v);
            }
          }
        } else if (h1[j]) {
          if (!hf2(s1.p, sk.p, sj.p)) {
            if (!hf2(s1.p, sj.p, sk.p)) {
              keep_tri(s, &sj, &sk);
              proj_plane(s1.p, sj.p, sk.p, v);
            } else {
              keep_seg(s, &sj);
              proj_line(s1.p, sj.p, v);
            }
          } else {
            if (!hf2(s1.p, sk.p, si.p)) {
              keep_tri(s, &si, &sk);
              proj_plane(s1.p, si.p, sk.p, v);
            } else {
              keep_seg(s, &sk);
              proj_line(s1.p, sk.p, v);
            }
          }
        }
      } else if (ndot == 3) {
        const int f_ik = hf2(s1.p, si.p, sk.p);
        const int f_jk = hf2(s1.p, sj.p, sk.p);
        const int f_ki = hf2(s1.p, sk.p, si.p);
        const int f_kj = hf2(s1.p, sk.p, sj.p);
        /* reference prints a warning when both f_ki and f_kj are 0 (:508) and
         * then continues into the final else arm; no print here. */
        if (f_ki == 1 && f_kj == 1) {
          keep_seg(s, &sk);
          proj_line(s1.p, sk.p, v);
        } else if (f_ki) {
          if (f_jk) {
            keep_seg(s, &sj);
            proj_line(s1.p, sj.p, v);
          } else {
            keep_tri(s, &sj, &sk);
            proj_plane(s1.p, sk.p, sj.p, v);
          }
        } else {
          if (f_ik) {
            keep_seg(s, &si);
            proj_line(s1.p, si.p, v);
          } else {
            keep_tri(s, &si, &sk);
            proj_plane(s1.p, sk.p, si.p, v);
          }
        }
      }
      break;

    case 0: /* :534-601 */
      if (ndot == 1) {
        if (line3) {
          k = 2; i = 1; j = 0;
        } else if (line4) {
          k = 1; i = 0; j = 2;
        } else {
          k = 0; i = 2; j = 1;
        }
        slot_get(s, i, &si);
        slot_get(s, j, &sj);
        slot_get(s, k, &sk);

        if (!hf2(s1.p, si.p, sj.p)) {
          keep_tri(s, &si, &sj);
          proj_plane(s1.p, si.p, sj.p, v);
        } else if (!hf2(s1.p, si.p, sk.p)) {
          keep_tri(s, &si, &sk);
          proj_plane(s1.p, si.p, sk.p, v);
        } else {
          keep_seg(s, &si);
          proj_line(s1.p, si.p, v);
        }
      } else if (ndot == 2) {
        s->nvrtx = 3;
        if (!line3) {
          k = 2; i = 1; j = 0;
        } else if (!line4) {
          k = 1; i = 0; j = 2;
        } else {
          k = 0; i = 2; j = 1;
        }
        slot_get(s, i, &si);
        slot_get(s, j, &sj);
        slot_get(s, k, &sk);

        if (!hf2(s1.p, sj.p, sk.p)) {
          if (!hf2(s1.p, sk.p, sj.p)) {
            keep_tri(s, &sj, &sk);
            proj_plane(s1.p, sj.p, sk.p, v);
          } else if (!hf2(s1.p, sk.p, si.p)) {
            keep_tri(s, &si, &sk);
            proj_plane(s1.p, sk.p, si.p, v);
          } else {
            keep_seg(s, &sk);
            proj_line(s1.p, sk.p, v);
          }
        } else if (!hf2(s1.p, sj.p, si.p)) {
          keep_tri(s, &si, &sj);
          proj_plane(s1.p, si.p, sj.p, v);
        } else {
          keep_seg(s, &sj);
          proj_line(s1.p, sj.p, v);
        }
      }
      break;
    default:
      break;
  }
}

/* ---- witnesses: openGJK.c:648-939, including the missing early returns --- */
#define VTX(c, i) ((c) + 3 * (size_t)(i))

static void wit_0d(const real* c1, const real* c2, osimplex* s) {
  const real* a = VTX(c1, s->vrtx_idx[0][0]);
  const real* b = VTX(c2, s->vrtx_idx[0][1]);
  for (int t = 0; t < 3; t++) {
    s->witnesses[0][t] = a[t];
    s->witnesses[1][t] = b[t];
  }
}

static void wit_1d(const real* c1, const real* c2, osimplex* s) {
  real pq[3], po[3];
  const real* p = s->vrtx[0];
  const real* q = s->vrtx[1];
  for (int t = 0; t < 3; t++) {
    pq[t] = q[t] - p[t];
    po[t] = -p[t];
  }
  const real det = dot3(pq, pq);
  if (det == 0.0) wit_0d(c1, c2, s); /* no return in the reference (:673-676) */

  const real a1 = dot3(pq, po) / det;
  const real a0 = 1.0 - a1;

  const real* w00 = VTX(c1, s->vrtx_idx[0][0]);
  const real* w01 = VTX(c2, s->vrtx_idx[0][1]);
  const real* w10 = VTX(c1, s->vrtx_idx[1][0]);
  const real* w11 = VTX(c2, s->vrtx_idx[1][1]);
  for (int t = 0; t < 3; t++) {
    s->witnesses[0][t] = w00[t] * a0 + w10[t] * a1;
    s->witnesses[1][t] = w01[t] * a0 + w11[t] * a1;
  }
}

static void wit_2d(const real* c1, const real* c2, osimplex* s) {
  real pq[3], pr[3], po[3];
  const real* p = s->vrtx[0];
  const real* q = s->vrtx[1];
  const real* r = s->vrtx[2];
  for (int t = 0; t < 3; t++) {
    pq[t] = q[t] - p[t];
    pr[t] = r[t] - p[t];
    po[t] = -p[t];
  }
  const real T00 = dot3(pq, pq);
  const real T01 = dot3(pq, pr);
  const real T11 = dot3(pr, pr);
  const real det = T00 * T11 - T01 * T01;
  if (det == 0.0) wit_1d(c1, c2, s); /* falls through (:735-738) */

  const real b0 = dot3(pq, po);
  const real b1 = dot3(pr, po);
  const real I00 = T11 / det;
  const real I01 = -T01 / det;
  const real I11 = T00 / det;
  const real a1 = I00 * b0 + I01 * b1;
  const real a2 = I01 * b0 + I11 * b1;
  const real a0 = 1.0 - a1 - a2; /* evaluated in double, rounded once */

  if (a0 < REAL_EPS) {
    s->nvrtx = 2;
    slot_copy(s, 0, 2);
    wit_1d(c1, c2, s);
  } else if (a1 < REAL_EPS) {
    s->nvrtx = 2;
    slot_copy(s, 1, 2);
    wit_1d(c1, c2, s);
  } else if (a2 < REAL_EPS) {
    s->nvrtx = 2;
    wit_1d(c1, c2, s);
  }

  const real* w00 = VTX(c1, s->vrtx_idx[0][0]);
  const real* w01 = VTX(c2, s->vrtx_idx[0][1]);
  const real* w10 = VTX(c1, s->vrtx_idx[1][0]);
  const real* w11 = VTX(c2, s->vrtx_idx[1][1]);
  const real* w20 = VTX(c1, s->vrtx_idx[2][0]);
  const real* w21 = VTX(c2, s->vrtx_idx[2][1]);
  for (int t = 0; t < 3; t++) {
    s->witnesses[0][t] = w00[t] * a0 + w10[t] * a1 + w20[t] * a2;
    s->witnesses[1][t] = w01[t] * a0 + w11[t] * a1 + w21[t] * a2;
  }
}

static void wit_3d(const real* c1, const real* c2, osimplex* s) {
  real pq[3], pr[3], ps[3], po[3];
  const real* p = s->vrtx[0];
  const real* q = s->vrtx[1];
  const real* r = s->vrtx[2];
  const real* u = s->vrtx[3];
  for (int t = 0; t < 3; t++) {
    pq[t] = q[t] - p[t];
    pr[t] = r[t] - p[t];
    ps[t] = u[t] - p[t];
    po[t] = -p[t];
  }
  const real T00 = dot3(pq, pq);
  const real T01 = dot3(pq, pr);
  const real T02 = dot3(pq, ps);
  const real T11 = dot3(pr, pr);
  const real T12 = dot3(pr, ps);
  const real T22 = dot3(ps, ps);
  const real det00 = T11 * T22 - T12 * T12;
  const real det01 = T01 * T22 - T02 * T12;
  const real det02 = T01 * T12 - T02 * T11;
  const real det = T00 * det00 - T01 * det01 + T02 * det02;
  if (det == 0.0) wit_2d(c1, c2, s); /* falls through (:845-848) */

  const real b0 = dot3(pq, po);
  const real b1 = dot3(pr, po);
  const real b2 = dot3(ps, po);

  const real det11 = T00 * T22 - T02 * T02;
  const real det12 = T00 * T12 - T01 * T02;
  const real det22 = T00 * T11 - T01 * T01;
  const real I00 = det00 / det;
  const real I01 = -det01 / det;
  const real I02 = det02 / det;
  const real I11 = det11 / det;
  const real I12 = -det12 / det;
  const real I22 = det22 / det;

  const real a1 = I00 * b0 + I01 * b1 + I02 * b2;
  const real a2 = I01 * b0 + I11 * b1 + I12 * b2;
  const real a3 = I02 * b0 + I12 * b1 + I22 * b2;
  const real a0 = 1.0 - a1 - a2 - a3; /* evaluated in double, rounded once */

  if (a0 < REAL_EPS) {
    s->nvrtx = 3;
    slot_copy(s, 0, 3);
    wit_2d(c1, c2, s);
  } else if (a1 < REAL_EPS) {
    s->nvrtx = 3;
    slot_copy(s, 1, 3);
    wit_2d(c1, c2, s);
  } else if (a2 < REAL_EPS) {
    s->nvrtx = 3;
    slot_copy(s, 2, 3);
    wit_2d(c1, c2, s);
  } else if (a3 < REAL_EPS) {
    s->nvrtx = 3;
    wit_2d(c1, c2, s);
  }

  const real* w00 = VTX(c1, s->vrtx_idx[0][0]);
  const real* w01 = VTX(c2, s->vrtx_idx[0][1]);
  const real* w10 = VTX(c1, s->vrtx_idx[1][0]);
  const real* w11 = VTX(c2, s->vrtx_idx[1][1]);
  const real* w20 = VTX(c1, s->vrtx_idx[2][0]);
  const real* w21 = VTX(c2, s->vrtx_idx[2][1]);
  const real* w30 = VTX(c1, s->vrtx_idx[3][0]);
  const real* w31 = VTX(c2, s->vrtx_idx[3][1]);
  for (int t = 0; t < 3; t++) {
    s->witnesses[0][t] =
        w00[t] * a0 + w10[t] * a1 + w20[t] * a2 + w30[t] * a3;
    s->witnesses[1][t] =
        w01[t] * a0 + w11[t] * a1 + w21[t] * a2 + w31[t] * a3;
  }
}

/* openGJK.c:607-630.  `cur`/`cur_idx` carry the previous support point. */
static inline void support_scan(int n, const real* c, const real* dir,
                                real* cur, int* cur_idx) {
  real best = dot3(cur, dir);
  int better = -1;
  for (int i = 0; i < n; ++i) {
    const real sc = dot3(VTX(c, i), dir);
    if (sc > best) {
      best = sc;
      better = i;
    }
  }
  if (better != -1) {
    cur[0] = VTX(c, better)[0];
    cur[1] = VTX(c, better)[1];
    cur[2] = VTX(c, better)[2];
    *cur_idx = better;
  }
}

/* openGJK.c:941-1046.  `s` must be zero- or caller-initialised: slots the
 * algorithm never writes keep their incoming bytes, exactly as the reference
 * does.  `iters` (optional) receives the iteration count k. */
real SUFFIX(oracle_gjk)(int n1, const real* c1, int n2, const real* c2,
                        osimplex* s, int* iters) {
  unsigned k = 0;
  const real eps_rel = EPS_REL;
  const real eps_tot = EPS_ABS;
  const real eps_rel2 = eps_rel * eps_rel;
  real w[3], v[3], vneg[3];
  real sup1[3], sup2[3];
  int sup1_idx = 0, sup2_idx = 0;
  real wmax2 = 0;

  for (int t = 0; t < 3; t++) v[t] = c1[t] - c2[t];
  s->nvrtx = 1;
  for (int t = 0; t < 3; t++) s->vrtx[0][t] = v[t];
  s->vrtx_idx[0][0] = 0;
  s->vrtx_idx[0][1] = 0;
  for (int t = 0; t < 3; t++) sup1[t] = c1[t];
  for (int t = 0; t < 3; t++) sup2[t] = c2[t];

  do {
    k++;
    for (int t = 0; t < 3; t++) vneg[t] = -v[t];
    support_scan(n1, c1, vneg, sup1, &sup1_idx);
    support_scan(n2, c2, v, sup2, &sup2_idx);
    for (int t = 0; t < 3; t++) w[t] = sup1[t] - sup2[t];

    const real gap = nrm2(v) - dot3(v, w);
    if (gap <= (eps_rel * nrm2(v)) || gap < EPS_ABS) break;
    if (nrm2(v) < eps_rel2) break;

    const int slot = s->nvrtx;
    for (int t = 0; t < 3; t++) s->vrtx[slot][t] = w[t];
    s->vrtx_idx[slot][0] = sup1_idx;
    s->vrtx_idx[slot][1] = sup2_idx;
    s->nvrtx++;

    switch (s->nvrtx) { /* openGJK.c:632-646 */
      case 4: sub_3d(s, v); break;
      case 3: sub_2d(s, v); break;
      case 2: sub_1d(s, v); break;
      default: break;
    }

    for (int j = 0; j < s->nvrtx; j++) {
      const real t2 = nrm2(s->vrtx[j]);
      if (t2 > wmax2) wmax2 = t2;
    }
    if (nrm2(v) <= (eps_tot * eps_tot * wmax2)) break;
  } while ((s->nvrtx != 4) && (k != MAX_ITER));

  switch (s->nvrtx) { /* openGJK.c:921-939 */
    case 4: wit_3d(c1, c2, s); break;
    case 3: wit_2d(c1, c2, s); break;
    case 2: wit_1d(c1, c2, s); break;
    case 1: wit_0d(c1, c2, s); break;
    default: break;
  }
  if (iters) *iters = (int)k;
  return sqrt(nrm2(v));
}

/* Batched driver over a polytope pool.
 *   coords : all vertices, xyz interleaved
 *   off    : off[h] = first vertex of polytope h (in vertices), cnt[h] = count
 *   pa/pb  : polytope index of body 1 / body 2 for each pair
 * Simplices are zero-filled before each query so never-written slots compare
 * equal across implementations.  nthreads <= 1 runs serially. */
void SUFFIX(oracle_gjk_batch)(int64_t npairs, const real* coords,
                              const int64_t* off, const int* cnt,
                              const int* pa, const int* pb, osimplex* simplices,
                              real* distances, int* iters, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t p = 0; p < npairs; p++) {
    const int a = pa[p], b = pb[p];
    osimplex* s = &simplices[p];
    memset(s, 0, sizeof(*s));
    int it = 0;
    distances[p] = SUFFIX(oracle_gjk)(cnt[a], coords + 3 * off[a], cnt[b],
                                      coords + 3 * off[b], s, &it);
    if (iters) iters[p] = it;
  }
}

/* Sub-solver entry points for the reference's known-answer vectors
 * (same role as opengjk_test_S{1,2,3}D, openGJK.c:1232-1248). */
void SUFFIX(oracle_sub)(osimplex* s, real* v) {
  switch (s->nvrtx) {
    case 4: sub_3d(s, v); break;
    case 3: sub_2d(s, v); break;
    case 2: sub_1d(s, v); break;
    default: break;
  }
}
int SUFFIX(oracle_hff1)(const real* p, const real* q) { return hf1(p, q); }
int SUFFIX(oracle_hff2)(const real* p, const real* q, const real* r) {
  return hf2(p, q, r);
}
int SUFFIX(oracle_hff3)(const real* p, const real* q, const real* r) {
  return hf3(p, q, r);
}
int SUFFIX(oracle_sizeof_simplex)(void) { return (int)sizeof(osimplex); }
