// gjk_core.cuh -- register-resident GJK building blocks for sm_100a.
//
// Device-side Signed-Volumes sub-solver (S1D/S2D/S3D), termination tests and
// witness reconstruction, templated on float/double.  Semantics follow
// /root/reference/scalar/openGJK.c (cited per function); the structure does
// not: the simplex lives in four named register slots (no dynamic indexing,
// so nothing spills to local memory), slot permutations are selects, and the
// S3D region tree is resolved to a small "decision" that is applied once.
//
// Bit-parity contract: this translation unit is compiled with -fmad=false and
// the default IEEE div/sqrt/denormal modes, and every expression keeps the
// reference's operand order, so fp32/fp64 results are bit-identical to the
// reference built for x86-64 without FMA (SURVEY.md section 8a checklist).
#pragma once
#include <cfloat>
#include <cstdint>

namespace ogjk {

template <typename T> struct Lim;
template <> struct Lim<float> {
  static __host__ __device__ constexpr float eps() { return FLT_EPSILON; }
};
template <> struct Lim<double> {
  static __host__ __device__ constexpr double eps() { return DBL_EPSILON; }
};

// openGJK.c:44-56: mk = 25, eps_rel = eps*1e4, eps_tot = eps*1e2 (both exact).
constexpr int kMaxIter = 25;
template <typename T> __device__ __forceinline__ T eps_rel() { return (T)(Lim<T>::eps() * 1e4); }
template <typename T> __device__ __forceinline__ T eps_abs() { return (T)(Lim<T>::eps() * 1e2); }

template <typename T> struct V3 { T x, y, z; };

// A Minkowski-difference vertex with the indices of its two source vertices.
template <typename T> struct MV { T x, y, z; int ia, ib; };

// Simplex in registers: q0 oldest ... q(n-1) newest (openGJK.h:113-118).
template <typename T> struct Splx { int n; MV<T> q0, q1, q2, q3; };

#define OGJK_DI __device__ __forceinline__

template <typename T> OGJK_DI T dot3(T ax, T ay, T az, T bx, T by, T bz) {
  return ax * bx + ay * by + az * bz;  // openGJK.c:59, left to right
}
template <typename T> OGJK_DI T dot3(const V3<T>& a, const V3<T>& b) { return dot3(a.x, a.y, a.z, b.x, b.y, b.z); }
template <typename T> OGJK_DI T nrm2(T x, T y, T z) { return x * x + y * y + z * z; }  // :58
template <typename T> OGJK_DI V3<T> pos(const MV<T>& m) { return V3<T>{m.x, m.y, m.z}; }
template <typename T> OGJK_DI V3<T> sub(const V3<T>& a, const V3<T>& b) { return V3<T>{a.x - b.x, a.y - b.y, a.z - b.z}; }

// openGJK.c:189-195
template <typename T> OGJK_DI V3<T> cross3(const V3<T>& a, const V3<T>& b) {
  return V3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// openGJK.c:181-187
template <typename T> OGJK_DI T det3(const V3<T>& p, const V3<T>& q, const V3<T>& r) {
  return p.x * ((q.y * r.z) - (r.y * q.z)) - p.y * (q.x * r.z - r.x * q.z) + p.z * (q.x * r.y - r.x * q.y);
}
// openGJK.c:197-210
template <typename T> OGJK_DI V3<T> proj_line(const V3<T>& p, const V3<T>& q) {
  const V3<T> pq = sub(p, q);
  const T t = dot3(p, pq) / dot3(pq, pq);
  return V3<T>{p.x - pq.x * t, p.y - pq.y * t, p.z - pq.z * t};
}
// openGJK.c:212-231
template <typename T> OGJK_DI V3<T> proj_plane(const V3<T>& p, const V3<T>& q, const V3<T>& r) {
  const V3<T> n = cross3(sub(p, q), sub(p, r));
  const T t = dot3(n, p) / dot3(n, n);
  return V3<T>{n.x * t, n.y * t, n.z * t};
}
// openGJK.c:233-244: sum from 0 in index order; (0 + a) == a up to the sign of zero.
template <typename T> OGJK_DI bool hf1(const V3<T>& p, const V3<T>& q) {
  T acc = (T)0;
  acc += (p.x * p.x - p.x * q.x);
  acc += (p.y * p.y - p.y * q.y);
  acc += (p.z * p.z - p.z * q.z);
  return acc > (T)0;
}
// openGJK.c:246-262
template <typename T> OGJK_DI bool hf2(const V3<T>& p, const V3<T>& q, const V3<T>& r) {
  const V3<T> pq = sub(q, p);
  const V3<T> m = cross3(pq, sub(r, p));
  const V3<T> n = cross3(pq, m);
  return dot3(p, n) < (T)0;
}
// openGJK.c:264-277
template <typename T> OGJK_DI bool hf3(const V3<T>& p, const V3<T>& q, const V3<T>& r) {
  const V3<T> n = cross3(sub(q, p), sub(r, p));
  return dot3(p, n) <= (T)0;
}

template <typename T> OGJK_DI MV<T> pick(bool c, const MV<T>& a, const MV<T>& b) {
  MV<T> r;
  r.x = c ? a.x : b.x; r.y = c ? a.y : b.y; r.z = c ? a.z : b.z;
  r.ia = c ? a.ia : b.ia; r.ib = c ? a.ib : b.ib;
  return r;
}
template <typename T> OGJK_DI V3<T> pick(bool c, const V3<T>& a, const V3<T>& b) {
  return V3<T>{c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z};
}

// ---- S1D: openGJK.c:279-290 (+ S1Dregion1 :132-141) ------------------------
template <typename T> OGJK_DI void sub_1d(Splx<T>& s, V3<T>& v) {
  const V3<T> a = pos(s.q1), b = pos(s.q0);
  if (hf1(a, b)) {
    v = proj_line(a, b);
  } else {
    v = a;
    s.n = 1;
    s.q0 = s.q1;
  }
}

// ---- S2D: openGJK.c:292-335 (+ region macros :143-168) ----------------------
// The tree is reduced to one of four outcomes so the projections are issued
// once per warp instead of once per leaf.
template <typename T> OGJK_DI void sub_2d(Splx<T>& s, V3<T>& v) {
  const V3<T> a = pos(s.q2), b = pos(s.q1), c = pos(s.q0);
  const bool ab = hf1(a, b);
  const bool ac = hf1(a, c);
  // outcome: 0 = face abc, 1 = edge ab (region12), 2 = edge ac (region13), 3 = vertex a
  int out;
  if (ab) {
    if (!hf2(a, b, c)) {
      out = (ac && hf2(a, c, b)) ? 2 : 0;
    } else {
      out = 1;
    }
  } else if (ac) {
    out = hf2(a, c, b) ? 2 : 0;
  } else {
    out = 3;
  }
  if (out == 0) {
    v = proj_plane(a, b, c);
  } else if (out == 3) {
    v = a;             // S2Dregion1
    s.n = 1;
    s.q0 = s.q2;
  } else {
    v = proj_line(a, out == 1 ? b : c);
    s.n = 2;
    if (out == 1) s.q0 = s.q2;  // S2Dregion12: newest into slot 0, b stays in slot 1
    else          s.q1 = s.q2;  // S2Dregion13: newest into slot 1, c stays in slot 0
  }
}

// ---- S3D: openGJK.c:337-605 --------------------------------------------------
// Decision produced by the region tree of the "one plane" / "no plane" arms:
//   kind 0: nothing (reference leaves simplex and v untouched)
//   kind 1: segment {newest, A}           -> v = proj_line(s1, A)
//   kind 2: triangle {newest, A, B}       -> v = proj_plane(s1, P, Q)  (P,Q = A,B or B,A)
template <typename T> OGJK_DI void sub_3d(Splx<T>& s, V3<T>& v) {
  const MV<T> m1 = s.q3, m2 = s.q2, m3 = s.q1, m4 = s.q0;
  const V3<T> s1 = pos(m1), s2 = pos(m2), s3 = pos(m3), s4 = pos(m4);

  const bool h2 = hf1(s1, s2);  // hff1_tests[2]
  const bool h3 = hf1(s1, s3);  // hff1_tests[1] == testLineThree
  const bool h4 = hf1(s1, s4);  // hff1_tests[0] == testLineFour
  const int ndot = (int)h2 + (int)h3 + (int)h4;

  if (ndot == 0) {  // S3Dregion1 :170-179
    v = s1;
    s.n = 1;
    s.q0 = m1;
    return;
  }

  const V3<T> e12 = sub(s2, s1), e13 = sub(s3, s1), e14 = sub(s4, s1);
  const int sss = (det3(e13, e14, e12) <= (T)0) ? 1 : 0;
  // (hff3 - sss)^2 == (hff3 != sss)
  const bool pl2 = ((int)hf3(s1, s3, s4) != sss);
  const bool pl3 = ((int)hf3(s1, s4, s2) != sss);
  const bool pl4 = ((int)hf3(s1, s2, s3) != sss);
  const int npl = (int)pl2 + (int)pl3 + (int)pl4;

  if (npl == 3) {  // S3Dregion1234 :61-65
    v = V3<T>{(T)0, (T)0, (T)0};
    s.n = 4;
    return;
  }
  if (npl == 2) {  // :382-413
    s.n = 3;
    if (!pl2) {
      s.q2 = m1;                    // q1 = s3, q0 = s4 stay
    } else if (!pl3) {
      s.q1 = m2; s.q2 = m1;         // q0 = s4 stays
    } else {
      s.q0 = m3; s.q1 = m2; s.q2 = m1;
    }
    sub_2d(s, v);
    return;
  }

  // slot order is [0]=s4, [1]=s3, [2]=s2.  (k,i,j) is one of three rotations:
  //   rot 0: k=2,i=1,j=0   rot 1: k=1,i=0,j=2   rot 2: k=0,i=2,j=1
  int kind = 0;
  bool swapPQ = false;      // projection argument order (B,A) instead of (A,B)
  MV<T> A = m2, B = m2;     // chosen survivors (besides the newest vertex)

  if (npl == 1) {  // :414-532
    s.n = 3;       // set before any region is known, as the reference does
    const int rot = pl2 ? 0 : (pl3 ? 1 : 2);
    const MV<T> mk = rot == 0 ? m2 : (rot == 1 ? m3 : m4);
    const MV<T> mi = rot == 0 ? m3 : (rot == 1 ? m4 : m2);
    const MV<T> mj = rot == 0 ? m4 : (rot == 1 ? m2 : m3);
    const bool hk = rot == 0 ? h2 : (rot == 1 ? h3 : h4);
    const bool hi = rot == 0 ? h3 : (rot == 1 ? h4 : h2);
    const bool hj = rot == 0 ? h4 : (rot == 1 ? h2 : h3);
    const V3<T> sk = pos(mk), si = pos(mi), sj = pos(mj);
    // what[]: 1 = tri(i,k) 2 = tri(j,k) 3 = seg(k) 4 = seg(i) 5 = seg(j)
    //         6 = tri(j,k) with plane(s1,sk,sj)   7 = tri(i,k) with plane(s1,sk,si)
    int what = 0;
    if (ndot == 1) {
      if (hk) {
        if (!hf2(s1, sk, si)) what = 1;
        else if (!hf2(s1, sk, sj)) what = 2;
        else what = 3;
      } else if (hi) {
        what = !hf2(s1, si, sk) ? 1 : 4;
      } else {
        what = !hf2(s1, sj, sk) ? 2 : 5;
      }
    } else if (ndot == 2) {
      if (hi) {
        if (!hf2(s1, sk, si)) what = !hf2(s1, si, sk) ? 1 : 3;
        else                  what = !hf2(s1, sk, sj) ? 2 : 3;
      } else if (hj) {
        if (!hf2(s1, sk, sj)) what = !hf2(s1, sj, sk) ? 2 : 5;
        else                  what = !hf2(s1, sk, si) ? 1 : 3;
      }
    } else {  // ndot == 3 :502-531
      const bool f_ik = hf2(s1, si, sk);
      const bool f_jk = hf2(s1, sj, sk);
      const bool f_ki = hf2(s1, sk, si);
      const bool f_kj = hf2(s1, sk, sj);
      if (f_ki && f_kj) what = 3;
      else if (f_ki)    what = f_jk ? 5 : 6;
      else              what = f_ik ? 4 : 7;
    }
    switch (what) {
      case 1: kind = 2; A = mi; B = mk; break;
      case 2: kind = 2; A = mj; B = mk; break;
      case 3: kind = 1; A = mk; break;
      case 4: kind = 1; A = mi; break;
      case 5: kind = 1; A = mj; break;
      case 6: kind = 2; A = mj; B = mk; swapPQ = true; break;
      case 7: kind = 2; A = mi; B = mk; swapPQ = true; break;
      default: break;
    }
  } else {  // npl == 0 :534-601
    if (ndot == 1) {
      const int rot = h3 ? 0 : (h4 ? 1 : 2);
      const MV<T> mk = rot == 0 ? m2 : (rot == 1 ? m3 : m4);
      const MV<T> mi = rot == 0 ? m3 : (rot == 1 ? m4 : m2);
      const MV<T> mj = rot == 0 ? m4 : (rot == 1 ? m2 : m3);
      const V3<T> sk = pos(mk), si = pos(mi), sj = pos(mj);
      if (!hf2(s1, si, sj))      { kind = 2; A = mi; B = mj; }
      else if (!hf2(s1, si, sk)) { kind = 2; A = mi; B = mk; }
      else                       { kind = 1; A = mi; }
    } else if (ndot == 2) {
      s.n = 3;
      const int rot = !h3 ? 0 : (!h4 ? 1 : 2);
      const MV<T> mk = rot == 0 ? m2 : (rot == 1 ? m3 : m4);
      const MV<T> mi = rot == 0 ? m3 : (rot == 1 ? m4 : m2);
      const MV<T> mj = rot == 0 ? m4 : (rot == 1 ? m2 : m3);
      const V3<T> sk = pos(mk), si = pos(mi), sj = pos(mj);
      if (!hf2(s1, sj, sk)) {
        if (!hf2(s1, sk, sj))      { kind = 2; A = mj; B = mk; }
        else if (!hf2(s1, sk, si)) { kind = 2; A = mi; B = mk; swapPQ = true; }
        else                       { kind = 1; A = mk; }
      } else if (!hf2(s1, sj, si)) { kind = 2; A = mi; B = mj; }
      else                         { kind = 1; A = mj; }
    }
    // ndot == 3: the reference does nothing (simplex stays at 4 vertices, v unchanged)
  }

  if (kind == 2) {          // select_1xy :67-92
    s.n = 3;
    s.q2 = m1; s.q1 = A; s.q0 = B;
    const V3<T> pa = pos(A), pb = pos(B);
    v = swapPQ ? proj_plane(s1, pb, pa) : proj_plane(s1, pa, pb);
  } else if (kind == 1) {   // select_1x :94-113
    s.n = 2;
    s.q1 = m1; s.q0 = A;
    v = proj_line(s1, pos(A));
  }
}

// openGJK.c:632-646 after the push at :1013-1019.  `w` goes to slot s.n.
template <typename T> OGJK_DI void push_and_solve(Splx<T>& s, const MV<T>& w, V3<T>& v) {
  if (s.n == 1)      { s.q1 = w; s.n = 2; sub_1d(s, v); }
  else if (s.n == 2) { s.q2 = w; s.n = 3; sub_2d(s, v); }
  else               { s.q3 = w; s.n = 4; sub_3d(s, v); }
}

// openGJK.c:1025-1030: running max of |vertex|^2 over the current simplex.
template <typename T> OGJK_DI T update_wmax(const Splx<T>& s, T wmax2) {
  T t = nrm2(s.q0.x, s.q0.y, s.q0.z);
  if (t > wmax2) wmax2 = t;
  if (s.n > 1) { t = nrm2(s.q1.x, s.q1.y, s.q1.z); if (t > wmax2) wmax2 = t; }
  if (s.n > 2) { t = nrm2(s.q2.x, s.q2.y, s.q2.z); if (t > wmax2) wmax2 = t; }
  if (s.n > 3) { t = nrm2(s.q3.x, s.q3.y, s.q3.z); if (t > wmax2) wmax2 = t; }
  return wmax2;
}

// ---- witnesses: openGJK.c:648-939 -------------------------------------------
// Barycentric weights of the origin on the final simplex.  The reference's
// W3D -> W2D -> W1D cascade has no early returns, so after a reduction the
// higher-dimensional blend overwrites the lower one using the already
// shuffled slots; only the slot shuffles and nvrtx survive.  That is what is
// reproduced here: `wts` are the weights of the *outermost* routine, applied
// to the final slot contents.
template <typename T> struct Wts { T a0, a1, a2, a3; int terms; };

template <typename T> OGJK_DI void wit_1d_shuffle(Splx<T>&) {}  // W1D never reshuffles

template <typename T> OGJK_DI Wts<T> wit_1d(const Splx<T>& s) {
  const V3<T> p = pos(s.q0), q = pos(s.q1);
  const V3<T> pq = sub(q, p);
  const V3<T> po = V3<T>{-p.x, -p.y, -p.z};
  const T det = dot3(pq, pq);
  const T a1 = dot3(pq, po) / det;
  const T a0 = 1.0 - a1;  // double arithmetic, one rounding (as the C source)
  return Wts<T>{a0, a1, (T)0, (T)0, 2};
}

// Returns weights; may reduce s to 2 vertices (slot shuffles :751-770).
template <typename T> OGJK_DI Wts<T> wit_2d(Splx<T>& s) {
  const V3<T> p = pos(s.q0), q = pos(s.q1), r = pos(s.q2);
  const V3<T> pq = sub(q, p), pr = sub(r, p);
  const V3<T> po = V3<T>{-p.x, -p.y, -p.z};
  const T T00 = dot3(pq, pq), T01 = dot3(pq, pr), T11 = dot3(pr, pr);
  const T det = T00 * T11 - T01 * T01;
  // det == 0: the reference calls W1D and falls through; W1D changes no slots.
  const T b0 = dot3(pq, po), b1 = dot3(pr, po);
  const T I00 = T11 / det, I01 = -T01 / det, I11 = T00 / det;
  const T a1 = I00 * b0 + I01 * b1;
  const T a2 = I01 * b0 + I11 * b1;
  const T a0 = 1.0 - a1 - a2;  // evaluated in double, rounded once
  if (a0 < Lim<T>::eps())      { s.n = 2; s.q0 = s.q2; }
  else if (a1 < Lim<T>::eps()) { s.n = 2; s.q1 = s.q2; }
  else if (a2 < Lim<T>::eps()) { s.n = 2; }
  return Wts<T>{a0, a1, a2, (T)0, 3};
}

template <typename T> OGJK_DI Wts<T> wit_3d(Splx<T>& s) {
  const V3<T> p = pos(s.q0), q = pos(s.q1), r = pos(s.q2), u = pos(s.q3);
  const V3<T> pq = sub(q, p), pr = sub(r, p), ps = sub(u, p);
  const V3<T> po = V3<T>{-p.x, -p.y, -p.z};
  const T T00 = dot3(pq, pq), T01 = dot3(pq, pr), T02 = dot3(pq, ps);
  const T T11 = dot3(pr, pr), T12 = dot3(pr, ps), T22 = dot3(ps, ps);
  const T det00 = T11 * T22 - T12 * T12;
  const T det01 = T01 * T22 - T02 * T12;
  const T det02 = T01 * T12 - T02 * T11;
  const T det = T00 * det00 - T01 * det01 + T02 * det02;
  bool reduced = false;
  if (det == 0.0) {  // :845-848 W2D is called for its side effects, then falls through
    (void)wit_2d(s);
    reduced = true;
  }
  const T b0 = dot3(pq, po), b1 = dot3(pr, po), b2 = dot3(ps, po);
  const T det11 = T00 * T22 - T02 * T02;
  const T det12 = T00 * T12 - T01 * T02;
  const T det22 = T00 * T11 - T01 * T01;
  const T I00 = det00 / det, I01 = -det01 / det, I02 = det02 / det;
  const T I11 = det11 / det, I12 = -det12 / det, I22 = det22 / det;
  const T a1 = I00 * b0 + I01 * b1 + I02 * b2;
  const T a2 = I01 * b0 + I11 * b1 + I12 * b2;
  const T a3 = I02 * b0 + I12 * b1 + I22 * b2;
  const T a0 = 1.0 - a1 - a2 - a3;  // evaluated in double, rounded once
  (void)reduced;
  if (a0 < Lim<T>::eps())      { s.n = 3; s.q0 = s.q3; (void)wit_2d(s); }
  else if (a1 < Lim<T>::eps()) { s.n = 3; s.q1 = s.q3; (void)wit_2d(s); }
  else if (a2 < Lim<T>::eps()) { s.n = 3; s.q2 = s.q3; (void)wit_2d(s); }
  else if (a3 < Lim<T>::eps()) { s.n = 3; (void)wit_2d(s); }
  return Wts<T>{a0, a1, a2, a3, 4};
}

// Dispatch of compute_witnesses (:921-939) on the final simplex size.
template <typename T> OGJK_DI Wts<T> witness_weights(Splx<T>& s) {
  if (s.n == 4) return wit_3d(s);
  if (s.n == 3) return wit_2d(s);
  if (s.n == 2) return wit_1d(s);
  return Wts<T>{(T)1, (T)0, (T)0, (T)0, 1};
}

// Same memory layout as the reference gkSimplex (openGJK.h:84-94):
// fp64 184 B {0,8,104,136}; fp32 108 B {0,4,52,84}.
template <typename T> struct GkSimplex {
  int nvrtx;
  T vrtx[4][3];
  int vrtx_idx[4][2];
  T witnesses[2][3];
};
static_assert(sizeof(GkSimplex<float>) == 108, "gkSimplex fp32 layout");
static_assert(sizeof(GkSimplex<double>) == 184, "gkSimplex fp64 layout");

}  // namespace ogjk
