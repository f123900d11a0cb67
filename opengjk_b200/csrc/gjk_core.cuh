// gjk_core.cuh -- register-resident GJK building blocks for sm_100a.
//
// Device-side Signed-Volumes sub-solver (S1D/S2D/S3D), termination tests and
// witness reconstruction, templated on float/double.  Semantics follow
// /root/reference/scalar/openGJK.c (cited per function); the structure does
// not: the simplex lives in four named register slots (no dynamic indexing,
// so nothing spills to local memory), slot permutations are selects, and the
// S3D region tree is a table-driven automaton with a single hff2 instance.
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
  // outcome: 0 = face abc, 1 = edge ab (region12), 2 = edge ac (region13), 3 = vertex a.
  // At most two hf2 tests are needed on any path: t1 = hf2(a,b,c) when ab, t2 = hf2(a,c,b) when ac.
  int out = 3;
  bool need2 = ac;
  if (ab) {
    if (hf2(a, b, c)) { out = 1; need2 = false; } else { out = 0; }
  } else if (ac) {
    out = 0;
  }
  if (need2 && hf2(a, c, b)) out = 2;
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
// The reference resolves the "one plane" / "no plane" arms with a tree of up to
// four hff2 tests whose operands depend on the branch taken.  Here the tree is a
// table-driven automaton: every node names the two operands (i, j or k) of one
// hff2(s1, X, Y) test and where to go on true / false; leaves carry the outcome
// (which vertices survive, which projection).  One hff2 instance serves every
// region, so lanes of a warp that sit in different regions still share the
// instruction stream, and the code stays small enough for the instruction cache.
//
// entry = x | y<<2 | on_true<<4 | on_false<<16 ; next values >= kLeaf are leaves:
//   leaf payload = kind | a<<2 | b<<4 | swap<<6   (kind 0 none, 1 segment{new,a}, 2 triangle{new,a,b})
constexpr unsigned kLeaf = 0x800;
constexpr unsigned S3I = 0, S3J = 1, S3K = 2;
constexpr unsigned leaf(unsigned kind, unsigned a, unsigned b, unsigned sw) { return kLeaf | kind | (a << 2) | (b << 4) | (sw << 6); }
constexpr unsigned node(unsigned x, unsigned y, unsigned t, unsigned f) { return x | (y << 2) | (t << 4) | (f << 16); }
constexpr unsigned L_TRI_IK = leaf(2, S3I, S3K, 0), L_TRI_JK = leaf(2, S3J, S3K, 0), L_TRI_IJ = leaf(2, S3I, S3J, 0);
constexpr unsigned L_TRI_IK_SW = leaf(2, S3I, S3K, 1), L_TRI_JK_SW = leaf(2, S3J, S3K, 1);
constexpr unsigned L_SEG_I = leaf(1, S3I, 0, 0), L_SEG_J = leaf(1, S3J, 0, 0), L_SEG_K = leaf(1, S3K, 0, 0);
constexpr unsigned L_NONE = leaf(0, 0, 0, 0);
// Branch is on r = hff2(s1, X, Y) (the reference mostly tests !hff2; true/false are swapped accordingly).
__device__ const unsigned kS3Table[20] = {
    /* 0  :436 */ node(S3K, S3I, 1, L_TRI_IK),        /* 1  :439 */ node(S3K, S3J, L_SEG_K, L_TRI_JK),
    /* 2  :447 */ node(S3I, S3K, L_SEG_I, L_TRI_IK),  /* 3  :455 */ node(S3J, S3K, L_SEG_J, L_TRI_JK),
    /* 4  :465 */ node(S3K, S3I, 6, 5),               /* 5  :466 */ node(S3I, S3K, L_SEG_K, L_TRI_IK),
    /* 6  :474 */ node(S3K, S3J, L_SEG_K, L_TRI_JK),  /* 7  :483 */ node(S3K, S3J, 9, 8),
    /* 8  :484 */ node(S3J, S3K, L_SEG_J, L_TRI_JK),  /* 9  :492 */ node(S3K, S3I, L_SEG_K, L_TRI_IK),
    /* 10 :505 */ node(S3K, S3I, 11, 13),             /* 11 :506 */ node(S3K, S3J, L_SEG_K, 12),
    /* 12 :504 */ node(S3J, S3K, L_SEG_J, L_TRI_JK_SW), /* 13 :503 */ node(S3I, S3K, L_SEG_I, L_TRI_IK_SW),
    /* 14 :553 */ node(S3I, S3J, 15, L_TRI_IJ),       /* 15 :556 */ node(S3I, S3K, L_SEG_I, L_TRI_IK),
    /* 16 :582 */ node(S3J, S3K, 19, 17),             /* 17 :583 */ node(S3K, S3J, 18, L_TRI_JK),
    /* 18 :586 */ node(S3K, S3I, L_SEG_K, L_TRI_IK_SW), /* 19 :593 */ node(S3J, S3I, L_SEG_J, L_TRI_IJ)};

template <typename T> OGJK_DI MV<T> pick3(unsigned c, const MV<T>& i, const MV<T>& j, const MV<T>& k) {
  MV<T> r;
  r.x = c == S3I ? i.x : (c == S3J ? j.x : k.x);
  r.y = c == S3I ? i.y : (c == S3J ? j.y : k.y);
  r.z = c == S3I ? i.z : (c == S3J ? j.z : k.z);
  r.ia = c == S3I ? i.ia : (c == S3J ? j.ia : k.ia);
  r.ib = c == S3I ? i.ib : (c == S3J ? j.ib : k.ib);
  return r;
}
template <typename T> OGJK_DI V3<T> pick3v(unsigned c, const MV<T>& i, const MV<T>& j, const MV<T>& k) {
  return V3<T>{c == S3I ? i.x : (c == S3J ? j.x : k.x), c == S3I ? i.y : (c == S3J ? j.y : k.y),
               c == S3I ? i.z : (c == S3J ? j.z : k.z)};
}

// Returns true when the reference falls through to S2D (plane-sum 2, :382-413).
template <typename T> OGJK_DI bool sub_3d(Splx<T>& s, V3<T>& v) {
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
    return false;
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
    return false;
  }
  if (npl == 2) {  // :382-413; S2D itself is run by the caller (single instance)
    s.n = 3;
    if (!pl2) {
      s.q2 = m1;                    // q1 = s3, q0 = s4 stay
    } else if (!pl3) {
      s.q1 = m2; s.q2 = m1;         // q0 = s4 stays
    } else {
      s.q0 = m3; s.q1 = m2; s.q2 = m1;
    }
    return true;
  }

  // slot order is [0]=s4, [1]=s3, [2]=s2.  (k,i,j) is one of three rotations:
  //   rot 0: k=2,i=1,j=0   rot 1: k=1,i=0,j=2   rot 2: k=0,i=2,j=1
  int rot;
  unsigned at;  // automaton start node, or a leaf
  if (npl == 1) {  // :414-532
    s.n = 3;       // set before any region is known, as the reference does
    rot = pl2 ? 0 : (pl3 ? 1 : 2);
    const bool hk = rot == 0 ? h2 : (rot == 1 ? h3 : h4);
    const bool hi = rot == 0 ? h3 : (rot == 1 ? h4 : h2);
    const bool hj = rot == 0 ? h4 : (rot == 1 ? h2 : h3);
    if (ndot == 1)      at = hk ? 0u : (hi ? 2u : 3u);
    else if (ndot == 2) at = hi ? 4u : (hj ? 7u : L_NONE);
    else                at = 10u;
  } else {         // npl == 0 :534-601
    if (ndot == 1)      { rot = h3 ? 0 : (h4 ? 1 : 2); at = 14u; }
    else if (ndot == 2) { s.n = 3; rot = !h3 ? 0 : (!h4 ? 1 : 2); at = 16u; }
    else                { rot = 0; at = L_NONE; }  // reference does nothing: simplex stays at 4, v unchanged
  }
  const MV<T> mk = rot == 0 ? m2 : (rot == 1 ? m3 : m4);
  const MV<T> mi = rot == 0 ? m3 : (rot == 1 ? m4 : m2);
  const MV<T> mj = rot == 0 ? m4 : (rot == 1 ? m2 : m3);

#pragma unroll 1
  while (at < kLeaf) {
    const unsigned e = kS3Table[at];
    const V3<T> X = pick3v(e & 3u, mi, mj, mk);
    const V3<T> Y = pick3v((e >> 2) & 3u, mi, mj, mk);
    at = hf2(s1, X, Y) ? ((e >> 4) & 0xfffu) : (e >> 16);
  }
  const unsigned kind = at & 3u;
  if (kind != 0) {
    const MV<T> A = pick3((at >> 2) & 3u, mi, mj, mk);
    if (kind == 2) {          // select_1xy :67-92
      const MV<T> B = pick3((at >> 4) & 3u, mi, mj, mk);
      const bool sw = (at >> 6) & 1u;
      s.n = 3;
      s.q2 = m1; s.q1 = A; s.q0 = B;
      const V3<T> pa = pos(A), pb = pos(B);
      v = proj_plane(s1, sw ? pb : pa, sw ? pa : pb);
    } else {                  // select_1x :94-113
      s.n = 2;
      s.q1 = m1; s.q0 = A;
      v = proj_line(s1, pos(A));
    }
  }
  return false;
}

// openGJK.c:632-646 after the push at :1013-1019.  `w` goes to slot s.n.  S3D
// runs first so that its fall-through into S2D shares the one S2D instance with
// the lanes that arrived with three vertices.
template <typename T> OGJK_DI void push_and_solve(Splx<T>& s, const MV<T>& w, V3<T>& v) {
  int todo;  // which lower-dimensional solver still has to run: 0 none, 1 S1D, 2 S2D
  if (s.n == 3)      { s.q3 = w; s.n = 4; todo = sub_3d(s, v) ? 2 : 0; }
  else if (s.n == 2) { s.q2 = w; s.n = 3; todo = 2; }
  else               { s.q1 = w; s.n = 2; todo = 1; }
  if (todo == 2) sub_2d(s, v);
  else if (todo == 1) sub_1d(s, v);
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
// to the final slot contents.  W1D never changes slots, so calls made into it
// only for their side effects vanish.
template <typename T> struct Wts { T a0, a1, a2, a3; int terms; };

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

// compute_witnesses (:921-939).  For a tetrahedron the reference may run W2D
// twice for its side effects (once when det == 0, :845-848, once after dropping
// the vertex whose weight is below eps, :873-900); both runs and the direct
// three-vertex case go through the single W2D instance in the loop below.
template <typename T> OGJK_DI Wts<T> witness_weights(Splx<T>& s) {
  Wts<T> w = Wts<T>{(T)1, (T)0, (T)0, (T)0, s.n};
  bool run0 = (s.n == 3), run1 = false;
  int drop = -1;
  if (s.n == 4) {
    const V3<T> p = pos(s.q0), q = pos(s.q1), r = pos(s.q2), u = pos(s.q3);
    const V3<T> pq = sub(q, p), pr = sub(r, p), ps = sub(u, p);
    const V3<T> po = V3<T>{-p.x, -p.y, -p.z};
    const T T00 = dot3(pq, pq), T01 = dot3(pq, pr), T02 = dot3(pq, ps);
    const T T11 = dot3(pr, pr), T12 = dot3(pr, ps), T22 = dot3(ps, ps);
    const T det00 = T11 * T22 - T12 * T12;
    const T det01 = T01 * T22 - T02 * T12;
    const T det02 = T01 * T12 - T02 * T11;
    const T det = T00 * det00 - T01 * det01 + T02 * det02;
    const T b0 = dot3(pq, po), b1 = dot3(pr, po), b2 = dot3(ps, po);
    const T det11 = T00 * T22 - T02 * T02;
    const T det12 = T00 * T12 - T01 * T02;
    const T det22 = T00 * T11 - T01 * T01;
    const T I00 = det00 / det, I01 = -det01 / det, I02 = det02 / det;
    const T I11 = det11 / det, I12 = -det12 / det, I22 = det22 / det;
    w.a1 = I00 * b0 + I01 * b1 + I02 * b2;
    w.a2 = I01 * b0 + I11 * b1 + I12 * b2;
    w.a3 = I02 * b0 + I12 * b1 + I22 * b2;
    w.a0 = 1.0 - w.a1 - w.a2 - w.a3;  // evaluated in double, rounded once
    run0 = (det == 0.0);
    drop = w.a0 < Lim<T>::eps() ? 0 : (w.a1 < Lim<T>::eps() ? 1 : (w.a2 < Lim<T>::eps() ? 2 : (w.a3 < Lim<T>::eps() ? 3 : -1)));
    run1 = drop >= 0;
  }
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && run1) {
      s.n = 3;
      if (drop == 0) s.q0 = s.q3;
      else if (drop == 1) s.q1 = s.q3;
      else if (drop == 2) s.q2 = s.q3;
    }
    if (pass == 0 ? run0 : run1) {
      const Wts<T> w2 = wit_2d(s);
      if (w.terms == 3) w = w2;
    }
  }
  if (w.terms == 2) w = wit_1d(s);
  return w;
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
