// gjk_kernels.cuh -- batched GJK distance kernels for sm_100a.
//
// Replaces the reference's compute_minimum_distance_kernel and
// compute_minimum_distance_indexed_kernel (gpu/opengjk_gpu.cu:1257-1631) with
// two kernels over one addressing scheme (pool + optional per-pair indices):
//
//   gjk_thread_kernel : one pair per thread  -- small hulls; the whole GJK
//                       state (simplex, search direction, carried supports)
//                       is register resident and nothing is exchanged
//                       between lanes.
//   gjk_warp_kernel   : one pair per warp    -- large hulls; the support scan
//                       is lane-strided over the vertex stream and reduced with
//                       a (value, lowest-index) butterfly of __shfl_xor; the
//                       sub-solver then runs redundantly (and convergently) in
//                       every lane, so no broadcast is needed.
//
// Both pull work from a device-side queue in warp-sized chunks so that the
// 1..25-iteration spread of GJK does not leave SMs idle at the tail.
#pragma once
#include <cuda_runtime.h>

#include "gjk_core.cuh"
#include "store.cuh"

namespace ogjk {

template <typename T> struct BatchArgs {
  const T* coords1; const long long* off1; const int* cnt1; const int* idx1;
  const T* coords2; const long long* off2; const int* cnt2; const int* idx2;
  GkSimplex<T>* simplices;  // may be null
  T* distances;
  long long npairs;
};

// ---- support functions (openGJK.c:607-630) ------------------------------------
// `cur`/`cur_i` carry the previous support point of the body: the scan starts
// from its dot product, takes vertex i only when strictly larger, lowest index
// first, and leaves the carried point alone when nothing beats it.

template <typename T> struct ThreadBody {
  const T* c; int n;
  OGJK_DI V3<T> vertex(int i) const {
    const T* p = c + 3 * (size_t)i;
    return V3<T>{p[0], p[1], p[2]};
  }
  OGJK_DI void support(T dx, T dy, T dz, V3<T>& cur, int& cur_i) const {
    T best = dot3(cur.x, cur.y, cur.z, dx, dy, dz);
    int better = -1;
    const T* p = c;
#pragma unroll 4
    for (int i = 0; i < n; ++i, p += 3) {
      const T sc = dot3(p[0], p[1], p[2], dx, dy, dz);
      if (sc > best) { best = sc; better = i; }
    }
    if (better >= 0) { cur = vertex(better); cur_i = better; }
  }
};

template <typename T> struct WarpBody {
  const T* c; int n; int lane;
  OGJK_DI V3<T> vertex(int i) const {
    const T* p = c + 3 * (size_t)i;
    return V3<T>{p[0], p[1], p[2]};
  }
  OGJK_DI void support(T dx, T dy, T dz, V3<T>& cur, int& cur_i) const {
    T best = dot3(cur.x, cur.y, cur.z, dx, dy, dz);
    int better = 0x7fffffff;
#pragma unroll 4
    for (int i = lane; i < n; i += 32) {
      const T* p = c + 3 * (size_t)i;
      const T sc = dot3(p[0], p[1], p[2], dx, dy, dz);
      if (sc > best) { best = sc; better = i; }
    }
    // butterfly argmax: larger value wins, equal values -> lower index
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const T ob = __shfl_xor_sync(0xffffffffu, best, m);
      const int oi = __shfl_xor_sync(0xffffffffu, better, m);
      if (ob > best || (ob == best && oi < better)) { best = ob; better = oi; }
    }
    if (better != 0x7fffffff) { cur = vertex(better); cur_i = better; }
  }
};

// Both support calls of one GJK iteration (openGJK.c:992-993): body 1 along -v,
// body 2 along +v.  (A fused single-stream scan of both bodies was measured and was
// slower than two 2-chunk-per-trip scans: less ILP per trip.)
template <typename T, typename Body>
OGJK_DI void support_both(const Body& b1, const Body& b2, const V3<T>& v, V3<T>& sa, int& ia, V3<T>& sb, int& ib) {
  b1.support(-v.x, -v.y, -v.z, sa, ia);
  b2.support(v.x, v.y, v.z, sb, ib);
}

#ifdef OGJK_PHASE_CLOCKS
// Debug build only (tools/phase_clocks.py): SM cycles summed over lane 0 of every
// group.  [0] support scans, [1] rest of the step, [2] witness phase, [3] steps, [4] pairs.
__device__ unsigned long long g_phase_clk[8];
#define OGJK_CLK_BEGIN() const long long _clk0 = clock64()
#define OGJK_CLK_MARK(name) const long long name = clock64()
#define OGJK_CLK_ADD(slot, cyc) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_clk[slot], (unsigned long long)(cyc)); } while (0)
#else
#define OGJK_CLK_BEGIN()
#define OGJK_CLK_MARK(name)
#define OGJK_CLK_ADD(slot, cyc)
#endif

// ---- one GJK query (openGJK.c:941-1046), split into init / step / finish --------
template <typename T> struct GjkState {
  V3<T> v, sa, sb;   // search direction, carried support points of body 1 / body 2
  int ia, ib;        // their vertex indices
  Splx<T> s;
  T wmax2;
  int k;
};

// :957-980
template <typename T, typename Body>
OGJK_DI void gjk_init(const Body& b1, const Body& b2, GjkState<T>& st) {
  st.sa = b1.vertex(0);
  st.sb = b2.vertex(0);
  st.ia = 0; st.ib = 0;
  st.v = sub(st.sa, st.sb);
  const MV<T> zero = MV<T>{(T)0, (T)0, (T)0, 0, 0};
  st.s.n = 1;
  st.s.q0 = MV<T>{st.v.x, st.v.y, st.v.z, 0, 0};
  st.s.q1 = zero; st.s.q2 = zero; st.s.q3 = zero;
  st.wmax2 = (T)0;
  st.k = 0;
}

// The part of one do-while pass (:983-1036) that follows the two support calls.
// Returns true when the query is finished.
template <typename T>
OGJK_DI bool gjk_step_tail(GjkState<T>& st) {
  const T er = eps_rel<T>(), ea = eps_abs<T>();
  const V3<T> w = sub(st.sa, st.sb);
  const T vv = nrm2(st.v.x, st.v.y, st.v.z);
  const T gap = vv - dot3(st.v, w);
  if (gap <= (er * vv) || gap < ea) return true;   // :1002-1006
  if (vv < er * er) return true;                     // :1008-1010
  push_and_solve(st.s, MV<T>{w.x, w.y, w.z, st.ia, st.ib}, st.v);
  st.wmax2 = update_wmax(st.s, st.wmax2);
  if (nrm2(st.v.x, st.v.y, st.v.z) <= (ea * ea * st.wmax2)) return true;  // :1032
  return st.s.n == 4 || st.k == kMaxIter;
}

// One pass of the do-while at :983-1036.  Returns true when the query is finished.
template <typename T, typename Body>
OGJK_DI bool gjk_step(const Body& b1, const Body& b2, GjkState<T>& st) {
  const T er = eps_rel<T>(), ea = eps_abs<T>();
  st.k++;
  OGJK_CLK_BEGIN();
  support_both(b1, b2, st.v, st.sa, st.ia, st.sb, st.ib);
  OGJK_CLK_MARK(_clk1);
  OGJK_CLK_ADD(0, _clk1 - _clk0);
  OGJK_CLK_ADD(3, 1);
#ifdef OGJK_PHASE_CLOCKS
  const bool done = gjk_step_tail(st);
  OGJK_CLK_ADD(1, clock64() - _clk1);
  return done;
#else
  const V3<T> w = sub(st.sa, st.sb);
  const T vv = nrm2(st.v.x, st.v.y, st.v.z);
  const T gap = vv - dot3(st.v, w);
  if (gap <= (er * vv) || gap < ea) return true;   // :1002-1006
  if (vv < er * er) return true;                     // :1008-1010
  push_and_solve(st.s, MV<T>{w.x, w.y, w.z, st.ia, st.ib}, st.v);
  st.wmax2 = update_wmax(st.s, st.wmax2);
  if (nrm2(st.v.x, st.v.y, st.v.z) <= (ea * ea * st.wmax2)) return true;  // :1032
  return st.s.n == 4 || st.k == kMaxIter;
#endif
}

// compute_witnesses (:921-939): may reduce `s`; blends the source vertices.
template <typename T, typename Body>
OGJK_DI void gjk_witnesses(const Body& b1, const Body& b2, Splx<T>& s, V3<T>& wit1, V3<T>& wit2) {
  const Wts<T> a = witness_weights(s);
  const V3<T> p0 = b1.vertex(s.q0.ia), r0 = b2.vertex(s.q0.ib);
  if (a.terms == 1) {
    wit1 = p0; wit2 = r0;
  } else {
    const V3<T> p1 = b1.vertex(s.q1.ia), r1 = b2.vertex(s.q1.ib);
    if (a.terms == 2) {
      wit1 = V3<T>{p0.x * a.a0 + p1.x * a.a1, p0.y * a.a0 + p1.y * a.a1, p0.z * a.a0 + p1.z * a.a1};
      wit2 = V3<T>{r0.x * a.a0 + r1.x * a.a1, r0.y * a.a0 + r1.y * a.a1, r0.z * a.a0 + r1.z * a.a1};
    } else {
      const V3<T> p2 = b1.vertex(s.q2.ia), r2 = b2.vertex(s.q2.ib);
      if (a.terms == 3) {
        wit1 = V3<T>{p0.x * a.a0 + p1.x * a.a1 + p2.x * a.a2, p0.y * a.a0 + p1.y * a.a1 + p2.y * a.a2,
                     p0.z * a.a0 + p1.z * a.a1 + p2.z * a.a2};
        wit2 = V3<T>{r0.x * a.a0 + r1.x * a.a1 + r2.x * a.a2, r0.y * a.a0 + r1.y * a.a1 + r2.y * a.a2,
                     r0.z * a.a0 + r1.z * a.a1 + r2.z * a.a2};
      } else {
        const V3<T> p3 = b1.vertex(s.q3.ia), r3 = b2.vertex(s.q3.ib);
        wit1 = V3<T>{p0.x * a.a0 + p1.x * a.a1 + p2.x * a.a2 + p3.x * a.a3,
                     p0.y * a.a0 + p1.y * a.a1 + p2.y * a.a2 + p3.y * a.a3,
                     p0.z * a.a0 + p1.z * a.a1 + p2.z * a.a2 + p3.z * a.a3};
        wit2 = V3<T>{r0.x * a.a0 + r1.x * a.a1 + r2.x * a.a2 + r3.x * a.a3,
                     r0.y * a.a0 + r1.y * a.a1 + r2.y * a.a2 + r3.y * a.a3,
                     r0.z * a.a0 + r1.z * a.a1 + r2.z * a.a2 + r3.z * a.a3};
      }
    }
  }
}

// Returns the distance; fills `s` (post-witness simplex) and the witnesses.
template <typename T, typename Body>
OGJK_DI T gjk_query(const Body& b1, const Body& b2, Splx<T>& s, V3<T>& wit1, V3<T>& wit2) {
  GjkState<T> st;
  gjk_init(b1, b2, st);
  while (!gjk_step(b1, b2, st)) {}
  s = st.s;
  OGJK_CLK_BEGIN();
  gjk_witnesses(b1, b2, s, wit1, wit2);
  OGJK_CLK_ADD(2, clock64() - _clk0);
  OGJK_CLK_ADD(4, 1);
  return sqrt(nrm2(st.v.x, st.v.y, st.v.z));  // :1045 (IEEE sqrt; double rounding via double is innocuous)
}

template <typename T>
OGJK_DI void store_simplex(GkSimplex<T>* o, const Splx<T>& s, const V3<T>& w1, const V3<T>& w2) {
  o->nvrtx = s.n;
  if (sizeof(T) == 8) reinterpret_cast<int*>(o)[1] = 0;  // fp64 layout pads nvrtx to 8 B: keep output bytes deterministic
  o->vrtx[0][0] = s.q0.x; o->vrtx[0][1] = s.q0.y; o->vrtx[0][2] = s.q0.z;
  o->vrtx[1][0] = s.q1.x; o->vrtx[1][1] = s.q1.y; o->vrtx[1][2] = s.q1.z;
  o->vrtx[2][0] = s.q2.x; o->vrtx[2][1] = s.q2.y; o->vrtx[2][2] = s.q2.z;
  o->vrtx[3][0] = s.q3.x; o->vrtx[3][1] = s.q3.y; o->vrtx[3][2] = s.q3.z;
  o->vrtx_idx[0][0] = s.q0.ia; o->vrtx_idx[0][1] = s.q0.ib;
  o->vrtx_idx[1][0] = s.q1.ia; o->vrtx_idx[1][1] = s.q1.ib;
  o->vrtx_idx[2][0] = s.q2.ia; o->vrtx_idx[2][1] = s.q2.ib;
  o->vrtx_idx[3][0] = s.q3.ia; o->vrtx_idx[3][1] = s.q3.ib;
  o->witnesses[0][0] = w1.x; o->witnesses[0][1] = w1.y; o->witnesses[0][2] = w1.z;
  o->witnesses[1][0] = w2.x; o->witnesses[1][1] = w2.y; o->witnesses[1][2] = w2.z;
}

// Work queue slots inside the context's counter block.
enum { Q_THREAD_HEAD = 0, Q_WARP_HEAD = 1, Q_SMALL_COUNT = 2, Q_LARGE_COUNT = 3, Q_SLOTS = 8 };

template <typename T>
__global__ void __launch_bounds__(128)
gjk_thread_kernel(BatchArgs<T> a, const int* __restrict__ list, const unsigned long long* __restrict__ list_count,
                  unsigned long long* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const long long total = list_count ? (long long)*list_count : a.npairs;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(head, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base >= total) break;
    const long long slot = (long long)base + lane;
    if (slot < total) {
      const long long p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      ThreadBody<T> b1{a.coords1 + 3 * a.off1[h1], a.cnt1[h1]};
      ThreadBody<T> b2{a.coords2 + 3 * a.off2[h2], a.cnt2[h2]};
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      a.distances[p] = d;
      if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
gjk_warp_kernel(BatchArgs<T> a, const int* __restrict__ list, const unsigned long long* __restrict__ list_count,
                unsigned long long* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const long long total = list_count ? (long long)*list_count : a.npairs;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(head, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((long long)slot >= total) break;
    const long long p = list ? (long long)list[slot] : (long long)slot;
    const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
    const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
    WarpBody<T> b1{a.coords1 + 3 * a.off1[h1], a.cnt1[h1], lane};
    WarpBody<T> b2{a.coords2 + 3 * a.off2[h2], a.cnt2[h2], lane};
    Splx<T> s;
    V3<T> w1, w2;
    const T d = gjk_query<T>(b1, b2, s, w1, w2);
    if (lane == 0) {
      a.distances[p] = d;
      if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
    }
  }
}

// =============================================================================
// Kernels over the packed polytope store (store.cuh)
// =============================================================================

template <typename T> struct StoreArgs {
  StoreView<T> s1, s2;
  const int* idx1; const int* idx2;
  GkSimplex<T>* simplices;  // may be null
  T* distances;
  long long npairs;
};

// One 16-byte chunk of a coordinate array: 4 fp32 or 2 fp64 vertices.
template <typename T> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int kVerts = 4;
  float v[4];
  OGJK_DI void load(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Chunk<double> {
  static constexpr int kVerts = 2;
  double v[2];
  OGJK_DI void load(const double* p) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
};

// G cooperating lanes (a power of two, 1..32) scan one polytope block:
// lane g takes chunks g, g+G, ... (ascending vertex index within the lane), then
// the group reduces with a (value, lowest index) butterfly.  Every lane of the
// group ends with the same support point, so the rest of the GJK iteration runs
// convergently inside the group and nothing has to be broadcast.
template <typename T, int G> struct GroupBody {
  const T* blk; int n; int np; int g; unsigned mask;
  OGJK_DI V3<T> vertex(int i) const { return V3<T>{blk[i], blk[np + i], blk[2 * np + i]}; }
  OGJK_DI void support(T dx, T dy, T dz, V3<T>& cur, int& cur_i) const {
    const int better = scan(dx, dy, dz, dot3(cur.x, cur.y, cur.z, dx, dy, dz));
    if (better != 0x7fffffff) { cur = vertex(better); cur_i = better; }
  }
  // Lowest index of the vertex with the largest dot product, provided it strictly exceeds
  // `best` (the carried support point's dot); 0x7fffffff when nothing does.  Same value in
  // every lane of the group.
  OGJK_DI int scan(T dx, T dy, T dz, T best) const {
    constexpr int CV = Chunk<T>::kVerts;
    int better = 0x7fffffff;
    const int nchunks = np / CV;
    const T* xp = blk;
    const T* yp = blk + np;
    const T* zp = blk + 2 * np;
    // Two chunks per trip.  The common case -- nothing in the chunk pair beats the
    // running best -- costs one max-tree and one compare; the exact "strictly
    // greater, ascending index" update runs only when the max says it must.
    int c = g;
    for (; c + G < nchunks; c += 2 * G) {
      Chunk<T> X0, Y0, Z0, X1, Y1, Z1;
      X0.load(xp + CV * c); Y0.load(yp + CV * c); Z0.load(zp + CV * c);
      X1.load(xp + CV * (c + G)); Y1.load(yp + CV * (c + G)); Z1.load(zp + CV * (c + G));
      T s0[CV], s1[CV];
#pragma unroll
      for (int t = 0; t < CV; ++t) {
        s0[t] = dot3(X0.v[t], Y0.v[t], Z0.v[t], dx, dy, dz);
        s1[t] = dot3(X1.v[t], Y1.v[t], Z1.v[t], dx, dy, dz);
      }
      T m = fmax(s0[0], s1[0]);
#pragma unroll
      for (int t = 1; t < CV; ++t) m = fmax(m, fmax(s0[t], s1[t]));
      if (m > best) {
#pragma unroll
        for (int t = 0; t < CV; ++t)
          if (s0[t] > best) { best = s0[t]; better = CV * c + t; }
#pragma unroll
        for (int t = 0; t < CV; ++t)
          if (s1[t] > best) { best = s1[t]; better = CV * (c + G) + t; }
      }
    }
    if (c < nchunks) {
      Chunk<T> X, Y, Z;
      X.load(xp + CV * c); Y.load(yp + CV * c); Z.load(zp + CV * c);
#pragma unroll
      for (int t = 0; t < CV; ++t) {
        const T sc = dot3(X.v[t], Y.v[t], Z.v[t], dx, dy, dz);
        if (sc > best) { best = sc; better = CV * c + t; }
      }
    }
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) {
      const T ob = __shfl_xor_sync(mask, best, m);
      const int oi = __shfl_xor_sync(mask, better, m);
      if (ob > best || (ob == best && oi < better)) { best = ob; better = oi; }
    }
    return better;
  }
};

// Warp-per-pair body with the store's cluster table (store.cuh, "cluster tables").
// Polytopes without a table have acc == null.
// ---- pruned support scans over the store's cluster tables (store.cuh) -------------
// Vector loads of N consecutive elements (N * sizeof <= 32 bytes, naturally aligned).
template <int N> OGJK_DI void load_vec(const float* p, float (&v)[N]) {
  if constexpr (N == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (N == 2) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
  else { v[0] = p[0]; }
}
template <int N> OGJK_DI void load_vec(const double* p, double (&v)[N]) {
  if constexpr (N == 4) {
    const double2 t = *reinterpret_cast<const double2*>(p), u = *reinterpret_cast<const double2*>(p + 2);
    v[0] = t.x; v[1] = t.y; v[2] = u.x; v[3] = u.y;
  } else if constexpr (N == 2) { const double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y; }
  else { v[0] = p[0]; }
}
template <int N> OGJK_DI void load_vec(const int* p, int (&v)[N]) {
  if constexpr (N == 4) { const int4 t = *reinterpret_cast<const int4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (N == 2) { const int2 t = *reinterpret_cast<const int2*>(p); v[0] = t.x; v[1] = t.y; }
  else { v[0] = p[0]; }
}

// Order-preserving key of a float / double (after -0 -> +0, so that equal values have equal
// keys); group-wide max and "lowest index among the lanes holding the max" then run on the
// redux unit instead of shuffle butterflies.  NaN keys sort above +inf / below -inf: a NaN
// bound only disables pruning, and a lane's best value is never NaN unless every lane's is.
OGJK_DI unsigned okey(float f) {
  const unsigned u = __float_as_uint(f + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
OGJK_DI float unkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
OGJK_DI unsigned long long okey(double f) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(f + 0.0);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
OGJK_DI double unkey(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
OGJK_DI unsigned group_max_key(unsigned gm, unsigned k) { return __reduce_max_sync(gm, k); }
OGJK_DI unsigned long long group_max_key(unsigned gm, unsigned long long k) {
  const unsigned hi = (unsigned)(k >> 32);
  const unsigned mh = __reduce_max_sync(gm, hi);
  const unsigned ml = __reduce_max_sync(gm, hi == mh ? (unsigned)k : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

// G-lane body (G = 8, 16, 32) with the polytope's cluster table; acc == null when it has none.
template <typename T, int G> struct AccelBody {
  GroupBody<T, G> base;
  const T* acc;
  OGJK_DI V3<T> vertex(int i) const { return base.vertex(i); }
};

// One body's half of the fused, pruned support scan below.  A lane owns VPL = 32/G
// consecutive clusters of every 32-cluster tile and VPL consecutive vertices of every
// cluster record, so all its loads are single 4..32-byte vectors.
template <typename T, int G> struct ScanSide {
  static constexpr int VPL = 32 / G;
  const T* acc; int nc, ncs;
  T dx, dy, dz;
  T bl, px, py, pz;  // this lane's best value and its vertex
  int il;            // original index of that vertex (-1: still the carried point)
  struct Hit { T sc[VPL], x[VPL], y[VPL], z[VPL]; int o[VPL]; };
  OGJK_DI void open(const AccelBody<T, G>& b, T ex, T ey, T ez, const V3<T>& cur) {
    acc = b.acc;
    nc = (b.base.n + kAccelCluster - 1) / kAccelCluster;
    ncs = accel_table_stride(nc);
    dx = ex; dy = ey; dz = ez;
    bl = dot3(cur.x, cur.y, cur.z, dx, dy, dz);
    px = cur.x; py = cur.y; pz = cur.z;
    il = -1;
  }
  // Upper bounds of this lane's clusters of tile t: the box corner picked by the signs of d,
  // multiplied and summed in the order of dot3 (rounding is monotone, so each bound holds
  // for the rounded dots of the cluster's vertices too).
  OGJK_DI void bounds(int t, int g, T (&ub)[VPL]) const {
    const int c0 = t * 32 + g * VPL;
    T cx[VPL], cy[VPL], cz[VPL];
    load_vec<VPL>(acc + (dx >= (T)0 ? 3 : 0) * ncs + c0, cx);
    load_vec<VPL>(acc + (dy >= (T)0 ? 4 : 1) * ncs + c0, cy);
    load_vec<VPL>(acc + (dz >= (T)0 ? 5 : 2) * ncs + c0, cz);
#pragma unroll
    for (int j = 0; j < VPL; ++j) ub[j] = dot3(cx[j], cy[j], cz[j], dx, dy, dz);
  }
  OGJK_DI void fetch(int c, int g, Hit& h) const {
    const T* r = acc + 6 * ncs + (long long)c * kAccelRecord<T> + g * VPL;
    load_vec<VPL>(r, h.x);
    load_vec<VPL>(r + kAccelCluster, h.y);
    load_vec<VPL>(r + 2 * kAccelCluster, h.z);
    load_vec<VPL>(reinterpret_cast<const int*>(r - g * VPL + 3 * kAccelCluster) + g * VPL, h.o);
#pragma unroll
    for (int j = 0; j < VPL; ++j) h.sc[j] = dot3(h.x[j], h.y[j], h.z[j], dx, dy, dz);
  }
  OGJK_DI void take(const Hit& h, bool valid) {
#pragma unroll
    for (int j = 0; j < VPL; ++j)
      if (valid && (h.sc[j] > bl || (h.sc[j] == bl && h.o[j] < il))) {
        bl = h.sc[j]; il = h.o[j]; px = h.x[j]; py = h.y[j]; pz = h.z[j];
      }
  }
  OGJK_DI T group_best(unsigned gm) const { return unkey(group_max_key(gm, okey(bl))); }
};

#ifndef OGJK_ACCEL_K
#define OGJK_ACCEL_K 0  // clusters per body per trip; 0 = by G (1, 2, 3 for G = 8, 16, 32: tools/accel_sweep.sh)
#endif

// Both support calls of one iteration for a G-lane query over two bodies with cluster
// tables, run in lockstep so that the two bodies' memory round trips overlap:
//   1. bounds of 32 clusters per body;
//   2. the vertices of the cluster with the largest bound (first tile only);
//   3. the clusters whose bound is not below the running maximum, a few per trip.
// All loads of a step are unconditional (indices clamped, results masked) so that they are
// issued back to back.  A skipped cluster cannot hold a vertex that beats or ties the
// maximum, so the result is the full scan's: the largest dot if it strictly exceeds the
// carried point's, lowest original index among equals.  The winner's coordinates travel by
// shuffle.  Without two tables both bodies take the plain scan.
template <typename T, int G>
OGJK_DI void support_both(const AccelBody<T, G>& b1, const AccelBody<T, G>& b2, const V3<T>& v, V3<T>& sa, int& ia,
                          V3<T>& sb, int& ib) {
  if (!(b1.acc && b2.acc)) {
    b1.base.support(-v.x, -v.y, -v.z, sa, ia);
    b2.base.support(v.x, v.y, v.z, sb, ib);
    return;
  }
  constexpr int VPL = 32 / G;
  constexpr int K = OGJK_ACCEL_K ? OGJK_ACCEL_K : (G == 8 ? 1 : G == 16 ? 2 : 3);
  using Side = ScanSide<T, G>;
  using Hit = typename Side::Hit;
  const int g = b1.base.g;
  const unsigned gm = b1.base.mask;
  Side s[2];
  s[0].open(b1, -v.x, -v.y, -v.z, sa);
  s[1].open(b2, v.x, v.y, v.z, sb);
  const int tiles = ((s[0].nc > s[1].nc ? s[0].nc : s[1].nc) + 31) >> 5;
  for (int t = 0; t < tiles; ++t) {
    T ub[2][VPL];
    unsigned has[2];  // bit j: this lane's j-th cluster of the tile exists
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int tt = t * 32 < s[k].nc ? t : 0;  // a body with fewer tiles re-reads tile 0, masked off
      s[k].bounds(tt, g, ub[k]);
      has[k] = 0;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (tt == t && t * 32 + g * VPL + j < s[k].nc) has[k] |= 1u << j;
    }
    int top[2] = {-1, -1};
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        auto mk = okey((T)-INFINITY);
#pragma unroll
        for (int j = 0; j < VPL; ++j)
          if (((has[k] >> j) & 1u) && okey(ub[k][j]) > mk) mk = okey(ub[k][j]);
        const auto m = group_max_key(gm, mk);
        unsigned mine = 0xffffffffu;
#pragma unroll
        for (int j = VPL - 1; j >= 0; --j)
          if (((has[k] >> j) & 1u) && okey(ub[k][j]) == m) mine = (unsigned)(g * VPL + j);
        const unsigned first = __reduce_min_sync(gm, mine);
        top[k] = first == 0xffffffffu ? 0 : (int)first;
      }
      Hit h[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) s[k].fetch(top[k], g, h[k]);
#pragma unroll
      for (int k = 0; k < 2; ++k) s[k].take(h[k], true);
    }
    unsigned live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const T wb = s[k].group_best(gm);
      unsigned bits = 0;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (((has[k] >> j) & 1u) && !(ub[k][j] < wb)) bits |= 1u << (g * VPL + j);
      live[k] = __reduce_or_sync(gm, bits);
      if (top[k] >= 0) live[k] &= ~(1u << top[k]);
    }
    while (live[0] | live[1]) {
      OGJK_CLK_ADD(5, __popc(live[0]) + __popc(live[1]));  // debug build: clusters still live at trip start
      OGJK_CLK_ADD(6, 1);                                   // trips
      int cl[2][K];
      Hit h[2][K];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          cl[k][j] = live[k] ? t * 32 + __ffs(live[k]) - 1 : -1;
          live[k] &= live[k] - 1;
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < K; ++j) s[k].fetch(cl[k][j] < 0 ? 0 : cl[k][j], g, h[k][j]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < K; ++j) s[k].take(h[k][j], cl[k][j] >= 0);
      }
      if (live[0] | live[1]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const T wb = s[k].group_best(gm);
          unsigned bits = 0;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (!(ub[k][j] < wb)) bits |= 1u << (g * VPL + j);
          live[k] &= __reduce_or_sync(gm, bits);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const auto key = okey(s[k].bl);
    const bool cand = key == group_max_key(gm, key);
    const unsigned fi = __reduce_min_sync(gm, cand ? (unsigned)s[k].il : 0xffffffffu);
    if (fi == 0xffffffffu) continue;  // nothing beats the carried point
    const int src = __ffs(__ballot_sync(gm, cand && (unsigned)s[k].il == fi)) - 1;
    const V3<T> w{__shfl_sync(gm, s[k].px, src), __shfl_sync(gm, s[k].py, src), __shfl_sync(gm, s[k].pz, src)};
    if (k == 0) { sa = w; ia = (int)fi; } else { sb = w; ib = (int)fi; }
  }
}

// Segment [*seg_begin, *seg_end) of the size-sorted pair list (or the identity
// list when `list` is null) is consumed through a device-side work queue, 32/G
// pairs per grab.
#ifndef OGJK_LB_GROUP_F32
#define OGJK_LB_GROUP_F32 4
#endif
#ifndef OGJK_LB_GROUP_F64
#define OGJK_LB_GROUP_F64 4
#endif
#ifndef OGJK_LB_TABLE_F64  // table kernels: the fused two-body pruned scan needs the registers in fp64
#define OGJK_LB_TABLE_F64 3
#endif
#ifndef OGJK_LB_TABLE_F32
#define OGJK_LB_TABLE_F32 4
#endif
#ifndef OGJK_LB_SMALL_F32
#define OGJK_LB_SMALL_F32 4
#endif
template <typename T, int G>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_GROUP_F64 : OGJK_LB_GROUP_F32))
gjk_group_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                 const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  constexpr int PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    OGJK_CLK_MARK(_grab0);
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)PER_WARP);
    got = __shfl_sync(0xffffffffu, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + lane / G;
    if (slot < end) {
      const long long p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
      GroupBody<T, G> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), g, mask};
      GroupBody<T, G> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), g, mask};
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      if (g == 0) {
        a.distances[p] = d;
        if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
      }
#ifdef OGJK_PHASE_CLOCKS
      OGJK_CLK_ADD(7, clock64() - _grab0);  // whole grab: queue, pair setup, query, stores
#endif
    }
  }
}

// ---- multi-pass scheduling of the group kernel ---------------------------------------
// The pairs of a warp finish after different numbers of GJK iterations (1..25, mean 5 on
// 8-64-vertex hulls), and a warp keeps iterating until its slowest pair is done: the plain
// kernel runs at 13 of 32 active lanes.  Here a pass runs at most `iter_cap` iterations; a pair
// that is not finished by then parks its iteration state (24 values + 12 ints) and its id in
// a compacted list, and the next pass resumes the parked pairs densely packed.  Every pair
// executes exactly the instruction sequence of the one-pass kernel, so results are unchanged.
constexpr int kParkF = 24;  // v, sa, sb, q0..q3 xyz, wmax2 (+2 pad: 16-byte rows in fp32 too)
constexpr int kParkI = 12;  // ia, ib, n, k, q0..q3 (ia, ib)

template <typename T> OGJK_DI void park_state(const GjkState<T>& st, T* f, int* i) {
  const MV<T>&q0 = st.s.q0, &q1 = st.s.q1, &q2 = st.s.q2, &q3 = st.s.q3;
  const T vals[kParkF] = {st.v.x, st.v.y, st.v.z, st.sa.x, st.sa.y, st.sa.z, st.sb.x, st.sb.y, st.sb.z,
                          q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, st.wmax2, (T)0, (T)0};
  constexpr int CV = Chunk<T>::kVerts;
#pragma unroll
  for (int c = 0; c < kParkF / CV; ++c) {
    if constexpr (CV == 4) *reinterpret_cast<float4*>(f + 4 * c) = make_float4(vals[4 * c], vals[4 * c + 1], vals[4 * c + 2], vals[4 * c + 3]);
    else *reinterpret_cast<double2*>(f + 2 * c) = make_double2(vals[2 * c], vals[2 * c + 1]);
  }
  int4* o = reinterpret_cast<int4*>(i);
  o[0] = make_int4(st.ia, st.ib, st.s.n, st.k);
  o[1] = make_int4(q0.ia, q0.ib, q1.ia, q1.ib);
  o[2] = make_int4(q2.ia, q2.ib, q3.ia, q3.ib);
}

template <typename T> OGJK_DI void unpark_state(GjkState<T>& st, const T* f, const int* i) {
  T vals[kParkF];
  constexpr int CV = Chunk<T>::kVerts;
#pragma unroll
  for (int c = 0; c < kParkF / CV; ++c) {
    Chunk<T> ch;
    ch.load(f + CV * c);
#pragma unroll
    for (int t = 0; t < CV; ++t) vals[CV * c + t] = ch.v[t];
  }
  const int4* in = reinterpret_cast<const int4*>(i);
  const int4 a = in[0], b = in[1], c = in[2];
  st.v = V3<T>{vals[0], vals[1], vals[2]};
  st.sa = V3<T>{vals[3], vals[4], vals[5]};
  st.sb = V3<T>{vals[6], vals[7], vals[8]};
  st.ia = a.x; st.ib = a.y; st.s.n = a.z; st.k = a.w;
  st.s.q0 = MV<T>{vals[9], vals[10], vals[11], b.x, b.y};
  st.s.q1 = MV<T>{vals[12], vals[13], vals[14], b.z, b.w};
  st.s.q2 = MV<T>{vals[15], vals[16], vals[17], c.x, c.y};
  st.s.q3 = MV<T>{vals[18], vals[19], vals[20], c.z, c.w};
  st.wmax2 = vals[21];
}

template <typename T> struct PassArgs {
  int resume;          // 0: fresh pairs of the class segment; 1: `list` holds parked pair ids, state row = list position
  int iter_cap;        // park pairs that are still running after this many iterations in total (kMaxIter: never park)
  const T* in_f; const int* in_i;          // parked state to resume from
  int* out_list; unsigned* out_count;      // where still-unfinished pairs are parked
  T* out_f; int* out_i;
};

template <typename T, int G>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_GROUP_F64 : OGJK_LB_GROUP_F32))
gjk_pass_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head, PassArgs<T> ps) {
  constexpr int PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)PER_WARP);
    got = __shfl_sync(0xffffffffu, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + lane / G;
    bool parked = false;
    long long p = 0;
    GjkState<T> st;
    if (slot < end) {
      p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
      GroupBody<T, G> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), g, mask};
      GroupBody<T, G> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), g, mask};
      if (ps.resume) unpark_state(st, ps.in_f + slot * kParkF, ps.in_i + slot * kParkI);
      else gjk_init(b1, b2, st);
      bool done;
      do { done = gjk_step(b1, b2, st); } while (!done && st.k < ps.iter_cap);
      if (done) {
        Splx<T> s = st.s;
        V3<T> w1, w2;
        gjk_witnesses(b1, b2, s, w1, w2);
        if (g == 0) {
          a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
          if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
        }
      } else {
        parked = true;
      }
    }
    // park the unfinished pairs of this warp: one atomic per warp, rows in lane order
    const unsigned pm = __ballot_sync(0xffffffffu, parked && g == 0);
    if (pm) {
      unsigned row0 = 0;
      if (lane == 0) row0 = atomicAdd(ps.out_count, (unsigned)__popc(pm));
      row0 = __shfl_sync(0xffffffffu, row0, 0);
      if (parked && g == 0) {
        const long long row = (long long)row0 + __popc(pm & ((1u << lane) - 1u));
        ps.out_list[row] = (int)p;
        park_state(st, ps.out_f + row * kParkF, ps.out_i + row * kParkI);
      }
    }
  }
}

// The group kernel for the pairs whose two bodies both carry a cluster table (persistent
// stores, large polytopes; the size sort puts them in a segment of their own): same queue
// and outputs, pruned support scans.  A pair without two tables would take the plain scans.
template <typename T, int G>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_TABLE_F64 : OGJK_LB_TABLE_F32))
gjk_table_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                 const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  constexpr int PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    OGJK_CLK_MARK(_grab0);
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)PER_WARP);
    got = __shfl_sync(0xffffffffu, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + lane / G;
    if (slot < end) {
      const long long p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
      const bool t1 = a.s1.accel && n1 >= a.s1.accel_min && n1 <= kAccelMaxVerts;
      const bool t2 = a.s2.accel && n2 >= a.s2.accel_min && n2 <= kAccelMaxVerts;
      AccelBody<T, G> b1{{a.s1.data + a.s1.boff[h1], n1, pad4(n1), g, mask}, t1 ? a.s1.accel + a.s1.aoff[h1] : nullptr};
      AccelBody<T, G> b2{{a.s2.data + a.s2.boff[h2], n2, pad4(n2), g, mask}, t2 ? a.s2.accel + a.s2.aoff[h2] : nullptr};
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      if (g == 0) {
        a.distances[p] = d;
        if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
      }
#ifdef OGJK_PHASE_CLOCKS
      OGJK_CLK_ADD(7, clock64() - _grab0);  // whole grab: queue, pair setup, query, stores
#endif
    }
  }
}

// Shared-memory staged flavour of the group kernel (north_star item 1: "staged into
// shared memory").  GJK re-reads both polytopes once per iteration (3..8 times per
// query); with tens of thousands of pairs in flight those re-reads overflow L1 and
// L2 and the first-generation kernels stalled on `long_scoreboard`.  Here every
// G-lane group owns a shared-memory slot: the pair's two store blocks (each one
// contiguous, 16-byte aligned range) are copied in once with 128-bit loads, and
// every support scan and witness lookup then reads shared memory.  Slots are
// `slot_bytes` apart with slot_bytes = 16*G (mod 128), which keeps the 128-bit
// shared loads of neighbouring groups on disjoint banks.  The host routes a size class
// here only when every pair of the class fits a slot.
template <typename T, int G>
__global__ void __launch_bounds__(128)
gjk_staged_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                  const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head, int slot_bytes) {
  extern __shared__ uint4 ogjk_stage[];
  constexpr int PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  uint4* slot = ogjk_stage + (size_t)(threadIdx.x / G) * (slot_bytes / 16);
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)PER_WARP);
    got = __shfl_sync(0xffffffffu, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long idx = base + lane / G;
    if (idx < end) {
      const long long p = list ? (long long)list[idx] : idx;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
      const int np1 = pad4(n1), np2 = pad4(n2);
      const T* src1 = a.s1.data + a.s1.boff[h1];
      const T* src2 = a.s2.data + a.s2.boff[h2];
      const int c1 = 3 * np1 * (int)sizeof(T) / 16, c2 = 3 * np2 * (int)sizeof(T) / 16;
      // the host only routes a class here when every pair of the class fits its slot
      // cp.async (LDGSTS): every 16-byte chunk of both blocks is put in flight at once, no
      // register staging; one wait covers the whole pair (the first version copied through
      // registers four chunks at a time and paid the L2 latency on every trip).
      const uint4* s1v = reinterpret_cast<const uint4*>(src1);
      const uint4* s2v = reinterpret_cast<const uint4*>(src2);
      const unsigned sbase = (unsigned)__cvta_generic_to_shared(slot);
      for (int c = g; c < c1; c += G)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * c), "l"(s1v + c) : "memory");
      for (int c = g; c < c2; c += G)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * (c1 + c)), "l"(s2v + c) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      const T* blk1 = reinterpret_cast<const T*>(slot);
      const T* blk2 = reinterpret_cast<const T*>(slot + c1);
      __syncwarp(mask);
      GroupBody<T, G> b1{blk1, n1, np1, g, mask};
      GroupBody<T, G> b2{blk2, n2, np2, g, mask};
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      if (g == 0) {
        a.distances[p] = d;
        if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
      }
      __syncwarp(mask);  // everyone is done reading the slot before it is refilled
    }
  }
}

// Scan-assist kernel for large hulls.  A warp owns P = 32/G pairs: the G lanes of a group
// carry one pair's GJK state and run its sub-solver (redundantly and convergently, as in
// gjk_group_kernel), but every support scan is done by ALL 32 lanes for one pair at a
// time (the group leader broadcasts the block pointers, direction and carried dot; the
// 32-lane butterfly result goes back to the group).  The warp-per-pair kernel spent ~40 %
// of its instructions in the per-iteration fixed part (sub-solver, termination tests,
// reductions); here that part is shared by P pairs while the scans keep all lanes busy
// with coalesced 512-byte reads.  Scan order and tie rule are those of GroupBody<T,32>,
// so results are unchanged.
template <typename T, int G>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? 2 : 4))
gjk_assist_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                  const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  constexpr int P = 32 / G;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)P);
    got = __shfl_sync(FULL, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + grp;
    bool active = slot < end;
    long long p = 0;
    const T* blk1 = nullptr; const T* blk2 = nullptr;
    int n1 = 0, n2 = 0;
    GjkState<T> st;
    if (active) {
      p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      n1 = a.s1.cnt[h1]; n2 = a.s2.cnt[h2];
      blk1 = a.s1.data + a.s1.boff[h1];
      blk2 = a.s2.data + a.s2.boff[h2];
    }
    const GroupBody<T, 32> own1{blk1, n1, pad4(n1), lane, FULL}, own2{blk2, n2, pad4(n2), lane, FULL};
    if (active) gjk_init(own1, own2, st);
    for (;;) {
      const unsigned live = __ballot_sync(FULL, active);
      if (!live) break;
      int win1 = 0x7fffffff, win2 = 0x7fffffff;
#pragma unroll 1
      for (int j = 0; j < P; ++j) {
        const int leader = j * G;
        if (!((live >> leader) & 1u)) continue;  // uniform
        // the leader of group j hands its pair to the whole warp
        const unsigned long long q1 = __shfl_sync(FULL, (unsigned long long)blk1, leader);
        const unsigned long long q2 = __shfl_sync(FULL, (unsigned long long)blk2, leader);
        const int m1 = __shfl_sync(FULL, n1, leader), m2 = __shfl_sync(FULL, n2, leader);
        const T vx = __shfl_sync(FULL, st.v.x, leader), vy = __shfl_sync(FULL, st.v.y, leader),
                vz = __shfl_sync(FULL, st.v.z, leader);
        const T c1 = __shfl_sync(FULL, dot3(st.sa.x, st.sa.y, st.sa.z, -st.v.x, -st.v.y, -st.v.z), leader);
        const T c2 = __shfl_sync(FULL, dot3(st.sb.x, st.sb.y, st.sb.z, st.v.x, st.v.y, st.v.z), leader);
        const GroupBody<T, 32> w1{reinterpret_cast<const T*>(q1), m1, pad4(m1), lane, FULL};
        const GroupBody<T, 32> w2{reinterpret_cast<const T*>(q2), m2, pad4(m2), lane, FULL};
        const int r1 = w1.scan(-vx, -vy, -vz, c1);
        const int r2 = w2.scan(vx, vy, vz, c2);
        if (grp == j) { win1 = r1; win2 = r2; }
      }
      if (active) {
        st.k++;
        if (win1 != 0x7fffffff) { st.sa = own1.vertex(win1); st.ia = win1; }
        if (win2 != 0x7fffffff) { st.sb = own2.vertex(win2); st.ib = win2; }
        if (gjk_step_tail(st)) {
          V3<T> w1, w2;
          Splx<T> s = st.s;
          gjk_witnesses(own1, own2, s, w1, w2);
          if (lane % G == 0) {
            a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
            if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
          }
          active = false;
        }
      }
      __syncwarp();
    }
  }
}

// CTA-per-pair kernel for large hulls (hundreds to thousands of vertices per pair).
// ncu on the warp-per-pair kernel (1000-vertex hulls): L1 hit 13 %, ~10 TB/s of L2->SM
// traffic -- every GJK iteration re-reads both hulls from L2, and that re-read, not HBM,
// bounds the kernel.  Here a CTA of WARPS warps owns one pair at a time:
//   * both store blocks are copied ONCE into shared memory with cp.async (all 16-byte
//     chunks in flight together), so HBM/L2 see each hull once per query;
//   * every support scan reads shared memory, thread t taking chunks t, t+NT, ...; a
//     warp butterfly and a WARPS-entry shared-memory pass produce the (value, lowest
//     index) winner -- the same rule as everywhere else, so results are unchanged;
//   * warp 0 alone runs the termination tests, the sub-solver and the witnesses while
//     the other warps wait at the barrier (no issue slots spent on redundant copies),
//     then publishes the next search direction.
// Shared memory per CTA is the class's largest pair, so 227 KB / slot pairs are resident
// per SM with WARPS warps each.
template <typename T, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
gjk_cta_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
               const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  extern __shared__ uint4 ogjk_stage[];
  constexpr int NT = 32 * WARPS;
  constexpr int CV = Chunk<T>::kVerts;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ long long s_slot;
  __shared__ T s_dir[3], s_carry[2];
  __shared__ int s_done;
  __shared__ T s_val[2][WARPS];
  __shared__ int s_idx[2][WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    if (tid == 0) s_slot = begin + (long long)atomicAdd(head, 1ull);
    __syncthreads();
    const long long slot = s_slot;
    if (slot >= end) break;
    const long long p = list ? (long long)list[slot] : slot;
    const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
    const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
    const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
    const int np1 = pad4(n1), np2 = pad4(n2);
    const uint4* s1v = reinterpret_cast<const uint4*>(a.s1.data + a.s1.boff[h1]);
    const uint4* s2v = reinterpret_cast<const uint4*>(a.s2.data + a.s2.boff[h2]);
    const int c1 = 3 * np1 * (int)sizeof(T) / 16, c2 = 3 * np2 * (int)sizeof(T) / 16;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(ogjk_stage);
    for (int c = tid; c < c1; c += NT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * c), "l"(s1v + c) : "memory");
    for (int c = tid; c < c2; c += NT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * (c1 + c)), "l"(s2v + c) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const T* blk1 = reinterpret_cast<const T*>(ogjk_stage);
    const T* blk2 = reinterpret_cast<const T*>(ogjk_stage + c1);
    const GroupBody<T, 32> own1{blk1, n1, np1, lane, FULL}, own2{blk2, n2, np2, lane, FULL};
    GjkState<T> st;
    if (warp == 0) {
      gjk_init(own1, own2, st);
      if (lane == 0) {
        s_dir[0] = st.v.x; s_dir[1] = st.v.y; s_dir[2] = st.v.z;
        s_carry[0] = dot3(st.sa.x, st.sa.y, st.sa.z, -st.v.x, -st.v.y, -st.v.z);
        s_carry[1] = dot3(st.sb.x, st.sb.y, st.sb.z, st.v.x, st.v.y, st.v.z);
        s_done = 0;
      }
    }
    for (;;) {
      __syncthreads();  // direction / carried dots / done flag of this round are visible
      if (s_done) break;
      const T vx = s_dir[0], vy = s_dir[1], vz = s_dir[2];
      T best1 = s_carry[0], best2 = s_carry[1];
      int bt1 = 0x7fffffff, bt2 = 0x7fffffff;
      // thread-strided scans of both blocks out of shared memory
      for (int c = tid; c < np1 / CV; c += NT) {
        Chunk<T> X, Y, Z;
        X.load(blk1 + CV * c); Y.load(blk1 + np1 + CV * c); Z.load(blk1 + 2 * np1 + CV * c);
#pragma unroll
        for (int t = 0; t < CV; ++t) {
          const T sc = dot3(X.v[t], Y.v[t], Z.v[t], -vx, -vy, -vz);
          if (sc > best1) { best1 = sc; bt1 = CV * c + t; }
        }
      }
      for (int c = tid; c < np2 / CV; c += NT) {
        Chunk<T> X, Y, Z;
        X.load(blk2 + CV * c); Y.load(blk2 + np2 + CV * c); Z.load(blk2 + 2 * np2 + CV * c);
#pragma unroll
        for (int t = 0; t < CV; ++t) {
          const T sc = dot3(X.v[t], Y.v[t], Z.v[t], vx, vy, vz);
          if (sc > best2) { best2 = sc; bt2 = CV * c + t; }
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const T o1 = __shfl_xor_sync(FULL, best1, m);
        const int i1 = __shfl_xor_sync(FULL, bt1, m);
        const T o2 = __shfl_xor_sync(FULL, best2, m);
        const int i2 = __shfl_xor_sync(FULL, bt2, m);
        if (o1 > best1 || (o1 == best1 && i1 < bt1)) { best1 = o1; bt1 = i1; }
        if (o2 > best2 || (o2 == best2 && i2 < bt2)) { best2 = o2; bt2 = i2; }
      }
      if (lane == 0) { s_val[0][warp] = best1; s_idx[0][warp] = bt1; s_val[1][warp] = best2; s_idx[1][warp] = bt2; }
      __syncthreads();
      if (warp == 0) {
        T b1 = s_carry[0], b2 = s_carry[1];
        int i1 = 0x7fffffff, i2 = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
          const T v1 = s_val[0][w]; const int k1 = s_idx[0][w];
          const T v2 = s_val[1][w]; const int k2 = s_idx[1][w];
          if (v1 > b1 || (v1 == b1 && k1 < i1)) { b1 = v1; i1 = k1; }
          if (v2 > b2 || (v2 == b2 && k2 < i2)) { b2 = v2; i2 = k2; }
        }
        st.k++;
        if (i1 != 0x7fffffff) { st.sa = own1.vertex(i1); st.ia = i1; }
        if (i2 != 0x7fffffff) { st.sb = own2.vertex(i2); st.ib = i2; }
        const bool done = gjk_step_tail(st);
        if (done) {
          V3<T> w1, w2;
          Splx<T> s = st.s;
          gjk_witnesses(own1, own2, s, w1, w2);
          if (lane == 0) {
            a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
            if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
          }
        }
        if (lane == 0) {
          s_dir[0] = st.v.x; s_dir[1] = st.v.y; s_dir[2] = st.v.z;
          s_carry[0] = dot3(st.sa.x, st.sa.y, st.sa.z, -st.v.x, -st.v.y, -st.v.z);
          s_carry[1] = dot3(st.sb.x, st.sb.y, st.sb.z, st.v.x, st.v.y, st.v.z);
          s_done = done ? 1 : 0;
        }
      }
    }
  }
}

// Thread-per-pair kernel for pairs whose two polytopes have at most 8 vertices
// each (boxes, tetrahedra, the reference's tiny random bodies): all 16 vertices
// live in registers for the whole query, so global memory is touched once per
// pair.  The support scan is fully unrolled over register slots; the winning
// vertex is carried by value, so no register array is ever indexed dynamically.
template <typename T> struct RegBody {
  T x[8], y[8], z[8];
  const T* blk; int np;
  OGJK_DI void load(const T* b, int np_) {
    constexpr int CV = Chunk<T>::kVerts;
    blk = b; np = np_;
#pragma unroll
    for (int c = 0; c < 8 / CV; ++c) {
      // a 4-vertex polytope has np == 4: its upper register slots repeat chunk 0
      const int cc = (CV * c < np_) ? c : 0;
      Chunk<T> X, Y, Z;
      X.load(b + CV * cc); Y.load(b + np_ + CV * cc); Z.load(b + 2 * np_ + CV * cc);
#pragma unroll
      for (int t = 0; t < CV; ++t) { x[CV * c + t] = X.v[t]; y[CV * c + t] = Y.v[t]; z[CV * c + t] = Z.v[t]; }
    }
  }
  OGJK_DI V3<T> vertex(int i) const { return V3<T>{blk[i], blk[np + i], blk[2 * np + i]}; }
  OGJK_DI void support(T dx, T dy, T dz, V3<T>& cur, int& cur_i) const {
    T best = dot3(cur.x, cur.y, cur.z, dx, dy, dz);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const T sc = dot3(x[i], y[i], z[i], dx, dy, dz);
      // slots >= np replicate lower-index vertices: equal value, never strictly larger
      if (sc > best) { best = sc; cur_i = i; cur.x = x[i]; cur.y = y[i]; cur.z = z[i]; }
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? 2 : OGJK_LB_SMALL_F32))
gjk_small_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                 const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, 32ull);
    got = __shfl_sync(0xffffffffu, got, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + lane;
    if (slot < end) {
      const long long p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      RegBody<T> b1, b2;
      b1.load(a.s1.data + a.s1.boff[h1], pad4(a.s1.cnt[h1]));
      b2.load(a.s2.data + a.s2.boff[h2], pad4(a.s2.cnt[h2]));
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      a.distances[p] = d;
      if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
    }
  }
}

// =============================================================================
// Persistent-lane kernels: refill on termination, witnesses in a second pass
// =============================================================================
// GJK needs 1..25 iterations per pair (mean ~5).  With one batch of pairs per warp,
// every lane waits for the slowest pair of its warp (ncu: 10-12 of 32 lanes active).
// Here a lane (or G-lane group) that finishes its pair writes the raw result -- the
// distance and the pre-witness simplex, straight into the output gkSimplex -- and
// takes the next pair from the queue at the top of the next round, so every round
// runs one GJK iteration for (almost) every lane.  The witness reconstruction,
// which is short and divergent, is left to gjk_witness_kernel over all pairs.

template <typename T>
OGJK_DI void store_raw(const StoreArgs<T>& a, long long p, const GjkState<T>& st) {
  a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
  if (a.simplices) {
    const V3<T> z{(T)0, (T)0, (T)0};
    store_simplex(a.simplices + p, st.s, z, z);
  }
}

// Warp-aggregated queue grab for the group leaders in `need` (a ballot of leader
// lanes).  Returns this group's queue position or -1.
OGJK_DI long long grab_slots(unsigned need, int leader_lane, int lane, long long begin, long long end,
                             unsigned long long* head) {
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(head, (unsigned long long)__popc(need));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (!((need >> leader_lane) & 1u)) return -1;
  const long long slot = begin + (long long)base + __popc(need & ((1u << leader_lane) - 1u));
  return slot < end ? slot : -1;
}

template <typename T, int G>
__global__ void __launch_bounds__(128, 4)
gjk_persist_group_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                         const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  GroupBody<T, G> b1{nullptr, 0, 0, g, mask}, b2{nullptr, 0, 0, g, mask};
  GjkState<T> st;
  long long p = -1;
  bool active = false, dry = false;
  for (;;) {
    __syncwarp();
    const unsigned need = __ballot_sync(0xffffffffu, !active && !dry && g == 0);
    if (need) {
      const long long slot = grab_slots(need, lane - g, lane, begin, end, head);
      if (!active && !dry) {
        if (slot >= 0) {
          p = list ? (long long)list[slot] : slot;
          const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
          const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
          const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
          b1.blk = a.s1.data + a.s1.boff[h1]; b1.n = n1; b1.np = pad4(n1);
          b2.blk = a.s2.data + a.s2.boff[h2]; b2.n = n2; b2.np = pad4(n2);
          gjk_init(b1, b2, st);
          active = true;
        } else {
          dry = true;
        }
      }
    }
    if (!__ballot_sync(0xffffffffu, active)) break;
    if (active && gjk_step(b1, b2, st)) {
      if (g == 0) store_raw(a, p, st);
      active = false;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? 2 : 4))
gjk_persist_small_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                         const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  RegBody<T> b1, b2;
  GjkState<T> st;
  long long p = -1;
  bool active = false, dry = false;
  for (;;) {
    __syncwarp();
    const unsigned need = __ballot_sync(0xffffffffu, !active && !dry);
    if (need) {
      const long long slot = grab_slots(need, lane, lane, begin, end, head);
      if (!active && !dry) {
        if (slot >= 0) {
          p = list ? (long long)list[slot] : slot;
          const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
          const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
          b1.load(a.s1.data + a.s1.boff[h1], pad4(a.s1.cnt[h1]));
          b2.load(a.s2.data + a.s2.boff[h2], pad4(a.s2.cnt[h2]));
          gjk_init(b1, b2, st);
          active = true;
        } else {
          dry = true;
        }
      }
    }
    if (!__ballot_sync(0xffffffffu, active)) break;
    if (active && gjk_step(b1, b2, st)) {
      store_raw(a, p, st);
      active = false;
    }
  }
}

// Second pass: witness reconstruction for every pair, thread per pair, in place on
// the raw simplices written by the persistent kernels.
template <typename T>
__global__ void __launch_bounds__(256)
gjk_witness_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                   const unsigned* __restrict__ seg_end) {
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
    const long long p = list ? (long long)list[i] : i;
    GkSimplex<T>* o = a.simplices + p;
    Splx<T> s;
    s.n = o->nvrtx;
    s.q0 = MV<T>{o->vrtx[0][0], o->vrtx[0][1], o->vrtx[0][2], o->vrtx_idx[0][0], o->vrtx_idx[0][1]};
    s.q1 = MV<T>{o->vrtx[1][0], o->vrtx[1][1], o->vrtx[1][2], o->vrtx_idx[1][0], o->vrtx_idx[1][1]};
    s.q2 = MV<T>{o->vrtx[2][0], o->vrtx[2][1], o->vrtx[2][2], o->vrtx_idx[2][0], o->vrtx_idx[2][1]};
    s.q3 = MV<T>{o->vrtx[3][0], o->vrtx[3][1], o->vrtx[3][2], o->vrtx_idx[3][0], o->vrtx_idx[3][1]};
    const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
    const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
    const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
    const GroupBody<T, 1> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), 0, 0u};
    const GroupBody<T, 1> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), 0, 0u};
    V3<T> w1, w2;
    gjk_witnesses(b1, b2, s, w1, w2);
    store_simplex(o, s, w1, w2);
  }
}

// Sub-solver test hook (reference: opengjk_test_S{1,2,3}D, scalar/openGJK.c:1232-1248):
// one thread runs S1D/S2D/S3D on a caller-supplied simplex, in place.
template <typename T>
__global__ void subsolver_kernel(GkSimplex<T>* io, T* v_out) {
  Splx<T> s;
  s.n = io->nvrtx;
  s.q0 = MV<T>{io->vrtx[0][0], io->vrtx[0][1], io->vrtx[0][2], io->vrtx_idx[0][0], io->vrtx_idx[0][1]};
  s.q1 = MV<T>{io->vrtx[1][0], io->vrtx[1][1], io->vrtx[1][2], io->vrtx_idx[1][0], io->vrtx_idx[1][1]};
  s.q2 = MV<T>{io->vrtx[2][0], io->vrtx[2][1], io->vrtx[2][2], io->vrtx_idx[2][0], io->vrtx_idx[2][1]};
  s.q3 = MV<T>{io->vrtx[3][0], io->vrtx[3][1], io->vrtx[3][2], io->vrtx_idx[3][0], io->vrtx_idx[3][1]};
  V3<T> v{v_out[0], v_out[1], v_out[2]};
  if (s.n == 4) { if (sub_3d(s, v)) sub_2d(s, v); }
  else if (s.n == 3) sub_2d(s, v);
  else if (s.n == 2) sub_1d(s, v);
  const V3<T> keep1{io->witnesses[0][0], io->witnesses[0][1], io->witnesses[0][2]};
  const V3<T> keep2{io->witnesses[1][0], io->witnesses[1][1], io->witnesses[1][2]};
  store_simplex(io, s, keep1, keep2);
  v_out[0] = v.x; v_out[1] = v.y; v_out[2] = v.z;
}

// Batch scheduler, stage 1: split the pair list by total vertex count so that
// thread-per-pair warps do not wait on one large hull (north_star item 5).
// Warp-aggregated appends keep the in-warp pair order.
template <typename T>
__global__ void __launch_bounds__(256)
classify_kernel(BatchArgs<T> a, int threshold, int* __restrict__ small_list, int* __restrict__ large_list,
                unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounded = (a.npairs + 31) & ~31ll;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < rounded; p += stride) {
    bool valid = p < a.npairs, big = false;
    if (valid) {
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      big = (a.cnt1[h1] + a.cnt2[h2]) >= threshold;
    }
    const unsigned mb = __ballot_sync(0xffffffffu, valid && big);
    const unsigned ms = __ballot_sync(0xffffffffu, valid && !big);
    unsigned long long bb = 0, bs = 0;
    if (lane == 0) {
      if (mb) bb = atomicAdd(&counters[Q_LARGE_COUNT], (unsigned long long)__popc(mb));
      if (ms) bs = atomicAdd(&counters[Q_SMALL_COUNT], (unsigned long long)__popc(ms));
    }
    bb = __shfl_sync(0xffffffffu, bb, 0);
    bs = __shfl_sync(0xffffffffu, bs, 0);
    const unsigned below = (1u << lane) - 1u;
    if (valid && big) large_list[bb + __popc(mb & below)] = (int)p;
    if (valid && !big) small_list[bs + __popc(ms & below)] = (int)p;
  }
}

}  // namespace ogjk
