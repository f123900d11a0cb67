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
#ifndef OGJK_SCAN_PREFETCH
#define OGJK_SCAN_PREFETCH 0  // 1: every block line of both bodies is requested into L1 before the scans (group bodies)
#endif
template <typename Body> OGJK_DI auto prefetch_block(const Body& b, int) -> decltype(b.prefetch(), void()) { b.prefetch(); }
template <typename Body> OGJK_DI void prefetch_block(const Body&, long) {}
template <typename T, typename Body>
OGJK_DI void support_both(const Body& b1, const Body& b2, const V3<T>& v, V3<T>& sa, int& ia, V3<T>& sb, int& ib) {
#if OGJK_SCAN_PREFETCH
  prefetch_block(b1, 0);
  prefetch_block(b2, 0);
#endif
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

// Group-wide maximum of a floating-point value.  fp32: sm_100a reduces floats on the redux unit
// (CREDUX.MAX.F32; NaN inputs are ignored unless every lane holds one, +0 > -0), one instruction
// instead of key conversion + integer reduction + conversion back; fp64 goes through the keys.
OGJK_DI float group_fmax(unsigned gm, float x) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "r"(gm));
  return r;
}

// G cooperating lanes (a power of two, 1..32) scan one polytope block:
// lane g takes chunks g, g+G, ... (ascending vertex index within the lane), then
// the group reduces with a (value, lowest index) butterfly.  Every lane of the
// group ends with the same support point, so the rest of the GJK iteration runs
// convergently inside the group and nothing has to be broadcast.
// WIDE (fp64 blocks in GLOBAL memory only): the scan reads 32-byte chunks (4 vertices) with 256-bit loads
// (ld.global.v4.f64 -> LDG.E.256 on sm_100a) instead of pairs of 16-byte chunks -- half the load instructions
// and half the L1TEX requests for the same bytes; with 16 pairs per warp every request touches 16 lines, and
// the L1TEX tag stage was 56 % busy on config 2.  Store blocks are 32-byte aligned in fp64 (np is a multiple of
// 4 doubles).  Which lane sees which vertex changes, the result does not (see scan()).
#ifndef OGJK_WIDE_F64
#define OGJK_WIDE_F64 1
#endif
OGJK_DI void load_wide(const double* p, double (&v)[4]) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
OGJK_DI void load_wide(const float*, float (&)[4]) {}  // never called (WIDE is fp64 only)

template <typename T, int G, bool WIDE = false> struct GroupBody {
  const T* blk; int n; int np; int g; unsigned mask;
  OGJK_DI V3<T> vertex(int i) const { return V3<T>{blk[i], blk[np + i], blk[2 * np + i]}; }
  // All 128-byte lines of the block [x | y | z] at once, lane g taking lines g, g+G, ...: the scan below
  // otherwise meets them one trip after the other, each a full L2 / DRAM round trip.
  OGJK_DI void prefetch() const {
    constexpr int kLine = 128 / (int)sizeof(T);
    const int last = 3 * np - 1;  // blocks are 16-byte, not line aligned: the last line is named by its last element
    for (int e = g * kLine; e < last + kLine; e += G * kLine) asm volatile("prefetch.global.L1 [%0];" ::"l"(blk + (e < last ? e : last)));
  }
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
    if constexpr (WIDE && sizeof(T) == 8) {
      // one 32-byte chunk (4 consecutive vertices) per trip and lane; np is a multiple of 4: no remainder
      for (int w = g; 4 * w < np; w += G) {
        T X[4], Y[4], Z[4], sc[4];
        load_wide(xp + 4 * w, X); load_wide(yp + 4 * w, Y); load_wide(zp + 4 * w, Z);
#pragma unroll
        for (int t = 0; t < 4; ++t) sc[t] = dot3(X[t], Y[t], Z[t], dx, dy, dz);
        const T m = fmax(fmax(sc[0], sc[1]), fmax(sc[2], sc[3]));
        if (m > best) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (sc[t] > best) { best = sc[t]; better = 4 * w + t; }
        }
      }
      return group_argmax(mask, best, better);
    }
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
    return group_argmax(mask, best, better);
  }
  // (largest value, lowest index among the lanes holding it) across the G lanes of the group.
  // fp32, full warp: two redux instructions (CREDUX.MAX.F32 + REDUX.MIN on the indices of the lanes that
  // hold the maximum) instead of five shuffle rounds (+4.7 % on 1000-vertex hulls).  With partial member
  // masks the redux path LOSES to the shuffles (G = 2/4/8: -17/-16/-9 %, profiles/r2j_redux_argmax_ab.jsonl),
  // so smaller groups keep the butterfly.  A lane without a candidate carries the carried point's dot and
  // index 0x7fffffff, which can only win when nobody beat the carried point.
  static OGJK_DI int group_argmax(unsigned mask, T best, int better) {
#ifndef OGJK_REDUX_ARGMAX_MIN_G
#define OGJK_REDUX_ARGMAX_MIN_G 32
#endif
    if constexpr (sizeof(T) == 4 && G >= OGJK_REDUX_ARGMAX_MIN_G) {
      const T m = group_fmax(mask, best);
      return (int)__reduce_min_sync(mask, best == m ? (unsigned)better : 0x7fffffffu);
    } else {
#pragma unroll
      for (int m = G / 2; m > 0; m >>= 1) {
        const T ob = __shfl_xor_sync(mask, best, m);
        const int oi = __shfl_xor_sync(mask, better, m);
        if (ob > best || (ob == best && oi < better)) { best = ob; better = oi; }
      }
      return better;
    }
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

OGJK_DI double group_fmax(unsigned gm, double x) { return unkey(group_max_key(gm, okey(x))); }


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
    const T* r = acc + accel_records_at(ncs) + (long long)c * kAccelRecord<T> + g * VPL;
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
  OGJK_DI T group_best(unsigned gm) const { return group_fmax(gm, bl); }
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
template <int KO = 0, typename T, int G>
OGJK_DI void support_both(const AccelBody<T, G>& b1, const AccelBody<T, G>& b2, const V3<T>& v, V3<T>& sa, int& ia,
                          V3<T>& sb, int& ib) {
  if (!(b1.acc && b2.acc)) {
    b1.base.support(-v.x, -v.y, -v.z, sa, ia);
    b2.base.support(v.x, v.y, v.z, sb, ib);
    return;
  }
  constexpr int VPL = 32 / G;
  constexpr int K = KO ? KO : OGJK_ACCEL_K ? OGJK_ACCEL_K : (G == 8 ? 1 : G == 16 ? 2 : 3);  // KO: the caller's choice
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
        T mv = (T)-INFINITY;  // NaN bounds are skipped here (they stay live below: a NaN bound never prunes)
#pragma unroll
        for (int j = 0; j < VPL; ++j)
          if (((has[k] >> j) & 1u) && ub[k][j] > mv) mv = ub[k][j];
        const T m = group_fmax(gm, mv);
        unsigned mine = 0xffffffffu;
#pragma unroll
        for (int j = VPL - 1; j >= 0; --j)
          if (((has[k] >> j) & 1u) && ub[k][j] == m) mine = (unsigned)(g * VPL + j);
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
    const bool cand = s[k].bl == group_fmax(gm, s[k].bl);  // all-NaN: no candidate, the carried point stays
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
      constexpr bool kWide = OGJK_WIDE_F64 && sizeof(T) == 8;
      GroupBody<T, G, kWide> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), g, mask};
      GroupBody<T, G, kWide> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), g, mask};
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

}  // namespace ogjk
