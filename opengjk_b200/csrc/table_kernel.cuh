// table_kernel.cuh -- warp-per-pair GJK over the store's cluster tables, second generation.
//
// The first table kernel (gjk_table_kernel<T,32>, gjk_kernels.cuh) was latency-bound on
// 1000-vertex hulls: every GJK iteration paid ~5 DEPENDENT memory round trips per body pair
// (cluster boxes -> best cluster -> its vertices -> re-prune -> survivors -> ...) on top of
// a 4-trip pair setup (queue -> list -> idx -> cnt/boff/aoff), at 16 warps per SM
// (profiles/r1m_hulls1k_f32_tables_ncu_summary.txt: long_scoreboard 2.85 of 7.4 cycles per
// issue, issue-active 54 %).  This kernel removes round trips instead of hiding them:
//
//   * the boxes of the first 32 clusters of each body live in registers for the whole
//     query (lane c holds cluster c: 6 values per body), so the bound of every cluster is
//     one dot product away in every iteration -- no memory trip;
//   * the cluster that held the last support vertex stays in registers too (lane v holds
//     vertex v of it).  GJK's search direction converges, so the next support vertex is
//     usually in the same cluster: evaluating it first costs no memory access and gives a
//     lower bound that prunes almost everything else;
//   * the surviving clusters are fetched highest-bound first, several per trip, with
//     unconditional loads issued back to back; the bound is re-tightened between trips;
//   * the NEXT pair's setup chain (queue grab -> list -> indices -> counts / offsets) is
//     advanced one dependent load per GJK iteration of the current pair, so its latency
//     is covered by work instead of being paid at the top of every pair.
//
// Exactness: a cluster is skipped only when its bound is below a value some vertex has
// already reached, so it cannot hold a vertex that beats or ties the final maximum; the
// order in which the other clusters are evaluated does not matter because the running
// best is a total order on (dot, lowest original index) with the carried support point
// ranked first among equals (openGJK.c:607-630).  Results are bit-identical to the plain
// scans (tests/test_gpu_accel.py).
#pragma once
#include "gjk_kernels.cuh"

namespace ogjk {

#ifndef OGJK_T2_K
#define OGJK_T2_K 3  // clusters fetched per body per trip
#endif
#ifndef OGJK_T2_CACHE
#define OGJK_T2_CACHE 1  // keep the last winner's cluster in registers
#endif
#ifndef OGJK_T2_PREFETCH
#define OGJK_T2_PREFETCH 1  // software-pipelined setup of the next pair
#endif
#ifndef OGJK_T2_PARK
#define OGJK_T2_PARK 0  // park the simplex in shared memory while the support scans run
#endif
#ifndef OGJK_LB_TABLE2_F32
#define OGJK_LB_TABLE2_F32 4
#endif
#ifndef OGJK_LB_TABLE2_F64
#define OGJK_LB_TABLE2_F64 3
#endif

// Per-body state of one query.
template <typename T> struct T2Side {
  const T* acc;    // table: lox loy loz hix hiy hiz, each ncs long; records follow
  int nc, ncs;
  T lo[3], hi[3];  // lane c: box of cluster c (first tile)
#if OGJK_T2_CACHE
  T cx, cy, cz;    // lane v: vertex v of the cached cluster
  int co, cc;      // its original index; cached cluster id (-1: none)
#endif
  OGJK_DI const T* record(int c) const { return acc + 6 * ncs + (long long)c * kAccelRecord<T>; }
  OGJK_DI void open(const T* table, int n, int lane) {
    acc = table;
    nc = (n + kAccelCluster - 1) / kAccelCluster;
    ncs = accel_table_stride(nc);
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = acc[a * ncs + lane]; hi[a] = acc[(3 + a) * ncs + lane]; }
#if OGJK_T2_CACHE
    cc = -1; cx = cy = cz = (T)0; co = 0;
#endif
  }
};

template <typename T> struct T2Hit { T sc, x, y, z; int o; };

// Running best of one lane: value, original index (-1 = the carried point), cluster, vertex.
template <typename T> struct T2Best {
  T bl, px, py, pz; int il, bc;
  OGJK_DI void take(const T2Hit<T>& h, int c, bool valid) {
    if (valid && (h.sc > bl || (h.sc == bl && h.o < il))) { bl = h.sc; il = h.o; bc = c; px = h.x; py = h.y; pz = h.z; }
  }
};

template <typename T>
OGJK_DI void t2_fetch(const T2Side<T>& s, int c, int lane, T dx, T dy, T dz, T2Hit<T>& h) {
  const T* r = s.record(c) + lane;
  h.x = r[0]; h.y = r[kAccelCluster]; h.z = r[2 * kAccelCluster];
  h.o = reinterpret_cast<const int*>(r - lane + 3 * kAccelCluster)[lane];
  h.sc = dot3(h.x, h.y, h.z, dx, dy, dz);
}

// Both support calls of one iteration (openGJK.c:992-993) in lockstep.
template <typename T>
OGJK_DI void t2_support_both(T2Side<T> (&S)[2], int lane, const V3<T>& v, V3<T>& sa, int& ia, V3<T>& sb, int& ib) {
  constexpr int K = OGJK_T2_K;
  constexpr unsigned FULL = 0xffffffffu;
  T dx[2] = {-v.x, v.x}, dy[2] = {-v.y, v.y}, dz[2] = {-v.z, v.z};
  T2Best<T> B[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const V3<T>& cur = k == 0 ? sa : sb;
    B[k].bl = dot3(cur.x, cur.y, cur.z, dx[k], dy[k], dz[k]);
    B[k].px = cur.x; B[k].py = cur.y; B[k].pz = cur.z;
    B[k].il = -1; B[k].bc = -1;
#if OGJK_T2_CACHE
    if (S[k].cc >= 0) {
      T2Hit<T> h{dot3(S[k].cx, S[k].cy, S[k].cz, dx[k], dy[k], dz[k]), S[k].cx, S[k].cy, S[k].cz, S[k].co};
      B[k].take(h, S[k].cc, true);
    }
#endif
  }
  const int tiles = ((S[0].nc > S[1].nc ? S[0].nc : S[1].nc) + 31) >> 5;
  for (int t = 0; t < tiles; ++t) {
    T ub[2];
    unsigned live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      T cx, cy, cz;
      if (t == 0) {
        cx = dx[k] >= (T)0 ? S[k].hi[0] : S[k].lo[0];
        cy = dy[k] >= (T)0 ? S[k].hi[1] : S[k].lo[1];
        cz = dz[k] >= (T)0 ? S[k].hi[2] : S[k].lo[2];
      } else {
        const int c0 = (t * 32 < S[k].nc ? t : 0) * 32 + lane;  // a body with fewer tiles re-reads tile 0, masked off
        cx = S[k].acc[(dx[k] >= (T)0 ? 3 : 0) * S[k].ncs + c0];
        cy = S[k].acc[(dy[k] >= (T)0 ? 4 : 1) * S[k].ncs + c0];
        cz = S[k].acc[(dz[k] >= (T)0 ? 5 : 2) * S[k].ncs + c0];
      }
      ub[k] = dot3(cx, cy, cz, dx[k], dy[k], dz[k]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool has = t * 32 + lane < S[k].nc;
      const T wb = unkey(group_max_key(FULL, okey(B[k].bl)));
      live[k] = __ballot_sync(FULL, has && !(ub[k] < wb));
#if OGJK_T2_CACHE
      if ((S[k].cc >> 5) == t && S[k].cc >= 0) live[k] &= ~(1u << (S[k].cc & 31));  // already evaluated
#endif
    }
    bool first = true;
    while (live[0] | live[1]) {
      int cl[2][K];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        int kmax = K;
        if (first) {
          // highest bound first: it most likely holds the winner and tightens the prune
          const auto key = ((live[k] >> lane) & 1u) ? okey(ub[k]) : decltype(okey(ub[k]))(0);
          const auto m = group_max_key(FULL, key);
          const unsigned at = __ballot_sync(FULL, ((live[k] >> lane) & 1u) && key == m);
          const int top = live[k] ? __ffs(at) - 1 : -1;
          cl[k][0] = top < 0 ? -1 : t * 32 + top;
          if (top >= 0) live[k] &= ~(1u << top);
#if OGJK_T2_CACHE
          if (S[k].cc < 0) kmax = 1;  // nothing is known yet about this body: do not speculate
#else
          kmax = 1;
#endif
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          if (first && j == 0) continue;
          if (j < kmax && live[k]) {
            cl[k][j] = t * 32 + __ffs(live[k]) - 1;
            live[k] &= live[k] - 1;
          } else {
            cl[k][j] = -1;
          }
        }
      }
      T2Hit<T> h[2][K];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < K; ++j) t2_fetch(S[k], cl[k][j] < 0 ? 0 : cl[k][j], lane, dx[k], dy[k], dz[k], h[k][j]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < K; ++j) B[k].take(h[k][j], cl[k][j], cl[k][j] >= 0);
      }
      if (live[0] | live[1]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const T wb = unkey(group_max_key(FULL, okey(B[k].bl)));
          live[k] &= __ballot_sync(FULL, !(ub[k] < wb));
        }
      }
      first = false;
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const auto key = okey(B[k].bl);
    const bool cand = key == group_max_key(FULL, key);
    const unsigned fi = __reduce_min_sync(FULL, cand ? (unsigned)B[k].il : 0xffffffffu);
    if (fi == 0xffffffffu) continue;  // nothing beats the carried point
    const int src = __ffs(__ballot_sync(FULL, cand && (unsigned)B[k].il == fi)) - 1;
    const V3<T> w{__shfl_sync(FULL, B[k].px, src), __shfl_sync(FULL, B[k].py, src), __shfl_sync(FULL, B[k].pz, src)};
    if (k == 0) { sa = w; ia = (int)fi; } else { sb = w; ib = (int)fi; }
#if OGJK_T2_CACHE
    const int wc = __shfl_sync(FULL, B[k].bc, src);
    if (wc != S[k].cc) {  // issued now, first used at the top of the next iteration
      const T* r = S[k].record(wc) + lane;
      S[k].cx = r[0]; S[k].cy = r[kAccelCluster]; S[k].cz = r[2 * kAccelCluster];
      S[k].co = reinterpret_cast<const int*>(r - lane + 3 * kAccelCluster)[lane];
      S[k].cc = wc;
    }
#endif
  }
}

// The simplex (warp-uniform: 20 coordinates/indices + count + running max) is dead weight
// during the support scans, which need their registers for the loads in flight; lane 0
// parks it in the warp's shared-memory slot and every lane reads it back (broadcast).
template <typename T> struct T2Park { T f[13]; int i[9]; };
template <typename T> OGJK_DI void t2_park(T2Park<T>* slot, const GjkState<T>& st, int lane) {
  if (lane == 0) {
    const Splx<T>& s = st.s;
    const T f[13] = {s.q0.x, s.q0.y, s.q0.z, s.q1.x, s.q1.y, s.q1.z, s.q2.x, s.q2.y, s.q2.z, s.q3.x, s.q3.y, s.q3.z, st.wmax2};
    const int i[9] = {s.n, s.q0.ia, s.q0.ib, s.q1.ia, s.q1.ib, s.q2.ia, s.q2.ib, s.q3.ia, s.q3.ib};
#pragma unroll
    for (int j = 0; j < 13; ++j) slot->f[j] = f[j];
#pragma unroll
    for (int j = 0; j < 9; ++j) slot->i[j] = i[j];
  }
  __syncwarp();
}
template <typename T> OGJK_DI void t2_unpark(const T2Park<T>* slot, GjkState<T>& st) {
  const T* f = slot->f; const int* i = slot->i;
  st.s.n = i[0];
  st.s.q0 = MV<T>{f[0], f[1], f[2], i[1], i[2]};
  st.s.q1 = MV<T>{f[3], f[4], f[5], i[3], i[4]};
  st.s.q2 = MV<T>{f[6], f[7], f[8], i[5], i[6]};
  st.s.q3 = MV<T>{f[9], f[10], f[11], i[7], i[8]};
  st.wmax2 = f[12];
  __syncwarp();
}

// Setup of one pair, produced by the lane-parallel pipeline below.
struct T2Desc {
  long long slot;  // position in the segment (>= end: no more work)
  int p;           // pair id
  int n1, n2;
  long long boff1, boff2, aoff1, aoff2;
};

// The next pair's setup chain, one dependent load per step():
//   0: queue grab   1: pair id   2: polytope index (even lanes body 1, odd lanes body 2)
//   3: lane 0/1 vertex count, lane 2/3 block offset, lane 4/5 table offset    4: done
template <typename T> struct T2Pipe {
  int stage;
  unsigned long long got;  // lane 0
  long long slot;
  int p;
  long long h;
  long long val;
  OGJK_DI void start() { stage = 0; }
  OGJK_DI void step(const StoreArgs<T>& a, const int* __restrict__ list, long long begin, long long end,
                    unsigned long long* __restrict__ head, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    if (stage == 0) {
      got = 0;
      if (lane == 0) got = atomicAdd(head, 1ull);
      stage = 1;
    } else if (stage == 1) {
      slot = begin + (long long)__shfl_sync(FULL, got, 0);
      p = 0;
      if (slot < end) { p = list ? list[slot] : (int)slot; stage = 2; } else { stage = 4; }
    } else if (stage == 2) {
      const int* idx = (lane & 1) ? a.idx2 : a.idx1;
      h = idx ? (long long)idx[p] : (long long)p;
      stage = 3;
    } else if (stage == 3) {
      const StoreView<T>& s = (lane & 1) ? a.s2 : a.s1;
      val = 0;
      if (lane < 2) val = s.cnt[h];
      else if (lane < 4) val = s.boff[h];
      else if (lane < 6) val = s.aoff[h];
      stage = 4;
    }
  }
  OGJK_DI T2Desc finish(const StoreArgs<T>& a, const int* __restrict__ list, long long begin, long long end,
                        unsigned long long* __restrict__ head, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    while (stage < 4) step(a, list, begin, end, head, lane);
    T2Desc d;
    d.slot = slot; d.p = p;
    d.n1 = d.n2 = 0; d.boff1 = d.boff2 = d.aoff1 = d.aoff2 = 0;
    if (slot < end) {
      d.n1 = (int)__shfl_sync(FULL, val, 0); d.n2 = (int)__shfl_sync(FULL, val, 1);
      d.boff1 = __shfl_sync(FULL, val, 2); d.boff2 = __shfl_sync(FULL, val, 3);
      d.aoff1 = __shfl_sync(FULL, val, 4); d.aoff2 = __shfl_sync(FULL, val, 5);
    }
    return d;
  }
};

template <typename T>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_TABLE2_F64 : OGJK_LB_TABLE2_F32))
gjk_table2_kernel(StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                  const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
#if OGJK_T2_PARK
  __shared__ T2Park<T> park[4];
  T2Park<T>* slot = &park[threadIdx.x >> 5];
#endif
  T2Pipe<T> pipe;
  pipe.start();
  T2Desc d = pipe.finish(a, list, begin, end, head, lane);
  while (d.slot < end) {
    pipe.start();
#if OGJK_T2_PREFETCH
    pipe.step(a, list, begin, end, head, lane);  // queue grab for the next pair
#endif
    const GroupBody<T, 32> b1{a.s1.data + d.boff1, d.n1, pad4(d.n1), lane, FULL};
    const GroupBody<T, 32> b2{a.s2.data + d.boff2, d.n2, pad4(d.n2), lane, FULL};
    T2Side<T> S[2];
    S[0].open(a.s1.accel + d.aoff1, d.n1, lane);
    S[1].open(a.s2.accel + d.aoff2, d.n2, lane);
    GjkState<T> st;
    gjk_init(b1, b2, st);
    bool done;
    do {
      st.k++;
#if OGJK_T2_PARK
      t2_park(slot, st, lane);
      t2_support_both(S, lane, st.v, st.sa, st.ia, st.sb, st.ib);
      t2_unpark(slot, st);
#else
      t2_support_both(S, lane, st.v, st.sa, st.ia, st.sb, st.ib);
#endif
      done = gjk_step_tail(st);
#if OGJK_T2_PREFETCH
      pipe.step(a, list, begin, end, head, lane);
#endif
    } while (!done);
    Splx<T> s = st.s;
    V3<T> w1, w2;
    gjk_witnesses(b1, b2, s, w1, w2);
    if (lane == 0) {
      a.distances[d.p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
      if (a.simplices) store_simplex(a.simplices + d.p, s, w1, w2);
    }
    d = pipe.finish(a, list, begin, end, head, lane);
  }
}

}  // namespace ogjk
