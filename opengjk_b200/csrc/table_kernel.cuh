// table_kernel.cuh -- warp-per-pair GJK over the store's cluster tables (store.cuh).
//
// History (profiles/r2a_table_gen2_ab.jsonl).  The first table kernel
// (gjk_table_kernel<T,32>, gjk_kernels.cuh) ran 1000-vertex hulls at 0.49-0.51 of the HBM
// roofline: per iteration and body it read the 32 cluster boxes, fetched the best cluster,
// re-pruned and fetched the survivors three per trip -- one cluster = 32 vertices = one
// 4-byte load per lane and array, ~25 warp instructions per cluster, ~7 clusters per body
// and iteration.  Its pruned scan therefore issued as many instructions as the plain scan
// it replaced (which reads 128 vertices per 3 vector loads); the kernel sat at 54 % issue
// utilisation with ~5 dependent round trips per iteration.  A second draft that only
// removed round trips (boxes in registers, last winner's cluster cached in registers,
// next pair's setup prefetched) was SLOWER (94 vs 129 M pairs/s): the added per-iteration
// instructions cost more than the trips saved.  The kernel is instruction-bound first.
//
// This version cuts instructions per vertex to the plain scan's level and keeps the
// pruning:
//   * cluster boxes of the first 32 clusters live in registers for the whole query
//     (lane c = cluster c), tile boxes (one per 32 clusters) likewise for bodies above
//     1024 vertices: bounds cost a dot product, no memory trip;
//   * a fetch moves VPL clusters at once with 16-byte vector loads: 32/VPL lanes per
//     cluster, VPL consecutive vertices per lane (fp32: 4 clusters = 128 vertices per
//     4 loads; fp64: 2 clusters).  The highest bound goes first, the rest in index order;
//   * the exact (value, lowest original index) update runs only when the max of a lane's
//     VPL scores reaches its running best (the plain scan's max-tree fast path);
//   * both bodies advance in lockstep so that their loads overlap.
//
// Exactness: a cluster (tile) is skipped only when its bound is below a value some vertex
// has already reached, so it cannot hold a vertex that beats or ties the final maximum;
// evaluation order does not matter because the running best is a total order on (dot,
// lowest original index) with the carried support point first among equals
// (openGJK.c:607-630).  Results are bit-identical to the plain scans (tests/test_gpu_accel.py).
#pragma once
#include "gjk_kernels.cuh"

namespace ogjk {

#ifndef OGJK_T3_VPL_F32
#define OGJK_T3_VPL_F32 4
#endif
#ifndef OGJK_T3_VPL_F64
#define OGJK_T3_VPL_F64 2
#endif
#ifndef OGJK_T3_PARK
#define OGJK_T3_PARK 1  // park the simplex in shared memory while the support scans run
#endif
#ifndef OGJK_T3_TOPFIRST
#define OGJK_T3_TOPFIRST 1  // crowded tile: the highest bound goes with the first trip
#endif
// OGJK_T3_CONE (store.cuh, default 0): min(box bound, cone bound) per cluster
#ifndef OGJK_LB_TABLE3_F32
#define OGJK_LB_TABLE3_F32 7
#endif
#ifndef OGJK_LB_TABLE3_F64
#define OGJK_LB_TABLE3_F64 4
#endif

template <typename T> struct T3Vpl { static constexpr int value = sizeof(T) == 4 ? OGJK_T3_VPL_F32 : OGJK_T3_VPL_F64; };

// Square root for the cone bound: only an upper estimate within ~1e-6 relative is needed
// (fp32: the 1-instruction approximation, scaled up; fp64: the correctly rounded one).
OGJK_DI float t3_sqrt_up(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 1.000001f;
}
OGJK_DI double t3_sqrt_up(double x) { return sqrt(x) * 1.000000001; }

// Per-body state of one query.
template <typename T> struct T3Side {
  const T* acc;      // table base (cluster boxes | tile boxes | records)
  int nc, ncs, ntiles;
  T lo[3], hi[3];    // lane c: box of cluster c (tile 0)
#if OGJK_T3_CONE
  T ox, oy, oz, slack;  // cone apex (centre of the body's box) and the bound's safety term per unit |d|
#endif
  OGJK_DI const T* record(int c) const { return acc + accel_records_at(ncs) + (long long)c * kAccelRecord<T>; }
  OGJK_DI void open(const T* table, int n, int lane) {
    acc = table;
    nc = (n + kAccelCluster - 1) / kAccelCluster;
    ncs = accel_table_stride(nc);
    ntiles = (nc + 31) >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = acc[a * ncs + lane]; hi[a] = acc[(3 + a) * ncs + lane]; }
#if OGJK_T3_CONE
    const T* ch = acc + accel_cone_at(ncs);
    ox = ch[0]; oy = ch[1]; oz = ch[2]; slack = ch[3];
#endif
  }
#if OGJK_T3_CONE
  // Cone bound of cluster c (this lane's cluster of the current tile) for direction d; od = o.d,
  // dd = |d|^2, dn = |d| (rounded, see store.cuh for the margins).  Returns +inf when unusable.
  OGJK_DI T cone_bound(int c, T dx, T dy, T dz, T od, T dd, T dn) const {
    const T* ca = acc + accel_cone_at(ncs) + 4 + c;
    const T nx = ca[0], ny = ca[ncs], nz = ca[2 * ncs], ct = ca[3 * ncs], A = ca[4 * ncs], B = ca[5 * ncs], R = ca[6 * ncs];
    const T nd = dot3(nx, ny, nz, dx, dy, dz);
    const T t = dd - nd * nd;
    const T st = t3_sqrt_up(t > (T)0 ? t : (T)0);
    const T off = A * nd + B * st;
    const T cone = nd >= ct * dn ? R * dn : (off > (T)0 ? off : (T)0);
    return od + cone + slack * dn;
  }
#endif
};

// Running best of one lane: value, original index (-1 = the carried point), vertex.
template <typename T> struct T3Best { T bl, px, py, pz; int il; };

template <typename T>
OGJK_DI T t3_corner_dot(const T (&lo)[3], const T (&hi)[3], T dx, T dy, T dz) {
  return dot3(dx >= (T)0 ? hi[0] : lo[0], dy >= (T)0 ? hi[1] : lo[1], dz >= (T)0 ? hi[2] : lo[2], dx, dy, dz);
}

// Warp-wide maximum.  fp32: one CREDUX.MAX.F32 (sm_100a has the float reduction; NaN inputs are
// ignored unless every lane holds one); fp64: the order-preserving key pair on the integer unit.
OGJK_DI float t3_wmax(float x) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
  return r;
}
OGJK_DI double t3_wmax(double x) { return unkey(group_max_key(0xffffffffu, okey(x))); }

// Lowest set lane of `bits` among those holding the largest bound (bounds are never NaN).
template <typename T> OGJK_DI int t3_argmax_lane(unsigned bits, int lane, T ub) {
  const bool mine = (bits >> lane) & 1u;
  const T m = t3_wmax(mine ? ub : (T)-INFINITY);
  return __ffs(__ballot_sync(0xffffffffu, mine && ub == m)) - 1;
}

// A NaN bound (of either sign) must never prune: it becomes +inf.
template <typename T> OGJK_DI T t3_bound(T ub) { return ub != ub ? (T)INFINITY : ub; }

// Both support calls of one iteration (openGJK.c:992-993) in lockstep.
//
// Cluster selection is lane-parallel: each trip the 32/VPL-lane group q takes the q-th lowest
// live cluster of the current tile (a few bit operations per lane; no serial "next cluster"
// chain per slot).
template <typename T>
OGJK_DI void t3_support_both(const T3Side<T> (&S)[2], int lane, const V3<T>& v, V3<T>& sa, int& ia, V3<T>& sb, int& ib) {
  constexpr int VPL = T3Vpl<T>::value;
  constexpr int LPC = 32 / VPL;  // lanes per cluster
  constexpr int NG = VPL;        // clusters per fetch
  constexpr unsigned FULL = 0xffffffffu;
  const int q = lane / LPC, sub = lane % LPC;
  const T dx[2] = {-v.x, v.x}, dy[2] = {-v.y, v.y}, dz[2] = {-v.z, v.z};
  T3Best<T> B[2];
#if OGJK_T3_CONE
  const T dd = v.x * v.x + v.y * v.y + v.z * v.z;
  const T dn = t3_sqrt_up(dd);
  const bool dd_ok = dd > (T)1e-30 && dd < (T)1e30;  // outside: under/overflow territory, box bounds only
  const T od[2] = {dot3(S[0].ox, S[0].oy, S[0].oz, dx[0], dy[0], dz[0]), dot3(S[1].ox, S[1].oy, S[1].oz, dx[1], dy[1], dz[1])};
  const bool cone_ok[2] = {dd_ok && S[0].slack < (T)INFINITY, dd_ok && S[1].slack < (T)INFINITY};
#endif
  T tub[2];        // this lane's tile bound
  unsigned tl[2];                // tiles still to visit
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const V3<T>& cur = k == 0 ? sa : sb;
    B[k].bl = dot3(cur.x, cur.y, cur.z, dx[k], dy[k], dz[k]);  // same value in every lane
    B[k].px = cur.x; B[k].py = cur.y; B[k].pz = cur.z;
    B[k].il = -1;
    tub[k] = (T)0;
    tl[k] = 1u;
    if (S[k].ntiles > 1) {
      // tile boxes (one per 32 clusters) are read per iteration: bodies above 1024 vertices only
      const T* tb = S[k].acc + 6 * S[k].ncs + (lane & (kAccelMaxTiles - 1));
      tub[k] = t3_bound(dot3(tb[(dx[k] >= (T)0 ? 3 : 0) * kAccelMaxTiles], tb[(dy[k] >= (T)0 ? 4 : 1) * kAccelMaxTiles],
                             tb[(dz[k] >= (T)0 ? 5 : 2) * kAccelMaxTiles], dx[k], dy[k], dz[k]));
      tl[k] = __ballot_sync(FULL, lane < S[k].ntiles && !(tub[k] < B[k].bl));
    }
  }
  while (tl[0] | tl[1]) {
    int t[2];
    T ukey[2];  // this lane's cluster bound (NaN-free)
    unsigned live[2];
    T thr[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      t[k] = -1;
      if (tl[k]) {
        t[k] = S[k].ntiles > 1 ? t3_argmax_lane(tl[k], lane, tub[k]) : 0;  // most promising tile first
        tl[k] &= ~(1u << t[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      T ub;
      if (t[k] <= 0) {
        ub = t3_corner_dot(S[k].lo, S[k].hi, dx[k], dy[k], dz[k]);
      } else {
        const T* bx = S[k].acc + t[k] * 32 + lane;
        ub = dot3(bx[(dx[k] >= (T)0 ? 3 : 0) * S[k].ncs], bx[(dy[k] >= (T)0 ? 4 : 1) * S[k].ncs],
                  bx[(dz[k] >= (T)0 ? 5 : 2) * S[k].ncs], dx[k], dy[k], dz[k]);
      }
#if OGJK_T3_CONE
      if (t[k] >= 0 && cone_ok[k]) {
        const T ubc = S[k].cone_bound(t[k] * 32 + lane, dx[k], dy[k], dz[k], od[k], dd, dn);
        ub = ubc < ub ? ubc : ub;  // a NaN cone bound leaves the box bound
      }
#endif
      ukey[k] = t3_bound(ub);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool has = t[k] >= 0 && t[k] * 32 + lane < S[k].nc;
      thr[k] = t3_wmax(B[k].bl);
      live[k] = __ballot_sync(FULL, has && !(ukey[k] < thr[k]));
    }
    bool first = true;
    while (live[0] | live[1]) {
      T x[2][VPL], y[2][VPL], z[2][VPL];
      int o[2][VPL];
      bool valid[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        // group q takes the q-th lowest live cluster (rank, not residue class: neighbouring clusters
        // of the k-d order differ by powers of two, so residue classes would pile up on one group)
        unsigned pool = live[k];
        int forced = -1, rank = q, take = NG;
#if OGJK_T3_TOPFIRST
        if (first && __popc(pool) > NG) {
          // more live clusters than one trip moves (typically the first iteration, when the carried
          // point bounds nothing): the highest bound most likely holds the winner and prunes the rest
          forced = t3_argmax_lane(pool, lane, ukey[k]);
          pool &= ~(1u << forced);
          rank = q - 1;
          take = NG - 1;
        }
#endif
        unsigned rest = pool;
#pragma unroll
        for (int g = 1; g < NG; ++g) rest = rank >= g ? rest & (rest - 1) : rest;  // drop the `rank` lowest bits
        const bool top = forced >= 0 && q == 0;
        valid[k] = top || rest != 0;
        unsigned left = pool;
#pragma unroll
        for (int g = 0; g < NG; ++g) left = g < take ? left & (left - 1) : left;   // what remains after this trip
        // idle groups re-read the lowest cluster taken this trip: its lines are in flight anyway
        const int bit = top ? forced : __ffs(rest ? rest : (live[k] ? live[k] : 1u)) - 1;
        live[k] = left;
        const T* r = S[k].record((t[k] < 0 ? 0 : t[k]) * 32 + bit) + sub * VPL;
        load_vec<VPL>(r, x[k]);
        load_vec<VPL>(r + kAccelCluster, y[k]);
        load_vec<VPL>(r + 2 * kAccelCluster, z[k]);
        load_vec<VPL>(reinterpret_cast<const int*>(r - sub * VPL + 3 * kAccelCluster) + sub * VPL, o[k]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        T sc[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) sc[j] = dot3(x[k][j], y[k][j], z[k][j], dx[k], dy[k], dz[k]);
        T m = sc[0];
#pragma unroll
        for (int j = 1; j < VPL; ++j) m = fmax(m, sc[j]);
        // only a vertex that reaches the best value seen by ANY lane so far can still win
        if (valid[k] && m >= thr[k]) {
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (sc[j] > B[k].bl || (sc[j] == B[k].bl && o[k][j] < B[k].il)) {
              B[k].bl = sc[j]; B[k].il = o[k][j]; B[k].px = x[k][j]; B[k].py = y[k][j]; B[k].pz = z[k][j];
            }
        }
      }
      if (live[0] | live[1]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          thr[k] = t3_wmax(B[k].bl);
          live[k] &= __ballot_sync(FULL, !(ukey[k] < thr[k]));
        }
      }
      first = false;
    }
    if (tl[0] | tl[1]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const T wb = t3_wmax(B[k].bl);
        tl[k] &= __ballot_sync(FULL, !(tub[k] < wb));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool cand = B[k].bl == t3_wmax(B[k].bl);  // +0 == -0; all-NaN leaves no candidate and the carried point stays
    const unsigned fi = __reduce_min_sync(FULL, cand ? (unsigned)B[k].il : 0xffffffffu);
    if (fi == 0xffffffffu) continue;  // nothing beats the carried point
    const int src = __ffs(__ballot_sync(FULL, cand && (unsigned)B[k].il == fi)) - 1;
    const V3<T> w{__shfl_sync(FULL, B[k].px, src), __shfl_sync(FULL, B[k].py, src), __shfl_sync(FULL, B[k].pz, src)};
    if (k == 0) { sa = w; ia = (int)fi; } else { sb = w; ib = (int)fi; }
  }
}

// The simplex (warp-uniform: 12 coordinates, 8 indices, count, running max) is dead weight
// during the support scans, which want their registers for loads in flight; lane 0 parks it
// in the warp's shared-memory slot and every lane reads it back (broadcast).
template <typename T> struct T3Park { T f[13]; int i[9]; };
template <typename T> OGJK_DI void t3_park(T3Park<T>* slot, const GjkState<T>& st, int lane) {
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
template <typename T> OGJK_DI void t3_unpark(const T3Park<T>* slot, GjkState<T>& st) {
  const T* f = slot->f; const int* i = slot->i;
  st.s.n = i[0];
  st.s.q0 = MV<T>{f[0], f[1], f[2], i[1], i[2]};
  st.s.q1 = MV<T>{f[3], f[4], f[5], i[3], i[4]};
  st.s.q2 = MV<T>{f[6], f[7], f[8], i[5], i[6]};
  st.s.q3 = MV<T>{f[9], f[10], f[11], i[7], i[8]};
  st.wmax2 = f[12];
  __syncwarp();
}

// Safety net: a pair routed here without two tables takes the plain warp-wide scans.
template <typename T>
__device__ __noinline__ void t3_plain_pair(const StoreArgs<T>& a, long long p, long long h1, long long h2, int n1, int n2,
                                           int lane) {
  const GroupBody<T, 32> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), lane, 0xffffffffu};
  const GroupBody<T, 32> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), lane, 0xffffffffu};
  Splx<T> s;
  V3<T> w1, w2;
  const T d = gjk_query<T>(b1, b2, s, w1, w2);
  if (lane == 0) {
    a.distances[p] = d;
    if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
  }
}

template <typename T>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_TABLE3_F64 : OGJK_LB_TABLE3_F32))
gjk_table3_kernel(const __grid_constant__ StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                  const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
#if OGJK_T3_PARK
  __shared__ T3Park<T> park[4];
  T3Park<T>* pslot = &park[threadIdx.x >> 5];
#endif
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, 1ull);
    got = __shfl_sync(FULL, got, 0);
    const long long slot = begin + (long long)got;
    if (slot >= end) break;
    const long long p = list ? (long long)list[slot] : slot;
    const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
    const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
    const int n1 = a.s1.cnt[h1], n2 = a.s2.cnt[h2];
    const bool t1 = a.s1.accel && n1 >= a.s1.accel_min && n1 <= kAccelMaxVerts;
    const bool t2 = a.s2.accel && n2 >= a.s2.accel_min && n2 <= kAccelMaxVerts;
    if (!(t1 && t2)) { t3_plain_pair(a, p, h1, h2, n1, n2, lane); continue; }
    const GroupBody<T, 32> b1{a.s1.data + a.s1.boff[h1], n1, pad4(n1), lane, FULL};
    const GroupBody<T, 32> b2{a.s2.data + a.s2.boff[h2], n2, pad4(n2), lane, FULL};
    T3Side<T> S[2];
    S[0].open(a.s1.accel + a.s1.aoff[h1], n1, lane);
    S[1].open(a.s2.accel + a.s2.aoff[h2], n2, lane);
    GjkState<T> st;
    gjk_init(b1, b2, st);
    bool done;
    do {
      st.k++;
#if OGJK_T3_PARK
      t3_park(pslot, st, lane);
      t3_support_both(S, lane, st.v, st.sa, st.ia, st.sb, st.ib);
      t3_unpark(pslot, st);
#else
      t3_support_both(S, lane, st.v, st.sa, st.ia, st.sb, st.ib);
#endif
      done = gjk_step_tail(st);
    } while (!done);
    Splx<T> s = st.s;
    V3<T> w1, w2;
    gjk_witnesses(b1, b2, s, w1, w2);
    if (lane == 0) {
      a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
      if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
    }
  }
}

// =============================================================================
// Pair-batched table kernel: 32 pairs per warp, supports one pair at a time
// =============================================================================
// ncu on both warp-per-pair table kernels (profiles/r2h_*): ~40 % of the ~4.2 k warp
// instructions per pair are the sub-solver, the termination tests and the witnesses --
// scalar work that all 32 lanes of the warp repeat for the ONE pair they share.  Here a
// warp owns 32 pairs: every lane keeps the GJK state of its own pair, so that part runs
// thread-per-pair (one pass serves 32 pairs), while the support scans -- the part that
// needs the lanes -- still run with the whole warp, one pair after the other: the pair's
// owner lane broadcasts its direction, carried support points and table pointers, all 32
// lanes run the pruned scan of the first-generation kernel (support_both<T,32>), and the
// owner keeps the result.  The pair setup (queue, list, indices, counts, offsets) becomes
// coalesced per-lane loads, amortised over 32 pairs.  Lanes whose pair has terminated cost
// nothing in the support phase; the warp leaves when its slowest pair is done.
// Every pair executes exactly the operation sequence of the warp-per-pair kernels, so the
// results are the same bytes.
#ifndef OGJK_LB_TABLEB_F32
#define OGJK_LB_TABLEB_F32 8
#endif
#ifndef OGJK_LB_TABLEB_F64
#define OGJK_LB_TABLEB_F64 4
#endif
#ifndef OGJK_TB_K
#define OGJK_TB_K 2  // clusters per body per trip of the scan (the warp-per-pair kernel takes 3 with its 128 registers)
#endif

template <typename T> OGJK_DI const T* t3_bcast_ptr(const T* p, int src) {
  return reinterpret_cast<const T*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(p), src));
}

template <typename T>
__global__ void __launch_bounds__(128, (sizeof(T) == 8 ? OGJK_LB_TABLEB_F64 : OGJK_LB_TABLEB_F32))
gjk_tablebatch_kernel(const __grid_constant__ StoreArgs<T> a, const int* __restrict__ list,
                      const unsigned* __restrict__ seg_begin, const unsigned* __restrict__ seg_end,
                      unsigned long long* __restrict__ head) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  // Guided batch sizes: a warp takes 32 pairs while the queue is long and fewer as it drains
  // (remaining / (2 x warps in the grid), down to single pairs), so the last wave of warps ends
  // together and a short queue still spreads over every warp.
  const long long nwarps2 = 2ll * gridDim.x * (blockDim.x >> 5);
  for (;;) {
    unsigned long long got = 0;
    int take = 32;
    if (lane == 0) {
      const long long left = end - begin - (long long)*(volatile unsigned long long*)head;
      take = (int)max(1ll, min(32ll, left / nwarps2));
      got = atomicAdd(head, (unsigned long long)take);
    }
    got = __shfl_sync(FULL, got, 0);
    take = __shfl_sync(FULL, take, 0);
    const long long base = begin + (long long)got;
    if (base >= end) break;
    const long long slot = base + lane;
    bool active = lane < take && slot < end;
    // this lane's pair
    long long p = 0;
    int n1 = 1, n2 = 1;
    const T *blk1 = a.s1.data, *blk2 = a.s2.data, *acc1 = nullptr, *acc2 = nullptr;
    if (active) {
      p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      n1 = a.s1.cnt[h1]; n2 = a.s2.cnt[h2];
      blk1 = a.s1.data + a.s1.boff[h1];
      blk2 = a.s2.data + a.s2.boff[h2];
      const bool t1 = a.s1.accel && n1 >= a.s1.accel_min && n1 <= kAccelMaxVerts;
      const bool t2 = a.s2.accel && n2 >= a.s2.accel_min && n2 <= kAccelMaxVerts;
      acc1 = t1 ? a.s1.accel + a.s1.aoff[h1] : nullptr;  // without two tables support_both takes the plain scans
      acc2 = t2 ? a.s2.accel + a.s2.aoff[h2] : nullptr;
    }
    const GroupBody<T, 1> own1{blk1, n1, pad4(n1), 0, 0u}, own2{blk2, n2, pad4(n2), 0, 0u};
    GjkState<T> st;
    gjk_init(own1, own2, st);  // idle lanes initialise from block 0 (valid memory) and never run
    for (;;) {
      const unsigned live = __ballot_sync(FULL, active);
      if (!live) break;
#pragma unroll 1
      for (unsigned m = live; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        // the owner lane hands its pair to the warp
        const int m1 = __shfl_sync(FULL, n1, j), m2 = __shfl_sync(FULL, n2, j);
        const AccelBody<T, 32> b1{{t3_bcast_ptr(blk1, j), m1, pad4(m1), lane, FULL}, t3_bcast_ptr(acc1, j)};
        const AccelBody<T, 32> b2{{t3_bcast_ptr(blk2, j), m2, pad4(m2), lane, FULL}, t3_bcast_ptr(acc2, j)};
        const V3<T> v{__shfl_sync(FULL, st.v.x, j), __shfl_sync(FULL, st.v.y, j), __shfl_sync(FULL, st.v.z, j)};
        V3<T> sa{__shfl_sync(FULL, st.sa.x, j), __shfl_sync(FULL, st.sa.y, j), __shfl_sync(FULL, st.sa.z, j)};
        V3<T> sb{__shfl_sync(FULL, st.sb.x, j), __shfl_sync(FULL, st.sb.y, j), __shfl_sync(FULL, st.sb.z, j)};
        int ia = __shfl_sync(FULL, st.ia, j), ib = __shfl_sync(FULL, st.ib, j);
        support_both<OGJK_TB_K>(b1, b2, v, sa, ia, sb, ib);
        if (lane == j) { st.sa = sa; st.ia = ia; st.sb = sb; st.ib = ib; }
      }
      if (active) {
        st.k++;
        if (gjk_step_tail(st)) {
          Splx<T> s = st.s;
          V3<T> w1, w2;
          gjk_witnesses(own1, own2, s, w1, w2);
          a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
          if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
          active = false;
        }
      }
    }
  }
}

}  // namespace ogjk
