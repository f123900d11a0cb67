// gjk_experimental.cuh -- kernel flavours that were measured on B200 and lost to the
// defaults (profiles/r1_lanes_sweep.txt): multi-pass parking,
// shared-memory staged groups, scan-assist, CTA-per-pair, persistent lanes.  All are
// bit-exact; they are compiled only with -DOGJK_EXPERIMENTAL (libopengjk_b200_exp.so,
// `make exp`) so that the product library carries the default path alone.
#pragma once
#include "gjk_kernels.cuh"

namespace ogjk {

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

}  // namespace ogjk
