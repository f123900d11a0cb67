// regroup_kernel.cuh -- small-hull group kernel with in-CTA regrouping.
//
// The group kernels (gjk_kernels.cuh) give every pair G lanes and a warp 32/G pairs; a warp
// runs until its slowest pair is done, and GJK iteration counts spread from 2 to ~12 around a
// mean of 5 (config 2: 63 % of the lane-iterations of a 16-pair warp are live).
// Here a CTA of 256 threads takes 256/G pairs, runs them iteration-synchronously in phases
// (up to r1 iterations, up to r2, to the end), and between phases packs the pairs that are
// still iterating into the lowest groups through shared memory (22 values + 12 indices per
// pair), so that whole warps fall idle instead of lanes.  A regroup is skipped when it would
// not free a warp.  Every pair executes exactly the operation sequence of the group kernel
// (same gjk_step, same lanes-per-pair scan), so the results are the same bytes; the pair's
// polytopes stay hot in the SM's L1 because the pair never leaves the CTA (the multi-pass
// scheduler of round 1 lost exactly there: the parked pairs were cold in the next pass).
#pragma once
#include "gjk_kernels.cuh"

namespace ogjk {

#ifndef OGJK_LB_REGROUP_F32
#define OGJK_LB_REGROUP_F32 2
#endif
#ifndef OGJK_LB_REGROUP_F64
#define OGJK_LB_REGROUP_F64 2
#endif

#ifndef OGJK_REGROUP_THREADS
#define OGJK_REGROUP_THREADS 256
#endif
constexpr int kRegroupThreads = OGJK_REGROUP_THREADS;
constexpr int kRegroupF = 22;  // v, sa, sb, q0..q3 xyz, wmax2
constexpr int kRegroupI = 12;  // ia, ib, n, k, q0..q3 (ia, ib)

template <typename T> struct RegroupRow {  // what a pair needs besides its iteration state
  long long p;
  const T* blk1;
  const T* blk2;
  int n1, n2;
};

template <typename T, int G>
__global__ void __launch_bounds__(kRegroupThreads, (sizeof(T) == 8 ? OGJK_LB_REGROUP_F64 : OGJK_LB_REGROUP_F32))
gjk_regroup_kernel(const __grid_constant__ StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
                   const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head, int r1, int r2) {
  constexpr int NP = kRegroupThreads / G;  // pairs per CTA batch
  constexpr int NW = kRegroupThreads / 32;
  constexpr int PER_WARP = 32 / G;
  // value e of pair j at sf[e * NP + j]: consecutive groups hit consecutive words
  __shared__ T sf[kRegroupF * NP];
  __shared__ int si[kRegroupI * NP];
  __shared__ RegroupRow<T> sx[NP];
  __shared__ int warp_live[NW];
  __shared__ long long batch_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane % G;
  const int gi = tid / G;  // this group's slot in the CTA
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - g));
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;
  for (;;) {
    __syncthreads();  // everybody is done with the previous batch's shared rows
    if (tid == 0) batch_base = (long long)atomicAdd(head, (unsigned long long)NP);
    __syncthreads();
    const long long base = begin + batch_base;
    if (base >= end) break;
    const long long slot = base + gi;
    bool live = slot < end;
    long long p = 0;
    int n1 = 1, n2 = 1;
    const T *blk1 = a.s1.data, *blk2 = a.s2.data;
    if (live) {
      p = list ? (long long)list[slot] : slot;
      const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
      const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
      n1 = a.s1.cnt[h1]; n2 = a.s2.cnt[h2];
      blk1 = a.s1.data + a.s1.boff[h1];
      blk2 = a.s2.data + a.s2.boff[h2];
    }
    GjkState<T> st;
    {
      const GroupBody<T, G> b1{blk1, n1, pad4(n1), g, mask}, b2{blk2, n2, pad4(n2), g, mask};
      gjk_init(b1, b2, st);
    }
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
      const int limit = phase == 0 ? r1 : phase == 1 ? r2 : kMaxIter + 1;
      if (live) {
        const GroupBody<T, G> b1{blk1, n1, pad4(n1), g, mask}, b2{blk2, n2, pad4(n2), g, mask};
        for (;;) {
          if (gjk_step(b1, b2, st)) {
            Splx<T> s = st.s;
            V3<T> w1, w2;
            gjk_witnesses(b1, b2, s, w1, w2);
            if (g == 0) {
              a.distances[p] = sqrt(nrm2(st.v.x, st.v.y, st.v.z));
              if (a.simplices) store_simplex(a.simplices + p, s, w1, w2);
            }
            live = false;
            break;
          }
          if (st.k >= limit) break;
        }
      }
      if (phase == 2) break;
      // ---- regroup: pack the pairs that are still iterating into groups 0 .. total-1 ----
      const unsigned leaders = __ballot_sync(0xffffffffu, live && g == 0);
      if (lane == 0) warp_live[wid] = __popc(leaders);
      __syncthreads();
      int before = 0, total = 0, busy = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int c = warp_live[w];
        before += w < wid ? c : 0;
        total += c;
        busy += c > 0;
      }
      if (total == 0) break;                                 // uniform: the batch is finished
      if ((total + PER_WARP - 1) / PER_WARP >= busy) {       // uniform: packing would not free a warp
        __syncthreads();                                     // warp_live is rewritten at the next regroup
        continue;
      }
      if (live && g == 0) {
        const int row = before + __popc(leaders & ((1u << lane) - 1u));
        const MV<T>&q0 = st.s.q0, &q1 = st.s.q1, &q2 = st.s.q2, &q3 = st.s.q3;
        const T vals[kRegroupF] = {st.v.x, st.v.y, st.v.z, st.sa.x, st.sa.y, st.sa.z, st.sb.x, st.sb.y, st.sb.z, q0.x, q0.y,
                                   q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, st.wmax2};
        const int ints[kRegroupI] = {st.ia, st.ib, st.s.n, st.k, q0.ia, q0.ib, q1.ia, q1.ib, q2.ia, q2.ib, q3.ia, q3.ib};
#pragma unroll
        for (int e = 0; e < kRegroupF; ++e) sf[e * NP + row] = vals[e];
#pragma unroll
        for (int e = 0; e < kRegroupI; ++e) si[e * NP + row] = ints[e];
        sx[row] = RegroupRow<T>{p, blk1, blk2, n1, n2};
      }
      __syncthreads();
      live = gi < total;
      if (live) {
        T vals[kRegroupF];
        int ints[kRegroupI];
#pragma unroll
        for (int e = 0; e < kRegroupF; ++e) vals[e] = sf[e * NP + gi];
#pragma unroll
        for (int e = 0; e < kRegroupI; ++e) ints[e] = si[e * NP + gi];
        const RegroupRow<T> x = sx[gi];
        p = x.p; blk1 = x.blk1; blk2 = x.blk2; n1 = x.n1; n2 = x.n2;
        st.v = V3<T>{vals[0], vals[1], vals[2]};
        st.sa = V3<T>{vals[3], vals[4], vals[5]};
        st.sb = V3<T>{vals[6], vals[7], vals[8]};
        st.ia = ints[0]; st.ib = ints[1]; st.s.n = ints[2]; st.k = ints[3];
        st.s.q0 = MV<T>{vals[9], vals[10], vals[11], ints[4], ints[5]};
        st.s.q1 = MV<T>{vals[12], vals[13], vals[14], ints[6], ints[7]};
        st.s.q2 = MV<T>{vals[15], vals[16], vals[17], ints[8], ints[9]};
        st.s.q3 = MV<T>{vals[18], vals[19], vals[20], ints[10], ints[11]};
        st.wmax2 = vals[21];
      }
      __syncthreads();  // rows are free again before anybody parks into them
    }
  }
}

}  // namespace ogjk
