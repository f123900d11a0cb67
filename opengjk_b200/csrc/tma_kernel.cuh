// tma_kernel.cuh -- group kernel with bulk-copy (TMA) staging, double buffered per warp.
//
// north_star item 1 asks for the polytope blocks to be "staged into shared memory with
// TMA/cp.async".  The staged flavour of round 1 (gjk_experimental.cuh) copied a pair's two blocks
// and waited: the copy latency was fully exposed.  Here every warp owns two shared-memory
// buffers and one mbarrier per buffer.  While the warp's 32/G pairs iterate out of buffer b,
// the NEXT batch's blocks are already on their way into buffer b^1: a store block is one
// contiguous 16-byte aligned range (store.cuh), so each is ONE cp.async.bulk
// (global -> shared, completion counted in bytes on the mbarrier), issued by the group's
// leader lane; blocks are packed back to back (exact sizes, prefix sum over the batch), and a
// pair that does not fit the buffer any more is read from global memory as in the plain group
// kernel.  The arithmetic is gjk_query over GroupBody<T,G>, i.e. byte-identical results.
// Experimental build only: measured against the plain group kernel in DESIGN.md section 4.
#pragma once
#include "gjk_kernels.cuh"

namespace ogjk {

constexpr int kTmaWarps = 4;  // warps per CTA

OGJK_DI void tma_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
OGJK_DI void tma_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
OGJK_DI bool tma_mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
OGJK_DI void tma_mbar_wait(unsigned bar, unsigned parity) {
  unsigned spins = 0;
  while (!tma_mbar_try(bar, parity))
    if (++spins > (1u << 28)) __trap();  // a lost copy must not hang the device
}
OGJK_DI void tma_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <typename T> struct TmaPair {  // one pair of a batch as its group sees it
  long long p;       // pair index (-1: no pair in this slot)
  const T* blk1;     // generic pointers: into the warp's buffer when staged, into the store otherwise
  const T* blk2;
  int n1, n2;
};

template <typename T, int G>
__global__ void __launch_bounds__(32 * kTmaWarps, (sizeof(T) == 8 ? 2 : 4))
gjk_tma_kernel(const __grid_constant__ StoreArgs<T> a, const int* __restrict__ list, const unsigned* __restrict__ seg_begin,
               const unsigned* __restrict__ seg_end, unsigned long long* __restrict__ head, int slot_bytes) {
  extern __shared__ uint4 ogjk_tma_stage[];
  __shared__ unsigned long long bars[kTmaWarps][2];
  constexpr int PER_WARP = 32 / G;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane % G;
  const unsigned mask = (G == 32) ? FULL : (((1u << G) - 1u) << (lane - g));
  char* const stage = reinterpret_cast<char*>(ogjk_tma_stage) + (size_t)wid * 2 * slot_bytes;
  const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bars[wid][0]);
  if (lane == 0) {
    tma_mbar_init(bar_s, 1);
    tma_mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long begin = seg_begin ? (long long)*seg_begin : 0;
  const long long end = seg_end ? (long long)*seg_end : a.npairs;

  // Takes the next batch from the queue and starts its copies into buffer `buf`.  Returns the
  // group's pair (p = -1: none) and, through `armed`, whether the buffer's barrier has to be waited for.
  auto issue = [&](int buf, bool& more, bool& armed) -> TmaPair<T> {
    TmaPair<T> r{-1, a.s1.data, a.s2.data, 1, 1};
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(head, (unsigned long long)PER_WARP);
    got = __shfl_sync(FULL, got, 0);
    const long long base = begin + (long long)got;
    more = base < end;
    armed = false;
    if (!more) return r;
    const long long idx = base + lane / G;
    unsigned bytes1 = 0, bytes2 = 0;
    if (idx < end) {
      r.p = list ? (long long)list[idx] : idx;
      const long long h1 = a.idx1 ? (long long)a.idx1[r.p] : r.p;
      const long long h2 = a.idx2 ? (long long)a.idx2[r.p] : r.p;
      r.n1 = a.s1.cnt[h1]; r.n2 = a.s2.cnt[h2];
      r.blk1 = a.s1.data + a.s1.boff[h1];
      r.blk2 = a.s2.data + a.s2.boff[h2];
      bytes1 = 3u * (unsigned)pad4(r.n1) * (unsigned)sizeof(T);  // multiples of 16 (np is a multiple of 4)
      bytes2 = 3u * (unsigned)pad4(r.n2) * (unsigned)sizeof(T);
    }
    // exclusive prefix of the pairs' byte counts over the group leaders, in slot order
    const unsigned mine = g == 0 ? bytes1 + bytes2 : 0u;
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += o;
    }
    const unsigned at = __shfl_sync(FULL, incl - mine, lane - g);  // the group's offset, known to all its lanes
    const bool staged = r.p >= 0 && at + bytes1 + bytes2 <= (unsigned)slot_bytes;
    // total bytes in flight on this buffer's barrier
    unsigned tot = (staged && g == 0) ? bytes1 + bytes2 : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(FULL, tot, d);
    armed = tot != 0;
    if (armed) {
      const unsigned bar = bar_s + 8u * (unsigned)buf;
      if (lane == 0) tma_mbar_expect(bar, tot);
      __syncwarp();
      if (staged) {
        const unsigned dst = stage_s + (unsigned)buf * (unsigned)slot_bytes + at;
        if (g == 0) {
          tma_bulk_load(dst, r.blk1, bytes1, bar);
          tma_bulk_load(dst + bytes1, r.blk2, bytes2, bar);
        }
        r.blk1 = reinterpret_cast<const T*>(stage + (size_t)buf * slot_bytes + at);
        r.blk2 = reinterpret_cast<const T*>(stage + (size_t)buf * slot_bytes + at + bytes1);
      }
    }
    return r;
  };

  unsigned parity = 0;  // bit b: phase parity of buffer b's barrier
  int cur = 0;
  bool more, armed;
  TmaPair<T> now = issue(0, more, armed);
  while (more) {
    bool more_next, armed_next;
    // the other buffer was read by the batch before this one, which the whole warp has left
    const TmaPair<T> next = issue(cur ^ 1, more_next, armed_next);
    if (armed) {
      tma_mbar_wait(bar_s + 8u * (unsigned)cur, (parity >> cur) & 1u);
      parity ^= 1u << cur;
    }
    if (now.p >= 0) {
      const GroupBody<T, G> b1{now.blk1, now.n1, pad4(now.n1), g, mask}, b2{now.blk2, now.n2, pad4(now.n2), g, mask};
      Splx<T> s;
      V3<T> w1, w2;
      const T d = gjk_query<T>(b1, b2, s, w1, w2);
      if (g == 0) {
        a.distances[now.p] = d;
        if (a.simplices) store_simplex(a.simplices + now.p, s, w1, w2);
      }
    }
    __syncwarp();  // every lane is done with buffer `cur` before the batch after next is copied into it
    now = next; more = more_next; armed = armed_next;
    cur ^= 1;
  }
}

}  // namespace ogjk
