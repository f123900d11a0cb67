// store.cuh -- the packed polytope store and the batch scheduler's sort.
//
// Polytope store (north_star item 1): every polytope becomes one contiguous,
// 16-byte aligned block
//
//     [ x[0..npad) | y[0..npad) | z[0..npad) ]        npad = round_up(n, 4)
//
// at element offset boff[h] (a multiple of 4) of one device array.  Padding
// slots repeat vertex 0, which can tie with but never beat vertex 0 under the
// reference's "strictly greater, lowest index first" support rule
// (scalar/openGJK.c:607-630), so they never change a result.  The layout lets a
// lane fetch four vertices with three 128-bit loads (fp32; six for fp64), keeps
// cooperating lanes coalesced, and makes a whole polytope one bulk-copyable range.
// It replaces the reference's per-polytope cudaMemcpy of xyz-interleaved blocks
// (gpu/opengjk_gpu.cu:3105-3166).
//
// Scheduler (north_star item 5): pairs are counting-sorted by their padded
// vertex total so that lanes of a warp scan hulls of equal length, and so that
// each size class can be routed to the kernel flavour that suits it.
#pragma once
#include <cuda_runtime.h>

namespace ogjk {

__host__ __device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }

template <typename T> struct StoreView {
  const T* data;           // packed blocks
  const long long* boff;   // element offset of each block (multiple of 4)
  const int* cnt;          // true vertex counts
  // support-scan accelerator (persistent stores only; null when absent), see "cluster tables" below
  const T* accel = nullptr;
  const long long* aoff = nullptr;  // element offset of each polytope's accelerator block
  int accel_min = 0x7fffffff;       // polytopes with accel_min <= n <= kAccelMaxVerts have one
  long long npoly = 0;              // polytopes in the store (only needed with tables)
};

// ---- cluster tables: exact pruning of the support scan for large polytopes --------
//   A polytope with many vertices gets, next to its block, a
// copy of its vertices grouped into spatially compact clusters of 32 (left-balanced k-d
// splits along the widest axis), the original index of every copied vertex, and each cluster's bounding box.
// For a search direction d the box corner picked by the signs of d gives
//   ub = (cx*dx + cy*dy) + cz*dz  >=  (x*dx + y*dy) + z*dz   for every vertex of the cluster,
// and the inequality holds for the ROUNDED values too (same operation order, IEEE
// rounding is monotone), so a cluster with ub < best cannot hold a vertex that beats or
// ties `best` and skipping it leaves the scan result bit-identical.
// Block layout in T elements, nc = ceil(n/32), ncs = nc rounded up to 32:
//   lox[ncs] loy[ncs] loz[ncs] hix[ncs] hiy[ncs] hiz[ncs]       cluster boxes
//   6 x kAccelMaxTiles tile boxes (same order)                   box of each run of 32 clusters
//   ox oy oz slack | nx[ncs] ny[ncs] nz[ncs] ct[ncs] A[ncs] B[ncs] R[ncs]   cluster cones (below; OGJK_T3_CONE builds only)
//   nc cluster records = x[32] y[32] z[32] | int orig[32]        (kAccelRecord<T> elements each)
// (slots past n in the last cluster repeat that cluster's first vertex; table entries
// past nc are zero and never used).  The k-d order makes 32 consecutive clusters a compact
// region, so a tile box prunes 1024 vertices with one dot product (bodies above 1024
// vertices; same monotone-rounding argument as for the cluster boxes).
//
// Cluster cones (experimental, -DOGJK_T3_CONE=1; off by default: on B200 the extra per-iteration
// arithmetic cost more than the saved fetches, profiles/r2g_table_cone_rank_ab.jsonl).
// A box bounds a patch of a curved hull loosely (on 1000-vertex sphere hulls 5-6
// of 32 boxes survive a search direction where the exact per-cluster support would keep 1-2),
// so every cluster also gets a cone seen from a point o inside the body (the centre of its
// bounding box): all its vertices lie within distance R of o and within angle theta of the unit
// axis n.  Then, with phi the angle between n and d,
//     p.d  <=  o.d + R |d| max(0, cos(max(0, phi - theta)))
// and the kernel takes min(box bound, cone bound).  The cone bound is evaluated in floating point
// with explicit margins instead of a rounding argument: ct = cos(theta) is lowered by 1e-4 at
// build time (absorbs the build's own rounding and |n| != 1), sin(theta) and R are rounded up,
// and `slack` * |d| = 2^-15 (|o|_1 + max R) |d| is added -- three orders of magnitude above the
// worst-case rounding error of both the bound and the vertices' dot products, five below the
// spread of the bounds.  Degenerate clusters (non-finite data, all vertices at o) get ct = -2,
// i.e. the ball bound o.d + R|d|; the kernel falls back to the box bound whenever |d|^2 is
// outside [1e-30, 1e30] or the cone bound is not finite.
// A = R * ct and B = R * sin(theta) are stored premultiplied: outside the cone the bound is
// o.d + max(0, A (n.d) + B sqrt(|d|^2 - (n.d)^2)) + slack |d|.
#ifndef OGJK_T3_CONE
#define OGJK_T3_CONE 0  // cluster cones in the tables (measured on B200: exact, fewer clusters fetched, slower: DESIGN.md section 4)
#endif
constexpr int kAccelCluster = 32;
constexpr int kAccelMaxVerts = 16384;
constexpr int kAccelMaxClusters = kAccelMaxVerts / kAccelCluster;
constexpr int kAccelMaxTiles = kAccelMaxClusters / 32;  // 16
template <typename T> constexpr int kAccelRecord = 3 * kAccelCluster + kAccelCluster * (int)sizeof(int) / (int)sizeof(T);

__host__ __device__ __forceinline__ int accel_table_stride(int nc) { return (nc + 31) & ~31; }
// Element offsets inside a polytope's table: cone header (ox oy oz slack) + cone arrays (only in
// -DOGJK_T3_CONE=1 builds), cluster records.
__host__ __device__ __forceinline__ int accel_cone_at(int ncs) { return 6 * ncs + 6 * (16384 / 32 / 32); }
#ifndef OGJK_TABLE_PAD
#define OGJK_TABLE_PAD 0  // tuning: extra elements between the box arrays and the records
#endif
__host__ __device__ __forceinline__ int accel_records_at(int ncs) {
  return (OGJK_T3_CONE ? 13 * ncs + 6 * (16384 / 32 / 32) + 4 : 6 * ncs + 6 * (16384 / 32 / 32)) + OGJK_TABLE_PAD;
}
template <typename T> __host__ __device__ __forceinline__ long long accel_elems(int n) {
  const int nc = (n + kAccelCluster - 1) / kAccelCluster;
  return (long long)accel_records_at(accel_table_stride(nc)) + (long long)nc * kAccelRecord<T>;
}

// Per-polytope element counts the three scan kernels below accumulate.
struct BlockElems {
  __device__ __forceinline__ long long operator()(int n) const { return 3ll * pad4(n); }
};
template <typename T> struct AccelElems {
  int amin;
  __device__ __forceinline__ long long operator()(int n) const {
    return (n >= amin && n <= kAccelMaxVerts) ? accel_elems<T>(n) : 0;
  }
};

// ---- exclusive scan of F(cnt[h]) over polytopes (three small kernels) ----
constexpr int kScanBlock = 1024;

template <typename F>
__global__ void __launch_bounds__(kScanBlock)
scan_block_sums_kernel(const int* __restrict__ cnt, long long n, long long* __restrict__ block_sums, F elems) {
  __shared__ long long warp_sums[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = i < n ? elems(cnt[i]) : 0;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long s = warp_sums[threadIdx.x];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
  }
}

// Single block: in-place exclusive scan of the block sums; total goes to *total.
__global__ void __launch_bounds__(1024)
scan_spine_kernel(long long* __restrict__ block_sums, long long nblocks, long long* __restrict__ total) {
  __shared__ long long carry;
  __shared__ long long warp_sums[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < nblocks; base += 1024) {
    const long long i = base + threadIdx.x;
    const long long x = i < nblocks ? block_sums[i] : 0;
    long long v = x;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, v, m);
      if ((threadIdx.x & 31) >= m) v += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long s = warp_sums[threadIdx.x];
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, s, m);
        if (threadIdx.x >= m) s += t;
      }
      warp_sums[threadIdx.x] = s;
    }
    __syncthreads();
    const long long wprefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const long long incl = v + wprefix + carry;
    if (i < nblocks) block_sums[i] = incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

template <typename F>
__global__ void __launch_bounds__(kScanBlock)
scan_finish_kernel(const int* __restrict__ cnt, long long n, const long long* __restrict__ block_sums,
                   long long* __restrict__ boff, F elems) {
  __shared__ long long warp_sums[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  const long long x = i < n ? elems(cnt[i]) : 0;
  long long v = x;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) >= m) v += t;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long s = warp_sums[threadIdx.x];
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, s, m);
      if (threadIdx.x >= m) s += t;
    }
    warp_sums[threadIdx.x] = s;
  }
  __syncthreads();
  const long long wprefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
  if (i < n) boff[i] = block_sums[blockIdx.x] + wprefix + v - x;
}

// Layout check for the chunked host pipeline: polytopes must be stored in pair order
// without overlap (off[p] >= off[p-1] + cnt[p-1]).  Sets *flag when violated.
__global__ void __launch_bounds__(256)
check_monotone_kernel(const long long* __restrict__ off, const int* __restrict__ cnt, long long n, int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; p < n; p += stride)
    if (off[p] < off[p - 1] + cnt[p - 1]) *flag = 1;
}

// ---- pool validation (raw pools come from the caller; nothing below may index out of bounds) ----
// The store keeps its OWN sanitised copy of the vertex counts: a count outside [1, nverts]
// becomes 1 and raises *flag (the sticky pool-error flag of the context).
__global__ void __launch_bounds__(256)
sanitize_counts_kernel(const int* __restrict__ cnt_in, long long npoly, long long nverts, int* __restrict__ cnt_out,
                       int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x; h < npoly; h += stride) {
    int n = cnt_in[h];
    if (n < 1 || (long long)n > nverts) { n = 1; *flag = 1; }
    cnt_out[h] = n;
  }
}
// Asynchronous builds size the block array from an upper bound (nverts + 3 per polytope), which
// holds only when the polytopes' vertex ranges do not add up to more than nverts (no sharing).
// When the scanned total exceeds the allocation every polytope is collapsed onto one dummy
// 4-vertex block at offset 0 (defined results, no out-of-bounds access) and *flag is raised.
__global__ void __launch_bounds__(256)
overflow_guard_kernel(const long long* __restrict__ total, long long capacity, long long npoly, int* __restrict__ cnt,
                      long long* __restrict__ boff, int* __restrict__ flag) {
  if (*total <= capacity) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x; h < npoly; h += stride) {
    cnt[h] = 1;
    boff[h] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1;
}

// ---- repack: xyz-interleaved blocks -> padded SoA blocks ------------------------
// Eight lanes per polytope, four polytopes per warp, warp-strided over polytopes.
// `cnt` is the store's sanitised copy.  A polytope whose source range leaves [0, nverts)
// is packed as zeros and raises *flag.
template <typename T>
__global__ void __launch_bounds__(256)
pack_kernel(const T* __restrict__ coords, const long long* __restrict__ off, const int* __restrict__ cnt,
            long long npoly, const long long* __restrict__ boff, T* __restrict__ out, long long nverts,
            int* __restrict__ flag) {
  const int sub = threadIdx.x & 7;
  const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long h = first; h < npoly; h += stride) {
    const int n = cnt[h];
    const int np = pad4(n);
    const long long o = off[h];
    const bool ok = o >= 0 && o + n <= nverts;
    if (!ok && sub == 0) *flag = 1;
    const T* src = coords + 3 * (ok ? o : 0);
    T* dst = out + boff[h];
    for (int i = sub; i < np; i += 8) {
      const int j = i < n ? i : 0;  // padding repeats vertex 0
      dst[i] = ok ? src[3 * j] : (T)0;
      dst[np + i] = ok ? src[3 * j + 1] : (T)0;
      dst[2 * np + i] = ok ? src[3 * j + 2] : (T)0;
    }
  }
}

// ---- cluster-table build (one CTA per large polytope, off the query path) ----------
__device__ __forceinline__ unsigned accel_ord(float f) {  // order-preserving float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float accel_unord(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// Segment of clusters [c0, c1) that position `i` belongs to after `level` splits of the
// cluster range [0, nc).  A segment of m clusters splits into the largest power of two below
// m and the rest (a left-balanced k-d tree), so every segment starts at a multiple of its
// size rounded up to a power of two: after `level` splits no segment has more than
// 2^(K - level) clusters (2^K >= nc) or crosses a boundary of that size, and the level's sort
// can run inside aligned blocks instead of over the whole polytope.
__device__ __forceinline__ void accel_segment(int i, int level, int nc, int& c0, int& c1) {
  c0 = 0; c1 = nc;
  const int c = i / kAccelCluster;
  for (int l = 0; l < level && c1 - c0 > 1; ++l) {
    const int mid = c0 + (1 << (31 - __clz(c1 - c0 - 1)));
    if (c < mid) c1 = mid; else c0 = mid;
  }
}

// Bitonic sort of every aligned block of `sortw` keys (a power of two >= 64) of keys[0, P) in
// shared memory, ascending.  The network's steps are grouped so that the keys make few round
// trips through shared memory (the sort is bound by its bandwidth): the steps with a partner
// distance below 32 run on one key per lane with shuffles, the others up to three at a time on
// 8 keys per thread.  Every lane of the CTA calls this (CTA widths are multiples of 32).
__device__ __forceinline__ unsigned long long accel_cx_lane(unsigned long long a, int j, bool asc) {
  const unsigned long long b = __shfl_xor_sync(0xffffffffu, a, j);
  const bool keep_min = ((threadIdx.x & j) == 0) == asc;
  return keep_min ? (a < b ? a : b) : (a > b ? a : b);
}
template <int R>
__device__ __forceinline__ void accel_sort_group(unsigned long long* keys, int P, int sortw, int k, int j, int tid, int nth) {
  constexpr int E = 1 << R;
  const int q = j >> (R - 1);  // smallest partner distance of the group
  const int qs = 31 - __clz(q);
  for (int t = tid; t < (P >> R); t += nth) {
    const int base = ((t >> qs) << (qs + R)) | (t & (q - 1));
    const bool asc = (base & k) == 0 || k == sortw;
    unsigned long long x[E];
#pragma unroll
    for (int m = 0; m < E; ++m) x[m] = keys[base + m * q];
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
#pragma unroll
      for (int m = 0; m < E; ++m) {
        if (m & (1 << r)) continue;
        const unsigned long long a = x[m], b = x[m | (1 << r)];
        if ((a > b) == asc) { x[m] = b; x[m | (1 << r)] = a; }
      }
    }
#pragma unroll
    for (int m = 0; m < E; ++m) keys[base + m * q] = x[m];
  }
}
__device__ __forceinline__ void accel_sort_blocks(unsigned long long* keys, int P, int sortw, int tid, int nth) {
  // k = 2 .. 32: whole sub-networks inside 32 consecutive keys
  for (int i = tid; i < P; i += nth) {
    unsigned long long a = keys[i];
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
      const bool asc = (i & k) == 0;  // sortw >= 64: never the final merge
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) a = accel_cx_lane(a, j, asc);
    }
    keys[i] = a;
  }
  __syncthreads();
  for (int k = 64; k <= sortw; k <<= 1) {
    int j = k >> 1;
    while (j >= 32) {
      const int left = 31 - __clz(j) - 4;  // steps with a partner distance >= 32 still to do
      if (left >= 3) { accel_sort_group<3>(keys, P, sortw, k, j, tid, nth); j >>= 3; }
      else if (left == 2) { accel_sort_group<2>(keys, P, sortw, k, j, tid, nth); j >>= 2; }
      else { accel_sort_group<1>(keys, P, sortw, k, j, tid, nth); j >>= 1; }
      __syncthreads();
    }
    for (int i = tid; i < P; i += nth) {
      unsigned long long a = keys[i];
      const bool asc = (i & k) == 0 || k == sortw;
#pragma unroll
      for (int jj = 16; jj > 0; jj >>= 1) a = accel_cx_lane(a, jj, asc);
      keys[i] = a;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
max_count_kernel(const int* __restrict__ cnt, long long n, int amin, int* __restrict__ out) {
  int m = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = cnt[i];
    if (c >= amin && c <= kAccelMaxVerts && c > m) m = c;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, k));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// Sort keys live in dynamic shared memory: (segment's first cluster) << 46 | ordered
// fp32 coordinate along the segment's widest axis << 14 | vertex index; one bitonic sort
// per k-d level, inside the aligned blocks that hold the level's segments.  The partition is a
// heuristic (any partition gives exact queries), so the split coordinate is compared in fp32
// even for fp64 stores.  One launch handles the polytopes with nlo < n <= pcap vertices
// (the host launches one per power-of-two size range, each with its own CTA shape).
template <typename T>
__global__ void __launch_bounds__(1024)
accel_build_kernel(StoreView<T> s, long long npoly, T* __restrict__ accel, int nlo, int pcap,
                   unsigned long long* __restrict__ next) {
  extern __shared__ unsigned long long accel_keys[];
  __shared__ unsigned seg_lo[3][kAccelMaxClusters], seg_hi[3][kAccelMaxClusters];
  __shared__ unsigned char seg_axis[kAccelMaxClusters];
#if OGJK_T3_CONE
  __shared__ T cone_o[3];
  __shared__ unsigned cone_rmax;  // ordered fp32 key of the largest cluster radius
#endif
  __shared__ long long next_h;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (;;) {
    // polytopes are handed out one at a time (build times differ by an order of magnitude
    // inside a launch's size range and most polytopes of a mixed pool are out of it)
    __syncthreads();
    if (tid == 0) next_h = (long long)atomicAdd(next, 1ull);
    __syncthreads();
    const long long h = next_h;
    if (h >= npoly) break;
    const int n = s.cnt[h];
    if (n < s.accel_min || n > kAccelMaxVerts || n <= nlo || n > pcap) continue;
    const int np = pad4(n);
    const T* blk = s.data + s.boff[h];
    const int nc = (n + kAccelCluster - 1) / kAccelCluster;
    int P = 32;
    while (P < n) P <<= 1;
    for (int i = tid; i < P; i += nth) accel_keys[i] = i < n ? (unsigned long long)i : ~0ull;
    int levels = 0;
    while ((1 << levels) < nc) ++levels;
    __syncthreads();
    for (int level = 0; level < levels; ++level) {
      for (int c = tid; c < nc; c += nth) {
        seg_lo[0][c] = seg_lo[1][c] = seg_lo[2][c] = 0xffffffffu;
        seg_hi[0][c] = seg_hi[1][c] = seg_hi[2][c] = 0u;
      }
      __syncthreads();
      // a warp covers one cluster (CTA widths are multiples of 32), i.e. one segment: its extent
      // is reduced on the redux unit and one lane per warp touches the segment's shared counters
      for (int i = tid; i < nc * kAccelCluster; i += nth) {
        int c0, c1;
        accel_segment(i, level, nc, c0, c1);
        if (c1 - c0 > 1) {  // warp-uniform
          const int v = (int)(accel_keys[i < n ? i : 0] & 0x3fffu);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const unsigned f = accel_ord((float)blk[a * np + v]);
            const unsigned lo = __reduce_min_sync(0xffffffffu, i < n ? f : 0xffffffffu);
            const unsigned hi = __reduce_max_sync(0xffffffffu, i < n ? f : 0u);
            if ((tid & 31) == 0) {
              atomicMin(&seg_lo[a][c0], lo);
              atomicMax(&seg_hi[a][c0], hi);
            }
          }
        }
      }
      __syncthreads();
      for (int c = tid; c < nc; c += nth) {
        int best = 0;
        float ext = -1.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float e = accel_unord(seg_hi[a][c]) - accel_unord(seg_lo[a][c]);
          if (e > ext) { ext = e; best = a; }  // NaN/empty extents leave axis 0
        }
        seg_axis[c] = (unsigned char)best;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nth) {
        int c0, c1;
        accel_segment(i, level, nc, c0, c1);
        const unsigned long long v = accel_keys[i] & 0x3fffu;
        const unsigned f = (c1 - c0 > 1) ? accel_ord((float)blk[seg_axis[c0] * np + (int)v]) : 0u;
        accel_keys[i] = ((unsigned long long)c0 << 46) | ((unsigned long long)f << 14) | v;
      }
      __syncthreads();
      const int sortw = max(P >> level, 2 * kAccelCluster);  // aligned block that holds any segment of this level
      accel_sort_blocks(accel_keys, P, sortw, tid, nth);
    }
    // cluster records + original indices; a warp per cluster reduces its bounding box
    const int ncs = accel_table_stride(nc);
    T* out = accel + s.aoff[h];
    T* rec = out + accel_records_at(ncs);
    for (int c = tid >> 5; c < ncs; c += nth >> 5) {
      const int lane = tid & 31;
      if (c >= nc) {  // table padding, never used by queries
        if (lane < 6) out[lane * ncs + c] = (T)0;
        continue;
      }
      const int i = c * kAccelCluster + lane;
      const int v = (int)(accel_keys[i < n ? i : c * kAccelCluster] & 0x3fffu);
      const T x = blk[v], y = blk[np + v], z = blk[2 * np + v];
      T* r = rec + (long long)c * kAccelRecord<T>;
      r[lane] = x; r[kAccelCluster + lane] = y; r[2 * kAccelCluster + lane] = z;
      reinterpret_cast<int*>(r + 3 * kAccelCluster)[lane] = v;
      T lx = x, ly = y, lz = z, hx = x, hy = y, hz = z;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        lx = fmin(lx, __shfl_xor_sync(0xffffffffu, lx, m)); hx = fmax(hx, __shfl_xor_sync(0xffffffffu, hx, m));
        ly = fmin(ly, __shfl_xor_sync(0xffffffffu, ly, m)); hy = fmax(hy, __shfl_xor_sync(0xffffffffu, hy, m));
        lz = fmin(lz, __shfl_xor_sync(0xffffffffu, lz, m)); hz = fmax(hz, __shfl_xor_sync(0xffffffffu, hz, m));
      }
      if (lane == 0) {
        out[c] = lx; out[ncs + c] = ly; out[2 * ncs + c] = lz;
        out[3 * ncs + c] = hx; out[4 * ncs + c] = hy; out[5 * ncs + c] = hz;
      }
    }
    __syncthreads();
    // tile boxes: a warp per run of 32 clusters reduces the cluster boxes just written
    T* tb = out + 6 * ncs;
    for (int t = tid >> 5; t < kAccelMaxTiles; t += nth >> 5) {
      const int lane = tid & 31;
      const int c = t * 32 + lane;
      const bool in = c < nc;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T l = in ? out[a * ncs + c] : (T)INFINITY, h = in ? out[(3 + a) * ncs + c] : (T)-INFINITY;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
          l = fmin(l, __shfl_xor_sync(0xffffffffu, l, m));
          h = fmax(h, __shfl_xor_sync(0xffffffffu, h, m));
        }
        if (lane == 0) {
          tb[a * kAccelMaxTiles + t] = t * 32 < nc ? l : (T)0;
          tb[(3 + a) * kAccelMaxTiles + t] = t * 32 < nc ? h : (T)0;
        }
      }
    }
    __syncthreads();
#if OGJK_T3_CONE
    // cluster cones about the centre of the body's box (see the layout comment above)
    if (tid < 32) {
      if (tid == 0) cone_rmax = 0u;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T l = (T)INFINITY, hh = (T)-INFINITY;
        for (int t = tid; t < kAccelMaxTiles && t * 32 < nc; t += 32) {
          l = fmin(l, tb[a * kAccelMaxTiles + t]);
          hh = fmax(hh, tb[(3 + a) * kAccelMaxTiles + t]);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
          l = fmin(l, __shfl_xor_sync(0xffffffffu, l, m));
          hh = fmax(hh, __shfl_xor_sync(0xffffffffu, hh, m));
        }
        if (tid == 0) cone_o[a] = l * (T)0.5 + hh * (T)0.5;
      }
    }
    __syncthreads();
    T* cone = out + accel_cone_at(ncs);
    T* carr = cone + 4;
    const T ox = cone_o[0], oy = cone_o[1], oz = cone_o[2];
    for (int c = tid >> 5; c < ncs; c += nth >> 5) {
      const int lane = tid & 31;
      if (c >= nc) {  // padding: never used
        if (lane < 7) carr[lane * ncs + c] = (T)0;
        continue;
      }
      const T* r = rec + (long long)c * kAccelRecord<T>;
      const T qx = r[lane] - ox, qy = r[kAccelCluster + lane] - oy, qz = r[2 * kAccelCluster + lane] - oz;
      const T rr = sqrt(qx * qx + qy * qy + qz * qz);
      const bool at_o = !(rr > (T)0);  // vertex at o (or NaN): contributes no direction
      T ux = at_o ? (T)0 : qx / rr, uy = at_o ? (T)0 : qy / rr, uz = at_o ? (T)0 : qz / rr;
      T sx = ux, sy = uy, sz = uz, rmax = rr;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, m);
        sy += __shfl_xor_sync(0xffffffffu, sy, m);
        sz += __shfl_xor_sync(0xffffffffu, sz, m);
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, m));
      }
      const T sn = sqrt(sx * sx + sy * sy + sz * sz);
      const bool has_axis = sn > (T)1e-6;
      const T nx = has_axis ? sx / sn : (T)0, ny = has_axis ? sy / sn : (T)0, nz = has_axis ? sz / sn : (T)0;
      T cmin = at_o ? (T)1 : ux * nx + uy * ny + uz * nz;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, m));
      const unsigned finite = __ballot_sync(0xffffffffu, rr == rr && rr < (T)INFINITY);
      T ct = cmin - (T)1e-4, R = rmax * (T)(1.0 + 1e-6);
      if (!has_axis || finite != 0xffffffffu || !(ct == ct)) { ct = (T)-2; R = (finite == 0xffffffffu) ? R : (T)INFINITY; }
      const T ctc = ct < (T)-1 ? (T)-1 : (ct > (T)1 ? (T)1 : ct);
      const T st = sqrt(fmax((T)0, (T)1 - ctc * ctc)) * (T)(1.0 + 1e-5) + (T)1e-7;
      if (lane == 0) {
        carr[0 * ncs + c] = nx; carr[1 * ncs + c] = ny; carr[2 * ncs + c] = nz;
        carr[3 * ncs + c] = ct;
        carr[4 * ncs + c] = R * ctc;
        carr[5 * ncs + c] = R * st;
        carr[6 * ncs + c] = R;
        atomicMax(&cone_rmax, accel_ord((float)fmin(R, (T)3.0e38)));
      }
    }
    __syncthreads();
    if (tid == 0) {
      const T rm = (T)accel_unord(cone_rmax) * (T)(1.0 + 1e-6);
      cone[0] = ox; cone[1] = oy; cone[2] = oz;
      cone[3] = (fabs(ox) + fabs(oy) + fabs(oz) + rm) * (T)(1.0 / 32768.0);
    }
    __syncthreads();
#endif
  }
}

// ---- device arrays of reference-shaped gkPolytope structs ------------------------
// (what the reference's mid-level API hands around: gpu/include/openGJK_GPU.h:71-76,
// built by allocate_and_copy_device_arrays).  Same memory layout for both precisions'
// struct; only numpoints and the device `coord` pointer are read.
template <typename T> struct DevPolytope { int numpoints; T s[3]; int s_idx; const T* coord; };

template <typename T>
__global__ void __launch_bounds__(256)
extract_counts_kernel(const DevPolytope<T>* __restrict__ bd, long long n, int* __restrict__ cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cnt[i] = bd[i].numpoints;
}

// Repack from per-polytope device pointers (eight lanes per polytope, as pack_kernel).
template <typename T>
__global__ void __launch_bounds__(256)
pack_ptr_kernel(const DevPolytope<T>* __restrict__ bd, long long npoly, const long long* __restrict__ boff,
                T* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long h = first; h < npoly; h += stride) {
    const int n = bd[h].numpoints;
    const int np = pad4(n);
    const T* src = bd[h].coord;
    T* dst = out + boff[h];
    for (int i = sub; i < np; i += 8) {
      const int j = i < n ? i : 0;
      dst[i] = src[3 * j];
      dst[np + i] = src[3 * j + 1];
      dst[2 * np + i] = src[3 * j + 2];
    }
  }
}

// ---- counting sort of pairs by padded vertex total ------------------------------
// size key tk = (pad4(n1) + pad4(n2)) / 4, clamped to kMaxKey: what the scheduler's class
// table is expressed in.  The sort itself uses a finer bin: for tk <= kFineKeys the
// bin also encodes pad4(n1)/4, so that the pairs of a warp agree not only on the
// total but on the split -- both support scans of every lane then have the same
// trip count (ncu: the scans ran at 12/32 lanes when only the total was sorted).
// Output order is DESCENDING bin (largest pairs first: better queue tail).
constexpr int kMaxKey = 2048;
constexpr int kSmallKeyMax = 4;
constexpr int kFineKeys = 64;   // totals up to 256 padded vertices get the (total, split) bin
constexpr int kFineSub = 32;    // sub-bins per total
constexpr int kNumBins = kFineKeys * kFineSub + kFineSub + (kMaxKey - kFineKeys);  // 4064
__host__ __device__ __forceinline__ int bin_of(int tk, int np1) {
  if (tk <= kFineKeys) {
    const int sub = (np1 >> 2) < kFineSub ? (np1 >> 2) : kFineSub - 1;
    return tk * kFineSub + sub;
  }
  return kFineKeys * kFineSub + kFineSub - 1 + (tk - kFineKeys);
}
// Largest bin that belongs to size key <= tk (class boundaries).
__host__ __device__ __forceinline__ int bin_upper(int tk) {
  return tk <= kFineKeys ? tk * kFineSub + kFineSub - 1 : kFineKeys * kFineSub + kFineSub - 1 + (tk - kFineKeys);
}
// Pairs whose two bodies both carry a cluster table (persistent stores) are sorted into
// 2 x (kMaxKey + 1) extra bins ABOVE the plain ones, so they come first in the sorted list and
// run as two segments through the table kernels: first the pairs with a body above
// kAccelOneTile vertices (several 32-cluster tiles: table_kernel.cuh), then the pairs whose two
// bodies fit one tile (first-generation table kernel, still the faster one there).  Inside a
// segment the bin is the first body's index (see pair_key).
constexpr int kAccelOneTile = 32 * kAccelCluster;  // 1024 vertices
constexpr int kTableBins = kMaxKey + 1;
constexpr int kNumBinsAll = kNumBins + 2 * kTableBins;
constexpr int kSortBlock = 256;
constexpr int kSortItems = 8;  // pairs per thread

template <typename T>
__device__ __forceinline__ int pair_key(const StoreView<T>& s1, const StoreView<T>& s2, const int* idx1, const int* idx2,
                                        long long p) {
  const long long h1 = idx1 ? (long long)idx1[p] : p;
  const long long h2 = idx2 ? (long long)idx2[p] : p;
  const int n1 = s1.cnt[h1], n2 = s2.cnt[h2];
  const int np1 = pad4(n1), np2 = pad4(n2);
  int k = (np1 + np2) >> 2;
  if (s1.accel && s2.accel && n1 >= s1.accel_min && n1 <= kAccelMaxVerts && n2 >= s2.accel_min && n2 <= kAccelMaxVerts) {
    // The table kernels run one pair per warp, so equal sizes inside a warp do not matter;
    // ordering a segment by the first body's index instead keeps the bodies a pair list
    // reuses (indexed batches) hot in L2.
    const int large = (n1 > kAccelOneTile || n2 > kAccelOneTile) ? kTableBins : 0;
    return kNumBins + large + (int)(h1 * (long long)kTableBins / (s1.npoly > 0 ? s1.npoly : 1));
  }
  // keys 2..4 are reserved for pairs whose two polytopes both fit 8 register slots
  if ((np1 > 8 || np2 > 8) && k < kSmallKeyMax + 1) k = kSmallKeyMax + 1;
  if (k > kMaxKey) k = kMaxKey;
  return bin_of(k, np1);
}

template <typename T>
__global__ void __launch_bounds__(kSortBlock)
sort_hist_kernel(StoreView<T> s1, StoreView<T> s2, const int* __restrict__ idx1, const int* __restrict__ idx2,
                 long long npairs, unsigned* __restrict__ hist, int nbins) {
  __shared__ unsigned sh[kNumBinsAll];  // nbins = kNumBins, or kNumBinsAll when both stores carry tables
  for (int i = threadIdx.x; i < nbins; i += kSortBlock) sh[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortBlock * kSortItems;
#pragma unroll
  for (int t = 0; t < kSortItems; ++t) {
    const long long p = base + (long long)t * kSortBlock + threadIdx.x;
    if (p < npairs) atomicAdd(&sh[pair_key(s1, s2, idx1, idx2, p)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += kSortBlock)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// Single block.  start[b] = number of pairs with bin > b (descending layout).  Also
// resets the scatter cursors.  Warp 0 does a chunked suffix scan over the bins.
__global__ void __launch_bounds__(1024)
sort_scan_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ start, unsigned* __restrict__ cursor, int nbins) {
  __shared__ unsigned sh[kNumBinsAll];
  for (int i = threadIdx.x; i < nbins; i += 1024) sh[i] = hist[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    // lane l owns the contiguous bins [l*per, (l+1)*per) counted from the TOP
    const int per = (nbins + 31) / 32;
    const int lane = threadIdx.x;
    unsigned local = 0;
    for (int j = 0; j < per; ++j) {
      const int b = nbins - 1 - (lane * per + j);
      if (b >= 0) local += sh[b];
    }
    unsigned incl = local;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, m);
      if (lane >= m) incl += t;
    }
    unsigned acc = incl - local;  // pairs in bins above this lane's range
    for (int j = 0; j < per; ++j) {
      const int b = nbins - 1 - (lane * per + j);
      if (b >= 0) {
        const unsigned c = sh[b];
        sh[b] = acc;
        acc += c;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += 1024) {
    start[i] = sh[i];
    cursor[i] = 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSortBlock)
sort_scatter_kernel(StoreView<T> s1, StoreView<T> s2, const int* __restrict__ idx1, const int* __restrict__ idx2,
                    long long npairs, const unsigned* __restrict__ start, unsigned* __restrict__ cursor,
                    int* __restrict__ sorted, int nbins) {
  __shared__ unsigned sh[kNumBinsAll];
  for (int i = threadIdx.x; i < nbins; i += kSortBlock) sh[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortBlock * kSortItems;
  int key[kSortItems];
  unsigned rank[kSortItems];
#pragma unroll
  for (int t = 0; t < kSortItems; ++t) {
    const long long p = base + (long long)t * kSortBlock + threadIdx.x;
    key[t] = -1;
    if (p < npairs) {
      key[t] = pair_key(s1, s2, idx1, idx2, p);
      rank[t] = atomicAdd(&sh[key[t]], 1u);
    }
  }
  __syncthreads();
  // reserve a contiguous range per key for this block
  for (int i = threadIdx.x; i < nbins; i += kSortBlock) {
    const unsigned c = sh[i];
    if (c) sh[i] = start[i] + atomicAdd(&cursor[i], c);
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kSortItems; ++t) {
    const long long p = base + (long long)t * kSortBlock + threadIdx.x;
    if (key[t] >= 0) sorted[sh[key[t]] + rank[t]] = (int)p;
  }
}

}  // namespace ogjk
