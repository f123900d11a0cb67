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
};

// ---- exclusive scan of 3*pad4(cnt[h]) over polytopes (three small kernels) ----
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock)
scan_block_sums_kernel(const int* __restrict__ cnt, long long n, long long* __restrict__ block_sums) {
  __shared__ long long warp_sums[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = i < n ? 3ll * pad4(cnt[i]) : 0;
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

__global__ void __launch_bounds__(kScanBlock)
scan_finish_kernel(const int* __restrict__ cnt, long long n, const long long* __restrict__ block_sums,
                   long long* __restrict__ boff) {
  __shared__ long long warp_sums[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  const long long x = i < n ? 3ll * pad4(cnt[i]) : 0;
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

// ---- repack: xyz-interleaved blocks -> padded SoA blocks ------------------------
// Eight lanes per polytope, four polytopes per warp, warp-strided over polytopes.
template <typename T>
__global__ void __launch_bounds__(256)
pack_kernel(const T* __restrict__ coords, const long long* __restrict__ off, const int* __restrict__ cnt,
            long long npoly, const long long* __restrict__ boff, T* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long h = first; h < npoly; h += stride) {
    const int n = cnt[h];
    const int np = pad4(n);
    const T* src = coords + 3 * off[h];
    T* dst = out + boff[h];
    for (int i = sub; i < np; i += 8) {
      const int j = i < n ? i : 0;  // padding repeats vertex 0
      dst[i] = src[3 * j];
      dst[np + i] = src[3 * j + 1];
      dst[2 * np + i] = src[3 * j + 2];
    }
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
constexpr int kSortBlock = 256;
constexpr int kSortItems = 8;  // pairs per thread

template <typename T>
__device__ __forceinline__ int pair_key(const StoreView<T>& s1, const StoreView<T>& s2, const int* idx1, const int* idx2,
                                        long long p) {
  const long long h1 = idx1 ? (long long)idx1[p] : p;
  const long long h2 = idx2 ? (long long)idx2[p] : p;
  const int np1 = pad4(s1.cnt[h1]), np2 = pad4(s2.cnt[h2]);
  int k = (np1 + np2) >> 2;
  // keys 2..4 are reserved for pairs whose two polytopes both fit 8 register slots
  if ((np1 > 8 || np2 > 8) && k < kSmallKeyMax + 1) k = kSmallKeyMax + 1;
  if (k > kMaxKey) k = kMaxKey;
  return bin_of(k, np1);
}

template <typename T>
__global__ void __launch_bounds__(kSortBlock)
sort_hist_kernel(StoreView<T> s1, StoreView<T> s2, const int* __restrict__ idx1, const int* __restrict__ idx2,
                 long long npairs, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[kNumBins];
  for (int i = threadIdx.x; i < kNumBins; i += kSortBlock) sh[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortBlock * kSortItems;
#pragma unroll
  for (int t = 0; t < kSortItems; ++t) {
    const long long p = base + (long long)t * kSortBlock + threadIdx.x;
    if (p < npairs) atomicAdd(&sh[pair_key(s1, s2, idx1, idx2, p)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumBins; i += kSortBlock)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// Single block.  start[b] = number of pairs with bin > b (descending layout).  Also
// resets the scatter cursors.  Warp 0 does a chunked suffix scan over the bins.
__global__ void __launch_bounds__(1024)
sort_scan_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ start, unsigned* __restrict__ cursor) {
  __shared__ unsigned sh[kNumBins];
  for (int i = threadIdx.x; i < kNumBins; i += 1024) sh[i] = hist[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    // lane l owns the contiguous bins [l*per, (l+1)*per) counted from the TOP
    constexpr int per = (kNumBins + 31) / 32;
    const int lane = threadIdx.x;
    unsigned local = 0;
    for (int j = 0; j < per; ++j) {
      const int b = kNumBins - 1 - (lane * per + j);
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
      const int b = kNumBins - 1 - (lane * per + j);
      if (b >= 0) {
        const unsigned c = sh[b];
        sh[b] = acc;
        acc += c;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumBins; i += 1024) {
    start[i] = sh[i];
    cursor[i] = 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSortBlock)
sort_scatter_kernel(StoreView<T> s1, StoreView<T> s2, const int* __restrict__ idx1, const int* __restrict__ idx2,
                    long long npairs, const unsigned* __restrict__ start, unsigned* __restrict__ cursor,
                    int* __restrict__ sorted) {
  __shared__ unsigned sh[kNumBins];
  for (int i = threadIdx.x; i < kNumBins; i += kSortBlock) sh[i] = 0;
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
  for (int i = threadIdx.x; i < kNumBins; i += kSortBlock) {
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
