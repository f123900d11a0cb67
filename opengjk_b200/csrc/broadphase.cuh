// Broad-phase-adjacent pair production on the device: the caller step
// just before the indexed GJK entry (gkCollisionPair lists, gpu/include/openGJK_GPU.h:89-92,
// 331-366).  Three pieces, all order-deterministic so that results can be compared index for
// index with a host restatement:
//   aabb_kernel        bounding box of every polytope of a packed store (one pass over HBM)
//   cull_*_kernel      stable compaction of a candidate pair list to the pairs whose boxes,
//                      grown by `margin`, overlap on all three axes
//   allpairs_*_kernel  every (i < j) of one pool with overlapping boxes, sorted by (i, j)
// A pair is KEPT unless on some axis loA - hiB > margin or loB - hiA > margin (comparisons with
// NaN are false, so a NaN box never culls).  Box separation along an axis is a lower bound of
// the distance, so culling never loses a pair with distance <= margin.
#pragma once
#include "store.cuh"

namespace ogjk {

// ---- per-polytope boxes: a few lanes per polytope, 16-byte chunk loads --------------------
template <typename T> struct BoxChunk;
template <> struct BoxChunk<float> {
  static constexpr int kVerts = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct BoxChunk<double> {
  static constexpr int kVerts = 2;
  static __device__ __forceinline__ void load(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
};

// out[6 h .. 6 h + 5] = {min x, min y, min z, max x, max y, max z}.  Padding slots repeat
// vertex 0, so whole chunks can be reduced without masking.  fmin/fmax skip NaN coordinates;
// a polytope whose every coordinate on an axis is NaN keeps the NaN of vertex 0.
#ifndef OGJK_AABB_LANES
#define OGJK_AABB_LANES 8  // lanes per polytope (tools/aabb_ab.sh: 8 / 4 / 2 lanes = 87 / 73 / 45 % of HBM peak on 1000-vertex bodies, 47-50 % on 8-64-vertex ones)
#endif
// 8 resident blocks per SM for fp32 (32 registers); fp64 needs 12 running extrema of two
// registers each plus the loads in flight and spilled 172 bytes at that cap, so it runs at 4.
template <typename T>
__global__ void __launch_bounds__(256, (sizeof(T) == 8 ? 4 : 8))
aabb_kernel(StoreView<T> s, long long npoly, T* __restrict__ out) {
  constexpr int CV = BoxChunk<T>::kVerts;
  constexpr int L = OGJK_AABB_LANES;
  const int sub = threadIdx.x & (L - 1);
  const unsigned gm = ((1u << L) - 1u) << ((threadIdx.x & 31) - sub);
  const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const long long stride = ((long long)gridDim.x * blockDim.x) / L;
  const long long rounds = (npoly + stride - 1) / stride;  // every lane of a warp loops equally often
  for (long long r = 0; r < rounds; ++r) {
    const long long h = first + r * stride;
    if (h >= npoly) continue;  // uniform per lane group; the shuffles below use the group mask
    const int np = pad4(s.cnt[h]);
    const T* blk = s.data + s.boff[h];
    const T* xs = blk;
    const T* ys = blk + np;
    const T* zs = blk + 2 * np;
    T lx = xs[0], ly = ys[0], lz = zs[0], hx = lx, hy = ly, hz = lz;
    for (int c = sub * CV; c < np; c += L * CV) {
      T vx[CV], vy[CV], vz[CV];
      BoxChunk<T>::load(xs + c, vx);
      BoxChunk<T>::load(ys + c, vy);
      BoxChunk<T>::load(zs + c, vz);
#pragma unroll
      for (int t = 0; t < CV; ++t) {
        lx = fmin(lx, vx[t]); hx = fmax(hx, vx[t]);
        ly = fmin(ly, vy[t]); hy = fmax(hy, vy[t]);
        lz = fmin(lz, vz[t]); hz = fmax(hz, vz[t]);
      }
    }
#pragma unroll
    for (int m = L / 2; m > 0; m >>= 1) {
      lx = fmin(lx, __shfl_xor_sync(gm, lx, m)); hx = fmax(hx, __shfl_xor_sync(gm, hx, m));
      ly = fmin(ly, __shfl_xor_sync(gm, ly, m)); hy = fmax(hy, __shfl_xor_sync(gm, hy, m));
      lz = fmin(lz, __shfl_xor_sync(gm, lz, m)); hz = fmax(hz, __shfl_xor_sync(gm, hz, m));
    }
    if (sub == 0) {
      T* o = out + 6 * h;
      o[0] = lx; o[1] = ly; o[2] = lz; o[3] = hx; o[4] = hy; o[5] = hz;
    }
  }
}

template <typename T>
__device__ __forceinline__ bool boxes_near(const T* __restrict__ a, const T* __restrict__ b, T margin) {
  bool keep = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    keep = keep && !(a[k] - b[3 + k] > margin) && !(b[k] - a[3 + k] > margin);
  }
  return keep;
}

// ---- stable compaction of a candidate list (flags + block counts, spine scan, scatter) ---
constexpr int kCullBlock = 256;
constexpr int kCullItems = 8;  // candidates per thread, consecutive => output keeps input order

template <typename T>
__global__ void __launch_bounds__(kCullBlock)
cull_count_kernel(const T* __restrict__ box1, const T* __restrict__ box2, const int* __restrict__ idx1,
                  const int* __restrict__ idx2, long long n, T margin, long long* __restrict__ block_sums) {
  __shared__ int warp_sums[kCullBlock / 32];
  const long long base = ((long long)blockIdx.x * kCullBlock + threadIdx.x) * kCullItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kCullItems; ++k) {
    const long long p = base + k;
    if (p < n && boxes_near(box1 + 6ll * idx1[p], box2 + 6ll * idx2[p], margin)) ++c;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < kCullBlock / 32; ++w) t += warp_sums[w];
    block_sums[blockIdx.x] = t;
  }
}

// block_sums holds the exclusive block offsets (scan_spine_kernel); *total the kept count.
template <typename T>
__global__ void __launch_bounds__(kCullBlock)
cull_scatter_kernel(const T* __restrict__ box1, const T* __restrict__ box2, const int* __restrict__ idx1,
                    const int* __restrict__ idx2, long long n, T margin, const long long* __restrict__ block_sums,
                    long long capacity, int* __restrict__ out1, int* __restrict__ out2) {
  __shared__ int warp_sums[kCullBlock / 32];
  const long long base = ((long long)blockIdx.x * kCullBlock + threadIdx.x) * kCullItems;
  unsigned keep = 0;
#pragma unroll
  for (int k = 0; k < kCullItems; ++k) {
    const long long p = base + k;
    if (p < n && boxes_near(box1 + 6ll * idx1[p], box2 + 6ll * idx2[p], margin)) keep |= 1u << k;
  }
  const int mine = __popc(keep);
  int incl = mine;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, m);
    if ((threadIdx.x & 31) >= m) incl += t;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  long long pos = block_sums[blockIdx.x] + incl - mine;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += warp_sums[w];
#pragma unroll
  for (int k = 0; k < kCullItems; ++k) {
    if ((keep >> k) & 1u) {
      if (pos < capacity) { out1[pos] = idx1[base + k]; out2[pos] = idx2[base + k]; }
      ++pos;
    }
  }
}

// ---- all (i < j) pairs of one pool whose boxes are near: a warp per row i ------------------
// Pass 1 counts row i; after the scan pass 2 writes row i's partners in ascending j, so the
// list is sorted by (i, j).  O(n^2 / 2) box tests out of L2 (24 or 48 bytes per polytope).
template <typename T, bool WRITE>
__global__ void __launch_bounds__(256)
allpairs_kernel(const T* __restrict__ box, long long n, T margin, long long* __restrict__ row_count,
                const long long* __restrict__ row_start, long long capacity, int* __restrict__ out1,
                int* __restrict__ out2) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  // rows are dealt round-robin from both ends so that every warp gets long and short rows
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const long long i = (r & 1) ? n - 1 - (r >> 1) : (r >> 1);
    T a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = box[6 * i + k];
    long long found = 0;
    long long pos = WRITE ? row_start[i] : 0;
    for (long long j0 = i + 1; j0 < n; j0 += 32) {
      const long long j = j0 + lane;
      const bool near = j < n && boxes_near(a, box + 6 * j, margin);
      const unsigned m = __ballot_sync(0xffffffffu, near);
      if (WRITE) {
        if (near) {
          const long long q = pos + __popc(m & ((1u << lane) - 1u));
          if (q < capacity) { out1[q] = (int)i; out2[q] = (int)j; }
        }
        pos += __popc(m);
      } else {
        found += __popc(m);
      }
    }
    if (!WRITE && lane == 0) row_count[i] = found;
  }
}

// ---- intersect flag ---------------------------------------------------------------------
// The reference has no flag field; its consumers derive one from the distance.  This is the
// EPA gate's rule (gpu/opengjk_gpu.cu:2053: `distance > gkEpsilon` means separated), so
// flag = !(distance > eps): 1 for touching or intersecting pairs (and for NaN distances).
template <typename T>
__global__ void __launch_bounds__(256)
intersect_flags_kernel(const T* __restrict__ distances, long long n, T eps, unsigned char* __restrict__ flags) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    flags[i] = !(distances[i] > eps) ? 1 : 0;
}

}  // namespace ogjk
