// epa_kernels.cuh -- EPA penetration depth / witness points for intersecting pairs.
//
// Replaces the reference's compute_epa_kernel (gpu/opengjk_gpu.cu:2022-2971).
// The reference keeps a full copy of the expanding polytope (128 faces, 132
// vertices, ~20 KB) in the LOCAL memory of every lane and rebuilds the horizon
// on lane 0.  Here one warp owns one pair and the polytope lives ONCE in shared
// memory (6.5 KB fp32 / 10 KB fp64 per warp); every phase is spread over the 32
// lanes in a way that reproduces the sequential order of the algorithm exactly:
//
//   face normals / distances   lane-strided over faces (independent per face)
//   closest face               per-lane ascending scan + (value, lowest index) butterfly
//   support                    chunked scan of the packed store + butterfly (first index wins)
//   duplicate / visibility     lane-strided, ballot
//   horizon                    invalid faces ranked by ballot prefix -> edge slots;
//                              an edge survives iff no other edge joins the same two
//                              vertices (order-free form of the reference's pairwise
//                              cancellation); the k-th surviving edge takes the k-th
//                              free face slot, as the reference's "first free slot" scan does
//
// so the result is bit-identical to the sequential algorithm evaluated with un-fused
// arithmetic (which is how the test suite checks it).
#pragma once
#include <cuda_runtime.h>

#include "gjk_kernels.cuh"

namespace ogjk {

constexpr int kEpaFaces = 128;               // MAX_EPA_FACES, opengjk_gpu.cu:49
constexpr int kEpaVerts = kEpaFaces + 4;     // :50
constexpr int kEpaEdges = kEpaFaces * 3;     // :2765
constexpr int kEpaMaxIter = 64;              // :2528
constexpr int kEpaWarps = 4;                 // warps (pairs in flight) per block

template <typename T> struct EpaArgs {
  StoreView<T> s1, s2;
  const int* idx1; const int* idx2;
  const GkSimplex<T>* simplices;  // post-witness GJK simplices (read only)
  T* distances;                   // in: GJK distance, out: -depth for penetrating pairs
  T* witness1; T* witness2;       // [npairs][3]
  T* normals;                     // [npairs][3], may be null
  long long npairs;
};

template <typename T> struct EpaShared {
  T vx[kEpaVerts], vy[kEpaVerts], vz[kEpaVerts];
  T fnx[kEpaFaces], fny[kEpaFaces], fnz[kEpaFaces], fd[kEpaFaces];
  int via[kEpaVerts], vib[kEpaVerts];
  unsigned char f0[kEpaFaces], f1[kEpaFaces], f2[kEpaFaces], fvalid[kEpaFaces];
  unsigned char e1[kEpaEdges], e2[kEpaEdges], evalid[kEpaEdges];
  unsigned char free_list[kEpaFaces];
};

template <typename T> OGJK_DI T epa_big() { return (T)1e10f; }

// Contact normal from witness1 to witness2 (:2061-2080).
template <typename T> OGJK_DI V3<T> normal_from_witnesses(const V3<T>& w1, const V3<T>& w2) {
  const V3<T> d = sub(w2, w1);
  T norm = 0.0f;
  norm += d.x * d.x;
  norm += d.y * d.y;
  norm += d.z * d.z;
  norm = sqrt(norm);
  if (norm > Lim<T>::eps()) return V3<T>{d.x / norm, d.y / norm, d.z / norm};
  return V3<T>{(T)1, (T)0, (T)0};
}

// Pre-pass over all pairs: fill the defaults (GJK witnesses + their normal, which is
// the complete answer for separated pairs) and append penetrating pairs
// (distance <= eps, the gate at :2053) to the EPA work list.
template <typename T>
__global__ void __launch_bounds__(256)
epa_prepass_kernel(EpaArgs<T> a, int* __restrict__ list, unsigned long long* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounded = (a.npairs + 31) & ~31ll;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < rounded; p += stride) {
    bool hit = false;
    if (p < a.npairs) {
      const GkSimplex<T>* s = a.simplices + p;
      const V3<T> w1{s->witnesses[0][0], s->witnesses[0][1], s->witnesses[0][2]};
      const V3<T> w2{s->witnesses[1][0], s->witnesses[1][1], s->witnesses[1][2]};
      a.witness1[3 * p] = w1.x; a.witness1[3 * p + 1] = w1.y; a.witness1[3 * p + 2] = w1.z;
      a.witness2[3 * p] = w2.x; a.witness2[3 * p + 1] = w2.y; a.witness2[3 * p + 2] = w2.z;
      if (a.normals) {
        const V3<T> n = normal_from_witnesses(w1, w2);
        a.normals[3 * p] = n.x; a.normals[3 * p + 1] = n.y; a.normals[3 * p + 2] = n.z;
      }
      hit = !(a.distances[p] > Lim<T>::eps());
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) list[base + __popc(m & ((1u << lane) - 1u))] = (int)p;
  }
}

// The reference reduces per-lane partial results with a __shfl_down tree (offsets 16..1,
// strict comparisons, gpu/opengjk_gpu.cu:1780-1801 and :2582-2593).  Lanes own ascending
// index chunks, but on an exact tie BETWEEN lanes the survivor is whatever that tree leaves
// in lane 0 -- not always the lowest index -- and overlapping boxes tie all the time.  The
// same partition and the same tree are used here, so ties resolve exactly as in the
// reference (the CPU checker emulates the tree lane for lane).
template <typename T>
OGJK_DI int epa_support_one(const T* blk, int n, int np, int ppt, T dx, T dy, T dz, int lane) {
  T best = -epa_big<T>();
  int bi = -1;
  const int start = lane * ppt;
  for (int i = start; i < n && i < start + ppt; ++i) {  // :1754-1764 (lane l scans [l*ppt, (l+1)*ppt))
    const T sc = dot3(blk[i], blk[np + i], blk[2 * np + i], dx, dy, dz);
    if (sc > best) { best = sc; bi = i; }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const T ob = __shfl_down_sync(0xffffffffu, best, m);
    const int oi = __shfl_down_sync(0xffffffffu, bi, m);
    if (ob > best) { best = ob; bi = oi; }
  }
  bi = __shfl_sync(0xffffffffu, bi, 0);
  return bi < 0 ? 0 : bi;
}

template <typename T> struct EpaBody {
  const T* blk; int n; int np;
  OGJK_DI V3<T> vertex(int i) const { return V3<T>{blk[i], blk[np + i], blk[2 * np + i]}; }
};

// Minkowski-difference support point along dir: body1 along +dir, body2 along -dir.
template <typename T>
OGJK_DI MV<T> epa_support(const EpaBody<T>& b1, const EpaBody<T>& b2, const V3<T>& dir, int lane) {
  const int ppt = ((b1.n > b2.n ? b1.n : b2.n) + 31) / 32;  // :1741-1743
  const int i1 = epa_support_one<T>(b1.blk, b1.n, b1.np, ppt, dir.x, dir.y, dir.z, lane);
  const int i2 = epa_support_one<T>(b2.blk, b2.n, b2.np, ppt, -dir.x, -dir.y, -dir.z, lane);
  const V3<T> p1 = b1.vertex(i1), p2 = b2.vertex(i2);
  return MV<T>{p1.x - p2.x, p1.y - p2.y, p1.z - p2.z, i1, i2};
}

template <typename T> OGJK_DI bool epa_far(const MV<T>& a, T px, T py, T pz) {
  const T dx = a.x - px, dy = a.y - py, dz = a.z - pz;
  const T d2 = dx * dx + dy * dy + dz * dz;
  return !(d2 < Lim<T>::eps() * Lim<T>::eps());
}

// Winding test shared by the initial tetrahedron (:1899-1934) and new faces (:2858-2891).
template <typename T>
OGJK_DI bool epa_faces_centroid(const V3<T>& v0, const V3<T>& v1, const V3<T>& v2, const V3<T>& c) {
  const V3<T> n = cross3(sub(v1, v0), sub(v2, v0));
  return dot3(n, sub(c, v0)) > (T)0;
}

// compute_barycentric_origin :1940-2011
template <typename T>
OGJK_DI void epa_barycentric(const V3<T>& v0, const V3<T>& v1, const V3<T>& v2, T& a0, T& a1, T& a2) {
  const V3<T> e0 = sub(v1, v0), e1 = sub(v2, v0);
  const V3<T> v0n{-v0.x, -v0.y, -v0.z};
  const T d00 = dot3(e0, e0), d01 = dot3(e0, e1), d11 = dot3(e1, e1);
  const T d20 = dot3(v0n, e0), d21 = dot3(v0n, e1);
  const T denom = d00 * d11 - d01 * d01;
  if (fabs(denom) < Lim<T>::eps()) {
    a0 = a1 = a2 = (T)1.0 / (T)3.0;
    return;
  }
  const T inv = (T)1.0 / denom;
  const T u = (d11 * d20 - d01 * d21) * inv;
  const T v = (d00 * d21 - d01 * d20) * inv;
  const T w = (T)1.0 - u - v;
  if (w < 0) {
    const V3<T> e12 = sub(v2, v1);
    const V3<T> v1n{-v1.x, -v1.y, -v1.z};
    T t = dot3(v1n, e12) / dot3(e12, e12);
    t = fmax((T)0.0, fmin((T)1.0, t));
    a0 = 0; a1 = (T)1.0 - t; a2 = t;
  } else if (u < 0) {
    T t = dot3(v0n, e1) / dot3(e1, e1);
    t = fmax((T)0.0, fmin((T)1.0, t));
    a0 = (T)1.0 - t; a1 = 0; a2 = t;
  } else if (v < 0) {
    T t = dot3(v0n, e0) / dot3(e0, e0);
    t = fmax((T)0.0, fmin((T)1.0, t));
    a0 = (T)1.0 - t; a1 = t; a2 = 0;
  } else {
    a0 = w; a1 = u; a2 = v;
  }
}

template <typename T>
OGJK_DI void epa_write(const EpaArgs<T>& a, long long p, T dist, const V3<T>& w1, const V3<T>& w2, const V3<T>& n) {
  a.distances[p] = dist;
  a.witness1[3 * p] = w1.x; a.witness1[3 * p + 1] = w1.y; a.witness1[3 * p + 2] = w1.z;
  a.witness2[3 * p] = w2.x; a.witness2[3 * p + 1] = w2.y; a.witness2[3 * p + 2] = w2.z;
  if (a.normals) { a.normals[3 * p] = n.x; a.normals[3 * p + 1] = n.y; a.normals[3 * p + 2] = n.z; }
}

// "no new support point": depth 0, the support vertices are the witnesses (:2136-2166).
template <typename T>
OGJK_DI void epa_no_progress(const EpaArgs<T>& a, long long p, const EpaBody<T>& b1, const EpaBody<T>& b2,
                             const MV<T>& w, int lane) {
  if (lane == 0) {
    const V3<T> w1 = b1.vertex(w.ia), w2 = b2.vertex(w.ib);
    epa_write(a, p, (T)0, w1, w2, normal_from_witnesses(w1, w2));
  }
}

// Result from face `fi` (:2626-2660).
template <typename T>
OGJK_DI void epa_finish_on_face(const EpaArgs<T>& a, long long p, const EpaShared<T>& S, int fi, T closest_distance,
                                const EpaBody<T>& b1, const EpaBody<T>& b2, int lane) {
  if (lane != 0) return;
  const int i0 = S.f0[fi], i1 = S.f1[fi], i2 = S.f2[fi];
  const V3<T> v0{S.vx[i0], S.vy[i0], S.vz[i0]}, v1{S.vx[i1], S.vy[i1], S.vz[i1]}, v2{S.vx[i2], S.vy[i2], S.vz[i2]};
  T a0, a1, a2;
  epa_barycentric(v0, v1, v2, a0, a1, a2);
  const V3<T> p0 = b1.vertex(S.via[i0]), p1 = b1.vertex(S.via[i1]), p2 = b1.vertex(S.via[i2]);
  const V3<T> q0 = b2.vertex(S.vib[i0]), q1 = b2.vertex(S.vib[i1]), q2 = b2.vertex(S.vib[i2]);
  const V3<T> w1{p0.x * a0 + p1.x * a1 + p2.x * a2, p0.y * a0 + p1.y * a1 + p2.y * a2, p0.z * a0 + p1.z * a1 + p2.z * a2};
  const V3<T> w2{q0.x * a0 + q1.x * a1 + q2.x * a2, q0.y * a0 + q1.y * a1 + q2.y * a2, q0.z * a0 + q1.z * a1 + q2.z * a2};
  epa_write(a, p, -closest_distance, w1, w2, V3<T>{S.fnx[fi], S.fny[fi], S.fnz[fi]});
}

// compute_face_normal_distance (:1659-1717) for every valid face, lane-strided.
template <typename T>
OGJK_DI void epa_face_pass(EpaShared<T>& S, int max_face, const V3<T>& centroid, int lane) {
  for (int f = lane; f < max_face; f += 32) {
    if (!S.fvalid[f]) continue;
    const int i0 = S.f0[f], i1 = S.f1[f], i2 = S.f2[f];
    const V3<T> v0{S.vx[i0], S.vy[i0], S.vz[i0]}, v1{S.vx[i1], S.vy[i1], S.vz[i1]}, v2{S.vx[i2], S.vy[i2], S.vz[i2]};
    V3<T> n = cross3(sub(v1, v0), sub(v2, v0));
    const T nsq = nrm2(n.x, n.y, n.z);
    if (nsq > Lim<T>::eps() * Lim<T>::eps()) {
      const T nrm = sqrt(nsq);
      n.x /= nrm; n.y /= nrm; n.z /= nrm;
      if (dot3(n, sub(centroid, v0)) > (T)0) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
      T d = dot3(n, v0);
      if (d < 0) { n.x = -n.x; n.y = -n.y; n.z = -n.z; d = -d; }
      S.fnx[f] = n.x; S.fny[f] = n.y; S.fnz[f] = n.z; S.fd[f] = d;
    } else {
      S.fnx[f] = n.x; S.fny[f] = n.y; S.fnz[f] = n.z;
      S.fvalid[f] = 0;
      S.fd[f] = epa_big<T>();
    }
  }
  __syncwarp();
}

// Smallest non-negative face distance (:2568-2597): lane l takes faces [l*fpt, (l+1)*fpt),
// first index inside the lane, then the reference's __shfl_down tree (see epa_support_one).
template <typename T>
OGJK_DI int epa_closest_face(const EpaShared<T>& S, int max_face, int lane, T& closest_distance) {
  T bd = epa_big<T>();
  int bi = -1;
  const int fpt = (max_face + 31) / 32;
  const int start = lane * fpt;
  const int end = start + fpt < max_face ? start + fpt : max_face;
  for (int f = start; f < end; ++f) {
    if (!S.fvalid[f]) continue;
    const T d = S.fd[f];
    if (d >= (T)0 && d < bd) { bd = d; bi = f; }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const T od = __shfl_down_sync(0xffffffffu, bd, m);
    const int oi = __shfl_down_sync(0xffffffffu, bi, m);
    if (oi >= 0 && (bi < 0 || od < bd)) { bd = od; bi = oi; }
  }
  closest_distance = __shfl_sync(0xffffffffu, bd, 0);
  return __shfl_sync(0xffffffffu, bi, 0);
}

// The iteration-cap exit (:2919-2970) is a lane-0 block in the reference: plain first index
// of the smallest non-negative distance.
template <typename T>
OGJK_DI int epa_closest_face_first(const EpaShared<T>& S, int max_face, int lane, T& closest_distance) {
  T bd = epa_big<T>();
  int bi = 0x7fffffff;
  for (int f = lane; f < max_face; f += 32) {
    if (!S.fvalid[f]) continue;
    const T d = S.fd[f];
    if (d >= (T)0 && d < bd) { bd = d; bi = f; }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const T od = __shfl_xor_sync(0xffffffffu, bd, m);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
  }
  closest_distance = bd;
  return bi == 0x7fffffff ? -1 : bi;
}

template <typename T>
__global__ void __launch_bounds__(32 * kEpaWarps)
epa_warp_kernel(EpaArgs<T> a, const int* __restrict__ list, const unsigned long long* __restrict__ count,
                unsigned long long* __restrict__ head) {
  __shared__ EpaShared<T> smem[kEpaWarps];
  EpaShared<T>& S = smem[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const long long total = (long long)*count;
  const T eps = Lim<T>::eps();
  const T tolerance = (T)eps * 1e2f;  // eps_tot22 :42, :2529

  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(head, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((long long)slot >= total) break;
    const long long p = list[slot];
    const long long h1 = a.idx1 ? (long long)a.idx1[p] : p;
    const long long h2 = a.idx2 ? (long long)a.idx2[p] : p;
    const EpaBody<T> b1{a.s1.data + a.s1.boff[h1], a.s1.cnt[h1], pad4(a.s1.cnt[h1])};
    const EpaBody<T> b2{a.s2.data + a.s2.boff[h2], a.s2.cnt[h2], pad4(a.s2.cnt[h2])};

    // private copy of the GJK simplex (the reference grows a local copy, :2049)
    const GkSimplex<T>* gs = a.simplices + p;
    int nv = gs->nvrtx;
    MV<T> q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      q[k] = MV<T>{gs->vrtx[k][0], gs->vrtx[k][1], gs->vrtx[k][2], gs->vrtx_idx[k][0], gs->vrtx_idx[k][1]};

    bool done = false;
    if (nv != 4) {  // :2086-2487 grow to a tetrahedron
      if (nv == 1) {
        const MV<T> w = epa_support(b1, b2, V3<T>{q[0].x, q[0].y, q[0].z}, lane);
        if (epa_far(q[0], w.x, w.y, w.z)) { q[1] = w; nv = 2; }
        else { epa_no_progress(a, p, b1, b2, w, lane); done = true; }
      }
      if (!done && nv == 2) {
        const V3<T> edge{q[1].x - q[0].x, q[1].y - q[0].y, q[1].z - q[0].z};
        V3<T> axis{(T)1, (T)0, (T)0};
        const T en = sqrt(edge.x * edge.x + edge.y * edge.y + edge.z * edge.z);
        if (en > eps && fabs(edge.x) > 0.9f * en) axis = V3<T>{(T)0, (T)1, (T)0};
        V3<T> dir = cross3(edge, axis);
        if (nrm2(dir.x, dir.y, dir.z) < eps) dir = cross3(edge, V3<T>{(T)0, (T)0, (T)1});
        const MV<T> w = epa_support(b1, b2, dir, lane);
        if (epa_far(q[0], w.x, w.y, w.z) && epa_far(q[1], w.x, w.y, w.z)) { q[2] = w; nv = 3; }
        else { epa_no_progress(a, p, b1, b2, w, lane); done = true; }
      }
      if (!done && nv == 3) {
        const V3<T> e0{q[1].x - q[0].x, q[1].y - q[0].y, q[1].z - q[0].z};
        const V3<T> e1{q[2].x - q[0].x, q[2].y - q[0].y, q[2].z - q[0].z};
        V3<T> dir = cross3(e0, e1);
        MV<T> w = epa_support(b1, b2, dir, lane);
        bool fresh = epa_far(q[0], w.x, w.y, w.z) && epa_far(q[1], w.x, w.y, w.z) && epa_far(q[2], w.x, w.y, w.z);
        if (!fresh) {
          dir = V3<T>{-dir.x, -dir.y, -dir.z};
          w = epa_support(b1, b2, dir, lane);
          fresh = epa_far(q[0], w.x, w.y, w.z) && epa_far(q[1], w.x, w.y, w.z) && epa_far(q[2], w.x, w.y, w.z);
        }
        if (fresh) { q[3] = w; nv = 4; }
        else { epa_no_progress(a, p, b1, b2, w, lane); done = true; }
      }
      if (!done && nv != 4) {  // :2461-2486 (nvrtx outside 1..4)
        if (lane == 0) a.distances[p] = (T)0;
        done = true;
      }
    }
    if (done) continue;

    // ---- init_epa_polytope :1828-1937 ------------------------------------------
    for (int f = lane; f < kEpaFaces; f += 32) { S.fvalid[f] = 0; S.fd[f] = epa_big<T>(); }
    if (lane < 4) {
      const MV<T> m = lane == 0 ? q[0] : (lane == 1 ? q[1] : (lane == 2 ? q[2] : q[3]));
      S.vx[lane] = m.x; S.vy[lane] = m.y; S.vz[lane] = m.z; S.via[lane] = m.ia; S.vib[lane] = m.ib;
    }
    V3<T> centroid{(T)0, (T)0, (T)0};
#pragma unroll
    for (int k = 0; k < 4; ++k) { centroid.x += q[k].x; centroid.y += q[k].y; centroid.z += q[k].z; }
    centroid.x /= 4.0f; centroid.y /= 4.0f; centroid.z /= 4.0f;
    __syncwarp();
    if (lane < 4) {
      // faces (0,1,2) (0,3,1) (0,2,3) (1,3,2)
      int t0 = lane == 3 ? 1 : 0;
      int t1 = lane == 0 ? 1 : (lane == 2 ? 2 : 3);
      int t2 = lane == 0 ? 2 : (lane == 1 ? 1 : (lane == 2 ? 3 : 2));
      const V3<T> v0{S.vx[t0], S.vy[t0], S.vz[t0]}, v1{S.vx[t1], S.vy[t1], S.vz[t1]}, v2{S.vx[t2], S.vy[t2], S.vz[t2]};
      if (epa_faces_centroid(v0, v1, v2, centroid)) { const int t = t1; t1 = t2; t2 = t; }
      S.f0[lane] = (unsigned char)t0; S.f1[lane] = (unsigned char)t1; S.f2[lane] = (unsigned char)t2;
      S.fvalid[lane] = 1;
    }
    __syncwarp();
    int nverts = 4, max_face = 4, iteration = 0;

    // ---- main loop :2533-2916 ---------------------------------------------------
    while (iteration < kEpaMaxIter && nverts < kEpaVerts - 1) {
      iteration++;
      epa_face_pass(S, max_face, centroid, lane);
      T closest_distance;
      const int closest = epa_closest_face(S, max_face, lane, closest_distance);
      if (closest < 0) break;  // :2605-2607 nothing written: defaults stand

      const V3<T> dir{S.fnx[closest], S.fny[closest], S.fnz[closest]};
      const MV<T> w = epa_support(b1, b2, dir, lane);
      const T dist_to_new = dir.x * w.x + dir.y * w.y + dir.z * w.z;
      const T improvement = dist_to_new - closest_distance;
      if (improvement < tolerance) {  // :2624 converged
        epa_finish_on_face(a, p, S, closest, closest_distance, b1, b2, lane);
        break;
      }
      bool dup = false;  // :2665-2677
      for (int i = lane; i < nverts; i += 32) {
        const T dx = w.x - S.vx[i], dy = w.y - S.vy[i], dz = w.z - S.vz[i];
        if (dx * dx + dy * dy + dz * dz < eps * eps) dup = true;
      }
      if (__any_sync(0xffffffffu, dup)) {
        epa_finish_on_face(a, p, S, closest, closest_distance, b1, b2, lane);
        break;
      }

      // add the vertex, incremental centroid :2719-2736
      const int new_id = nverts;
      if (lane == 0) { S.vx[new_id] = w.x; S.vy[new_id] = w.y; S.vz[new_id] = w.z; S.via[new_id] = w.ia; S.vib[new_id] = w.ib; }
      nverts++;
      {
        const T n = (T)nverts;
        centroid.x = centroid.x * (n - 1.0f) / n + w.x / n;
        centroid.y = centroid.y * (n - 1.0f) / n + w.y / n;
        centroid.z = centroid.z * (n - 1.0f) / n + w.z / n;
      }
      // faces that see the new vertex :1720-1732, :2758-2762
      for (int f = lane; f < max_face; f += 32) {
        if (!S.fvalid[f]) continue;
        const int i0 = S.f0[f];
        const T dx = w.x - S.vx[i0], dy = w.y - S.vy[i0], dz = w.z - S.vz[i0];
        if (S.fnx[f] * dx + S.fny[f] * dy + S.fnz[f] * dz > eps) S.fvalid[f] = 0;
      }
      __syncwarp();
      // edges of every invalid face below max_face, in face order :2765-2810
      int ninv = 0;
      for (int base = 0; base < max_face; base += 32) {
        const int f = base + lane;
        const bool inv = f < max_face && !S.fvalid[f];
        const unsigned m = __ballot_sync(0xffffffffu, inv);
        if (inv) {
          const int e = 3 * (ninv + __popc(m & lt));
          const unsigned char x0 = S.f0[f], x1 = S.f1[f], x2 = S.f2[f];
          S.e1[e] = x0; S.e2[e] = x1;
          S.e1[e + 1] = x1; S.e2[e + 1] = x2;
          S.e1[e + 2] = x2; S.e2[e + 2] = x0;
        }
        ninv += __popc(m);
      }
      const int ne = 3 * ninv;
      __syncwarp();
      // an edge survives iff no other edge joins the same two vertices :2813-2826
      for (int i = lane; i < ne; i += 32) {
        const int x = S.e1[i], y = S.e2[i];
        const int key = x < y ? (x << 8) | y : (y << 8) | x;
        bool alone = true;
        for (int j = 0; j < ne; ++j) {
          const int u = S.e1[j], v = S.e2[j];
          const int kj = u < v ? (u << 8) | v : (v << 8) | u;
          if (kj == key && j != i) alone = false;
        }
        S.evalid[i] = alone ? 1 : 0;
      }
      // free face slots in ascending order :2833-2840
      int nfree = 0;
      for (int base = 0; base < kEpaFaces; base += 32) {
        const int f = base + lane;
        const bool fr = !S.fvalid[f];
        const unsigned m = __ballot_sync(0xffffffffu, fr);
        if (fr) S.free_list[nfree + __popc(m & lt)] = (unsigned char)f;
        nfree += __popc(m);
      }
      __syncwarp();
      // k-th surviving edge -> k-th free slot :2829-2897
      int nsurv = 0, top = 0;
      for (int base = 0; base < ne; base += 32) {
        const int i = base + lane;
        const bool sv = i < ne && S.evalid[i];
        const unsigned m = __ballot_sync(0xffffffffu, sv);
        const int k = nsurv + __popc(m & lt);
        if (sv && k < nfree) {
          const int sl = S.free_list[k];
          int t0 = S.e1[i], t1 = S.e2[i], t2 = new_id;
          const V3<T> v0{S.vx[t0], S.vy[t0], S.vz[t0]}, v1{S.vx[t1], S.vy[t1], S.vz[t1]};
          const V3<T> v2{w.x, w.y, w.z};
          if (epa_faces_centroid(v0, v1, v2, centroid)) { const int t = t1; t1 = t2; t2 = t; }
          S.f0[sl] = (unsigned char)t0; S.f1[sl] = (unsigned char)t1; S.f2[sl] = (unsigned char)t2;
          S.fvalid[sl] = 1;
          top = sl + 1 > top ? sl + 1 : top;
        }
        nsurv += __popc(m);
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const int o = __shfl_xor_sync(0xffffffffu, top, m);
        top = o > top ? o : top;
      }
      if (top > max_face) max_face = top;
      __syncwarp();
    }

    if (iteration >= kEpaMaxIter) {  // :2919-2970
      epa_face_pass(S, max_face, centroid, lane);
      T closest_distance;
      const int closest = epa_closest_face_first(S, max_face, lane, closest_distance);
      if (closest >= 0) epa_finish_on_face(a, p, S, closest, closest_distance, b1, b2, lane);
    }
    __syncwarp();
  }
}

}  // namespace ogjk
