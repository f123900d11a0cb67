// compat_gpu.cu -- reference-named batched flavour: the 14 host symbols of the
// reference's GPU library (gpu/include/openGJK_GPU.h:153-399, export list
// gpu/examples/python/openGJK_GPU.def) with the reference's signatures, implemented
// over the C ABI of libopengjk_b200.so.  Built twice: libopenGJK_GPU.so (USE_32BITS,
// what pyopengjk_gpu assumes, opengjk_gpu.py:25) and libopenGJK_GPU_f64.so.
// The wrappers keep the reference's `void` returns and synchronous behaviour.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "../../include/opengjk_b200.h"

#ifdef USE_32BITS
typedef float gkFloat;
typedef ogjk_polytope_f32 gkPolytope;
typedef ogjk_simplex_f32 gkSimplex;
#define F(name) name##_f32
#else
typedef double gkFloat;
typedef ogjk_polytope_f64 gkPolytope;
typedef ogjk_simplex_f64 gkSimplex;
#define F(name) name##_f64
#endif
typedef ogjk_pair gkCollisionPair;

#define EXPORT extern "C" __attribute__((visibility("default")))

static ogjk_ctx* ctx() {
  static ogjk_ctx* c = nullptr;
  if (!c) {
    int dev = 0;
    cudaGetDevice(&dev);
    ogjk_ctx_create(dev, &c);
  }
  return c;
}

EXPORT void compute_minimum_distance(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                                     gkFloat* distances) {
  F(ogjk_compute_minimum_distance)(ctx(), n, bd1, bd2, simplices, distances);
}

EXPORT void compute_epa(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices, gkFloat* distances,
                        gkFloat* witness1, gkFloat* witness2, gkFloat* contact_normals) {
  F(ogjk_compute_epa)(ctx(), n, bd1, bd2, simplices, distances, witness1, witness2, contact_normals);
}

EXPORT void compute_gjk_epa(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                            gkFloat* distances, gkFloat* witness1, gkFloat* witness2) {
  F(ogjk_compute_gjk_epa)(ctx(), n, bd1, bd2, simplices, distances, witness1, witness2);
}

EXPORT void compute_minimum_distance_indexed(const int num_polytopes, const int num_pairs, const gkPolytope* polytopes,
                                             const gkCollisionPair* pairs, gkSimplex* simplices, gkFloat* distances) {
  F(ogjk_compute_minimum_distance_indexed)(ctx(), num_polytopes, num_pairs, polytopes, pairs, simplices, distances);
}

// Mid level.  Same device data shape as the reference (arrays of gkPolytope whose coord
// members point into one concatenated device block), but two bulk copies per side
// instead of one cudaMemcpy per polytope (gpu/opengjk_gpu.cu:3152-3153).
static void upload_side(int n, const gkPolytope* bd, gkPolytope** d_bd, gkFloat** d_coord) {
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += (size_t)bd[i].numpoints * 3;
  cudaMalloc((void**)d_bd, (size_t)n * sizeof(gkPolytope));
  cudaMalloc((void**)d_coord, (total ? total : 1) * sizeof(gkFloat));
  std::vector<gkFloat> flat(total);
  std::vector<gkPolytope> meta(bd, bd + n);
  size_t at = 0;
  for (int i = 0; i < n; ++i) {
    const size_t len = (size_t)bd[i].numpoints * 3;
    memcpy(flat.data() + at, bd[i].coord, len * sizeof(gkFloat));
    meta[i].coord = *d_coord + at;
    at += len;
  }
  cudaMemcpy(*d_coord, flat.data(), total * sizeof(gkFloat), cudaMemcpyHostToDevice);
  cudaMemcpy(*d_bd, meta.data(), (size_t)n * sizeof(gkPolytope), cudaMemcpyHostToDevice);
}

EXPORT void allocate_and_copy_device_arrays(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkPolytope** d_bd1,
                                            gkPolytope** d_bd2, gkFloat** d_coord1, gkFloat** d_coord2,
                                            gkSimplex** d_simplices, gkFloat** d_distances) {
  if (n <= 0) return;
  ctx();
  upload_side(n, bd1, d_bd1, d_coord1);
  upload_side(n, bd2, d_bd2, d_coord2);
  cudaMalloc((void**)d_simplices, (size_t)n * sizeof(gkSimplex));
  cudaMalloc((void**)d_distances, (size_t)n * sizeof(gkFloat));
  cudaMemset(*d_simplices, 0, (size_t)n * sizeof(gkSimplex));
}

EXPORT void allocate_epa_device_arrays(const int n, gkFloat** d_witness1, gkFloat** d_witness2, gkFloat** d_contact_normals) {
  cudaMalloc((void**)d_witness1, (size_t)n * 3 * sizeof(gkFloat));
  cudaMalloc((void**)d_witness2, (size_t)n * 3 * sizeof(gkFloat));
  if (d_contact_normals) cudaMalloc((void**)d_contact_normals, (size_t)n * 3 * sizeof(gkFloat));
}

EXPORT void compute_minimum_distance_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2,
                                            gkSimplex* d_simplices, gkFloat* d_distances) {
  F(ogjk_gjk_device_structs)(ctx(), n, d_bd1, d_bd2, d_simplices, d_distances, nullptr);
  cudaDeviceSynchronize();
}

EXPORT void compute_epa_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2, gkSimplex* d_simplices,
                               gkFloat* d_distances, gkFloat* d_witness1, gkFloat* d_witness2, gkFloat* d_contact_normals) {
  F(ogjk_epa_device_structs)(ctx(), n, d_bd1, d_bd2, d_simplices, d_distances, d_witness1, d_witness2, d_contact_normals,
                             nullptr);
  cudaDeviceSynchronize();
}

EXPORT void copy_results_from_device(const int n, const gkSimplex* d_simplices, const gkFloat* d_distances,
                                     gkSimplex* simplices, gkFloat* distances) {
  cudaMemcpy(distances, d_distances, (size_t)n * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  cudaMemcpy(simplices, d_simplices, (size_t)n * sizeof(gkSimplex), cudaMemcpyDeviceToHost);
}

EXPORT void copy_epa_results_from_device(const int n, const gkFloat* d_witness1, const gkFloat* d_witness2,
                                         const gkFloat* d_contact_normals, gkFloat* witness1, gkFloat* witness2,
                                         gkFloat* contact_normals) {
  cudaMemcpy(witness1, d_witness1, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  cudaMemcpy(witness2, d_witness2, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  if (contact_normals && d_contact_normals)
    cudaMemcpy(contact_normals, d_contact_normals, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost);
}

EXPORT void free_device_arrays(gkPolytope* d_bd1, gkPolytope* d_bd2, gkFloat* d_coord1, gkFloat* d_coord2,
                               gkSimplex* d_simplices, gkFloat* d_distances) {
  cudaFree(d_bd1); cudaFree(d_bd2); cudaFree(d_coord1); cudaFree(d_coord2); cudaFree(d_simplices); cudaFree(d_distances);
}

EXPORT void free_epa_device_arrays(gkFloat* d_witness1, gkFloat* d_witness2, gkFloat* d_contact_normals) {
  cudaFree(d_witness1); cudaFree(d_witness2);
  if (d_contact_normals) cudaFree(d_contact_normals);
}
