// compat_gpu.cu -- reference-named batched flavour: the 14 host symbols of the
// reference's GPU library (gpu/include/openGJK_GPU.h:153-399, export list
// gpu/examples/python/openGJK_GPU.def) with the reference's signatures, implemented
// over the C ABI of libopengjk_b200.so.  Built twice: libopenGJK_GPU.so (USE_32BITS,
// what pyopengjk_gpu assumes, opengjk_gpu.py:25) and libopenGJK_GPU_f64.so.
// The wrappers keep the reference's `void` returns and synchronous behaviour; failures go to
// stderr and turn the distances into NaN (the reference prints its CUDA errors and carries on).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
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

// One context for the library, created once (std::call_once) on the caller's current device;
// the context is not thread safe, so concurrent callers are serialised on a mutex (the
// reference's batched entry points use the legacy default stream + cudaDeviceSynchronize,
// gpu/opengjk_gpu.cu:3198-3201, i.e. they serialise on the device as well).
static ogjk_ctx* g_ctx = nullptr;
static std::once_flag g_once;
static std::mutex g_lock;

static ogjk_ctx* ctx() {
  std::call_once(g_once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    if (ogjk_ctx_create(dev, &g_ctx) != OGJK_OK) {
      g_ctx = nullptr;
      fprintf(stderr, "openGJK_GPU (opengjk_b200): %s\n", ogjk_last_error());
    }
  });
  return g_ctx;
}

// $OPENGJK_B200_DEVICES = "all" or a comma list of device ordinals: the two whole-batch entry points
// (compute_minimum_distance, compute_gjk_epa) spread every call over those GPUs (ogjk_multi_*: contiguous
// pair slice per device, results written straight into the caller's arrays).  Unset: the current device.
static ogjk_multi* g_multi = nullptr;
static std::once_flag g_multi_once;
static ogjk_multi* multi() {
  std::call_once(g_multi_once, [] {
    const char* env = getenv("OPENGJK_B200_DEVICES");
    if (!env || !*env) return;
    std::vector<int> devs;
    if (strcmp(env, "all") != 0)
      for (const char* p = env; *p;) {
        char* end = nullptr;
        const long d = strtol(p, &end, 10);
        if (end == p) break;
        devs.push_back((int)d);
        p = *end == ',' ? end + 1 : end;
      }
    if (strcmp(env, "all") != 0 && devs.empty()) {
      fprintf(stderr, "openGJK_GPU (opengjk_b200): OPENGJK_B200_DEVICES=%s is neither \"all\" nor a list of ordinals; using one device\n", env);
      return;
    }
    if (ogjk_multi_create(devs.empty() ? nullptr : devs.data(), (int)devs.size(), &g_multi) != OGJK_OK) {
      g_multi = nullptr;
      fprintf(stderr, "openGJK_GPU (opengjk_b200): OPENGJK_B200_DEVICES=%s: %s; using one device\n", env, ogjk_last_error());
    }
  });
  return g_multi;
}

// The reference API returns void: a failure is reported on stderr (as the reference prints its
// CUDA errors) and the outputs are filled with NaN so that it cannot pass for a result.
static void report(int rc, const char* what, int n, gkFloat* distances) {
  if (rc == OGJK_OK) return;
  fprintf(stderr, "openGJK_GPU (opengjk_b200): %s failed: %s\n", what, ogjk_last_error());
  if (distances)
    for (int i = 0; i < n; ++i) distances[i] = (gkFloat)NAN;
}
static bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  fprintf(stderr, "openGJK_GPU (opengjk_b200): %s: %s\n", what, cudaGetErrorString(e));
  return false;
}

EXPORT void compute_minimum_distance(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                                     gkFloat* distances) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  if (ogjk_multi* m = multi()) {
    report(F(ogjk_multi_compute_minimum_distance)(m, n, bd1, bd2, simplices, distances), "compute_minimum_distance", n, distances);
    return;
  }
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_compute_minimum_distance)(c, n, bd1, bd2, simplices, distances) : OGJK_ERR_CUDA,
         "compute_minimum_distance", n, distances);
}

EXPORT void compute_epa(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices, gkFloat* distances,
                        gkFloat* witness1, gkFloat* witness2, gkFloat* contact_normals) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_compute_epa)(c, n, bd1, bd2, simplices, distances, witness1, witness2, contact_normals) : OGJK_ERR_CUDA,
         "compute_epa", n, distances);
}

EXPORT void compute_gjk_epa(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                            gkFloat* distances, gkFloat* witness1, gkFloat* witness2) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  if (ogjk_multi* m = multi()) {
    report(F(ogjk_multi_compute_gjk_epa)(m, n, bd1, bd2, simplices, distances, witness1, witness2), "compute_gjk_epa", n, distances);
    return;
  }
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_compute_gjk_epa)(c, n, bd1, bd2, simplices, distances, witness1, witness2) : OGJK_ERR_CUDA,
         "compute_gjk_epa", n, distances);
}

EXPORT void compute_minimum_distance_indexed(const int num_polytopes, const int num_pairs, const gkPolytope* polytopes,
                                             const gkCollisionPair* pairs, gkSimplex* simplices, gkFloat* distances) {
  if (num_pairs <= 0 || num_polytopes <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_compute_minimum_distance_indexed)(c, num_polytopes, num_pairs, polytopes, pairs, simplices, distances)
           : OGJK_ERR_CUDA,
         "compute_minimum_distance_indexed", num_pairs, distances);
}

// Mid level.  Same device data shape as the reference (arrays of gkPolytope whose coord
// members point into one concatenated device block), but two bulk copies per side
// instead of one cudaMemcpy per polytope (gpu/opengjk_gpu.cu:3152-3153).
static void upload_side(int n, const gkPolytope* bd, gkPolytope** d_bd, gkFloat** d_coord) {
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += (size_t)bd[i].numpoints * 3;
  *d_bd = nullptr; *d_coord = nullptr;
  if (!cuda_ok(cudaMalloc((void**)d_bd, (size_t)n * sizeof(gkPolytope)), "cudaMalloc(polytopes)")) return;
  if (!cuda_ok(cudaMalloc((void**)d_coord, (total ? total : 1) * sizeof(gkFloat)), "cudaMalloc(coordinates)")) return;
  std::vector<gkFloat> flat(total);
  std::vector<gkPolytope> meta(bd, bd + n);
  size_t at = 0;
  for (int i = 0; i < n; ++i) {
    const size_t len = (size_t)bd[i].numpoints * 3;
    memcpy(flat.data() + at, bd[i].coord, len * sizeof(gkFloat));
    meta[i].coord = *d_coord + at;
    at += len;
  }
  cuda_ok(cudaMemcpy(*d_coord, flat.data(), total * sizeof(gkFloat), cudaMemcpyHostToDevice), "cudaMemcpy(coordinates)");
  cuda_ok(cudaMemcpy(*d_bd, meta.data(), (size_t)n * sizeof(gkPolytope), cudaMemcpyHostToDevice), "cudaMemcpy(polytopes)");
}

EXPORT void allocate_and_copy_device_arrays(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkPolytope** d_bd1,
                                            gkPolytope** d_bd2, gkFloat** d_coord1, gkFloat** d_coord2,
                                            gkSimplex** d_simplices, gkFloat** d_distances) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  ctx();
  upload_side(n, bd1, d_bd1, d_coord1);
  upload_side(n, bd2, d_bd2, d_coord2);
  if (cuda_ok(cudaMalloc((void**)d_simplices, (size_t)n * sizeof(gkSimplex)), "cudaMalloc(simplices)"))
    cuda_ok(cudaMemset(*d_simplices, 0, (size_t)n * sizeof(gkSimplex)), "cudaMemset(simplices)");
  cuda_ok(cudaMalloc((void**)d_distances, (size_t)n * sizeof(gkFloat)), "cudaMalloc(distances)");
}

EXPORT void allocate_epa_device_arrays(const int n, gkFloat** d_witness1, gkFloat** d_witness2, gkFloat** d_contact_normals) {
  cuda_ok(cudaMalloc((void**)d_witness1, (size_t)n * 3 * sizeof(gkFloat)), "cudaMalloc(witness1)");
  cuda_ok(cudaMalloc((void**)d_witness2, (size_t)n * 3 * sizeof(gkFloat)), "cudaMalloc(witness2)");
  if (d_contact_normals) cuda_ok(cudaMalloc((void**)d_contact_normals, (size_t)n * 3 * sizeof(gkFloat)), "cudaMalloc(normals)");
}

EXPORT void compute_minimum_distance_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2,
                                            gkSimplex* d_simplices, gkFloat* d_distances) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_gjk_device_structs)(c, n, d_bd1, d_bd2, d_simplices, d_distances, nullptr) : OGJK_ERR_CUDA,
         "compute_minimum_distance_device", 0, nullptr);
  cuda_ok(cudaDeviceSynchronize(), "compute_minimum_distance_device");
}

EXPORT void compute_epa_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2, gkSimplex* d_simplices,
                               gkFloat* d_distances, gkFloat* d_witness1, gkFloat* d_witness2, gkFloat* d_contact_normals) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(g_lock);
  ogjk_ctx* c = ctx();
  report(c ? F(ogjk_epa_device_structs)(c, n, d_bd1, d_bd2, d_simplices, d_distances, d_witness1, d_witness2, d_contact_normals,
                                        nullptr)
           : OGJK_ERR_CUDA,
         "compute_epa_device", 0, nullptr);
  cuda_ok(cudaDeviceSynchronize(), "compute_epa_device");
}

EXPORT void copy_results_from_device(const int n, const gkSimplex* d_simplices, const gkFloat* d_distances,
                                     gkSimplex* simplices, gkFloat* distances) {
  cuda_ok(cudaMemcpy(distances, d_distances, (size_t)n * sizeof(gkFloat), cudaMemcpyDeviceToHost), "copy_results_from_device");
  cuda_ok(cudaMemcpy(simplices, d_simplices, (size_t)n * sizeof(gkSimplex), cudaMemcpyDeviceToHost), "copy_results_from_device");
}

EXPORT void copy_epa_results_from_device(const int n, const gkFloat* d_witness1, const gkFloat* d_witness2,
                                         const gkFloat* d_contact_normals, gkFloat* witness1, gkFloat* witness2,
                                         gkFloat* contact_normals) {
  cuda_ok(cudaMemcpy(witness1, d_witness1, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost), "copy_epa_results_from_device");
  cuda_ok(cudaMemcpy(witness2, d_witness2, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost), "copy_epa_results_from_device");
  if (contact_normals && d_contact_normals)
    cuda_ok(cudaMemcpy(contact_normals, d_contact_normals, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost),
            "copy_epa_results_from_device");
}

EXPORT void free_device_arrays(gkPolytope* d_bd1, gkPolytope* d_bd2, gkFloat* d_coord1, gkFloat* d_coord2,
                               gkSimplex* d_simplices, gkFloat* d_distances) {
  cudaFree(d_bd1); cudaFree(d_bd2); cudaFree(d_coord1); cudaFree(d_coord2); cudaFree(d_simplices); cudaFree(d_distances);
}

EXPORT void free_epa_device_arrays(gkFloat* d_witness1, gkFloat* d_witness2, gkFloat* d_contact_normals) {
  cudaFree(d_witness1); cudaFree(d_witness2);
  if (d_contact_normals) cudaFree(d_contact_normals);
}
