/*
 * opengjk_b200.h -- C ABI of the B200-native batch GJK(+EPA) engine.
 *
 * Plain C, no CUDA or C++ types in any signature.  This is the drop-in
 * boundary for the reference's batched collision-query path; each entry point
 * cites the reference interface it replaces (paths under /root/reference).
 * Both precisions live in one shared library under _f32 / _f64 suffixes (the
 * reference selects one at build time with USE_32BITS,
 * scalar/include/openGJK/openGJK.h:54-60; the reference-named flavour
 * libraries in include/openGJK/ re-export the unsuffixed names).
 *
 * Data model ("polytope pool" + "pair list"):
 *   coords : xyz-interleaved vertices of all polytopes of a pool, contiguous
 *            per polytope -- the reference's flattened `coord` blocks
 *            (gpu/include/openGJK_GPU.h:71-76) laid end to end, exactly what
 *            its allocate_and_copy_device_arrays builds (gpu/opengjk_gpu.cu:3105-3166)
 *   off    : first vertex (not byte, not float) of each polytope, int64
 *   cnt    : vertex count of each polytope (gkPolytope.numpoints)
 *   idx    : per pair, which polytope of the pool is body 1 / body 2
 *            (gkCollisionPair, openGJK_GPU.h:89-92); NULL means "pair p uses
 *            polytope p" (the reference's parallel bd1[]/bd2[] arrays)
 *
 * All functions return 0 on success or a negative ogjk_status; the message is
 * available from ogjk_last_error() (thread local).  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with OGJK_ERR_CUDA.
 */
#ifndef OPENGJK_B200_H_
#define OPENGJK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OGJK_API __attribute__((visibility("default")))
#else
#define OGJK_API
#endif

typedef enum ogjk_status_ {
  OGJK_OK = 0,
  OGJK_ERR_ARG = -1,   /* bad argument (NULL pointer, negative count, ...) */
  OGJK_ERR_CUDA = -2,  /* CUDA runtime error or no device */
  OGJK_ERR_ALLOC = -3  /* host or device allocation failed */
} ogjk_status;

/* Layout-identical to the reference structs (offsets verified in SURVEY.md 8):
 * fp32 gkPolytope 32 B, gkSimplex 108 B; fp64 gkPolytope 48 B, gkSimplex 184 B. */
typedef struct ogjk_polytope_f32_ { int numpoints; float s[3]; int s_idx; float* coord; } ogjk_polytope_f32;
typedef struct ogjk_polytope_f64_ { int numpoints; double s[3]; int s_idx; double* coord; } ogjk_polytope_f64;
typedef struct ogjk_simplex_f32_ { int nvrtx; float vrtx[4][3]; int vrtx_idx[4][2]; float witnesses[2][3]; } ogjk_simplex_f32;
typedef struct ogjk_simplex_f64_ { int nvrtx; double vrtx[4][3]; int vrtx_idx[4][2]; double witnesses[2][3]; } ogjk_simplex_f64;
typedef struct ogjk_pair_ { int idx1; int idx2; } ogjk_pair; /* gkCollisionPair */

typedef struct ogjk_ctx_ ogjk_ctx;     /* one CUDA device + its scratch memory */
typedef struct ogjk_store_ ogjk_store; /* a packed, device-resident polytope pool */

/* Kernel selection for ogjk_gjk_device_*.  AUTO repacks the pools into the padded
 * SoA store and runs the size-binned scheduler; the other two run the
 * first-generation kernels straight off the xyz-interleaved arrays. */
enum { OGJK_MODE_AUTO = 0, OGJK_MODE_THREAD_PER_PAIR = 1, OGJK_MODE_WARP_PER_PAIR = 2 };

OGJK_API const char* ogjk_last_error(void);
OGJK_API const char* ogjk_version(void);

/* Context.  `device` is a CUDA ordinal.  Not thread safe per context; distinct
 * contexts may be used from distinct threads (replaces the reference's use of
 * the legacy default stream + cudaDeviceSynchronize, gpu/opengjk_gpu.cu:3198-3201). */
OGJK_API int ogjk_ctx_create(int device, ogjk_ctx** out);
OGJK_API void ogjk_ctx_destroy(ogjk_ctx* ctx);
OGJK_API int ogjk_ctx_device(const ogjk_ctx* ctx);
/* Raw polytope pools (coords/off/cnt) are validated on the device while they are repacked:
 * cnt[h] must lie in [1, nverts] and [off[h], off[h]+cnt[h]) inside [0, nverts).  The
 * DEVICE-array entry points (ogjk_gjk_device_* AUTO, ogjk_store_create_* with on_device = 1)
 * further require sum(cnt) <= nverts, i.e. polytopes that do not share vertices (the
 * asynchronous repack sizes its buffer from nverts); host-buffer entry points accept shared
 * and overlapping ranges.  An offending polytope is replaced by a dummy (never an
 * out-of-bounds access) and a sticky flag is raised: the synchronising entry points
 * (host-buffer calls, ogjk_store_create_*) report it as OGJK_ERR_ARG; after asynchronous
 * device calls, ogjk_ctx_pool_status() synchronises the device and returns OGJK_ERR_ARG if
 * any pool since the last report failed validation (results of that call are undefined),
 * else OGJK_OK.  Reporting clears the flag. */
OGJK_API int ogjk_ctx_pool_status(ogjk_ctx* ctx);
/* Number of kernels launched by this context so far (bench "gpu_launches"). */
OGJK_API int64_t ogjk_ctx_launch_count(const ogjk_ctx* ctx);
/* Per-phase device timing for the roofline report: when enabled, CUDA events
 * bracket every phase of the next device call on the caller's stream;
 * ogjk_ctx_last_timing fills ms8[8] = {repack, sort, class 0 .. class 5} in
 * milliseconds (0 = phase not run).  Raw modes report their kernel in class 0/1. */
OGJK_API int ogjk_ctx_set_timing(ogjk_ctx* ctx, int enable);
OGJK_API int ogjk_ctx_last_timing(ogjk_ctx* ctx, float* ms8);
/* Scheduler table (north_star item 5).  Pairs are counting-sorted by
 * key = (pad4(n1) + pad4(n2)) / 4; class i takes keys in (hi_keys[i-1], hi_keys[i]]
 * and runs with lanes[i] cooperating lanes per pair (1, 2, 4, 8, 16 or 32; add 100
 * for the flavour that stages both polytopes in shared memory first, 200 for the
 * persistent-lane flavour that refills a lane as soon as its pair terminates, 300 for the
 * scan-assist flavour: 32/lanes pairs per warp share the sub-solver pass while every
 * support scan uses all 32 lanes), or
 * lanes[i] = 0 for the register-resident thread-per-pair kernel (keys <= 4 only:
 * both polytopes have at most 8 vertices).  At most 6 classes; the last one
 * always extends to the largest key, so it cannot be the register-resident kernel.
 * The product library carries the plain flavours (lanes 0, 1..32) only; the 100/200/300/500
 * families are measured-and-lost experiments compiled into libopengjk_b200_exp.so
 * (-DOGJK_EXPERIMENTAL) and rejected with OGJK_ERR_ARG elsewhere. */
OGJK_API int ogjk_ctx_set_classes(ogjk_ctx* ctx, int n, const int* hi_keys, const int* lanes);
/* Size (in vertices of body1+body2) from which AUTO uses the warp-cooperative kernel. */
OGJK_API int ogjk_ctx_set_warp_threshold(ogjk_ctx* ctx, int nverts);

/* ---- device-resident batch (replaces compute_minimum_distance_device and the
 * indexed kernel launch: gpu/include/openGJK_GPU.h:236-242, gpu/opengjk_gpu.cu:3186-3202,
 * 3331-3336).  Every pointer is a DEVICE pointer.  `simplices` may be NULL
 * (distance only).  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  npoly/nverts are the pool sizes (polytopes and
 * vertices behind coords/off/cnt).  Asynchronous: returns after enqueueing. */
OGJK_API int ogjk_gjk_device_f32(ogjk_ctx* ctx, int64_t npairs,
                                 const float* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                                 const float* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                                 ogjk_simplex_f32* simplices, float* distances, int mode, void* stream);
OGJK_API int ogjk_gjk_device_f64(ogjk_ctx* ctx, int64_t npairs,
                                 const double* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                                 const double* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                                 ogjk_simplex_f64* simplices, double* distances, int mode, void* stream);

/* ---- packed polytope store: the engine's device-resident format.  Built once
 * (from host arrays, on_device = 0, or from device arrays, on_device = 1) and
 * queried many times -- the role of allocate_and_copy_device_arrays +
 * compute_minimum_distance_device + free_device_arrays in the reference
 * (gpu/include/openGJK_GPU.h:263-273, 236-242, 375-382), whose "static objects
 * that persist across frames" use case it serves.  idx1/idx2 are DEVICE arrays
 * of polytope indices (NULL = pair p uses polytope p); s1 and s2 may be the same
 * store (indexed mode).  Asynchronous on `stream`. */
OGJK_API int ogjk_store_create_f32(ogjk_ctx* ctx, int64_t npoly, int64_t nverts, const float* coords, const int64_t* off,
                                   const int* cnt, int on_device, void* stream, ogjk_store** out);
OGJK_API int ogjk_store_create_f64(ogjk_ctx* ctx, int64_t npoly, int64_t nverts, const double* coords, const int64_t* off,
                                   const int* cnt, int on_device, void* stream, ogjk_store** out);
/* Stores created AFTER this call give every polytope with at least `min_vertices`
 * vertices (default 512, the measured break-even on B200; clamped to >= 64; at most 16384 vertices; 0 = off) a cluster
 * table: a spatially clustered copy of its vertices with per-cluster bounding boxes that
 * lets the warp-per-pair kernel skip clusters which provably cannot change the support
 * point.  Results are bit-identical with and without the table. */
OGJK_API int ogjk_ctx_set_accel_min(ogjk_ctx* ctx, int min_vertices);
/* Multi-pass scheduling of the small-group size classes (1..8 lanes per pair): with passes = 2
 * or 3, pairs still iterating after `iters1` (then `iters1 + iters2`) GJK iterations are parked
 * and resumed densely packed in the next pass, so that warps do not idle on their slowest
 * pair.  passes = 1 is the one-pass kernel.  A tuning knob; results do not depend on it. */
OGJK_API int ogjk_ctx_set_passes(ogjk_ctx* ctx, int passes, int iters1, int iters2);
/* Kernels that use the cluster tables.  0 (default): pairs with a body above 1024 vertices run the
 * tile kernel (csrc/table_kernel.cuh: register-resident cluster boxes, tile boxes, vector cluster
 * fetches); pairs whose two bodies fit one 32-cluster tile run the pair-batched kernel (a warp
 * owns up to 32 pairs: per-lane sub-solver, warp-wide support scans one pair at a time) when the
 * batch has at least 131072 pairs, and the first-generation warp-per-pair kernel below that.
 * Tuning values: 1 = the pair-batched kernel for every table pair; 2 = tile kernel above one tile,
 * pair-batched kernel below, whatever the batch size; 3 = the tile kernel for every table pair;
 * 8, 16 or 32 = the first-generation kernel with that many lanes per pair for every table pair.
 * Results do not depend on the choice. */
OGJK_API int ogjk_ctx_set_table_lanes(ogjk_ctx* ctx, int lanes);
/* In-CTA regrouping of the 2-lane size classes (csrc/regroup_kernel.cuh): a CTA takes 128 pairs, runs
 * them iteration-synchronously and packs the pairs still iterating after `iters1` (and `iters2`) GJK
 * iterations into the lowest lanes through shared memory, so that whole warps fall idle instead of
 * lanes.  Bit-exact but measured 25 % slower than the plain group kernel on B200,
 * so it is compiled into libopengjk_b200_exp.so only: the product library accepts 0 (off) and rejects
 * anything else.  iters2 = 0 means one regroup point. */
OGJK_API int ogjk_ctx_set_regroup(ogjk_ctx* ctx, int iters1, int iters2);
OGJK_API void ogjk_store_destroy(ogjk_store* store);
OGJK_API int64_t ogjk_store_bytes(const ogjk_store* store);
OGJK_API int64_t ogjk_store_num_polytopes(const ogjk_store* store);
/* Bytes of cluster tables the store carries (0: none, queries use the plain scans). */
OGJK_API int64_t ogjk_store_table_bytes(const ogjk_store* store);
OGJK_API int ogjk_store_gjk_f32(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                int64_t npairs, ogjk_simplex_f32* simplices, float* distances, void* stream);
OGJK_API int ogjk_store_gjk_f64(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                int64_t npairs, ogjk_simplex_f64* simplices, double* distances, void* stream);

/* ---- broad-phase-adjacent pair production (device pointers, asynchronous on `stream`) ----
 * The caller step just before the indexed entries (gkCollisionPair lists,
 * gpu/include/openGJK_GPU.h:89-92, 331-366).  Output lists are deterministic: culling keeps
 * the candidates' order, the all-pairs list is sorted by (i, j).
 *
 * ogjk_store_aabbs_*: d_boxes[6*h .. 6*h+5] = {min x, min y, min z, max x, max y, max z} of
 *   polytope h of the store.
 * ogjk_pairs_cull_*: of the `ncand` candidates (d_idx1[p] into boxes1, d_idx2[p] into boxes2;
 *   indices are not range-checked) keep those not separated by more than `margin` on any axis
 *   (loA - hiB > margin or loB - hiA > margin culls; NaN never culls).  *d_count (device) receives
 *   the number kept; at most `capacity` of them are written to d_out1 / d_out2, which can be
 *   passed as idx1 / idx2 to ogjk_store_gjk_*.
 * ogjk_pairs_all_*: the same test over every i < j of one pool (warp per row, O(n^2) box tests). */
OGJK_API int ogjk_store_aabbs_f32(ogjk_ctx* ctx, const ogjk_store* store, float* d_boxes, void* stream);
OGJK_API int ogjk_store_aabbs_f64(ogjk_ctx* ctx, const ogjk_store* store, double* d_boxes, void* stream);
OGJK_API int ogjk_pairs_cull_f32(ogjk_ctx* ctx, const float* d_boxes1, const float* d_boxes2, int64_t ncand,
                                 const int* d_idx1, const int* d_idx2, float margin, int64_t capacity, int* d_out1,
                                 int* d_out2, int64_t* d_count, void* stream);
OGJK_API int ogjk_pairs_cull_f64(ogjk_ctx* ctx, const double* d_boxes1, const double* d_boxes2, int64_t ncand,
                                 const int* d_idx1, const int* d_idx2, double margin, int64_t capacity, int* d_out1,
                                 int* d_out2, int64_t* d_count, void* stream);
OGJK_API int ogjk_pairs_all_f32(ogjk_ctx* ctx, const float* d_boxes, int64_t npoly, float margin, int64_t capacity,
                                int* d_out1, int* d_out2, int64_t* d_count, void* stream);
OGJK_API int ogjk_pairs_all_f64(ogjk_ctx* ctx, const double* d_boxes, int64_t npoly, double margin, int64_t capacity,
                                int* d_out1, int* d_out2, int64_t* d_count, void* stream);

/* Intersect flag of every pair from its distance (device pointers): flag = !(distance > eps),
 * eps = FLT_EPSILON / DBL_EPSILON -- the reference's own rule for "separated" (the EPA gate,
 * gpu/opengjk_gpu.cu:2053; the reference has no flag field, openGJK.h:104-105 only documents
 * that intersecting or touching bodies return 0).  1 = touching or intersecting. */
OGJK_API int ogjk_intersect_flags_f32(ogjk_ctx* ctx, int64_t n, const float* d_distances, unsigned char* d_flags, void* stream);
OGJK_API int ogjk_intersect_flags_f64(ogjk_ctx* ctx, int64_t n, const double* d_distances, unsigned char* d_flags, void* stream);

/* Number of chunks (0..16, default 8) the un-indexed host-buffer entry points cut a
 * batch into so that H2D, kernels and D2H overlap; 0 or 1 = one shot. */
OGJK_API int ogjk_ctx_set_pipeline_chunks(ogjk_ctx* ctx, int chunks);

/* ---- host-buffer batch over flat pools (what a zero-copy binding calls; the
 * e2e number in bench.py goes through this).  All pointers are HOST pointers.
 * Copies inputs to the device, runs, copies results back, synchronises.
 * pool2 may alias pool1 (indexed mode: pass the same pointers). */
OGJK_API int ogjk_gjk_host_f32(ogjk_ctx* ctx, int64_t npairs,
                               const float* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                               const float* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                               ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_gjk_host_f64(ogjk_ctx* ctx, int64_t npairs,
                               const double* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                               const double* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                               ogjk_simplex_f64* simplices, double* distances);

/* ---- reference-shaped high-level API (arrays of gkPolytope structs in host
 * memory with flattened `coord`), replacing
 *   compute_minimum_distance          gpu/include/openGJK_GPU.h:153-159, gpu/opengjk_gpu.cu:2977-3005
 *   compute_minimum_distance_indexed  openGJK_GPU.h:213-220,             opengjk_gpu.cu:3284-3350
 * n <= 0 returns OGJK_OK without touching anything (reference: silent return, :2984). */
OGJK_API int ogjk_compute_minimum_distance_f32(ogjk_ctx* ctx, int n, const ogjk_polytope_f32* bd1, const ogjk_polytope_f32* bd2,
                                               ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_compute_minimum_distance_f64(ogjk_ctx* ctx, int n, const ogjk_polytope_f64* bd1, const ogjk_polytope_f64* bd2,
                                               ogjk_simplex_f64* simplices, double* distances);
OGJK_API int ogjk_compute_minimum_distance_indexed_f32(ogjk_ctx* ctx, int num_polytopes, int num_pairs, const ogjk_polytope_f32* polytopes,
                                                       const ogjk_pair* pairs, ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_compute_minimum_distance_indexed_f64(ogjk_ctx* ctx, int num_polytopes, int num_pairs, const ogjk_polytope_f64* polytopes,
                                                       const ogjk_pair* pairs, ogjk_simplex_f64* simplices, double* distances);

/* ---- single query with the scalar library's data layout (coord is an array
 * of row pointers), replacing compute_minimum_distance(gkPolytope, gkPolytope,
 * gkSimplex*) -- scalar/include/openGJK/openGJK.h:110, scalar/openGJK.c:941-1046.
 * Runs as a batch of one on the context's device. */
OGJK_API int ogjk_single_rows_f32(ogjk_ctx* ctx, int n1, float* const* rows1, int n2, float* const* rows2,
                                  ogjk_simplex_f32* s, float* distance);
OGJK_API int ogjk_single_rows_f64(ogjk_ctx* ctx, int n1, double* const* rows1, int n2, double* const* rows2,
                                  ogjk_simplex_f64* s, double* distance);

/* ---- EPA: penetration depth and witness points for intersecting pairs.
 * Replaces compute_epa_kernel / compute_epa_device / compute_epa / compute_gjk_epa
 * (gpu/opengjk_gpu.cu:2022-2971, 3204-3224, 3007-3053, 3055-3099;
 * gpu/include/openGJK_GPU.h:161-206, 244-264).  Semantics as the reference: pairs
 * whose GJK distance exceeds gkEpsilon keep their GJK witnesses; the others get
 * distance = -depth, witness points on the closest polytope face, and the face
 * normal.  `simplices` are the GJK results (not modified by EPA); `normals` may
 * be NULL.  Exits on which the reference writes nothing keep the GJK distance and
 * witnesses. */
OGJK_API int ogjk_store_epa_f32(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                int64_t npairs, const ogjk_simplex_f32* simplices, float* distances, float* witness1,
                                float* witness2, float* normals, void* stream);
OGJK_API int ogjk_store_epa_f64(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                int64_t npairs, const ogjk_simplex_f64* simplices, double* distances, double* witness1,
                                double* witness2, double* normals, void* stream);
OGJK_API int ogjk_store_gjk_epa_f32(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                    int64_t npairs, ogjk_simplex_f32* simplices, float* distances, float* witness1,
                                    float* witness2, float* normals, void* stream);
OGJK_API int ogjk_store_gjk_epa_f64(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,
                                    int64_t npairs, ogjk_simplex_f64* simplices, double* distances, double* witness1,
                                    double* witness2, double* normals, void* stream);
/* host buffers, flat pools (as ogjk_gjk_host_*) */
OGJK_API int ogjk_gjk_epa_host_f32(ogjk_ctx* ctx, int64_t npairs,
                                   const float* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                                   const float* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                                   ogjk_simplex_f32* simplices, float* distances, float* witness1, float* witness2, float* normals);
OGJK_API int ogjk_gjk_epa_host_f64(ogjk_ctx* ctx, int64_t npairs,
                                   const double* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1,
                                   const double* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2,
                                   ogjk_simplex_f64* simplices, double* distances, double* witness1, double* witness2, double* normals);
/* reference-shaped struct arrays: compute_gjk_epa (openGJK_GPU.h:198-206) and
 * compute_epa (openGJK_GPU.h:175-184; simplices/distances are inputs and outputs) */
OGJK_API int ogjk_compute_gjk_epa_f32(ogjk_ctx* ctx, int n, const ogjk_polytope_f32* bd1, const ogjk_polytope_f32* bd2,
                                      ogjk_simplex_f32* simplices, float* distances, float* witness1, float* witness2);
OGJK_API int ogjk_compute_gjk_epa_f64(ogjk_ctx* ctx, int n, const ogjk_polytope_f64* bd1, const ogjk_polytope_f64* bd2,
                                      ogjk_simplex_f64* simplices, double* distances, double* witness1, double* witness2);
OGJK_API int ogjk_compute_epa_f32(ogjk_ctx* ctx, int n, const ogjk_polytope_f32* bd1, const ogjk_polytope_f32* bd2,
                                  ogjk_simplex_f32* simplices, float* distances, float* witness1, float* witness2,
                                  float* contact_normals);
OGJK_API int ogjk_compute_epa_f64(ogjk_ctx* ctx, int n, const ogjk_polytope_f64* bd1, const ogjk_polytope_f64* bd2,
                                  ogjk_simplex_f64* simplices, double* distances, double* witness1, double* witness2,
                                  double* contact_normals);

/* ---- device arrays of gkPolytope structs whose `coord` members are DEVICE pointers:
 * the data shape of the reference's mid-level API.  Replaces
 *   compute_minimum_distance_device  gpu/include/openGJK_GPU.h:236-242, gpu/opengjk_gpu.cu:3186-3202
 *   compute_epa_device               openGJK_GPU.h:255-264,             opengjk_gpu.cu:3204-3224
 * The polytopes are repacked into the context's scratch stores on every call. */
OGJK_API int ogjk_gjk_device_structs_f32(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f32* d_bd1, const ogjk_polytope_f32* d_bd2,
                                         ogjk_simplex_f32* d_simplices, float* d_distances, void* stream);
OGJK_API int ogjk_gjk_device_structs_f64(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f64* d_bd1, const ogjk_polytope_f64* d_bd2,
                                         ogjk_simplex_f64* d_simplices, double* d_distances, void* stream);
OGJK_API int ogjk_epa_device_structs_f32(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f32* d_bd1, const ogjk_polytope_f32* d_bd2,
                                         ogjk_simplex_f32* d_simplices, float* d_distances, float* d_witness1,
                                         float* d_witness2, float* d_normals, void* stream);
OGJK_API int ogjk_epa_device_structs_f64(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f64* d_bd1, const ogjk_polytope_f64* d_bd2,
                                         ogjk_simplex_f64* d_simplices, double* d_distances, double* d_witness1,
                                         double* d_witness2, double* d_normals, void* stream);

/* ---- several GPUs of one box behind one call (north_star: contiguous pair slice per GPU,
 * no collective on the hot path, NCCL only for the optional final gather).  The multi-device
 * form of the one call a reference user makes, compute_minimum_distance(n, bd1, bd2, simplices,
 * distances) (gpu/opengjk_gpu.cu:2977-3004, gpu/include/openGJK_GPU.h:153-159).
 *
 * ogjk_multi_create: devices[ndev] are CUDA ordinals (devices = NULL: 0..ndev-1; ndev = 0 with
 *   NULL: every device of the box).  One context and one persistent host thread per device.
 * ogjk_multi_shard_bounds: THE partition -- device slot i of ndev owns pairs [lo, hi); sizes
 *   differ by at most one pair.  A pure function (callable without a GPU).
 * ogjk_multi_gjk_host_* / ogjk_multi_gjk_epa_host_*: the arguments of ogjk_gjk_host_* /
 *   ogjk_gjk_epa_host_*; every device computes its slice and copies its results straight into
 *   [lo, hi) of the caller's arrays.  Un-indexed batches send a device only the coordinate
 *   range its slice touches; indexed batches replicate the pool on every device.
 * ogjk_multi_store_*: a pool replicated on every device (via_nccl = 0: each device uploads the
 *   host arrays over its own PCIe link; 1: one upload + ncclBroadcast over NVLink), queried
 *   with HOST index lists; _gjk / _gjk_epa return results in host arrays, _gjk_resident leaves
 *   them on the devices: every device holds full-size result arrays (ogjk_multi_resident_*),
 *   of which it has filled [lo, hi); ogjk_multi_allgather_* completes them on every device
 *   (grouped ncclBroadcast per slice; NCCL is loaded with dlopen -- libnccl.so.2 or
 *   $OGJK_NCCL_LIB -- and only these two paths need it).
 * Not thread safe per multi context.  ogjk_multi_ctx exposes the per-device contexts for the
 * tuning knobs. */
typedef struct ogjk_multi_ ogjk_multi;
typedef struct ogjk_multi_store_ ogjk_multi_store;
OGJK_API int ogjk_multi_create(const int* devices, int ndev, ogjk_multi** out);
OGJK_API void ogjk_multi_destroy(ogjk_multi* m);
OGJK_API int ogjk_multi_num_devices(const ogjk_multi* m);
OGJK_API ogjk_ctx* ogjk_multi_ctx(ogjk_multi* m, int i);
OGJK_API int64_t ogjk_multi_launch_count(const ogjk_multi* m);
OGJK_API void ogjk_multi_shard_bounds(int64_t npairs, int ndev, int i, int64_t* lo, int64_t* hi);
OGJK_API void ogjk_multi_store_destroy(ogjk_multi_store* store);
OGJK_API int64_t ogjk_multi_store_bytes(const ogjk_multi_store* store);
/* The reference's own call shape over every device of the multi context (replaces
 * compute_minimum_distance / compute_gjk_epa, gpu/include/openGJK_GPU.h:153-159, :198-206): arrays of
 * gkPolytope structs that point at host vertex arrays; device i computes pairs [lo, hi) of
 * ogjk_multi_shard_bounds and writes that range of the caller's arrays. */
OGJK_API int ogjk_multi_compute_minimum_distance_f32(ogjk_multi* m, int n, const ogjk_polytope_f32* bd1,
        const ogjk_polytope_f32* bd2, ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_multi_compute_minimum_distance_f64(ogjk_multi* m, int n, const ogjk_polytope_f64* bd1,
        const ogjk_polytope_f64* bd2, ogjk_simplex_f64* simplices, double* distances);
OGJK_API int ogjk_multi_compute_gjk_epa_f32(ogjk_multi* m, int n, const ogjk_polytope_f32* bd1, const ogjk_polytope_f32* bd2,
        ogjk_simplex_f32* simplices, float* distances, float* witness1, float* witness2);
OGJK_API int ogjk_multi_compute_gjk_epa_f64(ogjk_multi* m, int n, const ogjk_polytope_f64* bd1, const ogjk_polytope_f64* bd2,
        ogjk_simplex_f64* simplices, double* distances, double* witness1, double* witness2);
OGJK_API int ogjk_multi_gjk_host_f32(ogjk_multi* m, int64_t npairs, const float* coords1, const int64_t* off1, const
        int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1, const float* coords2, const int64_t* off2, const
        int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_multi_gjk_epa_host_f32(ogjk_multi* m, int64_t npairs, const float* coords1, const int64_t* off1,
        const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1, const float* coords2, const int64_t* off2,
        const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f32* simplices, float*
        distances, float* witness1, float* witness2, float* normals);
OGJK_API int ogjk_multi_store_create_f32(ogjk_multi* m, int64_t npoly, int64_t nverts, const float* coords, const
        int64_t* off, const int* cnt, int via_nccl, ogjk_multi_store** out);
OGJK_API int ogjk_multi_store_gjk_f32(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs, ogjk_simplex_f32* simplices, float* distances);
OGJK_API int ogjk_multi_store_gjk_epa_f32(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs, ogjk_simplex_f32* simplices, float* distances, float*
        witness1, float* witness2, float* normals);
OGJK_API int ogjk_multi_store_gjk_resident_f32(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs);
OGJK_API int ogjk_multi_allgather_f32(ogjk_multi* m);
OGJK_API int ogjk_multi_resident_f32(ogjk_multi* m, int i, float** d_distances, ogjk_simplex_f32** d_simplices,
        int64_t* lo, int64_t* hi);
OGJK_API int ogjk_multi_gjk_host_f64(ogjk_multi* m, int64_t npairs, const double* coords1, const int64_t* off1, const
        int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1, const double* coords2, const int64_t* off2,
        const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f64* simplices, double*
        distances);
OGJK_API int ogjk_multi_gjk_epa_host_f64(ogjk_multi* m, int64_t npairs, const double* coords1, const int64_t* off1,
        const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1, const double* coords2, const int64_t*
        off2, const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f64* simplices, double*
        distances, double* witness1, double* witness2, double* normals);
OGJK_API int ogjk_multi_store_create_f64(ogjk_multi* m, int64_t npoly, int64_t nverts, const double* coords, const
        int64_t* off, const int* cnt, int via_nccl, ogjk_multi_store** out);
OGJK_API int ogjk_multi_store_gjk_f64(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs, ogjk_simplex_f64* simplices, double* distances);
OGJK_API int ogjk_multi_store_gjk_epa_f64(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs, ogjk_simplex_f64* simplices, double* distances,
        double* witness1, double* witness2, double* normals);
OGJK_API int ogjk_multi_store_gjk_resident_f64(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const
        ogjk_multi_store* s2, const int* idx2, int64_t npairs);
OGJK_API int ogjk_multi_allgather_f64(ogjk_multi* m);
OGJK_API int ogjk_multi_resident_f64(ogjk_multi* m, int i, double** d_distances, ogjk_simplex_f64** d_simplices,
        int64_t* lo, int64_t* hi);

/* ---- sub-solver test hook, replacing opengjk_test_S1D/S2D/S3D (scalar/include/openGJK/openGJK.h:120-122,
 * scalar/openGJK.c:1232-1248): runs the device Signed-Volumes sub-solver once on `simplex`
 * (nvrtx 2..4, newest vertex in the highest slot), in place; v[3] receives the closest point. */
OGJK_API int ogjk_test_subalgorithm_f32(ogjk_ctx* ctx, ogjk_simplex_f32* simplex, float* v);
OGJK_API int ogjk_test_subalgorithm_f64(ogjk_ctx* ctx, ogjk_simplex_f64* simplex, double* v);

#ifdef __cplusplus
}
#endif
#endif /* OPENGJK_B200_H_ */
