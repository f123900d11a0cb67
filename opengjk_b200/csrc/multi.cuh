// multi.cuh -- one call, one batch, several GPUs of one box (north_star: "the batch shards
// naturally by pair index across the 8 GPUs of one box, with each GPU owning a contiguous
// slice and no collective on the hot path; NCCL over NVLink is used only for an optional
// final gather of results").  Replaces, at the multi-device level, the single call a
// caller of the reference makes: compute_minimum_distance(n, bd1, bd2, simplices, distances)
// (gpu/opengjk_gpu.cu:2977-3004, gpu/include/openGJK_GPU.h:153-159).
//
// Included by abi.cu inside its anonymous namespace's translation unit (it reuses the
// single-device pipelines).  Design:
//   * one ogjk_ctx and one persistent host thread per device; a call posts one job per
//     device and waits for all of them (no thread creation on the call path);
//   * pairs [lo, hi) of device i come from ogjk_multi_shard_bounds -- the library, not the
//     caller, owns the partition;
//   * un-indexed batches (pair p = polytope p of each pool) send every device only the
//     coordinate range its slice touches; indexed batches replicate the pool (per-device
//     H2D over each GPU's own PCIe link, or one upload + ncclBroadcast over NVLink);
//   * results are copied device->host straight into the caller's arrays at [lo, hi): disjoint
//     ranges, no gather step.  The optional device-side gather (every GPU ends up with all
//     results) is a grouped ncclBroadcast per slice, NCCL loaded with dlopen so that the
//     library has no link-time dependency on it.
#pragma once
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>

#include <condition_variable>
#include <functional>
#include <mutex>

namespace {

struct MultiWorker {
  int device = 0;
  ogjk_ctx* ctx = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = 0;
  char err[512] = "";
  // per-device scratch of the multi layer
  std::vector<int64_t> off1, off2;  // rebased offsets of the slice
  std::vector<int> ident1, ident2;  // identity index slices (mixed indexed / un-indexed calls)
  DevBuf d_idx1, d_idx2;            // store path: index slices on the device
  DevBuf r_dist, r_simp;            // resident results (full batch size; this device fills [lo, hi))
};

// ---- NCCL through dlopen (optional) ---------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("OGJK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.CommInitAll = reinterpret_cast<int (*)(void**, int, const int*)>(dlsym(api.lib, "ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupEnd"));
    api.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
        dlsym(api.lib, "ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
    api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast;
  });
  return api;
}
constexpr int kNcclChar = 0;  // ncclInt8 / ncclChar

}  // namespace

struct ogjk_multi_ {
  int ndev = 0;
  std::vector<MultiWorker*> w;
  std::vector<void*> comms;  // NCCL communicators, created on first use
  bool comms_tried = false;
  int64_t resident_pairs = 0;  // batch size of the last *_resident call
  int resident_prec = 0;
};

struct ogjk_multi_store_ {
  ogjk_multi_* m = nullptr;
  std::vector<ogjk_store*> s;  // one replica per device
};

namespace {

// Best effort: run the device's host thread on the CPUs of the NUMA node its PCIe root hangs off
// (its staging allocations and its cudaMemcpyAsync submissions then stay local to the link).
void bind_thread_to_device_node(int device) {
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return; }
  for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return;
  int node = -1;
  const int got = fscanf(f, "%d", &node);
  fclose(f);
  if (got != 1 || node < 0) return;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return;
  cpu_set_t want, have;
  CPU_ZERO(&want);
  int lo = 0, hi = 0;
  for (;;) {
    if (fscanf(f, "%d", &lo) != 1) break;
    hi = lo;
    int ch = fgetc(f);
    if (ch == '-') { if (fscanf(f, "%d", &hi) != 1) break; ch = fgetc(f); }
    for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) CPU_SET(c, &want);
    if (ch != ',') break;
  }
  fclose(f);
  if (pthread_getaffinity_np(pthread_self(), sizeof(have), &have) != 0) return;
  cpu_set_t both;
  CPU_AND(&both, &want, &have);
  if (CPU_COUNT(&both) > 0) pthread_setaffinity_np(pthread_self(), sizeof(both), &both);
}

void multi_worker_main(MultiWorker* w) {
  cudaSetDevice(w->device);
  if (!getenv("OGJK_NO_NUMA_BIND")) bind_thread_to_device_node(w->device);
  std::unique_lock<std::mutex> lk(w->mu);
  for (;;) {
    w->cv.wait(lk, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    std::function<int()> job = std::move(w->job);
    w->has_job = false;
    lk.unlock();
    g_err[0] = 0;
    const int rc = job();
    lk.lock();
    w->rc = rc;
    snprintf(w->err, sizeof(w->err), "%s", g_err);
    w->done = true;
    w->cv.notify_all();
  }
}

// Runs fn(i) on every device's thread concurrently; returns the first failure.
int multi_run(ogjk_multi_* m, const std::function<int(int)>& fn) {
  for (int i = 0; i < m->ndev; ++i) {
    MultiWorker* w = m->w[i];
    std::lock_guard<std::mutex> lk(w->mu);
    w->job = [i, &fn] { return fn(i); };
    w->has_job = true;
    w->done = false;
    w->cv.notify_all();
  }
  int rc = OGJK_OK;
  for (int i = 0; i < m->ndev; ++i) {
    MultiWorker* w = m->w[i];
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->rc && !rc) {
      rc = w->rc;
      snprintf(g_err, sizeof(g_err), "device %d: %.400s", w->device, w->err);
    }
  }
  return rc;
}

void shard_bounds(int64_t npairs, int ndev, int i, int64_t* lo, int64_t* hi) {
  if (ndev < 1) ndev = 1;
  if (npairs < 0) npairs = 0;
  // contiguous slices whose sizes differ by at most one pair (the first npairs % ndev get the extra one)
  const int64_t q = npairs / ndev, r = npairs % ndev;
  *lo = q * i + (i < r ? i : r);
  *hi = *lo + q + (i < r ? 1 : 0);
}

// Coordinate range [lo_v, hi_v) the polytopes [lo, hi) of a pool touch, and their offsets rebased to it.
// One pass for pools stored in pair order (the common case: the first polytope starts the range);
// a pool that is not gets a second pass.
int slice_pool(const int64_t* off, const int* cnt, int64_t lo, int64_t hi, int64_t nverts, std::vector<int64_t>& reb,
               int64_t* lo_v, int64_t* nv) {
  const int64_t n = hi - lo;
  reb.resize((size_t)n);
  const int64_t first = off[lo];
  int64_t a = first, b = 0;
  bool bad = false;
  for (int64_t p = 0; p < n; ++p) {
    const int64_t o = off[lo + p], e = o + cnt[lo + p];
    bad |= cnt[lo + p] < 1 || o < 0 || e > nverts;
    a = o < a ? o : a;
    b = e > b ? e : b;
    reb[(size_t)p] = o - first;
  }
  if (bad) return fail(OGJK_ERR_ARG, "polytope range outside the pool%s");
  if (a != first)
    for (int64_t p = 0; p < n; ++p) reb[(size_t)p] += first - a;
  *lo_v = a;
  *nv = b - a;
  return OGJK_OK;
}

template <typename T, typename S>
int multi_host(ogjk_multi_* m, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1,
               int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2,
               int64_t nverts2, const int* idx2, S* simplices, T* distances, T* w1, T* w2, T* normals, bool run_epa) {
  if (!m) return fail(OGJK_ERR_ARG, "multi context is NULL%s");
  if (npairs < 0) return fail(OGJK_ERR_ARG, "npairs < 0%s");
  if (npairs == 0) return OGJK_OK;
  if (!coords1 || !off1 || !cnt1 || !coords2 || !off2 || !cnt2 || !distances) return fail(OGJK_ERR_ARG, "NULL host pointer%s");
  if ((!idx1 && npairs > npoly1) || (!idx2 && npairs > npoly2))
    return fail(OGJK_ERR_ARG, "npairs exceeds the pool size in un-indexed mode%s");
  return multi_run(m, [&](int i) -> int {
    MultiWorker* w = m->w[i];
    int64_t lo, hi;
    shard_bounds(npairs, m->ndev, i, &lo, &hi);
    const int64_t n = hi - lo;
    if (n <= 0) return OGJK_OK;
    S* so = simplices ? simplices + lo : nullptr;
    T* wo1 = w1 ? w1 + 3 * lo : nullptr;
    T* wo2 = w2 ? w2 + 3 * lo : nullptr;
    T* no = normals ? normals + 3 * lo : nullptr;
    if (!idx1 && !idx2) {
      // un-indexed: this device sees only its slice of both pools
      int64_t a1, nv1, a2, nv2;
      int rc = slice_pool(off1, cnt1, lo, hi, nverts1, w->off1, &a1, &nv1);
      if (rc) return rc;
      if ((rc = slice_pool(off2, cnt2, lo, hi, nverts2, w->off2, &a2, &nv2))) return rc;
      return host_pipeline<T, S>(w->ctx, n, coords1 + 3 * a1, w->off1.data(), cnt1 + lo, n, nv1, nullptr, coords2 + 3 * a2,
                                 w->off2.data(), cnt2 + lo, n, nv2, nullptr, so, distances + lo, wo1, wo2, no, true, run_epa);
    }
    // indexed: the pool is replicated, the pair list is sliced; a NULL index list on one
    // side means "pair p uses polytope p" and becomes an explicit identity slice
    const int* i1 = idx1 ? idx1 + lo : nullptr;
    const int* i2 = idx2 ? idx2 + lo : nullptr;
    if (!i1) { w->ident1.resize((size_t)n); for (int64_t p = 0; p < n; ++p) w->ident1[(size_t)p] = (int)(lo + p); i1 = w->ident1.data(); }
    if (!i2) { w->ident2.resize((size_t)n); for (int64_t p = 0; p < n; ++p) w->ident2[(size_t)p] = (int)(lo + p); i2 = w->ident2.data(); }
    return host_pipeline<T, S>(w->ctx, n, coords1, off1, cnt1, npoly1, nverts1, i1, coords2, off2, cnt2, npoly2, nverts2, i2, so,
                               distances + lo, wo1, wo2, no, true, run_epa);
  });
}

int multi_comms(ogjk_multi_* m) {
  if (!m->comms.empty()) return OGJK_OK;
  NcclApi& n = nccl_api();
  if (!n.ok) return fail(OGJK_ERR_CUDA, "NCCL is not loadable (libnccl.so.2 / OGJK_NCCL_LIB); the device-side gather and broadcast need it%s");
  std::vector<int> devs(m->ndev);
  for (int i = 0; i < m->ndev; ++i) devs[i] = m->w[i]->device;
  m->comms.assign(m->ndev, nullptr);
  const int rc = n.CommInitAll(m->comms.data(), m->ndev, devs.data());
  if (rc) {
    m->comms.clear();
    return fail(OGJK_ERR_CUDA, "ncclCommInitAll failed: %s", n.GetErrorString ? n.GetErrorString(rc) : "?");
  }
  return OGJK_OK;
}

// Replicated persistent store.  mode 0: every device uploads the host arrays over its own
// PCIe link (parallel); mode 1: device 0 uploads, the raw pool is broadcast over NVLink.
template <typename T>
int multi_store_create(ogjk_multi_* m, int64_t npoly, int64_t nverts, const T* coords, const int64_t* off, const int* cnt,
                       int via_nccl, ogjk_multi_store_** out) {
  if (!m || !out) return fail(OGJK_ERR_ARG, "NULL argument%s");
  *out = nullptr;
  if (npoly < 1 || nverts < 1 || !coords || !off || !cnt) return fail(OGJK_ERR_ARG, "empty polytope pool%s");
  auto* ms = new (std::nothrow) ogjk_multi_store_();
  if (!ms) return fail(OGJK_ERR_ALLOC, "out of host memory%s");
  ms->m = m;
  ms->s.assign(m->ndev, nullptr);
  int rc = OGJK_OK;
  if (via_nccl && m->ndev > 1) {
    if ((rc = multi_comms(m))) { delete ms; return rc; }
    NcclApi& n = nccl_api();
    const size_t cb = (size_t)nverts * 3 * sizeof(T), ob = (size_t)npoly * sizeof(int64_t), kb = (size_t)npoly * sizeof(int);
    std::vector<DevBuf> dc(m->ndev), dof(m->ndev), dk(m->ndev);
    rc = multi_run(m, [&](int i) -> int {
      int r;
      if ((r = dc[i].reserve(cb)) || (r = dof[i].reserve(ob)) || (r = dk[i].reserve(kb))) return r;
      if (i == 0) {
        cudaStream_t st = m->w[0]->ctx->own_stream;
        CU(cudaMemcpyAsync(dc[0].p, coords, cb, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dof[0].p, off, ob, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dk[0].p, cnt, kb, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
      }
      return OGJK_OK;
    });
    if (!rc) {
      // one thread issues the grouped broadcasts for all devices (single-process NCCL)
      int e = n.GroupStart();
      for (int i = 0; i < m->ndev && !e; ++i) {
        cudaStream_t st = m->w[i]->ctx->own_stream;
        e = n.Broadcast(dc[0].p, dc[i].p, cb, kNcclChar, 0, m->comms[i], st);
        if (!e) e = n.Broadcast(dof[0].p, dof[i].p, ob, kNcclChar, 0, m->comms[i], st);
        if (!e) e = n.Broadcast(dk[0].p, dk[i].p, kb, kNcclChar, 0, m->comms[i], st);
      }
      const int e2 = n.GroupEnd();
      if (e || e2) rc = fail(OGJK_ERR_CUDA, "ncclBroadcast failed: %s", n.GetErrorString ? n.GetErrorString(e ? e : e2) : "?");
    }
    if (!rc)
      rc = multi_run(m, [&](int i) -> int {
        MultiWorker* w = m->w[i];
        return store_create<T>(w->ctx, npoly, nverts, static_cast<const T*>(dc[i].p), static_cast<const int64_t*>(dof[i].p),
                               static_cast<const int*>(dk[i].p), 1, w->ctx->own_stream, &ms->s[i]);
      });
    multi_run(m, [&](int i) -> int {
      cudaStreamSynchronize(m->w[i]->ctx->own_stream);
      dc[i].release(); dof[i].release(); dk[i].release();
      return OGJK_OK;
    });
  } else {
    rc = multi_run(m, [&](int i) -> int {
      MultiWorker* w = m->w[i];
      return store_create<T>(w->ctx, npoly, nverts, coords, off, cnt, 0, w->ctx->own_stream, &ms->s[i]);
    });
  }
  if (rc) {
    multi_run(m, [&](int i) -> int {
      if (ms->s[i]) ogjk_store_destroy(ms->s[i]);
      return OGJK_OK;
    });
    delete ms;
    return rc;
  }
  *out = ms;
  return OGJK_OK;
}

// One batch over replicated stores.  host_out: results go to the caller's host arrays (disjoint
// ranges); otherwise they stay in per-device buffers of the full batch size (resident mode).
template <typename T, typename S>
int multi_store_run(ogjk_multi_* m, const ogjk_multi_store_* s1, const int* idx1, const ogjk_multi_store_* s2, const int* idx2,
                    int64_t npairs, S* simplices, T* distances, T* w1, T* w2, T* normals, bool run_epa, bool host_out) {
  if (!m || !s1 || !s2) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (s1->m != m || s2->m != m) return fail(OGJK_ERR_ARG, "store belongs to another multi context%s");
  if (npairs < 0) return fail(OGJK_ERR_ARG, "npairs < 0%s");
  if (npairs == 0) return OGJK_OK;
  if (!idx1 || !idx2) return fail(OGJK_ERR_ARG, "replicated stores are queried with explicit (host) index lists%s");
  if (host_out && !distances) return fail(OGJK_ERR_ARG, "NULL output array%s");
  if (run_epa && (!host_out || !w1 || !w2 || !simplices)) return fail(OGJK_ERR_ARG, "EPA needs host simplex and witness arrays%s");
  if (!host_out) { m->resident_pairs = npairs; m->resident_prec = (int)sizeof(T); }
  return multi_run(m, [&](int i) -> int {
    MultiWorker* w = m->w[i];
    ogjk_ctx* ctx = w->ctx;
    int64_t lo, hi;
    shard_bounds(npairs, m->ndev, i, &lo, &hi);
    const int64_t n = hi - lo;
    cudaStream_t st = ctx->own_stream;
    int rc;
    if (!host_out) {
      if ((rc = w->r_dist.reserve((size_t)npairs * sizeof(T)))) return rc;
      if ((rc = w->r_simp.reserve((size_t)npairs * sizeof(S)))) return rc;
    }
    if (n <= 0) return OGJK_OK;
    if ((rc = w->d_idx1.reserve((size_t)n * sizeof(int)))) return rc;
    if ((rc = w->d_idx2.reserve((size_t)n * sizeof(int)))) return rc;
    CU(cudaMemcpyAsync(w->d_idx1.p, idx1 + lo, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w->d_idx2.p, idx2 + lo, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    S* ds; T* dd;
    if (host_out) {
      if ((rc = ctx->d_dist.reserve((size_t)n * sizeof(T)))) return rc;
      if ((rc = ctx->d_simp.reserve((size_t)n * sizeof(S)))) return rc;
      dd = static_cast<T*>(ctx->d_dist.p);
      ds = static_cast<S*>(ctx->d_simp.p);
    } else {
      dd = static_cast<T*>(w->r_dist.p) + lo;
      ds = static_cast<S*>(w->r_simp.p) + lo;
    }
    const int* di1 = static_cast<const int*>(w->d_idx1.p);
    const int* di2 = static_cast<const int*>(w->d_idx2.p);
    if ((rc = store_gjk<T, S>(ctx, s1->s[i], di1, s2->s[i], di2, n, ds, dd, st))) return rc;
    if (run_epa) {
      if ((rc = ctx->d_w1.reserve((size_t)n * 3 * sizeof(T)))) return rc;
      if ((rc = ctx->d_w2.reserve((size_t)n * 3 * sizeof(T)))) return rc;
      if (normals && (rc = ctx->d_nrm.reserve((size_t)n * 3 * sizeof(T)))) return rc;
      if ((rc = store_epa<T, S>(ctx, s1->s[i], di1, s2->s[i], di2, n, ds, dd, static_cast<T*>(ctx->d_w1.p),
                                static_cast<T*>(ctx->d_w2.p), normals ? static_cast<T*>(ctx->d_nrm.p) : nullptr, st)))
        return rc;
      CU(cudaMemcpyAsync(w1 + 3 * lo, ctx->d_w1.p, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(w2 + 3 * lo, ctx->d_w2.p, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
      if (normals) CU(cudaMemcpyAsync(normals + 3 * lo, ctx->d_nrm.p, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    if (host_out) {
      CU(cudaMemcpyAsync(distances + lo, dd, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
      if (simplices) CU(cudaMemcpyAsync(simplices + lo, ds, (size_t)n * sizeof(S), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return OGJK_OK;
  });
}

// Device-side gather after a resident run: every device ends up with all results.
template <typename T, typename S>
int multi_allgather(ogjk_multi_* m) {
  if (!m) return fail(OGJK_ERR_ARG, "multi context is NULL%s");
  if (m->resident_pairs <= 0 || m->resident_prec != (int)sizeof(T)) return fail(OGJK_ERR_ARG, "no resident results of this precision%s");
  if (m->ndev == 1) return OGJK_OK;
  int rc = multi_comms(m);
  if (rc) return rc;
  NcclApi& n = nccl_api();
  int e = n.GroupStart();
  for (int root = 0; root < m->ndev && !e; ++root) {
    int64_t lo, hi;
    shard_bounds(m->resident_pairs, m->ndev, root, &lo, &hi);
    if (hi <= lo) continue;
    for (int i = 0; i < m->ndev && !e; ++i) {
      MultiWorker* w = m->w[i];
      T* dd = static_cast<T*>(w->r_dist.p) + lo;
      S* ds = static_cast<S*>(w->r_simp.p) + lo;
      e = n.Broadcast(dd, dd, (size_t)(hi - lo) * sizeof(T), kNcclChar, root, m->comms[i], w->ctx->own_stream);
      if (!e) e = n.Broadcast(ds, ds, (size_t)(hi - lo) * sizeof(S), kNcclChar, root, m->comms[i], w->ctx->own_stream);
    }
  }
  const int e2 = n.GroupEnd();
  if (e || e2) return fail(OGJK_ERR_CUDA, "ncclBroadcast failed: %s", n.GetErrorString ? n.GetErrorString(e ? e : e2) : "?");
  return multi_run(m, [&](int i) -> int {
    CU(cudaStreamSynchronize(m->w[i]->ctx->own_stream));
    return OGJK_OK;
  });
}

}  // namespace

// The reference's call shape (arrays of gkPolytope structs with host vertex arrays) over all devices: a
// struct array is per pair, so device i simply gets [lo, hi) of every argument and gathers its own slice
// with its context's host threads.
template <typename T, typename P, typename S>
int multi_structs(ogjk_multi_* m, int n, const P* bd1, const P* bd2, S* simplices, T* distances, T* w1, T* w2, T* normals,
                  bool epa) {
  if (!m) return fail(OGJK_ERR_ARG, "multi context is NULL%s");
  if (n <= 0) return OGJK_OK;  // reference: silent return (gpu/opengjk_gpu.cu:2984)
  if (!bd1 || !bd2 || !distances || (epa && (!simplices || !w1 || !w2))) return fail(OGJK_ERR_ARG, "NULL argument%s");
  return multi_run(m, [&](int i) -> int {
    int64_t lo, hi;
    shard_bounds(n, m->ndev, i, &lo, &hi);
    const int k = (int)(hi - lo);
    if (k <= 0) return OGJK_OK;
    S* so = simplices ? simplices + lo : nullptr;
    if (!epa) return compute_min_dist<T, P, S>(m->w[i]->ctx, k, bd1 + lo, bd2 + lo, so, distances + lo);
    return compute_epa_structs<T, P, S>(m->w[i]->ctx, k, bd1 + lo, bd2 + lo, so, distances + lo, w1 + 3 * lo, w2 + 3 * lo,
                                        normals ? normals + 3 * lo : nullptr, true);
  });
}

extern "C" {

void ogjk_multi_shard_bounds(int64_t npairs, int ndev, int i, int64_t* lo, int64_t* hi) {
  int64_t a = 0, b = 0;
  if (i >= 0 && i < (ndev < 1 ? 1 : ndev)) shard_bounds(npairs, ndev, i, &a, &b);
  if (lo) *lo = a;
  if (hi) *hi = b;
}

void ogjk_multi_destroy(ogjk_multi* m) {
  if (!m) return;
  if (!m->comms.empty()) {
    NcclApi& n = nccl_api();
    for (void* c : m->comms) if (c) n.CommDestroy(c);
  }
  for (MultiWorker* w : m->w) {
    if (!w) continue;
    if (w->th.joinable()) {
      { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
      w->th.join();
    }
    if (w->ctx) {  // a slot whose context was never created has nothing on a device (and maybe no valid ordinal)
      cudaSetDevice(w->device);
      w->d_idx1.release(); w->d_idx2.release(); w->r_dist.release(); w->r_simp.release();
      ogjk_ctx_destroy(w->ctx);
    }
    delete w;
  }
  delete m;
}

int ogjk_multi_create(const int* devices, int ndev, ogjk_multi** out) {
  if (!out) return fail(OGJK_ERR_ARG, "out is NULL%s");
  *out = nullptr;
  int have = 0;
  cudaError_t e = cudaGetDeviceCount(&have);
  if (e != cudaSuccess || have == 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device available (%s); this library has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return OGJK_ERR_CUDA;
  }
  if (ndev == 0 && !devices) ndev = have;  // all devices of the box
  if (ndev < 1 || ndev > 64) return fail(OGJK_ERR_ARG, "ndev must be 1..64 (or 0 with devices = NULL for all)%s");
  ogjk_multi* m = new (std::nothrow) ogjk_multi_();
  if (!m) return fail(OGJK_ERR_ALLOC, "out of host memory%s");
  m->ndev = ndev;
  m->w.assign(ndev, nullptr);
  for (int i = 0; i < ndev; ++i) {
    const int dev = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (m->w[j]->device == dev) { ogjk_multi_destroy(m); return fail(OGJK_ERR_ARG, "device listed twice%s"); }
    MultiWorker* w = new (std::nothrow) MultiWorker();
    if (!w) { ogjk_multi_destroy(m); return fail(OGJK_ERR_ALLOC, "out of host memory%s"); }
    m->w[i] = w;
    w->device = dev;
    const int rc = ogjk_ctx_create(dev, &w->ctx);
    if (rc) { ogjk_multi_destroy(m); return rc; }
    // the devices' gather / copy threads run at the same time: share the host's hardware threads out
    const unsigned hc = std::thread::hardware_concurrency();
    w->ctx->host_threads = (int)((hc ? hc : 16u) / (unsigned)ndev < 4u ? 4u : (hc ? hc : 16u) / (unsigned)ndev);
    w->th = std::thread(multi_worker_main, w);
  }
  *out = m;
  return OGJK_OK;
}

int ogjk_multi_num_devices(const ogjk_multi* m) { return m ? m->ndev : 0; }
ogjk_ctx* ogjk_multi_ctx(ogjk_multi* m, int i) { return (m && i >= 0 && i < m->ndev) ? m->w[i]->ctx : nullptr; }
int64_t ogjk_multi_launch_count(const ogjk_multi* m) {
  int64_t t = 0;
  if (m) for (MultiWorker* w : m->w) t += w->ctx->launches;
  return t;
}

#define OGJK_MULTI_IMPL(SFX, T, S)                                                                                        \
  int ogjk_multi_gjk_host_##SFX(ogjk_multi* m, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1,     \
                                int64_t npoly1, int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2,   \
                                const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, S* simplices,          \
                                T* distances) {                                                                          \
    return multi_host<T, S>(m, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2, npoly2, nverts2,   \
                            idx2, simplices, distances, nullptr, nullptr, nullptr, false);                                \
  }                                                                                                                      \
  int ogjk_multi_gjk_epa_host_##SFX(ogjk_multi* m, int64_t npairs, const T* coords1, const int64_t* off1,                  \
                                    const int* cnt1, int64_t npoly1, int64_t nverts1, const int* idx1, const T* coords2,   \
                                    const int64_t* off2, const int* cnt2, int64_t npoly2, int64_t nverts2,                \
                                    const int* idx2, S* simplices, T* distances, T* witness1, T* witness2, T* normals) {   \
    if (!witness1 || !witness2) return fail(OGJK_ERR_ARG, "EPA needs witness output arrays%s");                          \
    return multi_host<T, S>(m, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2, npoly2, nverts2,   \
                            idx2, simplices, distances, witness1, witness2, normals, true);                               \
  }                                                                                                                      \
  int ogjk_multi_compute_minimum_distance_##SFX(ogjk_multi* m, int n, const ogjk_polytope_##SFX* bd1,                     \
                                                const ogjk_polytope_##SFX* bd2, S* simplices, T* distances) {            \
    return multi_structs<T, ogjk_polytope_##SFX, S>(m, n, bd1, bd2, simplices, distances, nullptr, nullptr, nullptr,      \
                                                    false);                                                              \
  }                                                                                                                      \
  int ogjk_multi_compute_gjk_epa_##SFX(ogjk_multi* m, int n, const ogjk_polytope_##SFX* bd1,                              \
                                       const ogjk_polytope_##SFX* bd2, S* simplices, T* distances, T* witness1,          \
                                       T* witness2) {                                                                    \
    return multi_structs<T, ogjk_polytope_##SFX, S>(m, n, bd1, bd2, simplices, distances, witness1, witness2, nullptr,    \
                                                    true);                                                               \
  }                                                                                                                      \
  int ogjk_multi_store_create_##SFX(ogjk_multi* m, int64_t npoly, int64_t nverts, const T* coords, const int64_t* off,     \
                                    const int* cnt, int via_nccl, ogjk_multi_store** out) {                               \
    return multi_store_create<T>(m, npoly, nverts, coords, off, cnt, via_nccl, out);                                     \
  }                                                                                                                      \
  int ogjk_multi_store_gjk_##SFX(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1, const ogjk_multi_store* s2,   \
                                 const int* idx2, int64_t npairs, S* simplices, T* distances) {                           \
    return multi_store_run<T, S>(m, s1, idx1, s2, idx2, npairs, simplices, distances, nullptr, nullptr, nullptr, false,   \
                                 true);                                                                                  \
  }                                                                                                                      \
  int ogjk_multi_store_gjk_epa_##SFX(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1,                           \
                                     const ogjk_multi_store* s2, const int* idx2, int64_t npairs, S* simplices,           \
                                     T* distances, T* witness1, T* witness2, T* normals) {                                \
    return multi_store_run<T, S>(m, s1, idx1, s2, idx2, npairs, simplices, distances, witness1, witness2, normals, true,  \
                                 true);                                                                                  \
  }                                                                                                                      \
  int ogjk_multi_store_gjk_resident_##SFX(ogjk_multi* m, const ogjk_multi_store* s1, const int* idx1,                      \
                                          const ogjk_multi_store* s2, const int* idx2, int64_t npairs) {                  \
    return multi_store_run<T, S>(m, s1, idx1, s2, idx2, npairs, nullptr, nullptr, nullptr, nullptr, nullptr, false,       \
                                 false);                                                                                 \
  }                                                                                                                      \
  int ogjk_multi_allgather_##SFX(ogjk_multi* m) { return multi_allgather<T, S>(m); }                                       \
  int ogjk_multi_resident_##SFX(ogjk_multi* m, int i, T** d_distances, S** d_simplices, int64_t* lo, int64_t* hi) {        \
    if (!m || i < 0 || i >= m->ndev) return fail(OGJK_ERR_ARG, "bad device slot%s");                                      \
    if (m->resident_pairs <= 0 || m->resident_prec != (int)sizeof(T))                                                    \
      return fail(OGJK_ERR_ARG, "no resident results of this precision%s");                                              \
    if (d_distances) *d_distances = static_cast<T*>(m->w[i]->r_dist.p);                                                   \
    if (d_simplices) *d_simplices = static_cast<S*>(m->w[i]->r_simp.p);                                                   \
    int64_t a, b;                                                                                                        \
    shard_bounds(m->resident_pairs, m->ndev, i, &a, &b);                                                                 \
    if (lo) *lo = a;                                                                                                     \
    if (hi) *hi = b;                                                                                                     \
    return OGJK_OK;                                                                                                      \
  }
OGJK_MULTI_IMPL(f32, float, ogjk_simplex_f32)
OGJK_MULTI_IMPL(f64, double, ogjk_simplex_f64)

void ogjk_multi_store_destroy(ogjk_multi_store* ms) {
  if (!ms) return;
  for (ogjk_store* s : ms->s) if (s) ogjk_store_destroy(s);
  delete ms;
}
int64_t ogjk_multi_store_bytes(const ogjk_multi_store* ms) {
  int64_t t = 0;
  if (ms) for (ogjk_store* s : ms->s) t += ogjk_store_bytes(s);
  return t;
}

}  // extern "C"
