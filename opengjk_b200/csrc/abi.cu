// abi.cu -- the C ABI declared in include/opengjk_b200.h.
//
// Host logic only: context + scratch memory, the two-stage batch scheduler
// (classify by size, then one launch per kernel flavour), the host-buffer
// convenience entry points (one bulk H2D per array instead of the reference's
// one cudaMemcpy per polytope, gpu/opengjk_gpu.cu:3152-3153) and the
// reference-shaped struct-array wrappers.  No torch types, no C++ types across
// the boundary, every CUDA call checked.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/opengjk_b200.h"
#include "gjk_kernels.cuh"
#include "table_kernel.cuh"
#ifdef OGJK_EXPERIMENTAL
#include "gjk_experimental.cuh"
#include "regroup_kernel.cuh"
#include "tma_kernel.cuh"
#endif
#include "epa_kernels.cuh"
#include "broadphase.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      snprintf(g_err, sizeof(g_err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
      return OGJK_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return OGJK_OK;
    if (p) { CU(cudaFree(p)); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    CU(cudaMalloc(&p, want));
    cap = want;
    return OGJK_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return OGJK_OK;
    if (p) { CU(cudaFreeHost(p)); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    CU(cudaMallocHost(&p, want));
    cap = want;
    return OGJK_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// Persistent host threads of a context (created on first use, joined with the context): the
// struct-array entry points gather the callers' per-polytope vertex arrays with them while
// earlier chunks are already on their way to the device.  Jobs [0, njobs) are handed out in
// order through an atomic counter; every job raises its own flag when done, so the caller can
// wait for a prefix of the job list (one chunk) while later jobs are still running.
class HostPool {
 public:
  explicit HostPool(int nthreads) {
    for (int t = 0; t < nthreads; ++t) th_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(mu_); quit_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  // Starts njobs calls fn(job) on the pool; returns at once.  One batch at a time.
  void start(int njobs, std::function<void(int)> fn) {
    wait_all();
    auto b = std::make_shared<Batch>();
    b->fn = std::move(fn);
    b->njobs = njobs;
    b->done.reset(new std::atomic<char>[(size_t)(njobs > 0 ? njobs : 1)]);
    for (int j = 0; j < njobs; ++j) b->done[(size_t)j].store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      cur_ = b;
      ++epoch_;
    }
    cv_.notify_all();
  }
  // Blocks until jobs [0, upto) of the current batch are done.
  void wait_prefix(int upto) {
    std::shared_ptr<Batch> b;
    { std::lock_guard<std::mutex> lk(mu_); b = cur_; }
    if (!b) return;
    for (int j = 0; j < upto && j < b->njobs; ++j)
      while (!b->done[(size_t)j].load(std::memory_order_acquire)) std::this_thread::yield();
  }
  void wait_all() {
    std::shared_ptr<Batch> b;
    { std::lock_guard<std::mutex> lk(mu_); b = cur_; }
    if (!b) return;
    while (b->finished.load(std::memory_order_acquire) < b->njobs) std::this_thread::yield();
  }
  void run(int njobs, std::function<void(int)> fn) { start(njobs, std::move(fn)); wait_all(); }

 private:
  // Every batch owns its counters and flags, so a thread that is late leaving batch k can never
  // touch batch k+1's state.
  struct Batch {
    std::function<void(int)> fn;
    int njobs = 0;
    std::atomic<int> next{0}, finished{0};
    std::unique_ptr<std::atomic<char>[]> done;
  };
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      std::shared_ptr<Batch> b;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return quit_ || epoch_ != seen; });
        if (quit_) return;
        seen = epoch_;
        b = cur_;
      }
      for (;;) {
        const int j = b->next.fetch_add(1);
        if (j >= b->njobs) break;
        b->fn(j);
        b->done[(size_t)j].store(1, std::memory_order_release);
        b->finished.fetch_add(1, std::memory_order_release);
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::shared_ptr<Batch> cur_;
  unsigned long long epoch_ = 0;
  bool quit_ = false;
};

// "The host-side data of polytopes [0, upto) is complete" -- blocks until it is.
using HostReady = std::function<void(int64_t upto)>;

constexpr int kPhases = 8;   // timing slots: pack, sort, then up to six size classes
constexpr int kMaxClasses = 6;

struct ogjk_store_ {
  int device = 0;
  int prec = 0;  // 4 = fp32, 8 = fp64
  long long npoly = 0;
  long long nverts = 0;       // true vertices
  long long capacity = 0;     // elements allocated in `data`
  DevBuf data, boff, cnt, scan_tmp;
  DevBuf accel, aoff;            // cluster tables of the large polytopes (store.cuh); empty when none
  int accel_min = 0x7fffffff;    // smallest vertex count that has a table
  long long accel_elems = 0;
  void release_all() {
    data.release(); boff.release(); cnt.release(); scan_tmp.release(); accel.release(); aoff.release();
  }
};

// Batches at least this long send their one-tile table pairs to gjk_tablebatch_kernel: with fewer
// pairs than about 30 per resident warp the warp-per-pair kernel finishes sooner (measured
// cross-over on B200 between 60 k and 200 k pairs, profiles/r2p_table_batch_ab.jsonl).
static constexpr long long kBatchKernelMinPairs = 131072;

struct ogjk_ctx_ {
  int device = 0;
  int sm_count = 148;
  int warp_threshold = 192;  // raw modes only
  long long launches = 0;
  // scheduler classes: pairs with sort key <= hi[c] (and > hi[c-1]) run with lanes[c]
  // cooperating lanes per pair; lanes 0 = register-resident thread-per-pair kernel.
  // defaults from the B200 sweep (tools/tune_classes.py, profiles/r1_lanes_sweep.txt)
  int nclasses = 4;
  int class_hi[kMaxClasses] = {4, 32, 128, 2048, 2048, 2048};
  int class_lanes[kMaxClasses] = {0, 2, 8, 32, 32, 32};
  int accel_min = 512;  // persistent stores: cluster tables for polytopes with >= this many vertices (0 = none)
  // multi-pass scheduling of the small-group classes (lanes 1..8): pairs still running after
  // pass_iters[0] iterations are parked and resumed densely packed (gjk_pass_kernel)
  int passes = 1;
  int pass_iters[2] = {4, 3};
  DevBuf pass_list[2], pass_f[2], pass_i[2];
  int regroup[2] = {0, 0};       // > 0: the 2-lane classes regroup in the CTA after that many iterations (tuning)
  int batch_blocks[2] = {0, 0};  // resident CTAs of gjk_tablebatch_kernel<float|double> on this device
  int table_lanes = 0;   // 0: tile kernel (table_kernel.cuh) above 1024 vertices, first-generation kernel below; 3 / 8,16,32: force one
  bool timing = false;
  cudaEvent_t ev[2 * kPhases] = {};
  int timed_mask = 0;
  cudaStream_t own_stream = nullptr;  // used by the host-buffer entry points
  cudaStream_t in_stream = nullptr, out_stream = nullptr;  // H2D / D2H legs of the chunked host pipeline
  cudaEvent_t chunk_in[16] = {}, chunk_done[16] = {}, chunk_out[16] = {};
  PinBuf h_out;                       // pinned landing zone for results whose destination is pageable memory
  int pipeline_chunks = 8;            // 0/1 disables chunking
  DevBuf counters;                    // kCounterSlots x u64: work-queue heads (the first kQueueSlots are reset per batch)
  DevBuf sort_tmp;                    // hist | start | cursor (u32 each, kMaxKey+2)
  DevBuf sorted_list;                 // pair ids, size-sorted
  DevBuf epa_list;                    // pair ids that need EPA (distance <= eps)
  DevBuf bp_tmp;                      // broad phase: block sums / row counts
  DevBuf d_w1, d_w2, d_nrm;           // EPA outputs of the host-buffer paths
  DevBuf small_list, large_list;      // raw modes
  ogjk_store_ scratch1, scratch2;     // raw device arrays are repacked into these
  // staging for the host-buffer paths
  DevBuf d_coords1, d_coords2, d_off1, d_off2, d_cnt1, d_cnt2, d_idx1, d_idx2, d_simp, d_dist;
  PinBuf h_stage;
  HostPool* pool = nullptr;  // host threads for the struct-array gathers (lazy)
  int host_threads = 0;      // size of that pool; 0 = one per hardware thread (at most 32); a multi context shares them out
};

namespace {

using namespace ogjk;

void mark(ogjk_ctx* ctx, cudaStream_t st, int phase, int edge) {
  if (ctx->timing) {
    cudaEventRecord(ctx->ev[2 * phase + edge], st);
    ctx->timed_mask |= 1 << phase;
  }
}

// Counter slots: [0, kQueueSlots) are the work-queue heads run_store_batch / launch_raw reset
// at every batch (classes 0..5, table segment 6, EPA 8..9); kLayoutSlot holds the chunked
// pipeline's layout verdict, which must survive the per-chunk resets.
constexpr int kQueueSlots = 48;   // 16 + 4*c + {0,1,2,3}: parked-pair counts and queue heads of class c's later passes
constexpr int kLayoutSlot = 56;
constexpr int kPoolErrSlot = 57;  // sticky: a raw pool failed validation (store.cuh); cleared when reported
constexpr int kCounterSlots = 64;
int reset_counters(ogjk_ctx* ctx, cudaStream_t st, unsigned long long** out) {
  int rc = ctx->counters.reserve(kCounterSlots * sizeof(unsigned long long));
  if (rc) return rc;
  *out = static_cast<unsigned long long*>(ctx->counters.p);
  CU(cudaMemsetAsync(*out, 0, kQueueSlots * sizeof(unsigned long long), st));
  return OGJK_OK;
}

// Raw modes: the first-generation kernels that read xyz-interleaved blocks directly.
template <typename T>
int launch_raw(ogjk_ctx* ctx, const BatchArgs<T>& a, int mode, cudaStream_t st) {
  unsigned long long* counters = nullptr;
  int rc = reset_counters(ctx, st, &counters);
  if (rc) return rc;
  const int sm = ctx->sm_count;
  if (mode == OGJK_MODE_THREAD_PER_PAIR) {
    long long blocks = (a.npairs + 127) / 128, cap = (long long)sm * 16;
    mark(ctx, st, 2, 0);
    gjk_thread_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, nullptr, nullptr, counters + 0);
    mark(ctx, st, 2, 1);
  } else {
    long long blocks = (a.npairs + 7) / 8, cap = (long long)sm * 8;
    mark(ctx, st, 3, 0);
    gjk_warp_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(a, nullptr, nullptr, counters + 1);
    mark(ctx, st, 3, 1);
  }
  ctx->launches += 1;
  CU(cudaGetLastError());
  return OGJK_OK;
}

// Sticky pool-error flag (counter slot kPoolErrSlot): raised on the device by the pool
// validation in store.cuh, read by the synchronising entry points and ogjk_ctx_pool_status.
int* pool_flag(ogjk_ctx* ctx) {
  return reinterpret_cast<int*>(static_cast<unsigned long long*>(ctx->counters.p) + kPoolErrSlot);
}

// Reads the flag after the work queued on `st` has finished; clears it when set.
int pool_verdict(ogjk_ctx* ctx, cudaStream_t st) {
  int bad = 0;
  CU(cudaMemcpyAsync(&bad, pool_flag(ctx), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (!bad) return OGJK_OK;
  CU(cudaMemsetAsync(pool_flag(ctx), 0, sizeof(int), st));
  CU(cudaStreamSynchronize(st));
  return fail(OGJK_ERR_ARG,
              "polytope pool failed validation: a count outside [1, nverts], an offset range outside the coordinate "
              "array, or (device-array entry points) polytopes sharing vertices so that sum(cnt) > nverts%s");
}

long long host_sum(const int* cnt, long long n) {
  long long t = 0;
  for (long long i = 0; i < n; ++i) t += cnt[i] > 0 ? cnt[i] : 0;
  return t;
}

// Build a packed store from xyz-interleaved DEVICE arrays.  `exact` sizes the
// allocation from the scanned total (one stream sync); otherwise an upper bound
// (cap_verts vertices, every polytope padded by 3) keeps the call asynchronous.
// cap_verts >= nverts is the caller's bound on sum(cnt) (host paths know the sum;
// device paths pass nverts, i.e. assume polytopes do not share vertices -- a pool that
// breaks the bound is neutralised by overflow_guard_kernel and flagged, never overrun).
template <typename T>
int store_build(ogjk_ctx* ctx, ogjk_store_* s, long long npoly, long long nverts, const T* coords, const int64_t* off,
                const int* cnt, bool exact, cudaStream_t st, bool pack_now = true, long long cap_verts = 0) {
  s->device = ctx->device;
  s->prec = (int)sizeof(T);
  s->npoly = npoly;
  s->nverts = nverts;
  int rc;
  if ((rc = s->boff.reserve((size_t)npoly * sizeof(long long)))) return rc;
  if ((rc = s->cnt.reserve((size_t)npoly * sizeof(int)))) return rc;
  const long long nblocks = (npoly + kScanBlock - 1) / kScanBlock;
  if ((rc = s->scan_tmp.reserve((size_t)(nblocks + 1) * sizeof(long long)))) return rc;
  int* flag = pool_flag(ctx);
  int* scnt = static_cast<int*>(s->cnt.p);
  const long long gb = (npoly + 255) / 256, gcap = (long long)ctx->sm_count * 32;
  const unsigned ggrid = (unsigned)(gb < gcap ? gb : gcap);
  sanitize_counts_kernel<<<ggrid, 256, 0, st>>>(cnt, npoly, nverts, scnt, flag);
  auto* sums = static_cast<long long*>(s->scan_tmp.p);
  auto* boff = static_cast<long long*>(s->boff.p);
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(scnt, npoly, sums, BlockElems{});
  scan_spine_kernel<<<1, 1024, 0, st>>>(sums, nblocks, sums + nblocks);
  scan_finish_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(scnt, npoly, sums, boff, BlockElems{});
  if (cap_verts < nverts) cap_verts = nverts;
  long long elems = 3 * (cap_verts + 3 * npoly);
  if (exact) {
    long long total = 0;
    CU(cudaMemcpyAsync(&total, sums + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    elems = total;
  }
  if (elems < 12) elems = 12;  // room for the guard's dummy block
  if ((rc = s->data.reserve((size_t)elems * sizeof(T)))) return rc;
  s->capacity = elems;
  ctx->launches += 4;
  if (!exact) {
    overflow_guard_kernel<<<ggrid, 256, 0, st>>>(sums + nblocks, elems, npoly, scnt, boff, flag);
    ctx->launches += 1;
  }
  if (pack_now) {
    long long pblocks = (npoly * 8 + 255) / 256, cap = (long long)ctx->sm_count * 32;
    pack_kernel<T><<<(unsigned)(pblocks < cap ? pblocks : cap), 256, 0, st>>>(
        coords, reinterpret_cast<const long long*>(off), scnt, npoly, boff, static_cast<T*>(s->data.p), nverts, flag);
    ctx->launches += 1;
  }
  CU(cudaGetLastError());
  return OGJK_OK;
}

// Repack polytopes [lo, hi) of a prepared store (offsets already scanned).
template <typename T>
int store_pack_range(ogjk_ctx* ctx, ogjk_store_* s, long long lo, long long hi, const T* coords, const int64_t* off,
                     cudaStream_t st) {
  const long long n = hi - lo;
  if (n <= 0) return OGJK_OK;
  long long pblocks = (n * 8 + 255) / 256, cap = (long long)ctx->sm_count * 32;
  pack_kernel<T><<<(unsigned)(pblocks < cap ? pblocks : cap), 256, 0, st>>>(
      coords, reinterpret_cast<const long long*>(off) + lo, static_cast<const int*>(s->cnt.p) + lo, n,
      static_cast<const long long*>(s->boff.p) + lo, static_cast<T*>(s->data.p), s->nverts, pool_flag(ctx));
  ctx->launches += 1;
  CU(cudaGetLastError());
  return OGJK_OK;
}

// Cluster tables for the polytopes of a packed store with at least ctx->accel_min vertices
// (persistent stores only: the build sorts every such polytope once, queries then prune
// their support scans, see store.cuh).  Synchronises `st`.
template <typename T>
int store_build_accel(ogjk_ctx* ctx, ogjk_store_* s, cudaStream_t st) {
  s->accel_min = 0x7fffffff;
  s->accel_elems = 0;
  const int amin = ctx->accel_min;
  if (amin <= 0) return OGJK_OK;
  const long long npoly = s->npoly;
  const long long nblocks = (npoly + kScanBlock - 1) / kScanBlock;
  int rc;
  if ((rc = s->aoff.reserve((size_t)npoly * sizeof(long long)))) return rc;
  if ((rc = s->scan_tmp.reserve((size_t)(nblocks + 2 + 8) * sizeof(long long)))) return rc;
  const int* cnt = static_cast<const int*>(s->cnt.p);
  auto* sums = static_cast<long long*>(s->scan_tmp.p);
  auto* aoff = static_cast<long long*>(s->aoff.p);
  int* d_max = reinterpret_cast<int*>(sums + nblocks + 1);
  auto* build_next = reinterpret_cast<unsigned long long*>(sums + nblocks + 2);  // one work counter per build launch
  CU(cudaMemsetAsync(d_max, 0, 9 * sizeof(long long), st));
  const AccelElems<T> f{amin};
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(cnt, npoly, sums, f);
  scan_spine_kernel<<<1, 1024, 0, st>>>(sums, nblocks, sums + nblocks);
  scan_finish_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(cnt, npoly, sums, aoff, f);
  {
    long long mb = (npoly + 255) / 256, cap = (long long)ctx->sm_count * 8;
    max_count_kernel<<<(unsigned)(mb < cap ? mb : cap), 256, 0, st>>>(cnt, npoly, amin, d_max);
  }
  long long tail[2] = {0, 0};  // total elements, largest eligible vertex count
  CU(cudaMemcpyAsync(tail, sums + nblocks, sizeof(tail), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  ctx->launches += 4;
  const long long total = tail[0];
  const int maxn = (int)(tail[1] & 0xffffffffll);
  if (total <= 0 || maxn <= 0) return OGJK_OK;
  if ((rc = s->accel.reserve((size_t)total * sizeof(T)))) return rc;
  StoreView<T> v{static_cast<const T*>(s->data.p), static_cast<const long long*>(s->boff.p), cnt};
  v.accel = static_cast<const T*>(s->accel.p);
  v.aoff = aoff;
  v.accel_min = amin;
  // One launch per power-of-two size range, so that small polytopes get narrow CTAs, many per SM
  // (the per-level sorts are barrier-latency bound), and only the largest ones a whole SM.
  int pmax = 2048;
  while (pmax < maxn) pmax <<= 1;
  CU(cudaFuncSetAttribute(accel_build_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)((size_t)pmax * sizeof(unsigned long long))));
  int launch = 0;
  for (int pcap = 2048, nlo = 0; nlo < maxn; nlo = pcap, pcap <<= 1, ++launch) {
    const int threads = pcap <= 2048 ? 256 : pcap <= 4096 ? 512 : 1024;
    const int per_sm = pcap <= 2048 ? 8 : pcap <= 4096 ? 4 : pcap <= 8192 ? 2 : 1;
    const long long cap = (long long)ctx->sm_count * per_sm;
    accel_build_kernel<T><<<(unsigned)(npoly < cap ? npoly : cap), threads, (size_t)pcap * sizeof(unsigned long long), st>>>(
        v, npoly, static_cast<T*>(s->accel.p), nlo, pcap, build_next + launch);
    ctx->launches += 1;
  }
  CU(cudaGetLastError());
  s->accel_min = amin;
  s->accel_elems = total;
  return OGJK_OK;
}

template <typename T> StoreView<T> view_of(const ogjk_store_* s) {
  StoreView<T> v{static_cast<const T*>(s->data.p), static_cast<const long long*>(s->boff.p),
                 static_cast<const int*>(s->cnt.p)};
  if (s->accel_elems > 0) {
    v.accel = static_cast<const T*>(s->accel.p);
    v.aoff = static_cast<const long long*>(s->aoff.p);
    v.accel_min = s->accel_min;
    v.npoly = s->npoly;
  }
  return v;
}

template <typename T, int G>
void launch_group(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                  unsigned long long* head, cudaStream_t st) {
  const char* cap_env = getenv("OGJK_GRID_CAP");  // tuning hook (blocks per SM), read per call
  const int cap_per_sm = cap_env ? atoi(cap_env) : 16;
  long long blocks = (a.npairs * G + 127) / 128, cap = (long long)ctx->sm_count * cap_per_sm;
#ifdef OGJK_EXPERIMENTAL
  if constexpr (G == 2) {
    if (ctx->regroup[0] > 0) {  // in-CTA regrouping after regroup[0] and regroup[1] iterations (regroup_kernel.cuh)
      blocks = (a.npairs * G + kRegroupThreads - 1) / kRegroupThreads;
      cap = (long long)ctx->sm_count * 8;
      gjk_regroup_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), kRegroupThreads, 0, st>>>(a, list, b, e, head,
                                                                                                   ctx->regroup[0], ctx->regroup[1]);
      return;
    }
  }
#endif
  gjk_group_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, list, b, e, head);
}

#ifdef OGJK_EXPERIMENTAL
// One size class in 2 or 3 passes (see gjk_pass_kernel).  Falls back to the one-pass kernel
// when the parking buffers cannot be allocated.
template <typename T, int G>
void launch_passes(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                   unsigned long long* counters, int c, cudaStream_t st) {
  const int passes = ctx->passes > 3 ? 3 : ctx->passes;
  const size_t rows = (size_t)a.npairs;
  bool ok = true;
  for (int j = 0; j + 1 < passes && ok; ++j)
    ok = ctx->pass_list[j].reserve(rows * sizeof(int)) == OGJK_OK &&
         ctx->pass_f[j].reserve(rows * kParkF * sizeof(T)) == OGJK_OK &&
         ctx->pass_i[j].reserve(rows * kParkI * sizeof(int)) == OGJK_OK;
  if (!ok) {
    cudaGetLastError();
    launch_group<T, G>(ctx, a, list, b, e, counters + c, st);
    return;
  }
  long long blocks = (a.npairs * G + 127) / 128, cap = (long long)ctx->sm_count * 16;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  unsigned long long* extra = counters + 16 + 4 * c;  // count A, head A, count B, head B
  int* list_a = static_cast<int*>(ctx->pass_list[0].p);
  T* f_a = static_cast<T*>(ctx->pass_f[0].p);
  int* i_a = static_cast<int*>(ctx->pass_i[0].p);
  unsigned* count_a = reinterpret_cast<unsigned*>(extra + 0);
  const int k1 = ctx->pass_iters[0], k2 = k1 + ctx->pass_iters[1];
  gjk_pass_kernel<T, G><<<grid, 128, 0, st>>>(a, list, b, e, counters + c, PassArgs<T>{0, k1, nullptr, nullptr, list_a, count_a, f_a, i_a});
  if (passes == 2) {
    gjk_pass_kernel<T, G><<<grid, 128, 0, st>>>(a, list_a, nullptr, count_a, extra + 1,
                                                PassArgs<T>{1, kMaxIter, f_a, i_a, nullptr, nullptr, nullptr, nullptr});
    ctx->launches += 1;
    return;
  }
  int* list_b = static_cast<int*>(ctx->pass_list[1].p);
  T* f_b = static_cast<T*>(ctx->pass_f[1].p);
  int* i_b = static_cast<int*>(ctx->pass_i[1].p);
  unsigned* count_b = reinterpret_cast<unsigned*>(extra + 2);
  gjk_pass_kernel<T, G><<<grid, 128, 0, st>>>(a, list_a, nullptr, count_a, extra + 1, PassArgs<T>{1, k2, f_a, i_a, list_b, count_b, f_b, i_b});
  gjk_pass_kernel<T, G><<<grid, 128, 0, st>>>(a, list_b, nullptr, count_b, extra + 3,
                                              PassArgs<T>{1, kMaxIter, f_b, i_b, nullptr, nullptr, nullptr, nullptr});
  ctx->launches += 2;
}

#endif  // OGJK_EXPERIMENTAL

template <typename T, int G>
void launch_table(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                  unsigned long long* head, cudaStream_t st) {
  long long blocks = (a.npairs * G + 127) / 128, cap = (long long)ctx->sm_count * 16;
  gjk_table_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, list, b, e, head);
}

#ifdef OGJK_EXPERIMENTAL
// Shared-memory slot size for a staged class: room for 4*hi_key vertices, rounded so
// that slot_bytes = 16*G (mod 128) (bank spreading, see gjk_staged_kernel).
template <typename T> int staged_slot_bytes(int hi_key, int G) {
  long long bytes = 4ll * hi_key * 3 * (long long)sizeof(T);
  bytes = (bytes + 127) & ~127ll;
  if (G < 8) bytes += 16 * G;
  const long long limit = 227ll * 1024 / (128 / G);  // slots per block = 128 / G
  return bytes <= limit ? (int)bytes : 0;             // 0: does not fit, do not stage
}

template <typename T, int G>
int launch_staged(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                  unsigned long long* head, int hi_key, cudaStream_t st) {
  const int slot = staged_slot_bytes<T>(hi_key, G);
  if (slot == 0 || hi_key >= kMaxKey) {  // open-ended or oversized class: read global memory instead
    launch_group<T, G>(ctx, a, list, b, e, head, st);
    return OGJK_OK;
  }
  const size_t smem = (size_t)(128 / G) * slot;
  // per device and cheap: set on every launch (a process-wide cache would miss other devices)
  CU(cudaFuncSetAttribute(gjk_staged_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (a.npairs * G + 127) / 128, cap = (long long)ctx->sm_count * 16;
  gjk_staged_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), 128, smem, st>>>(a, list, b, e, head, slot);
  return OGJK_OK;
}

// Bulk-copy (TMA) staged flavour, double buffered per warp (tma_kernel.cuh).  The buffer size is a tuning
// hook ($OGJK_TMA_SLOT bytes per buffer, default 14336: 4 warps x 2 buffers = 112 KB per CTA, 2 CTAs per SM).
template <typename T, int G>
int launch_tma(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
               unsigned long long* head, cudaStream_t st) {
  const char* env = getenv("OGJK_TMA_SLOT");
  int slot = env ? atoi(env) : 14336;
  slot = (slot < 1024 ? 1024 : slot > 26624 ? 26624 : slot) & ~127;
  const size_t smem = (size_t)kTmaWarps * 2 * slot;
  CU(cudaFuncSetAttribute(gjk_tma_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk_tma_kernel<T, G>, 32 * kTmaWarps, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  long long blocks = (a.npairs * G + 32 * kTmaWarps - 1) / (32 * kTmaWarps), cap = (long long)ctx->sm_count * per_sm;
  gjk_tma_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), 32 * kTmaWarps, smem, st>>>(a, list, b, e, head, slot);
  return OGJK_OK;
}

// CTA-per-pair flavour: dynamic shared memory = the class's largest pair.
template <typename T, int WARPS>
int launch_cta(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
               unsigned long long* head, int hi_key, cudaStream_t st) {
  const long long bytes = (4ll * hi_key * 3 * (long long)sizeof(T) + 127) & ~127ll;
  if (bytes > 200 * 1024 || hi_key >= kMaxKey) {  // does not fit shared memory: warp-per-pair over global memory
    launch_group<T, 32>(ctx, a, list, b, e, head, st);
    return OGJK_OK;
  }
  CU(cudaFuncSetAttribute(gjk_cta_kernel<T, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  long long per_sm = (220ll * 1024) / (bytes + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2048 / (32 * WARPS)) per_sm = 2048 / (32 * WARPS);
  long long blocks = a.npairs, cap = (long long)ctx->sm_count * per_sm;
  gjk_cta_kernel<T, WARPS><<<(unsigned)(blocks < cap ? blocks : cap), 32 * WARPS, (size_t)bytes, st>>>(a, list, b, e, head);
  return OGJK_OK;
}

template <typename T, int G>
void launch_assist(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                   unsigned long long* head, cudaStream_t st) {
  long long blocks = (a.npairs * G + 127) / 128, cap = (long long)ctx->sm_count * 16;
  gjk_assist_kernel<T, G><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, list, b, e, head);
}

// Persistent-lane flavour: iterate kernel, then (when simplices are wanted) the witness pass
// over the same list segment.
template <typename T, int G>
void launch_persist(ogjk_ctx* ctx, const StoreArgs<T>& a, const int* list, const unsigned* b, const unsigned* e,
                    unsigned long long* head, cudaStream_t st) {
  long long blocks = (a.npairs * (G ? G : 1) + 127) / 128, cap = (long long)ctx->sm_count * 16;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  if (G == 0) gjk_persist_small_kernel<T><<<grid, 128, 0, st>>>(a, list, b, e, head);
  else gjk_persist_group_kernel<T, (G ? G : 1)><<<grid, 128, 0, st>>>(a, list, b, e, head);
  if (a.simplices) {
    long long wb = (a.npairs + 255) / 256, wcap = (long long)ctx->sm_count * 16;
    gjk_witness_kernel<T><<<(unsigned)(wb < wcap ? wb : wcap), 256, 0, st>>>(a, list, b, e);
    ctx->launches += 1;
  }
}

#endif  // OGJK_EXPERIMENTAL

// The batch scheduler: counting sort by size, then one launch per size class.
template <typename T>
int run_store_batch(ogjk_ctx* ctx, const StoreArgs<T>& a, cudaStream_t st) {
  if (a.npairs <= 0) return OGJK_OK;
  if (a.npairs > 0x7fffffffll) return fail(OGJK_ERR_ARG, "npairs exceeds 2^31-1 per call%s");
  unsigned long long* counters = nullptr;
  int rc = reset_counters(ctx, st, &counters);
  if (rc) return rc;
  const size_t nkeys = kNumBinsAll + 1;
  if ((rc = ctx->sort_tmp.reserve(3 * nkeys * sizeof(unsigned)))) return rc;
  if ((rc = ctx->sorted_list.reserve((size_t)a.npairs * sizeof(int)))) return rc;
  unsigned* hist = static_cast<unsigned*>(ctx->sort_tmp.p);
  unsigned* start = hist + nkeys;
  unsigned* cursor = start + nkeys;
  int* sorted = static_cast<int*>(ctx->sorted_list.p);

  mark(ctx, st, 1, 0);
  CU(cudaMemsetAsync(hist, 0, nkeys * sizeof(unsigned), st));
  const long long per_block = (long long)kSortBlock * kSortItems;
  const unsigned sblocks = (unsigned)((a.npairs + per_block - 1) / per_block);
  const bool tables = a.s1.accel && a.s2.accel;
  const int nbins = tables ? kNumBinsAll : kNumBins;
  sort_hist_kernel<T><<<sblocks, kSortBlock, 0, st>>>(a.s1, a.s2, a.idx1, a.idx2, a.npairs, hist, nbins);
  sort_scan_kernel<<<1, 1024, 0, st>>>(hist, start, cursor, nbins);
  sort_scatter_kernel<T><<<sblocks, kSortBlock, 0, st>>>(a.s1, a.s2, a.idx1, a.idx2, a.npairs, start, cursor, sorted, nbins);
  mark(ctx, st, 1, 1);
  ctx->launches += 3;

  // start[b] = number of pairs with bin > b  =>  the class of size keys (lo, hi] occupies
  // [start[bin_upper(hi)], start[bin_upper(lo)]) of the sorted list
  int last_run = 0;  // last class that launches (classes whose range is empty are skipped)
  for (int c = 0, l = 0; c < ctx->nclasses; ++c) {
    const int h = ctx->class_hi[c] > kMaxKey ? kMaxKey : ctx->class_hi[c];
    if (h > l) { last_run = c; l = h; }
  }
  int lo = 0;
  for (int c = 0; c < ctx->nclasses; ++c) {
    const int hi = ctx->class_hi[c] > kMaxKey ? kMaxKey : ctx->class_hi[c];
    if (hi <= lo) continue;
    const unsigned* b = start + bin_upper(hi);  // class = sorted[*b, *e)
    const unsigned* e = start + bin_upper(lo);
    unsigned long long* head = counters + c;
    mark(ctx, st, 2 + c, 0);
    int lanes = ctx->class_lanes[c];
    // Persistent stores with cluster tables: the pairs whose two bodies both have one were
    // sorted in front of everything else (store.cuh, kNumBinsAll) and run through the
    // pruned-scan kernel; timed with the last class.
    if (tables && c == last_run) {
      // [0, *t_mid): pairs with a body above one tile (1024 vertices); [*t_mid, *t_end): both bodies within one
      const unsigned* t_beg = start + (kNumBinsAll - 1);          // = 0 pairs above the top bin
      const unsigned* t_mid = start + (kNumBins + kTableBins - 1);  // = pairs in the multi-tile bins
      const unsigned* t_end = start + (kNumBins - 1);             // = pairs in all table bins
      auto gen1 = [&](int lanes1, const unsigned* sb, const unsigned* se, unsigned long long* h) {
        switch (lanes1) {
          case 8: launch_table<T, 8>(ctx, a, sorted, sb, se, h, st); break;
          case 16: launch_table<T, 16>(ctx, a, sorted, sb, se, h, st); break;
          default: launch_table<T, 32>(ctx, a, sorted, sb, se, h, st); break;
        }
      };
      auto gen3 = [&](const unsigned* sb, const unsigned* se, unsigned long long* h) {
        long long blocks = (a.npairs * 32 + 127) / 128, cap = (long long)ctx->sm_count * 16;
        gjk_table3_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, sorted, sb, se, h);
      };
      auto batch = [&](const unsigned* sb, const unsigned* se, unsigned long long* h) {
        // one resident wave: the kernel sizes its batches from the number of warps in the grid
        int& per_sm = ctx->batch_blocks[sizeof(T) == 8];
        if (!per_sm && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk_tablebatch_kernel<T>, 128, 0) != cudaSuccess)
          per_sm = 4;
        long long blocks = (a.npairs + 15) / 16, cap = (long long)ctx->sm_count * per_sm;
        gjk_tablebatch_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, sorted, sb, se, h);
      };
      const int tl = ctx->table_lanes;
      if (tl == 1) {          // tuning: the pair-batched kernel (32 pairs per warp) for everything
        batch(t_beg, t_mid, counters + 6);
        batch(t_mid, t_end, counters + 7);
      } else if (tl == 2) {   // tuning: tile kernel above one tile, pair-batched kernel below
        gen3(t_beg, t_mid, counters + 6);
        batch(t_mid, t_end, counters + 7);
      } else if (tl == 0) {   // default: the tile kernel above one tile; below, the pair-batched kernel for
        gen3(t_beg, t_mid, counters + 6);  // long batches and the first-generation kernel for short ones
        if (a.npairs >= kBatchKernelMinPairs) batch(t_mid, t_end, counters + 7);
        else gen1(32, t_mid, t_end, counters + 7);
      } else if (tl == 3) {   // tuning: the tile kernel for everything
        gen3(t_beg, t_mid, counters + 6);
        gen3(t_mid, t_end, counters + 7);
      } else {                // tuning: the first-generation kernel with 8 / 16 / 32 lanes per pair
        gen1(tl, t_beg, t_mid, counters + 6);
        gen1(tl, t_mid, t_end, counters + 7);
      }
      ctx->launches += 2;
    }
    switch (lanes) {
      case 0: {
        long long blocks = (a.npairs + 127) / 128, cap = (long long)ctx->sm_count * 16;
        gjk_small_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(a, sorted, b, e, head);
      } break;
#ifdef OGJK_EXPERIMENTAL
      case 1: if (ctx->passes > 1) launch_passes<T, 1>(ctx, a, sorted, b, e, counters, c, st); else launch_group<T, 1>(ctx, a, sorted, b, e, head, st); break;
      case 2: if (ctx->passes > 1) launch_passes<T, 2>(ctx, a, sorted, b, e, counters, c, st); else launch_group<T, 2>(ctx, a, sorted, b, e, head, st); break;
      case 4: if (ctx->passes > 1) launch_passes<T, 4>(ctx, a, sorted, b, e, counters, c, st); else launch_group<T, 4>(ctx, a, sorted, b, e, head, st); break;
      case 8: if (ctx->passes > 1) launch_passes<T, 8>(ctx, a, sorted, b, e, counters, c, st); else launch_group<T, 8>(ctx, a, sorted, b, e, head, st); break;
#else
      case 1: launch_group<T, 1>(ctx, a, sorted, b, e, head, st); break;
      case 2: launch_group<T, 2>(ctx, a, sorted, b, e, head, st); break;
      case 4: launch_group<T, 4>(ctx, a, sorted, b, e, head, st); break;
      case 8: launch_group<T, 8>(ctx, a, sorted, b, e, head, st); break;
#endif
      case 16: launch_group<T, 16>(ctx, a, sorted, b, e, head, st); break;
      case 32: launch_group<T, 32>(ctx, a, sorted, b, e, head, st); break;
#ifdef OGJK_EXPERIMENTAL
      case 502: rc = launch_cta<T, 2>(ctx, a, sorted, b, e, head, hi, st); break;
      case 504: rc = launch_cta<T, 4>(ctx, a, sorted, b, e, head, hi, st); break;
      case 508: rc = launch_cta<T, 8>(ctx, a, sorted, b, e, head, hi, st); break;
      case 516: rc = launch_cta<T, 16>(ctx, a, sorted, b, e, head, hi, st); break;
      case 302: launch_assist<T, 2>(ctx, a, sorted, b, e, head, st); break;
      case 304: launch_assist<T, 4>(ctx, a, sorted, b, e, head, st); break;
      case 308: launch_assist<T, 8>(ctx, a, sorted, b, e, head, st); break;
      case 316: launch_assist<T, 16>(ctx, a, sorted, b, e, head, st); break;
      case 200: launch_persist<T, 0>(ctx, a, sorted, b, e, head, st); break;
      case 201: launch_persist<T, 1>(ctx, a, sorted, b, e, head, st); break;
      case 202: launch_persist<T, 2>(ctx, a, sorted, b, e, head, st); break;
      case 204: launch_persist<T, 4>(ctx, a, sorted, b, e, head, st); break;
      case 208: launch_persist<T, 8>(ctx, a, sorted, b, e, head, st); break;
      case 216: launch_persist<T, 16>(ctx, a, sorted, b, e, head, st); break;
      case 232: launch_persist<T, 32>(ctx, a, sorted, b, e, head, st); break;
      case 602: rc = launch_tma<T, 2>(ctx, a, sorted, b, e, head, st); break;
      case 604: rc = launch_tma<T, 4>(ctx, a, sorted, b, e, head, st); break;
      case 608: rc = launch_tma<T, 8>(ctx, a, sorted, b, e, head, st); break;
      case 101: rc = launch_staged<T, 1>(ctx, a, sorted, b, e, head, hi, st); break;
      case 102: rc = launch_staged<T, 2>(ctx, a, sorted, b, e, head, hi, st); break;
      case 104: rc = launch_staged<T, 4>(ctx, a, sorted, b, e, head, hi, st); break;
      case 108: rc = launch_staged<T, 8>(ctx, a, sorted, b, e, head, hi, st); break;
      case 116: rc = launch_staged<T, 16>(ctx, a, sorted, b, e, head, hi, st); break;
      case 132: rc = launch_staged<T, 32>(ctx, a, sorted, b, e, head, hi, st); break;
#endif
      default: return fail(OGJK_ERR_ARG, "kernel flavour of the class table is not compiled into this build%s");
    }
    if (rc) return rc;
    mark(ctx, st, 2 + c, 1);
    ctx->launches += 1;
    lo = hi;
  }
  CU(cudaGetLastError());
  return OGJK_OK;
}

// EPA over a batch whose GJK results are on the device: pre-pass (defaults + work
// list of penetrating pairs), then one warp per listed pair.  Timing slot 7.
template <typename T>
int run_store_epa(ogjk_ctx* ctx, const EpaArgs<T>& a, cudaStream_t st) {
  if (a.npairs <= 0) return OGJK_OK;
  if (a.npairs > 0x7fffffffll) return fail(OGJK_ERR_ARG, "npairs exceeds 2^31-1 per call%s");
  int rc = ctx->counters.reserve(kCounterSlots * sizeof(unsigned long long));
  if (rc) return rc;
  if ((rc = ctx->epa_list.reserve((size_t)a.npairs * sizeof(int)))) return rc;
  auto* counters = static_cast<unsigned long long*>(ctx->counters.p);
  CU(cudaMemsetAsync(counters + 8, 0, 2 * sizeof(unsigned long long), st));
  int* list = static_cast<int*>(ctx->epa_list.p);
  mark(ctx, st, 7, 0);
  long long pb = (a.npairs + 255) / 256, pcap = (long long)ctx->sm_count * 16;
  epa_prepass_kernel<T><<<(unsigned)(pb < pcap ? pb : pcap), 256, 0, st>>>(a, list, counters + 8);
  long long eb = (a.npairs + kEpaWarps - 1) / kEpaWarps, ecap = (long long)ctx->sm_count * 8;
  epa_warp_kernel<T><<<(unsigned)(eb < ecap ? eb : ecap), 32 * kEpaWarps, 0, st>>>(a, list, counters + 8, counters + 9);
  mark(ctx, st, 7, 1);
  ctx->launches += 2;
  CU(cudaGetLastError());
  return OGJK_OK;
}

template <typename T, typename S>
int gjk_device(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1,
               int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2,
               int64_t nverts2, const int* idx2, S* simplices, T* distances, int mode, void* stream) {
  if (!ctx) return fail(OGJK_ERR_ARG, "ctx is NULL%s");
  if (npairs < 0) return fail(OGJK_ERR_ARG, "npairs < 0%s");
  if (npairs == 0) return OGJK_OK;
  if (!coords1 || !off1 || !cnt1 || !coords2 || !off2 || !cnt2 || !distances)
    return fail(OGJK_ERR_ARG, "NULL device pointer%s");
  if (mode < OGJK_MODE_AUTO || mode > OGJK_MODE_WARP_PER_PAIR) return fail(OGJK_ERR_ARG, "bad mode%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ctx->timed_mask = 0;
  if (mode != OGJK_MODE_AUTO) {
    BatchArgs<T> a;
    a.coords1 = coords1; a.off1 = reinterpret_cast<const long long*>(off1); a.cnt1 = cnt1; a.idx1 = idx1;
    a.coords2 = coords2; a.off2 = reinterpret_cast<const long long*>(off2); a.cnt2 = cnt2; a.idx2 = idx2;
    a.simplices = reinterpret_cast<GkSimplex<T>*>(simplices);
    a.distances = distances;
    a.npairs = npairs;
    if (npairs > 0x7fffffffll) return fail(OGJK_ERR_ARG, "npairs exceeds 2^31-1 per call%s");
    return launch_raw<T>(ctx, a, mode, st);
  }
  // AUTO: repack the raw pools into the context's scratch stores, then schedule.
  if (npoly1 <= 0 || npoly2 <= 0) return fail(OGJK_ERR_ARG, "AUTO mode needs the pool sizes (npoly, nverts)%s");
  const bool shared_pool = (coords1 == coords2 && off1 == off2 && cnt1 == cnt2);
  int rc;
  mark(ctx, st, 0, 0);
  if ((rc = store_build<T>(ctx, &ctx->scratch1, npoly1, nverts1, coords1, off1, cnt1, false, st))) return rc;
  if (!shared_pool && (rc = store_build<T>(ctx, &ctx->scratch2, npoly2, nverts2, coords2, off2, cnt2, false, st))) return rc;
  mark(ctx, st, 0, 1);
  StoreArgs<T> sa;
  sa.s1 = view_of<T>(&ctx->scratch1);
  sa.s2 = shared_pool ? sa.s1 : view_of<T>(&ctx->scratch2);
  sa.idx1 = idx1; sa.idx2 = idx2;
  sa.simplices = reinterpret_cast<GkSimplex<T>*>(simplices);
  sa.distances = distances;
  sa.npairs = npairs;
  return run_store_batch<T>(ctx, sa, st);
}

// Upload one host array into a grow-only device buffer on the context stream.
int upload(ogjk_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
  int rc = b.reserve(bytes ? bytes : 1);
  if (rc) return rc;
  if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->own_stream));
  return OGJK_OK;
}

HostPool* host_pool(ogjk_ctx* ctx);

// Pinned (page-locked or registered) host memory?  Device-to-host copies into pageable memory are
// staged by the driver through small bounce buffers and block the issuing thread: ~4x slower than
// the link, and they serialise the pipeline.
bool host_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Sequential writer for the gather: collects the small per-polytope blocks in a cache-resident
// buffer and writes whole lines to the pinned stage with streaming stores, so that the stage is
// written once instead of read (for ownership) and written -- the gather runs beside the H2D
// copy and both are bound by host memory bandwidth.
class StreamOut {
 public:
  explicit StreamOut(void* dst) : dst_(static_cast<char*>(dst)) {}
  void put(const void* src, size_t n) {
    const char* s = static_cast<const char*>(src);
    while (n) {
      if (!fill_ && (reinterpret_cast<uintptr_t>(dst_) & 63)) {  // head: up to the next line
        size_t h = 64 - (reinterpret_cast<uintptr_t>(dst_) & 63);
        if (h > n) h = n;
        memcpy(dst_, s, h);
        dst_ += h; s += h; n -= h;
        continue;
      }
      size_t take = sizeof(buf_) - fill_;
      if (take > n) take = n;
      memcpy(buf_ + fill_, s, take);
      fill_ += take; s += take; n -= take;
      if (fill_ == sizeof(buf_)) { lines(fill_); fill_ = 0; }
    }
  }
  void finish() {
    const size_t whole = fill_ & ~(size_t)63;
    lines(whole);
    if (fill_ > whole) memcpy(dst_, buf_ + whole, fill_ - whole);
    dst_ += fill_ - whole;
    fill_ = 0;
#if defined(__SSE2__)
    _mm_sfence();
#endif
  }

 private:
  void lines(size_t bytes) {  // dst_ is 64-byte aligned here
#if defined(__SSE2__)
    for (size_t i = 0; i < bytes; i += 16)
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst_ + i), _mm_load_si128(reinterpret_cast<const __m128i*>(buf_ + i)));
#else
    memcpy(dst_, buf_, bytes);
#endif
    dst_ += bytes;
  }
  char* dst_;
  size_t fill_ = 0;
  alignas(64) char buf_[4096];
};

// One large block through the same streaming stores (pageable -> pinned staging and back).
inline void stream_copy(void* dst, const void* src, size_t n) {
#if defined(__SSE2__)
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
  if (head > n) head = n;
  memcpy(d, s, head);
  d += head; s += head; n -= head;
  const size_t body = n & ~(size_t)63;
  for (size_t i = 0; i < body; i += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)));
  memcpy(d + body, s + body, n - body);
  _mm_sfence();
#else
  memcpy(dst, src, n);
#endif
}

// Chunked variant of the host pipeline for the un-indexed form (pair p uses polytope p
// of each pool -- the reference's bd1[] / bd2[] arrays): the pair range is cut into
// chunks whose coordinate ranges are contiguous, and H2D of chunk c+1, the kernels
// of chunk c and D2H of chunk c-1 overlap on three streams (PCIe is full duplex and
// the transfer is ~10x longer than the kernels, so e2e approaches the H2D time).
// Returns 1 when the layout is not chunkable (caller falls back), 0 on success.
template <typename T, typename S>
int host_pipeline_chunked(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1,
                          int64_t nverts1, const T* coords2, const int64_t* off2, const int* cnt2, int64_t nverts2,
                          S* simplices, T* distances, T* w1, T* w2, T* normals, bool run_epa,
                          const HostReady* ready = nullptr) {
  const int K = ctx->pipeline_chunks > 16 ? 16 : ctx->pipeline_chunks;
  if (K < 2 || npairs < 4096 * (int64_t)K) return 1;
  // chunk coordinate ranges; require monotone, gap-free-ish layouts
  struct Range { int64_t lo, hi, a1, b1, a2, b2; };
  Range r[16];
  int64_t covered1 = 0, covered2 = 0;
  for (int c = 0; c < K; ++c) {
    r[c].lo = npairs * c / K;
    r[c].hi = npairs * (c + 1) / K;
    const int64_t last = r[c].hi - 1;
    r[c].a1 = off1[r[c].lo]; r[c].b1 = off1[last] + cnt1[last];
    r[c].a2 = off2[r[c].lo]; r[c].b2 = off2[last] + cnt2[last];
    if (r[c].b1 < r[c].a1 || r[c].b2 < r[c].a2 || r[c].b1 > nverts1 || r[c].b2 > nverts2) return 1;
    if (c && (r[c].a1 < r[c - 1].b1 || r[c].a2 < r[c - 1].b2)) return 1;  // not sorted by pair
    covered1 += r[c].b1 - r[c].a1;
    covered2 += r[c].b2 - r[c].a2;
  }
  (void)covered1; (void)covered2;

  cudaStream_t sc = ctx->own_stream, si = ctx->in_stream, so = ctx->out_stream;
  int rc;
  // small metadata first, then offsets scan for both pools (no coordinates needed)
  if ((rc = ctx->d_coords1.reserve((size_t)nverts1 * 3 * sizeof(T)))) return rc;
  if ((rc = ctx->d_coords2.reserve((size_t)nverts2 * 3 * sizeof(T)))) return rc;
  if ((rc = upload(ctx, ctx->d_off1, off1, (size_t)npairs * sizeof(int64_t)))) return rc;
  if ((rc = upload(ctx, ctx->d_cnt1, cnt1, (size_t)npairs * sizeof(int)))) return rc;
  if ((rc = upload(ctx, ctx->d_off2, off2, (size_t)npairs * sizeof(int64_t)))) return rc;
  if ((rc = upload(ctx, ctx->d_cnt2, cnt2, (size_t)npairs * sizeof(int)))) return rc;
  if ((rc = ctx->d_dist.reserve((size_t)npairs * sizeof(T)))) return rc;
  if ((rc = ctx->d_simp.reserve((size_t)npairs * sizeof(S)))) return rc;
  if (run_epa) {
    if ((rc = ctx->d_w1.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
    if ((rc = ctx->d_w2.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
    if (normals && (rc = ctx->d_nrm.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
  }
  const T* dc1 = static_cast<const T*>(ctx->d_coords1.p);
  const T* dc2 = static_cast<const T*>(ctx->d_coords2.p);
  const int64_t* do1 = static_cast<const int64_t*>(ctx->d_off1.p);
  const int64_t* do2 = static_cast<const int64_t*>(ctx->d_off2.p);
  const int* dk1 = static_cast<const int*>(ctx->d_cnt1.p);
  const int* dk2 = static_cast<const int*>(ctx->d_cnt2.p);
  // results go straight to the caller's arrays when those are pinned, through the pinned landing zone otherwise
  const bool stage_out = !(host_pinned(distances) && (!simplices || host_pinned(simplices)) &&
                           (!run_epa || (host_pinned(w1) && host_pinned(w2) && (!normals || host_pinned(normals)))));
  T *o_dist = distances, *o_w1 = w1, *o_w2 = w2, *o_nrm = normals;
  S* o_simp = simplices;
  if (stage_out) {
    const size_t bs = ((size_t)npairs * sizeof(S) + 255) & ~(size_t)255, bd = ((size_t)npairs * sizeof(T) + 255) & ~(size_t)255;
    const size_t bw = ((size_t)npairs * 3 * sizeof(T) + 255) & ~(size_t)255;
    if ((rc = ctx->h_out.reserve(bs + bd + 3 * bw))) return rc;
    char* z = static_cast<char*>(ctx->h_out.p);
    o_simp = simplices ? reinterpret_cast<S*>(z) : nullptr;
    o_dist = reinterpret_cast<T*>(z + bs);
    o_w1 = reinterpret_cast<T*>(z + bs + bd);
    o_w2 = reinterpret_cast<T*>(z + bs + bd + bw);
    o_nrm = normals ? reinterpret_cast<T*>(z + bs + bd + 2 * bw) : nullptr;
  }
  ctx->timed_mask = 0;
  if ((rc = store_build<T>(ctx, &ctx->scratch1, npairs, nverts1, dc1, do1, dk1, false, sc, false, host_sum(cnt1, npairs)))) return rc;
  if ((rc = store_build<T>(ctx, &ctx->scratch2, npairs, nverts2, dc2, do2, dk2, false, sc, false, host_sum(cnt2, npairs)))) return rc;
  // Every polytope of a chunk must lie inside the chunk's coordinate range: checked on
  // the device (optimistically -- the verdict is read after the pipeline has drained).
  if ((rc = ctx->counters.reserve(kCounterSlots * sizeof(unsigned long long)))) return rc;
  int* bad_layout = reinterpret_cast<int*>(static_cast<unsigned long long*>(ctx->counters.p) + kLayoutSlot);
  CU(cudaMemsetAsync(bad_layout, 0, sizeof(int), sc));
  {
    long long cb = (npairs + 255) / 256, ccap = (long long)ctx->sm_count * 8;
    const unsigned grid = (unsigned)(cb < ccap ? cb : ccap);
    check_monotone_kernel<<<grid, 256, 0, sc>>>(reinterpret_cast<const long long*>(do1), dk1, npairs, bad_layout);
    check_monotone_kernel<<<grid, 256, 0, sc>>>(reinterpret_cast<const long long*>(do2), dk2, npairs, bad_layout);
    ctx->launches += 2;
  }

  // H2D of chunk c is issued as soon as the caller says its host data is complete (struct-array
  // entry points gather chunk c+1 on the context's host threads while chunk c is in flight)
  auto h2d = [&](int c) -> int {
    if (ready) (*ready)(r[c].hi);
    CU(cudaMemcpyAsync(static_cast<T*>(ctx->d_coords1.p) + 3 * r[c].a1, coords1 + 3 * r[c].a1,
                       (size_t)(r[c].b1 - r[c].a1) * 3 * sizeof(T), cudaMemcpyHostToDevice, si));
    CU(cudaMemcpyAsync(static_cast<T*>(ctx->d_coords2.p) + 3 * r[c].a2, coords2 + 3 * r[c].a2,
                       (size_t)(r[c].b2 - r[c].a2) * 3 * sizeof(T), cudaMemcpyHostToDevice, si));
    CU(cudaEventRecord(ctx->chunk_in[c], si));
    return OGJK_OK;
  };
  if (!ready)
    for (int c = 0; c < K; ++c)
      if ((rc = h2d(c))) return rc;
  for (int c = 0; c < K; ++c) {
    const int64_t lo = r[c].lo, n = r[c].hi - r[c].lo;
    if (ready && (rc = h2d(c))) return rc;
    CU(cudaStreamWaitEvent(sc, ctx->chunk_in[c], 0));
    if ((rc = store_pack_range<T>(ctx, &ctx->scratch1, lo, r[c].hi, dc1, do1, sc))) return rc;
    if ((rc = store_pack_range<T>(ctx, &ctx->scratch2, lo, r[c].hi, dc2, do2, sc))) return rc;
    StoreArgs<T> sa;
    sa.s1 = view_of<T>(&ctx->scratch1); sa.s1.boff += lo; sa.s1.cnt += lo;
    sa.s2 = view_of<T>(&ctx->scratch2); sa.s2.boff += lo; sa.s2.cnt += lo;
    sa.idx1 = nullptr; sa.idx2 = nullptr;
    sa.simplices = static_cast<GkSimplex<T>*>(ctx->d_simp.p) + lo;
    sa.distances = static_cast<T*>(ctx->d_dist.p) + lo;
    sa.npairs = n;
    if ((rc = run_store_batch<T>(ctx, sa, sc))) return rc;
    if (run_epa) {
      EpaArgs<T> ea;
      ea.s1 = sa.s1; ea.s2 = sa.s2; ea.idx1 = nullptr; ea.idx2 = nullptr;
      ea.simplices = sa.simplices; ea.distances = sa.distances;
      ea.witness1 = static_cast<T*>(ctx->d_w1.p) + 3 * lo;
      ea.witness2 = static_cast<T*>(ctx->d_w2.p) + 3 * lo;
      ea.normals = normals ? static_cast<T*>(ctx->d_nrm.p) + 3 * lo : nullptr;
      ea.npairs = n;
      if ((rc = run_store_epa<T>(ctx, ea, sc))) return rc;
    }
    CU(cudaEventRecord(ctx->chunk_done[c], sc));
    CU(cudaStreamWaitEvent(so, ctx->chunk_done[c], 0));
    CU(cudaMemcpyAsync(o_dist + lo, static_cast<T*>(ctx->d_dist.p) + lo, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, so));
    if (simplices)
      CU(cudaMemcpyAsync(o_simp + lo, static_cast<S*>(ctx->d_simp.p) + lo, (size_t)n * sizeof(S), cudaMemcpyDeviceToHost, so));
    if (run_epa) {
      CU(cudaMemcpyAsync(o_w1 + 3 * lo, static_cast<T*>(ctx->d_w1.p) + 3 * lo, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, so));
      CU(cudaMemcpyAsync(o_w2 + 3 * lo, static_cast<T*>(ctx->d_w2.p) + 3 * lo, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, so));
      if (normals)
        CU(cudaMemcpyAsync(o_nrm + 3 * lo, static_cast<T*>(ctx->d_nrm.p) + 3 * lo, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, so));
    }
    if (stage_out) CU(cudaEventRecord(ctx->chunk_out[c], so));
  }
  if (stage_out) {
    // results landed in the pinned zone: hand every chunk to the caller's (pageable) arrays as soon as its
    // copy has finished, with the context's host threads, while later chunks are still in flight
    HostPool* pool = host_pool(ctx);
    for (int c = 0; c < K; ++c) {
      CU(cudaEventSynchronize(ctx->chunk_out[c]));
      const int64_t lo = r[c].lo, n = r[c].hi - r[c].lo;
      constexpr int64_t kPart = 16384;  // pairs per copy job
      const int parts = (int)((n + kPart - 1) / kPart);
      auto job = [&](int j) {
        const int64_t a0 = lo + (int64_t)j * kPart, m = (a0 + kPart < lo + n ? a0 + kPart : lo + n) - a0;
        stream_copy(distances + a0, o_dist + a0, (size_t)m * sizeof(T));
        if (simplices) stream_copy(simplices + a0, o_simp + a0, (size_t)m * sizeof(S));
        if (run_epa) {
          stream_copy(w1 + 3 * a0, o_w1 + 3 * a0, (size_t)m * 3 * sizeof(T));
          stream_copy(w2 + 3 * a0, o_w2 + 3 * a0, (size_t)m * 3 * sizeof(T));
          if (normals) stream_copy(normals + 3 * a0, o_nrm + 3 * a0, (size_t)m * 3 * sizeof(T));
        }
      };
      if (pool && parts > 1) pool->run(parts, job);
      else for (int j = 0; j < parts; ++j) job(j);
    }
  }
  int bad = 0;
  CU(cudaMemcpyAsync(&bad, bad_layout, sizeof(int), cudaMemcpyDeviceToHost, sc));
  CU(cudaStreamSynchronize(so));
  CU(cudaStreamSynchronize(si));
  if ((rc = pool_verdict(ctx, sc))) return rc;
  return bad ? 1 : OGJK_OK;  // 1: polytopes not in pair order -> caller reruns the one-shot path
}

// Pageable coordinate pools in the un-indexed form: a host-to-device copy from pageable memory
// goes through the driver's small bounce buffers and blocks the issuing thread, so the pools are
// copied into the pinned stage by the context's host threads, chunk by chunk in pipeline order,
// and the chunked pipeline uploads from there (`ready` gates each chunk).  Returns false when
// the layout is not chunkable or there are no host threads (the caller passes the arrays as is).
template <typename T>
bool stage_flat_inputs(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1, int64_t nverts1,
                       const T* coords2, const int64_t* off2, const int* cnt2, int64_t nverts2, const T** s1, const T** s2,
                       HostReady* ready) {
  const int K = ctx->pipeline_chunks > 16 ? 16 : ctx->pipeline_chunks;
  if (K < 2 || npairs < 4096 * (int64_t)K) return false;
  HostPool* pool = host_pool(ctx);
  if (!pool) return false;
  const size_t b1 = (((size_t)nverts1 * 3 * sizeof(T)) + 255) & ~(size_t)255, b2 = (size_t)nverts2 * 3 * sizeof(T);
  if (ctx->h_stage.reserve(b1 + b2)) { cudaGetLastError(); return false; }
  char* z = static_cast<char*>(ctx->h_stage.p);
  struct Piece { char* dst; const char* src; size_t bytes; };
  auto pieces = std::make_shared<std::vector<Piece>>();
  auto ends = std::make_shared<std::vector<std::pair<int64_t, int>>>();  // (pairs up to, jobs up to)
  constexpr size_t kPiece = (size_t)1 << 20;
  int64_t pa1 = 0, pa2 = 0;
  for (int c = 0; c < K; ++c) {
    const int64_t lo = npairs * c / K, hi = npairs * (c + 1) / K, last = hi - 1;
    const int64_t a1 = off1[lo], e1 = off1[last] + cnt1[last], a2 = off2[lo], e2 = off2[last] + cnt2[last];
    if (a1 < pa1 || a2 < pa2 || e1 < a1 || e2 < a2 || e1 > nverts1 || e2 > nverts2) return false;  // not chunkable
    pa1 = e1; pa2 = e2;
    const size_t o1 = (size_t)a1 * 3 * sizeof(T), n1 = (size_t)(e1 - a1) * 3 * sizeof(T);
    const size_t o2 = (size_t)a2 * 3 * sizeof(T), n2 = (size_t)(e2 - a2) * 3 * sizeof(T);
    for (size_t at = 0; at < n1; at += kPiece)
      pieces->push_back({z + o1 + at, reinterpret_cast<const char*>(coords1) + o1 + at, n1 - at < kPiece ? n1 - at : kPiece});
    for (size_t at = 0; at < n2; at += kPiece)
      pieces->push_back({z + b1 + o2 + at, reinterpret_cast<const char*>(coords2) + o2 + at, n2 - at < kPiece ? n2 - at : kPiece});
    ends->push_back({hi, (int)pieces->size()});
  }
  pool->start((int)pieces->size(), [pieces](int j) { const Piece& p = (*pieces)[(size_t)j]; stream_copy(p.dst, p.src, p.bytes); });
  *s1 = reinterpret_cast<const T*>(z);
  *s2 = reinterpret_cast<const T*>(z + b1);
  *ready = HostReady([pool, ends](int64_t upto) {
    int jobs = ends->empty() ? 0 : ends->back().second;
    for (const auto& e : *ends)
      if (e.first >= upto) { jobs = e.second; break; }
    pool->wait_prefix(jobs);
  });
  return true;
}

// Host-buffer pipeline shared by every host entry point: upload the pools, repack
// them into the scratch stores, run GJK and/or EPA, download.  With run_gjk ==
// false the caller's simplices/distances are inputs (the reference's compute_epa).
template <typename T, typename S>
int host_pipeline(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1,
                  int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2,
                  int64_t nverts2, const int* idx2, S* simplices, T* distances, T* w1, T* w2, T* normals, bool run_gjk,
                  bool run_epa, const HostReady* ready = nullptr) {
  if (!ctx) return fail(OGJK_ERR_ARG, "ctx is NULL%s");
  if (npairs < 0 || npoly1 < 0 || npoly2 < 0 || nverts1 < 0 || nverts2 < 0) return fail(OGJK_ERR_ARG, "negative count%s");
  if (npairs == 0) return OGJK_OK;
  if (!coords1 || !off1 || !cnt1 || !coords2 || !off2 || !cnt2 || !distances) return fail(OGJK_ERR_ARG, "NULL host pointer%s");
  if (run_epa && (!w1 || !w2)) return fail(OGJK_ERR_ARG, "EPA needs witness output arrays%s");
  if (!run_gjk && !simplices) return fail(OGJK_ERR_ARG, "EPA-only needs the GJK simplices%s");
  if (npoly1 < 1 || npoly2 < 1) return fail(OGJK_ERR_ARG, "empty polytope pool%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  const bool shared_pool = (coords1 == coords2 && off1 == off2 && cnt1 == cnt2);
  const bool need_simp = simplices != nullptr || run_epa;
  if (run_gjk && !idx1 && !idx2 && !shared_pool && npoly1 >= npairs && npoly2 >= npairs) {
    const T *c1 = coords1, *c2 = coords2;
    HostReady staged_ready;
    const bool staged = !ready && !(host_pinned(coords1) && host_pinned(coords2)) &&
                        stage_flat_inputs<T>(ctx, npairs, coords1, off1, cnt1, nverts1, coords2, off2, cnt2, nverts2, &c1, &c2,
                                             &staged_ready);
    const int rcc = host_pipeline_chunked<T, S>(ctx, npairs, c1, off1, cnt1, nverts1, c2, off2, cnt2, nverts2, simplices,
                                                distances, w1, w2, normals, run_epa, staged ? &staged_ready : ready);
    if (staged) staged_ready(npairs);  // the copy jobs read the caller's arrays: drained before returning
    if (rcc <= 0) return rcc;  // done (0) or error (<0); 1 = layout not chunkable, use the one-shot path
  }
  if (ready) (*ready)(npoly1 > npoly2 ? npoly1 : npoly2);  // one-shot path: everything must be in place
  int rc;
  if ((rc = upload(ctx, ctx->d_coords1, coords1, (size_t)nverts1 * 3 * sizeof(T)))) return rc;
  if ((rc = upload(ctx, ctx->d_off1, off1, (size_t)npoly1 * sizeof(int64_t)))) return rc;
  if ((rc = upload(ctx, ctx->d_cnt1, cnt1, (size_t)npoly1 * sizeof(int)))) return rc;
  if (idx1 && (rc = upload(ctx, ctx->d_idx1, idx1, (size_t)npairs * sizeof(int)))) return rc;
  if (!shared_pool) {
    if ((rc = upload(ctx, ctx->d_coords2, coords2, (size_t)nverts2 * 3 * sizeof(T)))) return rc;
    if ((rc = upload(ctx, ctx->d_off2, off2, (size_t)npoly2 * sizeof(int64_t)))) return rc;
    if ((rc = upload(ctx, ctx->d_cnt2, cnt2, (size_t)npoly2 * sizeof(int)))) return rc;
  }
  if (idx2 && (rc = upload(ctx, ctx->d_idx2, idx2, (size_t)npairs * sizeof(int)))) return rc;
  if ((rc = ctx->d_dist.reserve((size_t)npairs * sizeof(T)))) return rc;
  if (need_simp && (rc = ctx->d_simp.reserve((size_t)npairs * sizeof(S)))) return rc;
  if (!run_gjk) {
    CU(cudaMemcpyAsync(ctx->d_simp.p, simplices, (size_t)npairs * sizeof(S), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_dist.p, distances, (size_t)npairs * sizeof(T), cudaMemcpyHostToDevice, st));
  }
  ctx->timed_mask = 0;
  if ((rc = store_build<T>(ctx, &ctx->scratch1, npoly1, nverts1, static_cast<const T*>(ctx->d_coords1.p),
                           static_cast<const int64_t*>(ctx->d_off1.p), static_cast<const int*>(ctx->d_cnt1.p), false, st, true,
                           host_sum(cnt1, npoly1))))
    return rc;
  if (!shared_pool &&
      (rc = store_build<T>(ctx, &ctx->scratch2, npoly2, nverts2, static_cast<const T*>(ctx->d_coords2.p),
                           static_cast<const int64_t*>(ctx->d_off2.p), static_cast<const int*>(ctx->d_cnt2.p), false, st, true,
                           host_sum(cnt2, npoly2))))
    return rc;
  const int* di1 = idx1 ? static_cast<const int*>(ctx->d_idx1.p) : nullptr;
  const int* di2 = idx2 ? static_cast<const int*>(ctx->d_idx2.p) : nullptr;
  if (run_gjk) {
    StoreArgs<T> sa;
    sa.s1 = view_of<T>(&ctx->scratch1);
    sa.s2 = shared_pool ? sa.s1 : view_of<T>(&ctx->scratch2);
    sa.idx1 = di1; sa.idx2 = di2;
    sa.simplices = need_simp ? static_cast<GkSimplex<T>*>(ctx->d_simp.p) : nullptr;
    sa.distances = static_cast<T*>(ctx->d_dist.p);
    sa.npairs = npairs;
    if ((rc = run_store_batch<T>(ctx, sa, st))) return rc;
  }
  if (run_epa) {
    if ((rc = ctx->d_w1.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
    if ((rc = ctx->d_w2.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
    if (normals && (rc = ctx->d_nrm.reserve((size_t)npairs * 3 * sizeof(T)))) return rc;
    EpaArgs<T> ea;
    ea.s1 = view_of<T>(&ctx->scratch1);
    ea.s2 = shared_pool ? ea.s1 : view_of<T>(&ctx->scratch2);
    ea.idx1 = di1; ea.idx2 = di2;
    ea.simplices = static_cast<const GkSimplex<T>*>(ctx->d_simp.p);
    ea.distances = static_cast<T*>(ctx->d_dist.p);
    ea.witness1 = static_cast<T*>(ctx->d_w1.p);
    ea.witness2 = static_cast<T*>(ctx->d_w2.p);
    ea.normals = normals ? static_cast<T*>(ctx->d_nrm.p) : nullptr;
    ea.npairs = npairs;
    if ((rc = run_store_epa<T>(ctx, ea, st))) return rc;
    CU(cudaMemcpyAsync(w1, ctx->d_w1.p, (size_t)npairs * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(w2, ctx->d_w2.p, (size_t)npairs * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (normals) CU(cudaMemcpyAsync(normals, ctx->d_nrm.p, (size_t)npairs * 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(distances, ctx->d_dist.p, (size_t)npairs * sizeof(T), cudaMemcpyDeviceToHost, st));
  if (simplices && run_gjk)
    CU(cudaMemcpyAsync(simplices, ctx->d_simp.p, (size_t)npairs * sizeof(S), cudaMemcpyDeviceToHost, st));
  return pool_verdict(ctx, st);  // synchronises
}

template <typename T, typename S>
int gjk_host(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1, int64_t npoly1,
             int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2, const int* cnt2, int64_t npoly2,
             int64_t nverts2, const int* idx2, S* simplices, T* distances) {
  return host_pipeline<T, S>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2, npoly2, nverts2,
                             idx2, simplices, distances, nullptr, nullptr, nullptr, true, false);
}

HostPool* host_pool(ogjk_ctx* ctx) {
  if (!ctx->pool) {
    unsigned nt = ctx->host_threads > 0 ? (unsigned)ctx->host_threads : std::thread::hardware_concurrency();
    nt = nt > 32 ? 32 : (nt < 2 ? 2 : nt);
    ctx->pool = new (std::nothrow) HostPool((int)nt);
  }
  return ctx->pool;
}

// An array of reference-shaped polytope structs (each with its own host `coord` block) laid
// out as one pinned pool [coords | off | cnt] inside ctx->h_stage (the reference copies each
// block to the device with its own cudaMemcpy, gpu/opengjk_gpu.cu:3105-3166).
template <typename T> struct FlatSide { T* coords; int64_t* off; int* cnt; int64_t nverts; size_t bytes; };

// Pass 1 (two short parallel sweeps over the structs): counts and offsets of every polytope --
// the scheduler needs them for the whole batch before the first chunk leaves.  `sums` receives
// one partial vertex total per job of kGatherJob polytopes; returns the vertex total or -1 when a
// polytope has no vertices.
constexpr int kGatherJob = 8192;
template <typename P>
int64_t struct_totals(ogjk_ctx* ctx, int n, const P* bd, std::vector<int64_t>& sums) {
  const int njobs = (n + kGatherJob - 1) / kGatherJob;
  sums.assign((size_t)njobs, 0);
  std::atomic<int> bad{0};
  auto job = [&](int j) {
    const int lo = j * kGatherJob, hi = lo + kGatherJob < n ? lo + kGatherJob : n;
    int64_t t = 0;
    for (int i = lo; i < hi; ++i) {
      if (bd[i].numpoints < 1 || !bd[i].coord) { bad.store(1); return; }
      t += bd[i].numpoints;
    }
    sums[(size_t)j] = t;
  };
  HostPool* pool = host_pool(ctx);
  if (pool && njobs >= 4) pool->run(njobs, job);
  else for (int j = 0; j < njobs; ++j) job(j);
  if (bad.load()) return -1;
  int64_t total = 0;
  for (int j = 0; j < njobs; ++j) { const int64_t t = sums[(size_t)j]; sums[(size_t)j] = total; total += t; }  // exclusive
  return total;
}
inline size_t flat_bytes(int64_t total, int n, size_t tsize) {
  return (((((size_t)total * 3 * tsize + 15) & ~(size_t)15) + (size_t)n * 12 + 128) + 63) & ~(size_t)63;
}
template <typename T, typename P>
void flat_layout(ogjk_ctx* ctx, size_t base, int n, const P* bd, int64_t total, const std::vector<int64_t>& sums,
                 FlatSide<T>* out) {
  char* p = static_cast<char*>(ctx->h_stage.p) + base;
  const size_t cbytes = ((size_t)total * 3 * sizeof(T) + 15) & ~(size_t)15;
  out->coords = reinterpret_cast<T*>(p);
  out->off = reinterpret_cast<int64_t*>(p + cbytes);
  out->cnt = reinterpret_cast<int*>(p + cbytes + (size_t)n * sizeof(int64_t));
  out->nverts = total;
  out->bytes = ((cbytes + (size_t)n * (sizeof(int64_t) + sizeof(int))) + 63) & ~(size_t)63;
  const int njobs = (n + kGatherJob - 1) / kGatherJob;
  int64_t* off = out->off;
  int* cnt = out->cnt;
  auto job = [&, off, cnt](int j) {
    const int lo = j * kGatherJob, hi = lo + kGatherJob < n ? lo + kGatherJob : n;
    int64_t at = sums[(size_t)j];
    for (int i = lo; i < hi; ++i) { off[i] = at; cnt[i] = bd[i].numpoints; at += bd[i].numpoints; }
  };
  HostPool* pool = host_pool(ctx);
  if (pool && njobs >= 4) pool->run(njobs, job);
  else for (int j = 0; j < njobs; ++j) job(j);
}

// Pass 2: the coordinate gather of up to two sides, cut into jobs of kGatherJob polytopes in
// polytope order and started on the context's host threads.  Returns the HostReady callback
// that waits for a prefix of the polytopes.
template <typename T, typename P>
HostReady start_gather(ogjk_ctx* ctx, int n, const P* bd1, const FlatSide<T>& f1, const P* bd2, const FlatSide<T>* f2) {
  HostPool* pool = host_pool(ctx);
  const int njobs = (n + kGatherJob - 1) / kGatherJob;
  auto job = [=](int j) {
    const int lo = j * kGatherJob, hi = lo + kGatherJob < n ? lo + kGatherJob : n;
    // a job's polytopes are consecutive in the stage (off[i + 1] = off[i] + cnt[i])
    StreamOut o1(f1.coords + 3 * f1.off[lo]);
    for (int i = lo; i < hi; ++i) o1.put(bd1[i].coord, (size_t)bd1[i].numpoints * 3 * sizeof(T));
    o1.finish();
    if (f2) {
      StreamOut o2(f2->coords + 3 * f2->off[lo]);
      for (int i = lo; i < hi; ++i) o2.put(bd2[i].coord, (size_t)bd2[i].numpoints * 3 * sizeof(T));
      o2.finish();
    }
  };
  if (!pool || n < 2 * kGatherJob) {  // small batch (or no threads): gather inline
    for (int j = 0; j < njobs; ++j) job(j);
    return HostReady([](int64_t) {});
  }
  pool->start(njobs, job);
  return HostReady([pool](int64_t upto) { pool->wait_prefix((int)((upto + kGatherJob - 1) / kGatherJob)); });
}

// Lays both sides out in the pinned stage and starts the gather.  On success *ready waits for a
// prefix of the polytopes; the caller MUST drain it (ready(n)) before returning.
template <typename T, typename P>
int stage_structs(ogjk_ctx* ctx, int n, const P* bd1, const P* bd2, size_t extra, FlatSide<T>* f1, FlatSide<T>* f2,
                  HostReady* ready, size_t* extra_at) {
  std::vector<int64_t> s1, s2;
  const int64_t t1 = struct_totals(ctx, n, bd1, s1);
  const int64_t t2 = bd2 ? struct_totals(ctx, n, bd2, s2) : 0;
  if (t1 < 0 || t2 < 0) return fail(OGJK_ERR_ARG, "polytope with no vertices%s");
  const size_t b1 = flat_bytes(t1, n, sizeof(T)), b2 = bd2 ? flat_bytes(t2, n, sizeof(T)) : 0;
  int rc = ctx->h_stage.reserve(b1 + b2 + extra);
  if (rc) return rc;
  flat_layout<T, P>(ctx, 0, n, bd1, t1, s1, f1);
  if (bd2) flat_layout<T, P>(ctx, b1, n, bd2, t2, s2, f2);
  if (extra_at) *extra_at = b1 + b2;
  *ready = start_gather<T, P>(ctx, n, bd1, *f1, bd2, bd2 ? f2 : nullptr);
  return OGJK_OK;
}

template <typename T, typename P, typename S>
int compute_min_dist(ogjk_ctx* ctx, int n, const P* bd1, const P* bd2, S* simplices, T* distances) {
  if (n <= 0) return OGJK_OK;  // reference: silent return (gpu/opengjk_gpu.cu:2984)
  if (!ctx || !bd1 || !bd2 || !distances) return fail(OGJK_ERR_ARG, "NULL argument%s");
  CU(cudaSetDevice(ctx->device));
  FlatSide<T> f1, f2;
  HostReady ready;
  int rc = stage_structs<T, P>(ctx, n, bd1, bd2, 0, &f1, &f2, &ready, nullptr);
  if (rc) return rc;
  rc = host_pipeline<T, S>(ctx, n, f1.coords, f1.off, f1.cnt, n, f1.nverts, nullptr, f2.coords, f2.off, f2.cnt, n, f2.nverts,
                           nullptr, simplices, distances, nullptr, nullptr, nullptr, true, false, &ready);
  ready(n);  // never leave gather threads writing into the stage
  return rc;
}

// compute_gjk_epa (gpu/opengjk_gpu.cu:3055-3099) and compute_epa (:3007-3053) over struct arrays.
template <typename T, typename P, typename S>
int compute_epa_structs(ogjk_ctx* ctx, int n, const P* bd1, const P* bd2, S* simplices, T* distances, T* w1, T* w2,
                        T* normals, bool run_gjk) {
  if (n <= 0) return OGJK_OK;
  if (!ctx || !bd1 || !bd2 || !simplices || !distances || !w1 || !w2) return fail(OGJK_ERR_ARG, "NULL argument%s");
  CU(cudaSetDevice(ctx->device));
  FlatSide<T> f1, f2;
  HostReady ready;
  int rc = stage_structs<T, P>(ctx, n, bd1, bd2, 0, &f1, &f2, &ready, nullptr);
  if (rc) return rc;
  rc = host_pipeline<T, S>(ctx, n, f1.coords, f1.off, f1.cnt, n, f1.nverts, nullptr, f2.coords, f2.off, f2.cnt, n, f2.nverts,
                           nullptr, simplices, distances, w1, w2, normals, run_gjk, true, &ready);
  ready(n);
  return rc;
}

template <typename T, typename P, typename S>
int compute_min_dist_indexed(ogjk_ctx* ctx, int npoly, int npairs, const P* polys, const ogjk_pair* pairs, S* simplices,
                             T* distances) {
  if (npairs <= 0 || npoly <= 0) return OGJK_OK;  // reference: silent return (:3292)
  if (!ctx || !polys || !pairs || !distances) return fail(OGJK_ERR_ARG, "NULL argument%s");
  CU(cudaSetDevice(ctx->device));
  const size_t ib = ((size_t)npairs * sizeof(int) + 63) & ~(size_t)63;
  FlatSide<T> f, unused;
  HostReady ready;
  size_t at = 0;
  int rc = stage_structs<T, P>(ctx, npoly, polys, static_cast<const P*>(nullptr), 2 * ib, &f, &unused, &ready, &at);
  if (rc) return rc;
  int* i1 = reinterpret_cast<int*>(static_cast<char*>(ctx->h_stage.p) + at);
  int* i2 = reinterpret_cast<int*>(static_cast<char*>(ctx->h_stage.p) + at + ib);
  for (int p = 0; p < npairs; ++p) {
    if (pairs[p].idx1 < 0 || pairs[p].idx1 >= npoly || pairs[p].idx2 < 0 || pairs[p].idx2 >= npoly) {
      ready(npoly);
      return fail(OGJK_ERR_ARG, "pair index out of range%s");
    }
    i1[p] = pairs[p].idx1;
    i2[p] = pairs[p].idx2;
  }
  rc = host_pipeline<T, S>(ctx, npairs, f.coords, f.off, f.cnt, npoly, f.nverts, i1, f.coords, f.off, f.cnt, npoly, f.nverts, i2,
                           simplices, distances, nullptr, nullptr, nullptr, true, false, &ready);
  ready(npoly);
  return rc;
}

template <typename T, typename S>
int single_rows(ogjk_ctx* ctx, int n1, T* const* rows1, int n2, T* const* rows2, S* s, T* distance) {
  if (!ctx || !rows1 || !rows2 || !s || !distance) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (n1 < 1 || n2 < 1) return fail(OGJK_ERR_ARG, "polytope with no vertices%s");
  std::vector<T> c((size_t)(n1 + n2) * 3);
  for (int i = 0; i < n1; ++i) for (int t = 0; t < 3; ++t) c[3 * (size_t)i + t] = rows1[i][t];
  for (int i = 0; i < n2; ++i) for (int t = 0; t < 3; ++t) c[3 * (size_t)(n1 + i) + t] = rows2[i][t];
  int64_t off[2] = {0, n1};
  int cnt[2] = {n1, n2};
  int i1 = 0, i2 = 1;
  return gjk_host<T, S>(ctx, 1, c.data(), off, cnt, 2, n1 + n2, &i1, c.data(), off, cnt, 2, n1 + n2, &i2, s, distance);
}


template <typename T>
int store_create(ogjk_ctx* ctx, int64_t npoly, int64_t nverts, const T* coords, const int64_t* off, const int* cnt,
                 int on_device, void* stream, ogjk_store** out) {
  if (!ctx || !out) return fail(OGJK_ERR_ARG, "NULL argument%s");
  *out = nullptr;
  if (npoly < 1 || nverts < 1 || !coords || !off || !cnt) return fail(OGJK_ERR_ARG, "empty polytope pool%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ogjk_store_* s = new (std::nothrow) ogjk_store_();
  if (!s) return fail(OGJK_ERR_ALLOC, "out of host memory%s");
  int rc = OGJK_OK;
  if (on_device) {
    rc = store_build<T>(ctx, s, npoly, nverts, coords, off, cnt, true, st);
    if (!rc) rc = store_build_accel<T>(ctx, s, st);
    if (!rc) rc = pool_verdict(ctx, st);  // synchronises: the store is complete when the call returns
  } else {
    DevBuf dc, doff, dcnt;
    if (!(rc = dc.reserve((size_t)nverts * 3 * sizeof(T))) && !(rc = doff.reserve((size_t)npoly * sizeof(int64_t))) &&
        !(rc = dcnt.reserve((size_t)npoly * sizeof(int)))) {
      cudaError_t e = cudaMemcpyAsync(dc.p, coords, (size_t)nverts * 3 * sizeof(T), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(doff.p, off, (size_t)npoly * sizeof(int64_t), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dcnt.p, cnt, (size_t)npoly * sizeof(int), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "store upload failed: %s", cudaGetErrorString(e));
        rc = OGJK_ERR_CUDA;
      } else {
        rc = store_build<T>(ctx, s, npoly, nverts, static_cast<const T*>(dc.p), static_cast<const int64_t*>(doff.p),
                            static_cast<const int*>(dcnt.p), true, st);
        if (!rc) rc = store_build_accel<T>(ctx, s, st);
        if (!rc) rc = pool_verdict(ctx, st);  // synchronises
      }
    }
    dc.release(); doff.release(); dcnt.release();
  }
  if (rc) {
    s->release_all();
    delete s;
    return rc;
  }
  *out = s;
  return OGJK_OK;
}

template <typename T>
int store_aabbs(ogjk_ctx* ctx, const ogjk_store* s, T* d_boxes, void* stream) {
  if (!ctx || !s || !d_boxes) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (s->prec != (int)sizeof(T)) return fail(OGJK_ERR_ARG, "store precision mismatch%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (s->npoly * OGJK_AABB_LANES + 255) / 256, cap = (long long)ctx->sm_count * 64;
  aabb_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(view_of<T>(s), s->npoly, d_boxes);
  ctx->launches += 1;
  CU(cudaGetLastError());
  return OGJK_OK;
}

template <typename T>
int pairs_cull(ogjk_ctx* ctx, const T* box1, const T* box2, int64_t n, const int* idx1, const int* idx2, T margin,
               int64_t capacity, int* out1, int* out2, int64_t* d_count, void* stream) {
  if (!ctx || !box1 || !box2 || !d_count || n < 0 || capacity < 0) return fail(OGJK_ERR_ARG, "bad argument%s");
  if (n && (!idx1 || !idx2)) return fail(OGJK_ERR_ARG, "NULL candidate list%s");
  if (capacity && (!out1 || !out2)) return fail(OGJK_ERR_ARG, "NULL output list%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) { CU(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st)); return OGJK_OK; }
  const long long per_block = (long long)kCullBlock * kCullItems;
  const long long nblocks = (n + per_block - 1) / per_block;
  int rc = ctx->bp_tmp.reserve((size_t)nblocks * sizeof(long long));
  if (rc) return rc;
  auto* sums = static_cast<long long*>(ctx->bp_tmp.p);
  cull_count_kernel<T><<<(unsigned)nblocks, kCullBlock, 0, st>>>(box1, box2, idx1, idx2, n, margin, sums);
  scan_spine_kernel<<<1, 1024, 0, st>>>(sums, nblocks, reinterpret_cast<long long*>(d_count));
  cull_scatter_kernel<T><<<(unsigned)nblocks, kCullBlock, 0, st>>>(box1, box2, idx1, idx2, n, margin, sums, capacity, out1,
                                                                   out2);
  ctx->launches += 3;
  CU(cudaGetLastError());
  return OGJK_OK;
}

template <typename T>
int pairs_all(ogjk_ctx* ctx, const T* box, int64_t n, T margin, int64_t capacity, int* out1, int* out2, int64_t* d_count,
              void* stream) {
  if (!ctx || !box || !d_count || n < 0 || capacity < 0) return fail(OGJK_ERR_ARG, "bad argument%s");
  if (n > 0x7fffffffll) return fail(OGJK_ERR_ARG, "more than 2^31-1 polytopes%s");
  if (capacity && (!out1 || !out2)) return fail(OGJK_ERR_ARG, "NULL output list%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 2) { CU(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st)); return OGJK_OK; }
  int rc = ctx->bp_tmp.reserve((size_t)n * sizeof(long long));
  if (rc) return rc;
  auto* rows = static_cast<long long*>(ctx->bp_tmp.p);
  long long blocks = (n * 32 + 255) / 256, cap = (long long)ctx->sm_count * 8;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  allpairs_kernel<T, false><<<grid, 256, 0, st>>>(box, n, margin, rows, nullptr, 0, nullptr, nullptr);
  scan_spine_kernel<<<1, 1024, 0, st>>>(rows, n, reinterpret_cast<long long*>(d_count));
  allpairs_kernel<T, true><<<grid, 256, 0, st>>>(box, n, margin, nullptr, rows, capacity, out1, out2);
  ctx->launches += 3;
  CU(cudaGetLastError());
  return OGJK_OK;
}

template <typename T>
int intersect_flags(ogjk_ctx* ctx, int64_t n, const T* d_distances, unsigned char* d_flags, void* stream, T eps) {
  if (!ctx || n < 0 || (n && (!d_distances || !d_flags))) return fail(OGJK_ERR_ARG, "bad argument%s");
  if (n == 0) return OGJK_OK;
  CU(cudaSetDevice(ctx->device));
  long long blocks = (n + 255) / 256, cap = (long long)ctx->sm_count * 16;
  intersect_flags_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_distances, n, eps, d_flags);
  ctx->launches += 1;
  CU(cudaGetLastError());
  return OGJK_OK;
}
template <typename T, typename S>
int store_gjk(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2, int64_t npairs,
              S* simplices, T* distances, void* stream) {
  if (!ctx || !s1 || !s2 || !distances) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (npairs < 0) return fail(OGJK_ERR_ARG, "npairs < 0%s");
  if (npairs == 0) return OGJK_OK;
  if (s1->prec != (int)sizeof(T) || s2->prec != (int)sizeof(T)) return fail(OGJK_ERR_ARG, "store precision mismatch%s");
  if (s1->device != ctx->device || s2->device != ctx->device) return fail(OGJK_ERR_ARG, "store lives on another device%s");
  if ((!idx1 && npairs > s1->npoly) || (!idx2 && npairs > s2->npoly))
    return fail(OGJK_ERR_ARG, "npairs exceeds the pool size in un-indexed mode%s");
  CU(cudaSetDevice(ctx->device));
  ctx->timed_mask = 0;
  StoreArgs<T> sa;
  sa.s1 = view_of<T>(s1);
  sa.s2 = view_of<T>(s2);
  sa.idx1 = idx1; sa.idx2 = idx2;
  sa.simplices = reinterpret_cast<GkSimplex<T>*>(simplices);
  sa.distances = distances;
  sa.npairs = npairs;
  return run_store_batch<T>(ctx, sa, static_cast<cudaStream_t>(stream));
}

// Build a scratch store from a DEVICE array of gkPolytope structs.  The vertex total is
// not known on the host, so the data buffer is sized from the scanned total (one sync).
template <typename T>
int store_build_structs(ogjk_ctx* ctx, ogjk_store_* s, long long npoly, const void* d_bd, cudaStream_t st) {
  const auto* bd = static_cast<const DevPolytope<T>*>(d_bd);
  s->device = ctx->device;
  s->prec = (int)sizeof(T);
  s->npoly = npoly;
  int rc;
  if ((rc = s->boff.reserve((size_t)npoly * sizeof(long long)))) return rc;
  if ((rc = s->cnt.reserve((size_t)npoly * sizeof(int)))) return rc;
  const long long nblocks = (npoly + kScanBlock - 1) / kScanBlock;
  if ((rc = s->scan_tmp.reserve((size_t)(nblocks + 1) * sizeof(long long)))) return rc;
  int* cnt = static_cast<int*>(s->cnt.p);
  auto* sums = static_cast<long long*>(s->scan_tmp.p);
  auto* boff = static_cast<long long*>(s->boff.p);
  long long eb = (npoly + 255) / 256, cap = (long long)ctx->sm_count * 32;
  extract_counts_kernel<T><<<(unsigned)(eb < cap ? eb : cap), 256, 0, st>>>(bd, npoly, cnt);
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(cnt, npoly, sums, BlockElems{});
  scan_spine_kernel<<<1, 1024, 0, st>>>(sums, nblocks, sums + nblocks);
  scan_finish_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(cnt, npoly, sums, boff, BlockElems{});
  long long total = 0;
  CU(cudaMemcpyAsync(&total, sums + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((rc = s->data.reserve((size_t)(total > 0 ? total : 4) * sizeof(T)))) return rc;
  s->capacity = total;
  s->nverts = total / 3;
  long long pblocks = (npoly * 8 + 255) / 256;
  pack_ptr_kernel<T><<<(unsigned)(pblocks < cap ? pblocks : cap), 256, 0, st>>>(bd, npoly, boff, static_cast<T*>(s->data.p));
  ctx->launches += 5;
  CU(cudaGetLastError());
  return OGJK_OK;
}

// compute_minimum_distance_device / compute_epa_device over device gkPolytope arrays.
template <typename T, typename S>
int device_structs(ogjk_ctx* ctx, int64_t n, const void* d_bd1, const void* d_bd2, S* d_simplices, T* d_distances,
                   T* d_w1, T* d_w2, T* d_normals, bool run_gjk, bool run_epa, void* stream) {
  if (n <= 0) return OGJK_OK;
  if (!ctx || !d_bd1 || !d_bd2 || !d_simplices || !d_distances) return fail(OGJK_ERR_ARG, "NULL device pointer%s");
  if (run_epa && (!d_w1 || !d_w2)) return fail(OGJK_ERR_ARG, "EPA needs witness arrays%s");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ctx->timed_mask = 0;
  int rc;
  if ((rc = store_build_structs<T>(ctx, &ctx->scratch1, n, d_bd1, st))) return rc;
  if ((rc = store_build_structs<T>(ctx, &ctx->scratch2, n, d_bd2, st))) return rc;
  if (run_gjk) {
    StoreArgs<T> sa;
    sa.s1 = view_of<T>(&ctx->scratch1);
    sa.s2 = view_of<T>(&ctx->scratch2);
    sa.idx1 = nullptr; sa.idx2 = nullptr;
    sa.simplices = reinterpret_cast<GkSimplex<T>*>(d_simplices);
    sa.distances = d_distances;
    sa.npairs = n;
    if ((rc = run_store_batch<T>(ctx, sa, st))) return rc;
  }
  if (run_epa) {
    EpaArgs<T> ea;
    ea.s1 = view_of<T>(&ctx->scratch1);
    ea.s2 = view_of<T>(&ctx->scratch2);
    ea.idx1 = nullptr; ea.idx2 = nullptr;
    ea.simplices = reinterpret_cast<const GkSimplex<T>*>(d_simplices);
    ea.distances = d_distances;
    ea.witness1 = d_w1; ea.witness2 = d_w2; ea.normals = d_normals;
    ea.npairs = n;
    if ((rc = run_store_epa<T>(ctx, ea, st))) return rc;
  }
  return OGJK_OK;
}

template <typename T, typename S>
int test_subalgorithm(ogjk_ctx* ctx, S* simplex, T* v) {
  if (!ctx || !simplex || !v) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (simplex->nvrtx < 2 || simplex->nvrtx > 4) return fail(OGJK_ERR_ARG, "sub-solver needs 2..4 vertices%s");
  CU(cudaSetDevice(ctx->device));
  int rc = ctx->d_simp.reserve(sizeof(S));
  if (rc) return rc;
  if ((rc = ctx->d_dist.reserve(3 * sizeof(T)))) return rc;
  cudaStream_t st = ctx->own_stream;
  CU(cudaMemcpyAsync(ctx->d_simp.p, simplex, sizeof(S), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_dist.p, v, 3 * sizeof(T), cudaMemcpyHostToDevice, st));
  subsolver_kernel<T><<<1, 1, 0, st>>>(static_cast<GkSimplex<T>*>(ctx->d_simp.p), static_cast<T*>(ctx->d_dist.p));
  ctx->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(simplex, ctx->d_simp.p, sizeof(S), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(v, ctx->d_dist.p, 3 * sizeof(T), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return OGJK_OK;
}

template <typename T, typename S>
int store_epa(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2, int64_t npairs,
              const S* simplices, T* distances, T* w1, T* w2, T* normals, void* stream) {
  if (!ctx || !s1 || !s2 || !simplices || !distances || !w1 || !w2) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (npairs < 0) return fail(OGJK_ERR_ARG, "npairs < 0%s");
  if (npairs == 0) return OGJK_OK;
  if (s1->prec != (int)sizeof(T) || s2->prec != (int)sizeof(T)) return fail(OGJK_ERR_ARG, "store precision mismatch%s");
  if (s1->device != ctx->device || s2->device != ctx->device) return fail(OGJK_ERR_ARG, "store lives on another device%s");
  if ((!idx1 && npairs > s1->npoly) || (!idx2 && npairs > s2->npoly))
    return fail(OGJK_ERR_ARG, "npairs exceeds the pool size in un-indexed mode%s");
  CU(cudaSetDevice(ctx->device));
  EpaArgs<T> ea;
  ea.s1 = view_of<T>(s1);
  ea.s2 = view_of<T>(s2);
  ea.idx1 = idx1; ea.idx2 = idx2;
  ea.simplices = reinterpret_cast<const GkSimplex<T>*>(simplices);
  ea.distances = distances;
  ea.witness1 = w1; ea.witness2 = w2; ea.normals = normals;
  ea.npairs = npairs;
  return run_store_epa<T>(ctx, ea, static_cast<cudaStream_t>(stream));
}

}  // namespace

#include "multi.cuh"

extern "C" {

const char* ogjk_last_error(void) { return g_err; }
const char* ogjk_version(void) { return "opengjk_b200 0.1 (sm_100a)"; }

int ogjk_ctx_create(int device, ogjk_ctx** out) {
  if (!out) return fail(OGJK_ERR_ARG, "out is NULL%s");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device available (%s); this library has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return OGJK_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return fail(OGJK_ERR_ARG, "device ordinal out of range%s");
  CU(cudaSetDevice(device));
  ogjk_ctx* c = new (std::nothrow) ogjk_ctx_();
  if (!c) return fail(OGJK_ERR_ALLOC, "out of host memory%s");
  c->device = device;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 16 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&c->chunk_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->chunk_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->chunk_out[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess && c->counters.reserve(kCounterSlots * sizeof(unsigned long long)) != OGJK_OK) e = cudaErrorMemoryAllocation;
  if (e == cudaSuccess) e = cudaMemset(c->counters.p, 0, kCounterSlots * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    ogjk_ctx_destroy(c);
    snprintf(g_err, sizeof(g_err), "context setup failed: %s", cudaGetErrorString(e));
    return OGJK_ERR_CUDA;
  }
  *out = c;
  return OGJK_OK;
}

void ogjk_ctx_destroy(ogjk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  ogjk_store_* ss[] = {&c->scratch1, &c->scratch2};
  for (ogjk_store_* q : ss) { q->data.release(); q->boff.release(); q->cnt.release(); q->scan_tmp.release(); }
  c->sort_tmp.release(); c->sorted_list.release(); c->epa_list.release(); c->bp_tmp.release();
  for (int j = 0; j < 2; ++j) { c->pass_list[j].release(); c->pass_f[j].release(); c->pass_i[j].release(); }
  c->d_w1.release(); c->d_w2.release(); c->d_nrm.release();
  DevBuf* bufs[] = {&c->counters, &c->small_list, &c->large_list, &c->d_coords1, &c->d_coords2, &c->d_off1, &c->d_off2,
                    &c->d_cnt1,   &c->d_cnt2,     &c->d_idx1,     &c->d_idx2,    &c->d_simp,    &c->d_dist};
  for (DevBuf* b : bufs) b->release();
  c->h_stage.release();
  c->h_out.release();
  delete c->pool;
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->in_stream) cudaStreamDestroy(c->in_stream);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  for (int i = 0; i < 16; ++i) {
    if (c->chunk_in[i]) cudaEventDestroy(c->chunk_in[i]);
    if (c->chunk_done[i]) cudaEventDestroy(c->chunk_done[i]);
    if (c->chunk_out[i]) cudaEventDestroy(c->chunk_out[i]);
  }
  for (int i = 0; i < 2 * kPhases; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  delete c;
}

int ogjk_ctx_device(const ogjk_ctx* c) { return c ? c->device : -1; }
int ogjk_ctx_pool_status(ogjk_ctx* c) {
  if (!c) return fail(OGJK_ERR_ARG, "ctx is NULL%s");
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  return pool_verdict(c, c->own_stream);
}
int64_t ogjk_ctx_launch_count(const ogjk_ctx* c) { return c ? c->launches : 0; }
int ogjk_ctx_set_timing(ogjk_ctx* c, int enable) {
  if (!c) return fail(OGJK_ERR_ARG, "ctx is NULL%s");
  CU(cudaSetDevice(c->device));
  if (enable && !c->ev[0])
    for (int i = 0; i < 2 * kPhases; ++i) CU(cudaEventCreate(&c->ev[i]));
  c->timing = enable != 0;
  c->timed_mask = 0;
  return OGJK_OK;
}
int ogjk_ctx_last_timing(ogjk_ctx* c, float* ms8) {
  if (!c || !ms8) return fail(OGJK_ERR_ARG, "NULL argument%s");
  if (!c->timing) return fail(OGJK_ERR_ARG, "timing not enabled%s");
  for (int k = 0; k < kPhases; ++k) {
    ms8[k] = 0.f;
    if (c->timed_mask & (1 << k)) {
      CU(cudaEventSynchronize(c->ev[2 * k + 1]));
      CU(cudaEventElapsedTime(&ms8[k], c->ev[2 * k], c->ev[2 * k + 1]));
    }
  }
  return OGJK_OK;
}
int ogjk_ctx_set_classes(ogjk_ctx* c, int n, const int* hi_keys, const int* lanes) {
  if (!c || !hi_keys || !lanes || n < 1 || n > kMaxClasses) return fail(OGJK_ERR_ARG, "bad class table%s");
  int prev = 0;
  for (int i = 0; i < n; ++i) {
    const int l = lanes[i];
    // plain G or 0; 100 + G: shared-memory staged; 200 + G: persistent lanes (200 = register-resident);
    // 300 + G: scan-assist (G = 2, 4, 8, 16); 500 + W: CTA-per-pair with W = 2, 4, 8, 16 warps;
    // 600 + G: bulk-copy (TMA) staging, double buffered per warp (G = 2, 4, 8)
    const int fam = l / 100, gl = l % 100;
    const bool pow2 = gl == 1 || gl == 2 || gl == 4 || gl == 8 || gl == 16 || gl == 32;
    bool ok = false;
    if (fam == 0) ok = gl == 0 || pow2;
#ifdef OGJK_EXPERIMENTAL
    else if (fam == 1) ok = pow2;
    else if (fam == 2) ok = gl == 0 || pow2;
    else if (fam == 3 || fam == 5) ok = gl == 2 || gl == 4 || gl == 8 || gl == 16;
    else if (fam == 6) ok = gl == 2 || gl == 4 || gl == 8;
#else
    else if (l >= 0 && (fam == 1 || fam == 2 || fam == 3 || fam == 5 || fam == 6))
      return fail(OGJK_ERR_ARG, "experimental kernel flavour: only libopengjk_b200_exp.so (-DOGJK_EXPERIMENTAL) carries it%s");
#endif
    if (l < 0 || !ok) return fail(OGJK_ERR_ARG, "unknown kernel flavour in class table%s");
    // the last class is extended to the largest key below, so it is validated against that
    const int hi = i == n - 1 ? kMaxKey : hi_keys[i];
    if (gl == 0 && hi > kSmallKeyMax)
      return fail(OGJK_ERR_ARG, "register-resident class is limited to key <= 4 and cannot be the last (open-ended) class%s");
    if (hi_keys[i] <= prev) return fail(OGJK_ERR_ARG, "class keys must increase%s");
    prev = hi_keys[i];
  }
  c->nclasses = n;
  for (int i = 0; i < n; ++i) { c->class_hi[i] = hi_keys[i]; c->class_lanes[i] = lanes[i]; }
  c->class_hi[n - 1] = kMaxKey;  // the last class always takes everything above
  return OGJK_OK;
}
int ogjk_ctx_set_warp_threshold(ogjk_ctx* c, int nverts) {
  if (!c || nverts < 2) return fail(OGJK_ERR_ARG, "bad threshold%s");
  c->warp_threshold = nverts;
  return OGJK_OK;
}

int ogjk_gjk_device_f32(ogjk_ctx* ctx, int64_t npairs, const float* coords1, const int64_t* off1, const int* cnt1,
                        int64_t npoly1, int64_t nverts1, const int* idx1, const float* coords2, const int64_t* off2,
                        const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f32* simplices,
                        float* distances, int mode, void* stream) {
  return gjk_device<float, ogjk_simplex_f32>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2,
                                             npoly2, nverts2, idx2, simplices, distances, mode, stream);
}
int ogjk_gjk_device_f64(ogjk_ctx* ctx, int64_t npairs, const double* coords1, const int64_t* off1, const int* cnt1,
                        int64_t npoly1, int64_t nverts1, const int* idx1, const double* coords2, const int64_t* off2,
                        const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f64* simplices,
                        double* distances, int mode, void* stream) {
  return gjk_device<double, ogjk_simplex_f64>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2,
                                              npoly2, nverts2, idx2, simplices, distances, mode, stream);
}

// ---- packed polytope store ------------------------------------------------------

#define OGJK_STORE_IMPL(SFX, T, S)                                                                                   \
  int ogjk_store_create_##SFX(ogjk_ctx* ctx, int64_t npoly, int64_t nverts, const T* coords, const int64_t* off,        \
                              const int* cnt, int on_device, void* stream, ogjk_store** out) {                           \
    return store_create<T>(ctx, npoly, nverts, coords, off, cnt, on_device, stream, out);                               \
  }                                                                                                                  \
  int ogjk_store_gjk_##SFX(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,  \
                           int64_t npairs, S* simplices, T* distances, void* stream) {                                  \
    return store_gjk<T, S>(ctx, s1, idx1, s2, idx2, npairs, simplices, distances, stream);                              \
  }
OGJK_STORE_IMPL(f32, float, ogjk_simplex_f32)
OGJK_STORE_IMPL(f64, double, ogjk_simplex_f64)

void ogjk_store_destroy(ogjk_store* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  s->release_all();
  delete s;
}
int64_t ogjk_store_bytes(const ogjk_store* s) {
  return s ? (int64_t)((s->capacity + s->accel_elems) * s->prec + s->npoly * (s->accel_elems ? 20 : 12)) : 0;
}
int64_t ogjk_store_num_polytopes(const ogjk_store* s) { return s ? s->npoly : 0; }
int64_t ogjk_store_table_bytes(const ogjk_store* s) { return s ? (int64_t)(s->accel_elems * s->prec) : 0; }

int ogjk_gjk_host_f32(ogjk_ctx* ctx, int64_t npairs, const float* coords1, const int64_t* off1, const int* cnt1,
                      int64_t npoly1, int64_t nverts1, const int* idx1, const float* coords2, const int64_t* off2,
                      const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f32* simplices,
                      float* distances) {
  return gjk_host<float, ogjk_simplex_f32>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2,
                                           npoly2, nverts2, idx2, simplices, distances);
}
int ogjk_gjk_host_f64(ogjk_ctx* ctx, int64_t npairs, const double* coords1, const int64_t* off1, const int* cnt1,
                      int64_t npoly1, int64_t nverts1, const int* idx1, const double* coords2, const int64_t* off2,
                      const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, ogjk_simplex_f64* simplices,
                      double* distances) {
  return gjk_host<double, ogjk_simplex_f64>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2,
                                            npoly2, nverts2, idx2, simplices, distances);
}

int ogjk_compute_minimum_distance_f32(ogjk_ctx* ctx, int n, const ogjk_polytope_f32* bd1, const ogjk_polytope_f32* bd2,
                                      ogjk_simplex_f32* simplices, float* distances) {
  return compute_min_dist<float, ogjk_polytope_f32, ogjk_simplex_f32>(ctx, n, bd1, bd2, simplices, distances);
}
int ogjk_compute_minimum_distance_f64(ogjk_ctx* ctx, int n, const ogjk_polytope_f64* bd1, const ogjk_polytope_f64* bd2,
                                      ogjk_simplex_f64* simplices, double* distances) {
  return compute_min_dist<double, ogjk_polytope_f64, ogjk_simplex_f64>(ctx, n, bd1, bd2, simplices, distances);
}
int ogjk_compute_minimum_distance_indexed_f32(ogjk_ctx* ctx, int num_polytopes, int num_pairs,
                                              const ogjk_polytope_f32* polytopes, const ogjk_pair* pairs,
                                              ogjk_simplex_f32* simplices, float* distances) {
  return compute_min_dist_indexed<float, ogjk_polytope_f32, ogjk_simplex_f32>(ctx, num_polytopes, num_pairs, polytopes,
                                                                              pairs, simplices, distances);
}
int ogjk_compute_minimum_distance_indexed_f64(ogjk_ctx* ctx, int num_polytopes, int num_pairs,
                                              const ogjk_polytope_f64* polytopes, const ogjk_pair* pairs,
                                              ogjk_simplex_f64* simplices, double* distances) {
  return compute_min_dist_indexed<double, ogjk_polytope_f64, ogjk_simplex_f64>(ctx, num_polytopes, num_pairs, polytopes,
                                                                               pairs, simplices, distances);
}

int ogjk_single_rows_f32(ogjk_ctx* ctx, int n1, float* const* rows1, int n2, float* const* rows2, ogjk_simplex_f32* s,
                         float* distance) {
  return single_rows<float, ogjk_simplex_f32>(ctx, n1, rows1, n2, rows2, s, distance);
}
int ogjk_single_rows_f64(ogjk_ctx* ctx, int n1, double* const* rows1, int n2, double* const* rows2,
                         ogjk_simplex_f64* s, double* distance) {
  return single_rows<double, ogjk_simplex_f64>(ctx, n1, rows1, n2, rows2, s, distance);
}

#define OGJK_EPA_IMPL(SFX, T, P, S)                                                                                    \
  int ogjk_store_epa_##SFX(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2, const int* idx2,   \
                           int64_t npairs, const S* simplices, T* distances, T* witness1, T* witness2, T* normals,       \
                           void* stream) {                                                                              \
    return store_epa<T, S>(ctx, s1, idx1, s2, idx2, npairs, simplices, distances, witness1, witness2, normals, stream);  \
  }                                                                                                                    \
  int ogjk_store_gjk_epa_##SFX(ogjk_ctx* ctx, const ogjk_store* s1, const int* idx1, const ogjk_store* s2,                \
                               const int* idx2, int64_t npairs, S* simplices, T* distances, T* witness1, T* witness2,    \
                               T* normals, void* stream) {                                                              \
    if (!simplices) return fail(OGJK_ERR_ARG, "GJK+EPA needs the simplex array%s");                                     \
    int rc = store_gjk<T, S>(ctx, s1, idx1, s2, idx2, npairs, simplices, distances, stream);                            \
    if (rc) return rc;                                                                                                 \
    return store_epa<T, S>(ctx, s1, idx1, s2, idx2, npairs, simplices, distances, witness1, witness2, normals, stream);  \
  }                                                                                                                    \
  int ogjk_gjk_epa_host_##SFX(ogjk_ctx* ctx, int64_t npairs, const T* coords1, const int64_t* off1, const int* cnt1,      \
                              int64_t npoly1, int64_t nverts1, const int* idx1, const T* coords2, const int64_t* off2,   \
                              const int* cnt2, int64_t npoly2, int64_t nverts2, const int* idx2, S* simplices,           \
                              T* distances, T* witness1, T* witness2, T* normals) {                                      \
    return host_pipeline<T, S>(ctx, npairs, coords1, off1, cnt1, npoly1, nverts1, idx1, coords2, off2, cnt2, npoly2,     \
                               nverts2, idx2, simplices, distances, witness1, witness2, normals, true, true);           \
  }                                                                                                                    \
  int ogjk_compute_gjk_epa_##SFX(ogjk_ctx* ctx, int n, const P* bd1, const P* bd2, S* simplices, T* distances,            \
                                 T* witness1, T* witness2) {                                                            \
    return compute_epa_structs<T, P, S>(ctx, n, bd1, bd2, simplices, distances, witness1, witness2, nullptr, true);     \
  }                                                                                                                    \
  int ogjk_compute_epa_##SFX(ogjk_ctx* ctx, int n, const P* bd1, const P* bd2, S* simplices, T* distances, T* witness1,   \
                             T* witness2, T* contact_normals) {                                                         \
    return compute_epa_structs<T, P, S>(ctx, n, bd1, bd2, simplices, distances, witness1, witness2, contact_normals,    \
                                        false);                                                                         \
  }
OGJK_EPA_IMPL(f32, float, ogjk_polytope_f32, ogjk_simplex_f32)
OGJK_EPA_IMPL(f64, double, ogjk_polytope_f64, ogjk_simplex_f64)

int ogjk_gjk_device_structs_f32(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f32* d_bd1, const ogjk_polytope_f32* d_bd2,
                                ogjk_simplex_f32* d_simplices, float* d_distances, void* stream) {
  return device_structs<float, ogjk_simplex_f32>(ctx, n, d_bd1, d_bd2, d_simplices, d_distances, nullptr, nullptr, nullptr,
                                                 true, false, stream);
}
int ogjk_gjk_device_structs_f64(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f64* d_bd1, const ogjk_polytope_f64* d_bd2,
                                ogjk_simplex_f64* d_simplices, double* d_distances, void* stream) {
  return device_structs<double, ogjk_simplex_f64>(ctx, n, d_bd1, d_bd2, d_simplices, d_distances, nullptr, nullptr, nullptr,
                                                  true, false, stream);
}
int ogjk_epa_device_structs_f32(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f32* d_bd1, const ogjk_polytope_f32* d_bd2,
                                ogjk_simplex_f32* d_simplices, float* d_distances, float* d_w1, float* d_w2,
                                float* d_normals, void* stream) {
  return device_structs<float, ogjk_simplex_f32>(ctx, n, d_bd1, d_bd2, d_simplices, d_distances, d_w1, d_w2, d_normals, false,
                                                 true, stream);
}
int ogjk_epa_device_structs_f64(ogjk_ctx* ctx, int64_t n, const ogjk_polytope_f64* d_bd1, const ogjk_polytope_f64* d_bd2,
                                ogjk_simplex_f64* d_simplices, double* d_distances, double* d_w1, double* d_w2,
                                double* d_normals, void* stream) {
  return device_structs<double, ogjk_simplex_f64>(ctx, n, d_bd1, d_bd2, d_simplices, d_distances, d_w1, d_w2, d_normals,
                                                  false, true, stream);
}

int ogjk_intersect_flags_f32(ogjk_ctx* ctx, int64_t n, const float* d_distances, unsigned char* d_flags, void* stream) {
  return intersect_flags<float>(ctx, n, d_distances, d_flags, stream, 1.1920928955078125e-07f);   // FLT_EPSILON = gkEpsilon
}
int ogjk_intersect_flags_f64(ogjk_ctx* ctx, int64_t n, const double* d_distances, unsigned char* d_flags, void* stream) {
  return intersect_flags<double>(ctx, n, d_distances, d_flags, stream, 2.220446049250313e-16);   // DBL_EPSILON = gkEpsilon
}

// ---- broad-phase-adjacent pair production (broadphase.cuh) -------------------------------
#define OGJK_BROADPHASE_IMPL(SFX, T)                                                                                    \
  int ogjk_store_aabbs_##SFX(ogjk_ctx* ctx, const ogjk_store* s, T* d_boxes, void* stream) {                            \
    return store_aabbs<T>(ctx, s, d_boxes, stream);                                                                    \
  }                                                                                                                    \
  int ogjk_pairs_cull_##SFX(ogjk_ctx* ctx, const T* d_boxes1, const T* d_boxes2, int64_t ncand, const int* d_idx1,      \
                            const int* d_idx2, T margin, int64_t capacity, int* d_out1, int* d_out2, int64_t* d_count, \
                            void* stream) {                                                                            \
    return pairs_cull<T>(ctx, d_boxes1, d_boxes2, ncand, d_idx1, d_idx2, margin, capacity, d_out1, d_out2, d_count,    \
                         stream);                                                                                      \
  }                                                                                                                    \
  int ogjk_pairs_all_##SFX(ogjk_ctx* ctx, const T* d_boxes, int64_t npoly, T margin, int64_t capacity, int* d_out1,     \
                           int* d_out2, int64_t* d_count, void* stream) {                                              \
    return pairs_all<T>(ctx, d_boxes, npoly, margin, capacity, d_out1, d_out2, d_count, stream);                       \
  }
OGJK_BROADPHASE_IMPL(f32, float)
OGJK_BROADPHASE_IMPL(f64, double)

int ogjk_ctx_set_accel_min(ogjk_ctx* c, int min_vertices) {
  if (!c || min_vertices < 0) return fail(OGJK_ERR_ARG, "min_vertices must be >= 0%s");
  c->accel_min = (min_vertices > 0 && min_vertices < 64) ? 64 : min_vertices;
  return OGJK_OK;
}

int ogjk_ctx_set_passes(ogjk_ctx* c, int passes, int iters1, int iters2) {
  if (!c || passes < 1 || passes > 3 || iters1 < 1 || iters2 < 1 || iters1 + iters2 >= kMaxIter)
    return fail(OGJK_ERR_ARG, "passes must be 1..3 with 1 <= iters1, iters2 and iters1 + iters2 < 25%s");
#ifndef OGJK_EXPERIMENTAL
  if (passes > 1) return fail(OGJK_ERR_ARG, "multi-pass scheduling is compiled into libopengjk_b200_exp.so only%s");
#endif
  c->passes = passes;
  c->pass_iters[0] = iters1;
  c->pass_iters[1] = iters2;
  return OGJK_OK;
}

int ogjk_ctx_set_table_lanes(ogjk_ctx* c, int lanes) {
  if (!c || (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 3 && lanes != 8 && lanes != 16 && lanes != 32))
    return fail(OGJK_ERR_ARG, "table lanes must be 0 (default), 1, 2, 3, 8, 16 or 32%s");
  c->table_lanes = lanes;
  return OGJK_OK;
}

int ogjk_ctx_set_regroup(ogjk_ctx* c, int iters1, int iters2) {
  if (!c || iters1 < 0 || iters2 < 0 || (iters1 > 0 && iters2 > 0 && iters2 <= iters1) || iters1 > kMaxIter || iters2 > kMaxIter)
    return fail(OGJK_ERR_ARG, "regroup points must be 0 (off) or 1 <= iters1 < iters2 <= 25%s");
#ifndef OGJK_EXPERIMENTAL
  if (iters1 > 0) return fail(OGJK_ERR_ARG, "in-CTA regrouping lives in libopengjk_b200_exp.so (measured slower, see DESIGN.md)%s");
#endif
  c->regroup[0] = iters1;
  c->regroup[1] = iters1 > 0 ? (iters2 > 0 ? iters2 : kMaxIter + 1) : 0;
  return OGJK_OK;
}

int ogjk_ctx_set_pipeline_chunks(ogjk_ctx* c, int chunks) {
  if (!c || chunks < 0 || chunks > 16) return fail(OGJK_ERR_ARG, "chunks must be 0..16%s");
  c->pipeline_chunks = chunks;
  return OGJK_OK;
}

int ogjk_test_subalgorithm_f32(ogjk_ctx* ctx, ogjk_simplex_f32* simplex, float* v) {
  return test_subalgorithm<float, ogjk_simplex_f32>(ctx, simplex, v);
}
int ogjk_test_subalgorithm_f64(ogjk_ctx* ctx, ogjk_simplex_f64* simplex, double* v) {
  return test_subalgorithm<double, ogjk_simplex_f64>(ctx, simplex, v);
}

#ifdef OGJK_PHASE_CLOCKS
// Debug build only: copy out and clear the phase-cycle counters (tools/phase_clocks.py).
OGJK_API int ogjk_debug_phase_clocks(unsigned long long* out8) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out8, ogjk::g_phase_clk, 8 * sizeof(unsigned long long)));
  unsigned long long zero[8] = {0};
  CU(cudaMemcpyToSymbol(ogjk::g_phase_clk, zero, sizeof(zero)));
  return OGJK_OK;
}
#endif

}  // extern "C"
