// runtime.cu - the C-ABI of include/tpp_xsmm_abi.h on top of the sm_100a kernels.
//
// B200-native counterpart of runtime/Xsmm/XsmmRunnerUtils.cpp (the libxsmm shim)
// and runtime/PerfRunnerUtils.cpp. Dispatch validates the shape, picks a kernel
// family + tile configuration and returns the address of an immortal KernelDesc;
// invoke resolves the operands to device memory and launches. There is no CPU
// execution path in this file: every invoke ends in a CUDA kernel launch, and a
// process without a usable sm_100 device dies in the first dispatch.
#include "tpp_xsmm_abi.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "kernel_desc.h"
#include "kernels.h"

using namespace tpp;

namespace {

// ---- process-wide state -----------------------------------------------------
std::atomic<int64_t> g_launches{0};
std::atomic<bool> g_cuda_ready{false};
std::mutex g_init_mutex;

void fail(const char *what) {
  fprintf(stderr, "tpp-xsmm-cuda: %s\n", what);
  exit(-1);
}

// First CUDA use: there must be an sm_100 device. No fallback of any kind.
void ensure_cuda() {
  if (g_cuda_ready.load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (g_cuda_ready.load()) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr,
            "tpp-xsmm-cuda: no CUDA device available (%s). This backend has no CPU path; "
            "run on a B200 (sm_100a).\n",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    exit(-1);
  }
  int dev = 0;
  if (const char *env = getenv("TPP_XSMM_DEVICE")) {
    dev = atoi(env);
    TPP_CUDA_CHECK(cudaSetDevice(dev));
  } else {
    TPP_CUDA_CHECK(cudaGetDevice(&dev));
  }
  cudaDeviceProp prop;
  TPP_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    fprintf(stderr, "tpp-xsmm-cuda: device %d (%s) is sm_%d%d; kernels are built for sm_100a only\n", dev, prop.name,
            prop.major, prop.minor);
    exit(-1);
  }
  g_cuda_ready.store(true, std::memory_order_release);
}

// ---- per-thread execution context ---------------------------------------------
struct Staging {
  void *ptr = nullptr;
  size_t cap = 0;
  void *get(size_t bytes) {
    if (bytes > cap) {
      if (ptr) TPP_CUDA_CHECK(cudaFree(ptr));
      size_t want = bytes < (1u << 20) ? (1u << 20) : bytes + bytes / 4;
      TPP_CUDA_CHECK(cudaMalloc(&ptr, want));
      cap = want;
    }
    return ptr;
  }
};

struct PendingGemm {
  const KernelDesc *d;
  GemmArgs g;
};

// a data-movement unary invoke (identity copy or transpose of one tile) recorded during graph capture
struct PendingTile {
  const KernelDesc *d;
  char *in, *out;
};

constexpr size_t kLazyQueueMax = 16384;   // lazy mode: invokes queued before a flush is forced
constexpr int kDownloadRing = 64;   // downloads xsmm_cuda_wait_host can still find (callers fall back to a stream sync)

struct ThreadCtx {
  // BRGEMM invokes recorded during graph capture and not launched yet: consecutive layers whose C is the next
  // layer's A are fused into one persistent kernel when the sequence is flushed (see flush_pending)
  std::vector<PendingGemm> pending;
  // run of tile moves with ONE descriptor and no hazards among them (a tensor.pack / unpack lowered tile by tile):
  // launched as one batched kernel by flush_tiles(). At most one of the two pending lists is non-empty.
  std::vector<PendingTile> pending_tiles;
  // the run's source / destination rectangles by first byte (every tile of a run has the same extents): a new tile is
  // tested only against the tiles whose byte range can reach its own - O(tiles of one row block) instead of O(run)
  std::multimap<const char *, uint32_t> tiles_by_in, tiles_by_out;
  // a unary zero(C) recorded during capture and not launched yet (runtime CombineXsmmOp: see pending_producer_of)
  struct HeldZero { const KernelDesc *d = nullptr; char *out = nullptr; } held_zero;
  std::vector<void *> capture_tables;   // device tables baked into the graph being captured (freed with it)
  std::vector<std::pair<cudaStream_t, Staging *>> vnni_scratch;   // VNNI-2 un-interleave scratch, one per stream
  cudaStream_t stream = nullptr; // legacy default stream unless xsmm_cuda_set_stream was called
  int device = -1;               // device this thread last launched on
  const char *last_kernel = "";
  Staging stage[4];
  // lazy mode (xsmm_cuda_set_lazy / TPP_XSMM_LAZY=1): outside a capture BRGEMM / tile-move invokes on device operands are
  // queued exactly as during a capture and launched - folded into layers, chained, batched - at the next flush point:
  // any sync entry point, a host-visible operation of this ABI, an invoke that cannot be queued, a full queue. An
  // unmodified invoke loop (no graph calls) then gets the fused kernels. The caller must drain through this ABI
  // (xsmm_cuda_sync / perf_stop_timer / xsmm_cuda_update_host ...) before it reads or frees operands by other means.
  bool lazy = false;
  bool lazy_init = false;
  bool recording() {
    if (!lazy_init) {
      lazy_init = true;
      const char *e = getenv("TPP_XSMM_LAZY");
      lazy = e && e[0] == '1';
    }
    return capturing || lazy;
  }
  // device tables of lazily launched fused kernels: freed once the event recorded after their launch has completed
  struct Retired { cudaEvent_t ev = nullptr; std::vector<void *> ptrs; };
  std::vector<Retired> retired;
  // graph capture (xsmm_cuda_graph_begin/end)
  bool capturing = false;
  cudaStream_t capture_stream = nullptr;   // internal stream used when the thread is on the legacy stream
  cudaStream_t saved_stream = nullptr;
  int64_t captured_launches = 0;
  // device ranges written by the last few kernels of this thread (dependency tracking for PDL)
  // kPdlWindow entries: every kernel that can still be running when a new one starts its prologue is in here, because
  // after kPdlWindow programmatic launches in a row the next one is issued in plain stream order (pdl_run)
  struct Range { const char *lo = nullptr, *hi = nullptr; } recent_out[kPdlWindow];
  int recent_pos = 0;
  int pdl_run = 0;               // programmatic launches since the last launch in plain stream order
  void note_output(const char *lo, size_t bytes) {
    recent_out[recent_pos % kPdlWindow] = {lo, lo + bytes};
    ++recent_pos;
  }
  bool recently_written(const char *lo, size_t bytes) const {
    for (const Range &r : recent_out)
      if (r.lo && lo < r.hi && r.lo < lo + bytes) return true;
    return false;
  }
  // asynchronous residency (xsmm_cuda_upload_async / download_async): copies run on two side streams so that
  // they overlap kernels; events order them against the compute stream
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  cudaEvent_t ev_up = nullptr, ev_compute = nullptr, ev_fork = nullptr, ev_join = nullptr;
  bool up_pending = false;       // the compute stream must wait for ev_up before its next launch
  bool up_in_capture = false, down_in_capture = false;   // side streams forked into the running capture
  struct Download { void *host = nullptr; cudaEvent_t ev = nullptr; } downloads[kDownloadRing];
  int download_pos = 0;
};

struct GraphHandle {
  uint32_t magic = 0x47525048u; // "GRPH"
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;         // kernels per replay (for xsmm_cuda_launch_count)
  char last_kernel[96] = {0};   // name of the last kernel captured (what xsmm_cuda_last_kernel reports after a replay)
  std::vector<void *> device_allocs;   // descriptor tables the captured kernels read (freed with the graph)
};
thread_local ThreadCtx t_ctx;

inline void count_launch() {
  if (t_ctx.capturing) ++t_ctx.captured_launches;   // recorded, not executed yet
  else g_launches.fetch_add(1, std::memory_order_relaxed);
}

// NVTX range per ABI entry (TPP_XSMM_NVTX=1; off by default: one predictable branch per invoke). The range names are the
// invoke kinds, so `ncu --nvtx --nvtx-include "xsmm_fused_brgemm_invoke/"` (or an nsys timeline) selects the launches of
// one kind of TPP; without a profiler attached the NVTX calls are no-ops of the header-only library.
struct NvtxRange {
  static bool enabled() {
    static const bool on = [] { const char *e = getenv("TPP_XSMM_NVTX"); return e && e[0] == '1'; }();
    return on;
  }
  const bool on;
  explicit NvtxRange(const char *name) : on(enabled()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

// ---- registered host ranges -> device mirrors -----------------------------------
struct Mirror {
  char *host;
  size_t bytes;
  char *dev;
  bool pinned_here;
};
std::shared_mutex g_mirror_mutex;
std::map<uintptr_t, Mirror> g_mirrors; // keyed by host start
std::atomic<int> g_mirror_count{0};

bool find_mirror(const void *p, Mirror *out) {
  if (g_mirror_count.load(std::memory_order_relaxed) == 0) return false;
  std::shared_lock<std::shared_mutex> lock(g_mirror_mutex);
  auto it = g_mirrors.upper_bound(reinterpret_cast<uintptr_t>(p));
  if (it == g_mirrors.begin()) return false;
  --it;
  const Mirror &m = it->second;
  if (reinterpret_cast<const char *>(p) >= m.host && reinterpret_cast<const char *>(p) < m.host + m.bytes) {
    *out = m;
    return true;
  }
  return false;
}

enum class Where { Device, HostMirrored, HostPlain };

struct Resolved {
  Where where;
  char *dev;    // device address of the element the host/device pointer designates (Device / HostMirrored)
  int device;   // owning device ordinal, -1 if unknown
};

// Classification of a base pointer is cached per thread: the JIT passes the same
// few memref base pointers over and over.
struct ClassCacheEntry {
  const void *base = nullptr;
  bool is_device = false;
  int device = -1;
};
constexpr int kClassCacheSize = 256;
thread_local ClassCacheEntry t_class_cache[kClassCacheSize];

Resolved resolve(const void *base, const void *elem) {
  Mirror mir;
  if (find_mirror(elem, &mir)) {
    return {Where::HostMirrored, mir.dev + (reinterpret_cast<const char *>(elem) - mir.host), -1};
  }
  const uintptr_t bits = reinterpret_cast<uintptr_t>(base);
  const size_t slot = ((bits >> 8) ^ (bits >> 17) ^ (bits >> 26)) & (kClassCacheSize - 1);
  ClassCacheEntry &ce = t_class_cache[slot];
  if (ce.base != base) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, base);
    bool is_dev = false;
    int dev = -1;
    if (e == cudaSuccess) {
      is_dev = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
      dev = attr.device;
    } else {
      cudaGetLastError(); // plain malloc memory on older drivers reports an error: treat as host
    }
    ce.base = base;
    ce.is_device = is_dev;
    ce.device = dev;
  }
  if (ce.is_device) return {Where::Device, const_cast<char *>(reinterpret_cast<const char *>(elem)), ce.device};
  return {Where::HostPlain, nullptr, -1};
}

void use_device(int dev) {
  if (dev < 0 && t_ctx.device < 0) {   // a thread's first launch without a device operand: TPP_XSMM_DEVICE, in EVERY thread
    static const int env_dev = [] { const char *e = getenv("TPP_XSMM_DEVICE"); return e ? atoi(e) : -1; }();
    dev = env_dev;
  }
  if (dev >= 0 && dev != t_ctx.device) {
    TPP_CUDA_CHECK(cudaSetDevice(dev));
    t_ctx.device = dev;
  }
}

inline size_t esize(int64_t dtype) { return dtype == kF32 ? 4 : 2; }

inline char *elem_ptr(int64_t dtype, void *aligned, int64_t offset) {
  // runtime/Xsmm/XsmmRunnerUtils.cpp:63-75 (get_base_ptr)
  if (dtype != kF32 && dtype != kBF16) {
    fprintf(stderr, "Unhandled data type in get_data_pointer_from_memref_desc:%lld", (long long)dtype);
    return nullptr;
  }
  return static_cast<char *>(aligned) + offset * (int64_t)esize(dtype);
}

// One operand of an invoke: where it lives and, for plain host memory, the 2-D
// footprint (rows x width elements, pitch ld) that has to be staged.
struct Operand {
  void *aligned = nullptr;
  char *elem = nullptr;
  int64_t rows = 0, width = 0, ld = 0; // footprint in elements
  bool is_input = false, is_output = false;
  char *dev = nullptr; // resolved device address
  Where where = Where::Device;
};

struct StagedCall {
  Operand *ops;
  int nops;
  size_t es;
  bool any_host = false;
};

void stage_in(StagedCall &sc, cudaStream_t stream) {
  int known_dev = -1;
  for (int i = 0; i < sc.nops; ++i) {
    Operand &o = sc.ops[i];
    if (!o.elem) continue;
    Resolved r = resolve(o.aligned, o.elem);
    o.where = r.where;
    o.dev = r.dev;
    if (r.where == Where::Device && r.device >= 0) known_dev = r.device;
    if (r.where == Where::HostPlain) sc.any_host = true;
  }
  use_device(known_dev);
  if (!sc.any_host) return;
  if (t_ctx.capturing)
    fail("graph capture needs device or registered host operands: plain host pointers are staged synchronously");
  for (int i = 0; i < sc.nops; ++i) {
    Operand &o = sc.ops[i];
    if (!o.elem || o.where != Where::HostPlain) continue;
    const size_t pitch = (size_t)o.ld * sc.es;
    const size_t bytes = o.rows > 0 ? (size_t)(o.rows - 1) * pitch + (size_t)o.width * sc.es : 0;
    o.dev = static_cast<char *>(t_ctx.stage[i].get(bytes ? bytes : 16));
    if (o.is_input && bytes) {
      if (o.rows == 1 || o.ld == o.width)
        TPP_CUDA_CHECK(cudaMemcpyAsync(o.dev, o.elem, bytes, cudaMemcpyHostToDevice, stream));
      else
        TPP_CUDA_CHECK(cudaMemcpy2DAsync(o.dev, pitch, o.elem, pitch, (size_t)o.width * sc.es, (size_t)o.rows,
                                         cudaMemcpyHostToDevice, stream));
    }
  }
}

void stage_out(StagedCall &sc, cudaStream_t stream) {
  if (!sc.any_host) return;
  for (int i = 0; i < sc.nops; ++i) {
    Operand &o = sc.ops[i];
    if (!o.elem || o.where != Where::HostPlain || !o.is_output) continue;
    const size_t pitch = (size_t)o.ld * sc.es;
    if (o.rows == 1 || o.ld == o.width) {
      const size_t bytes = (size_t)(o.rows - 1) * pitch + (size_t)o.width * sc.es;
      TPP_CUDA_CHECK(cudaMemcpyAsync(o.elem, o.dev, bytes, cudaMemcpyDeviceToHost, stream));
    } else {
      TPP_CUDA_CHECK(cudaMemcpy2DAsync(o.elem, pitch, o.dev, pitch, (size_t)o.width * sc.es, (size_t)o.rows,
                                       cudaMemcpyDeviceToHost, stream));
    }
  }
  // strict mode keeps the reference's synchronous semantics: the result is
  // visible to the host when the invoke returns
  TPP_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ---- dispatch cache -------------------------------------------------------------
struct Key {
  int64_t v[16];
  bool operator==(const Key &o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key &k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) {
      h ^= (uint64_t)k.v[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};
std::mutex g_cache_mutex;
std::unordered_map<Key, KernelDesc *, KeyHash> g_cache;

template <typename Build> int64_t cached(const Key &key, Build build) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) return reinterpret_cast<int64_t>(it->second);
  KernelDesc *d = new KernelDesc(); // immortal, like libxsmm's code registry
  build(*d);
  g_cache.emplace(key, d);
  return reinterpret_cast<int64_t>(d);
}

const KernelDesc *desc_of(int64_t addr, OpClass a, OpClass b = OpClass::TileConfig, OpClass c = OpClass::TileConfig) {
  const KernelDesc *d = reinterpret_cast<const KernelDesc *>(addr);
  if (!d || d->magic != kDescMagic || (d->op != a && d->op != b && d->op != c)) {
    fail("invoke called with a handle that was not returned by the matching dispatch");
  }
  return d;
}

void print_gemm_shape(const char *what, int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                      int64_t ldc, int64_t sa, int64_t sb, int64_t flags) {
  // same information the reference prints before exit(-1) (XsmmRunnerUtils.cpp:352-358),
  // in the row-major view
  fprintf(stderr, "%s\n", what);
  fprintf(stderr, "dtype: %lld\nM: %lld\nN: %lld\nK: %lld\nlda: %lld\nldb: %lld\nldc: %lld\n", (long long)dtype,
          (long long)m, (long long)n, (long long)k, (long long)lda, (long long)ldb, (long long)ldc);
  fprintf(stderr, "stride_a: %lld\nstride_b: %lld\nflags: %lld\n", (long long)sa, (long long)sb, (long long)flags);
}

int64_t gemm_family_dispatch(OpClass op, int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                             int64_t ldc, int64_t sa, int64_t sb, int64_t gflags, int64_t uflags, int64_t ukind,
                             int64_t bflags, int64_t bkind) {
  ensure_cuda();
  const char *what = op == OpClass::Gemm     ? "failed to generate matmul func"
                     : op == OpClass::Brgemm ? "failed to generate brgemm func"
                                             : "failed to generate fused brgemm func";
  bool ok = (dtype == kF32 || dtype == kBF16) && m > 0 && n > 0 && k > 0;
  const bool vnni_b = (gflags & XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI) != 0;
  // op verifier rules (lib/TPP/Dialect/Xsmm/XsmmOps.cpp:319-344): lda >= k, ldb >= n, ldc >= n
  ok = ok && lda >= k && ldb >= n && ldc >= n && sa >= 0 && sb >= 0;
  // VNNI blocking factor of B: what the compiler asked libxsmm_cpuid_dot_pack_factor (2; 4 with TPP_XSMM_VNNI=4)
  const int64_t vfac = vnni_b ? (int64_t)libxsmm_cpuid_dot_pack_factor((int)kBF16) : 0;
  if (vnni_b) ok = ok && dtype == kBF16 && (vfac == 2 || vfac == 4) && (k % vfac) == 0;
  if (gflags & XSMM_GEMM_FLAG_VNNI_C) ok = false; // never produced by the pipeline (XsmmVerify.cpp:91-95)
  if (op == OpClass::FusedBrgemm) {
    ok = ok && bkind >= 0 && bkind <= 4 && (ukind == XSMM_UNARY_NONE || ukind == XSMM_UNARY_RELU);
    ok = ok && (bflags == 0 || bflags == 1 || bflags == 4 || bflags == 16) && uflags == 0;
  }
  if (!ok) {
    print_gemm_shape(what, dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags);
    exit(-1);
  }
  Key key{{(int64_t)op, dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags, uflags, ukind, bflags, bkind, vfac}};
  return cached(key, [&](KernelDesc &d) {
    d.op = op;
    d.dtype = dtype;
    d.m = m; d.n = n; d.k = k; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.stride_a = sa; d.stride_b = sb;
    d.gemm_flags = gflags;
    d.unary_flags = uflags; d.unary_kind = ukind; d.binary_flags = bflags; d.binary_kind = bkind;
    d.vnni_factor = (int32_t)vfac;
    const char *force = getenv("TPP_XSMM_FORCE_SIMT");
    if (brgemm_tc_supported(d) && !(force && force[0] == '1')) {
      d.impl = KernelImpl::BrgemmTC;
      brgemm_tc_configure(d);
    } else {
      d.impl = KernelImpl::BrgemmSimt;
      snprintf(d.name, sizeof(d.name), "brgemm_simt_%s_64x64x16", dtype == kF32 ? "f32" : "bf16");
      if (vnni_b && !(force && force[0] == '1')) {
        // VNNI-2 B ([K/2][N][2], the reference's default bf16 weight layout) is not a canonical UMMA operand
        // layout: large shapes un-interleave B into a scratch buffer (one HBM-bound pass) and run the tcgen05
        // kernel on the flat twin; small ones stay on the generic kernel.
        KernelDesc twin = d;
        twin.gemm_flags &= ~(int64_t)XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI;
        twin.vnni_factor = 0;
        if (brgemm_tc_supported(twin)) {
          twin.impl = KernelImpl::BrgemmTC;
          brgemm_tc_configure(twin);
          d.flat_twin = new KernelDesc(twin);
          snprintf(d.name, sizeof(d.name), "brgemm_bf16_vnni(unpack+tc | simt)");
        }
      }
    }
  });
}

// Launch one resolved BRGEMM (tensor-core kernel, VNNI un-interleave + tensor-core twin, or the generic kernel).
void issue_gemm(const KernelDesc *d, const GemmArgs &g, cudaStream_t stream) {
  const int64_t batch = g.batch;
  bool launched = false;
  if (d->impl == KernelImpl::BrgemmTC) {
    launched = launch_brgemm_tc(*d, g, stream);
    if (launched) t_ctx.last_kernel = brgemm_tc_last_name();
  } else if (d->flat_twin && d->vnni_factor == 2 && batch > 0 && (double)d->m * d->n * d->k * batch >= 2097152.0 && aligned16(g.B) &&
             [&] {
               // VNNI-2 B consumed as it is: the CTA-pair kernel rewrites the raw rows in shared memory (short reductions;
               // launch_brgemm_tc declines the long ones)
               GemmArgs gv = g;
               gv.b_vnni2 = true;
               return launch_brgemm_tc(*d->flat_twin, gv, stream);
             }()) {
    launched = true;
    t_ctx.last_kernel = brgemm_tc_last_name();
  } else if (d->flat_twin && batch > 0 && (double)d->m * d->n * d->k * batch >= 2097152.0 &&
             (batch == 1 || d->stride_b == d->k * d->ldb) && aligned16(g.B)) {
    // VNNI-B -> flat B in a scratch buffer (batches are contiguous: one tall [batch*k][ldb] un-interleave). The scratch
    // belongs to this (thread, stream): launches of one stream are serialised and may share it, kernels on different
    // streams never do. A captured launch gets a buffer owned by the graph instead, so that no later, larger direct
    // invoke can free memory a graph still references.
    const int64_t rows = batch * d->k;
    const size_t need = (size_t)rows * d->ldb * 2;
    void *flat = nullptr;
    if (t_ctx.capturing) {
      TPP_CUDA_CHECK(cudaMalloc(&flat, need));
      t_ctx.capture_tables.push_back(flat);
    } else {
      Staging *scratch = nullptr;
      for (auto &e : t_ctx.vnni_scratch)
        if (e.first == stream) scratch = e.second;
      if (!scratch) {
        scratch = new Staging();
        t_ctx.vnni_scratch.emplace_back(stream, scratch);
      }
      if (need > scratch->cap) TPP_CUDA_CHECK(cudaStreamSynchronize(stream));   // the old buffer may still be in use
      flat = scratch->get(need);
    }
    {
      if (d->vnni_factor == 4) launch_vnni4_unpack(g.B, flat, rows, d->n, d->ldb, d->ldb, stream);
      else launch_vnni2_unpack(g.B, flat, rows, d->n, d->ldb, d->ldb, stream);
      count_launch();
      GemmArgs gf = g;
      gf.B = flat;
      gf.b_independent = false;   // produced by the kernel just launched
      launched = launch_brgemm_tc(*d->flat_twin, gf, stream);
      if (launched) t_ctx.last_kernel = d->vnni_factor == 4 ? "vnni4_unpack+brgemm_tc_bf16" : "vnni2_unpack+brgemm_tc_bf16";
    }
  }
  if (!launched) {
    launch_brgemm_simt(*d, g, stream);
    t_ctx.last_kernel = d->dtype == kF32 ? "brgemm_simt_f32_64x64x16" : "brgemm_simt_bf16_64x64x16";
  }
  count_launch();
}

// ---- regrouping of captured tile invokes into layers ------------------------------------------------------------
// The reference tiles every layer into (iN, iK) output blocks and emits one small BRGEMM per block on block-packed
// operands (benchmarks/config/omp/mlir-bf16.json:37 runs --tiles=32,32,32: 256 invokes of a
// 32 x 32 x 32 x batch-32 BRGEMM per layer, scf.parallel over OpenMP threads). One GPU launch per block is hopeless, so
// runs of recorded invokes that share ONE descriptor and walk a regular grid - A + i a_step, B + j b_step,
// C + i c_step_n + j c_step_k, D + j d_step - are folded back into one layer-sized work item (GemmArgs::grid_*), which
// the pair-per-chain kernel addresses through 4-D tensor maps. Anything irregular stays a plain invoke.
struct Layer {
  const KernelDesc *d;
  GemmArgs g;
  size_t first, count;   // the invokes of `list` this layer was folded from
};

inline int64_t elem_diff(const void *a, const void *b, int64_t es) {   // (a - b) in elements of es bytes
  return (static_cast<const char *>(a) - static_cast<const char *>(b)) / es;
}

// the longest grid that starts at list[i]; returns the number of invokes folded (>= 1)
size_t fold_grid(const std::vector<PendingGemm> &list, size_t i, Layer *out) {
  const PendingGemm &p0 = list[i];
  out->d = p0.d;
  out->g = p0.g;
  out->first = i;
  out->count = 1;
  static const bool off = [] { const char *e = getenv("TPP_XSMM_REGROUP"); return e && e[0] == '0'; }();
  if (off) return 1;
  const int64_t es = (int64_t)esize(p0.d->dtype);
  auto elem_diff = [es](const void *a, const void *b) { return ::elem_diff(a, b, es); };
  size_t run = 1;   // invokes with the same descriptor / batch count / bias-ness
  while (i + run < list.size() && run < (1u << 20) && list[i + run].d == p0.d && list[i + run].g.batch == p0.g.batch &&
         (list[i + run].g.D != nullptr) == (p0.g.D != nullptr))
    ++run;
  if (run < 2) return 1;
  const GemmArgs &g0 = p0.g, &g1 = list[i + 1].g;
  // the inner loop runs over output-column blocks (same A, next B) or over row blocks (same B, next A)
  const bool inner_k = g1.A == g0.A && g1.B != g0.B;
  const bool inner_n = g1.B == g0.B && g1.A != g0.A;
  if (!inner_k && !inner_n) return 1;
  size_t inner = 1;
  while (inner < run && (inner_k ? list[i + inner].g.A == g0.A : list[i + inner].g.B == g0.B)) ++inner;
  // steps from the first two invokes of the inner loop and the first invoke of the second outer iteration
  const int64_t in_a = inner_k ? 0 : elem_diff(g1.A, g0.A), in_b = inner_k ? elem_diff(g1.B, g0.B) : 0;
  const int64_t in_c = elem_diff(g1.C, g0.C), in_d = g0.D ? elem_diff(g1.D, g0.D) : 0;
  int64_t out_a = 0, out_b = 0, out_c = 0, out_d = 0;
  size_t outer = 1;
  if (inner < run) {
    const GemmArgs &gn = list[i + inner].g;
    out_a = elem_diff(gn.A, g0.A); out_b = elem_diff(gn.B, g0.B); out_c = elem_diff(gn.C, g0.C);
    out_d = g0.D ? elem_diff(gn.D, g0.D) : 0;
    if (inner_k ? (out_b != 0 || out_a == 0) : (out_a != 0 || out_b == 0)) outer = 1;   // second outer iteration is not one
    else outer = run / inner;
  }
  auto matches = [&](size_t o, size_t t) {
    const GemmArgs &g = list[i + o * inner + t].g;
    const int64_t io = (int64_t)o, it = (int64_t)t;
    return elem_diff(g.A, g0.A) == io * out_a + it * in_a && elem_diff(g.B, g0.B) == io * out_b + it * in_b &&
           elem_diff(g.C, g0.C) == io * out_c + it * in_c && (!g0.D || elem_diff(g.D, g0.D) == io * out_d + it * in_d);
  };
  size_t good_outer = 0;
  for (size_t o = 0; o < outer; ++o) {
    bool ok = true;
    for (size_t t = 0; t < inner && ok; ++t) ok = matches(o, t);
    if (!ok) break;
    ++good_outer;
  }
  if (good_outer == 0) return 1;   // not even the first inner loop is regular
  GemmArgs g = g0;
  if (inner_k) {
    g.grid_k = (int32_t)inner; g.grid_n = (int32_t)good_outer;
    g.b_step = in_b; g.c_step_k = in_c; g.d_step = in_d;
    g.a_step = good_outer > 1 ? out_a : 0; g.c_step_n = good_outer > 1 ? out_c : 0;
  } else {
    g.grid_n = (int32_t)inner; g.grid_k = (int32_t)good_outer;
    g.a_step = in_a; g.c_step_n = in_c;
    g.b_step = good_outer > 1 ? out_b : 0; g.c_step_k = good_outer > 1 ? out_c : 0; g.d_step = good_outer > 1 ? out_d : 0;
  }
  // only forward-walking grids whose tiles cannot touch each other or the layer's inputs
  const KernelDesc &d = *p0.d;
  const int64_t c_tile = (d.m - 1) * d.ldc + d.n;
  bool ok = g.a_step >= 0 && g.b_step >= 0 && g.c_step_n >= 0 && g.c_step_k >= 0 && g.d_step >= 0;
  if (g.grid_k > 1) ok = ok && g.c_step_k >= d.n && (g.c_step_k >= c_tile || d.ldc >= (g.grid_k - 1) * g.c_step_k + d.n);
  if (g.grid_n > 1) ok = ok && g.c_step_n >= (g.grid_k - 1) * g.c_step_k + c_tile;
  if (g.grid_k > 1 && g.D) ok = ok && g.d_step == d.n;   // the bias slices of neighbouring column blocks are contiguous
  if (ok) {

    // bounding ranges: the outputs of the grid must not overlap anything the grid reads
    const char *c_lo = static_cast<const char *>(g.C);
    const char *c_hi = c_lo + ((g.grid_n - 1) * g.c_step_n + (g.grid_k - 1) * g.c_step_k + c_tile) * es;
    const int64_t nb = g.batch > 0 ? g.batch : 1;
    const bool vnni = (d.gemm_flags & XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI) != 0;
    const char *a_lo = static_cast<const char *>(g.A);
    const char *a_hi = a_lo + ((g.grid_n - 1) * g.a_step + (nb - 1) * d.stride_a + (d.m - 1) * d.lda + d.k) * es;
    const char *b_lo = static_cast<const char *>(g.B);
    const int64_t vf = d.vnni_factor > 0 ? d.vnni_factor : 2;
    const int64_t b_tile = vnni ? (nb - 1) * d.stride_b + ((d.k / vf - 1) * d.ldb + d.n) * vf
                                : (nb - 1) * d.stride_b + (d.k - 1) * d.ldb + d.n;
    const char *b_hi = b_lo + ((g.grid_k - 1) * g.b_step + b_tile) * es;
    ok = !(c_lo < a_hi && a_lo < c_hi) && !(c_lo < b_hi && b_lo < c_hi);
    if (ok && g.D) {
      const char *d_lo = static_cast<const char *>(g.D), *d_hi = d_lo + ((g.grid_k - 1) * g.d_step + d.n) * es;
      ok = !(c_lo < d_hi && d_lo < c_hi);
    }
  }
  if (!ok) return 1;
  out->g = g;
  out->count = inner * good_outer;
  return out->count;
}

// Launch the BRGEMMs recorded during graph capture, in order. Tile invokes are first folded into layers (above); runs
// of 2..4 consecutive layers that form a chain (C of layer l is A of layer l+1) go to the persistent fused kernels
// ("whole-MLP fusion ... or CUDA-graph capture of the invoke sequence"); everything else is launched
// exactly as a direct invoke would.
void flush_tiles();
void flush_held_zero();
const KernelDesc *fused_variant(const KernelDesc *d, bool add_bias, bool relu, bool beta0);

void retire_lazy_allocs(bool all_done);

void flush_pending_impl() {
  if (t_ctx.up_pending) {   // kernels issued from here on see every upload_async issued before them
    TPP_CUDA_CHECK(cudaStreamWaitEvent(t_ctx.stream, t_ctx.ev_up, 0));
    t_ctx.up_pending = false;
  }
  flush_tiles();
  // program order: the pending BRGEMMs were recorded BEFORE a held zero (a zero is held only until the next BRGEMM)
  struct ZeroLast { ~ZeroLast() { flush_held_zero(); } } zero_last;
  if (t_ctx.pending.empty()) return;
  std::vector<PendingGemm> list;
  list.swap(t_ctx.pending);
  cudaStream_t stream = t_ctx.stream;
  std::vector<Layer> layers;
  for (size_t i = 0; i < list.size();) {
    Layer L;
    i += fold_grid(list, i, &L);
    layers.push_back(L);
  }
  // segment the layers: maximal chains (2..4 layers, C of one = A of the next) and single layers
  const size_t n = layers.size();
  std::vector<const KernelDesc *> descs(n);
  std::vector<GemmArgs> args(n);
  for (size_t k = 0; k < n; ++k) { descs[k] = layers[k].d; args[k] = layers[k].g; }
  std::vector<int> seg_first, seg_len;     // seg_len 1 = not a chain
  for (size_t i = 0; i < n;) {
    int L = 1;
    for (int t = 4; t >= 2; --t)
      if (i + t <= n && brgemm_chain_linked(descs.data() + i, args.data() + i, t)) { L = t; break; }
    seg_first.push_back((int)i);
    seg_len.push_back(L);
    i += L;
  }
  // a layer no fused kernel took. A grid of SMALL tile invokes (f32 tiles, shapes the tcgen05 chain kernels do not take:
  // a batch that is no multiple of 256 rows, odd widths, ...) still goes out as ONE launch of the generic kernel, one
  // z-slice per tile, instead of hundreds of launches of a few microseconds each; big tiles and plain invokes are launched
  // one by one, exactly as they were recorded (the per-layer tcgen05 kernels)
  static const bool batch_off = [] { const char *e = getenv("TPP_XSMM_GRID_SIMT"); return e && e[0] == '0'; }();
  auto issue_layer = [&](size_t l) {
    const KernelDesc *d = descs[l];
    const GemmArgs &g = args[l];
    // (f32 has no other kernel: its grids are batched whatever the tile size - 64^3 tiles x batch 64, the reference's
    // --layers=4096,1024 --tiles=64,64,64 fp32 configs, are 2^24 MACs each)
    if (!batch_off && g.is_grid() && g.batch > 0 &&
        (d->dtype == kF32 || (double)d->m * d->n * d->k * g.batch < 16777216.0)) {
      launch_brgemm_simt(*d, g, stream);
      static thread_local char name[96];
      const int ct = (d->m <= 32 && d->n <= 32) ? 32 : 64;   // CTA tile of the generic kernel (brgemm_simt.cu)
      snprintf(name, sizeof(name), "brgemm_simt_%s_%dx%dx16_grid%dx%d", d->dtype == kF32 ? "f32" : "bf16", ct, ct, g.grid_n, g.grid_k);
      t_ctx.last_kernel = name;
      count_launch();
      return;
    }
    for (size_t k = layers[l].first; k < layers[l].first + layers[l].count; ++k) issue_gemm(list[k].d, list[k].g, stream);
  };
  // layers only the pair-per-chain kernel can run on the tensor cores in one launch: grids of tile invokes, VNNI-2 weights
  auto pair_only = [&](size_t l) {
    return (args[l].is_grid() || descs[l]->impl != KernelImpl::BrgemmTC) && brgemm_layer_chainable(*descs[l], args[l]);
  };
  // a single layer (no chain) goes to the pair-per-chain kernel only when it is a grid of tile invokes (the alternative is
  // one launch per tile); a plain invoke with VNNI-2 weights has the per-layer kernels (issue_gemm: the CTA-pair GEMM
  // converts VNNI-2 B in shared memory), which spread one big layer over the whole machine
  auto single_pair = [&](size_t l) { return args[l].is_grid() && brgemm_layer_chainable(*descs[l], args[l]); };
  for (size_t sidx = 0; sidx < seg_first.size();) {
    // a run of consecutive segments of the same kind: chains, or single layers that need the pair kernel
    const bool chains = seg_len[sidx] > 1;
    if (!chains && !single_pair((size_t)seg_first[sidx])) {
      issue_layer((size_t)seg_first[sidx]);
      ++sidx;
      continue;
    }
    size_t run = sidx;
    while (run < seg_first.size() && (chains ? seg_len[run] > 1 : (seg_len[run] == 1 && single_pair((size_t)seg_first[run]))))
      ++run;
    bool force = !chains;
    for (size_t q = sidx; q < run && !force; ++q)
      for (int l = 0; l < seg_len[q]; ++l) force = force || pair_only((size_t)(seg_first[q] + l));
    // many independent chains: one CTA pair per chain (4x less L2 -> SM traffic per layer than the pass kernels)
    int took = launch_brgemm_chains_pair(descs.data(), args.data(), seg_first.data() + sidx, seg_len.data() + sidx,
                                         (int)(run - sidx), stream, /*force=*/false);
    // few chains (a lone forward pass above all, or the exact repeats of an unrolled loop): the pass kernels spread ONE
    // layer over 128 SMs. Flat chains always; block-packed / VNNI-2 chains (GEN instantiations) decline more than a
    // handful of independent chains - the pair kernel with column-split items is faster there.
    // (the few-chain kernels own per-launch counters: an allocation and a sync each time, which only a captured graph
    // amortises - a lazily flushed lone chain goes out as PDL-chained per-layer launches instead)
    if (took == 0 && chains && t_ctx.capturing)
      took = launch_brgemm_chains_ft(descs.data(), args.data(), seg_first.data() + sidx, seg_len.data() + sidx,
                                     (int)(run - sidx), stream);
    if (took == 0 && force)
      took = launch_brgemm_chains_pair(descs.data(), args.data(), seg_first.data() + sidx, seg_len.data() + sidx,
                                       (int)(run - sidx), stream, /*force=*/true);
    if (took > 0) {
      t_ctx.last_kernel = brgemm_tc_last_name();
      count_launch();
      for (int x = brgemm_tc_take_extra_launches(); x > 0; --x) count_launch();   // e.g. the VNNI-2 -> flat weight copy
      sidx += took;
      continue;
    }
    const int f = seg_first[sidx], L = seg_len[sidx];
    if (chains && t_ctx.capturing && launch_brgemm_chain(descs.data() + f, args.data() + f, L, stream)) {
      t_ctx.last_kernel = brgemm_tc_last_name();
      count_launch();
    } else {
      for (int l = 0; l < L; ++l) issue_layer((size_t)(f + l));
    }
    ++sidx;
  }
}

// Device tables the fused kernels of a LAZY flush read (outside a capture nothing owns them): parked with an event and
// freed when it has completed; all_done = the device has been drained.
void retire_lazy_allocs(bool all_done) {
  for (size_t i = 0; i < t_ctx.retired.size();) {
    ThreadCtx::Retired &r = t_ctx.retired[i];
    if (all_done || cudaEventQuery(r.ev) == cudaSuccess) {
      for (void *p : r.ptrs) cudaFree(p);
      cudaEventDestroy(r.ev);
      t_ctx.retired[i] = std::move(t_ctx.retired.back());
      t_ctx.retired.pop_back();
    } else {
      ++i;
    }
  }
  cudaGetLastError();   // cudaEventQuery's cudaErrorNotReady is not an error
}

void flush_pending() {
  flush_pending_impl();
  if (t_ctx.capturing) return;
  // lazy mode: whatever the launches allocated belongs to nobody - retire it behind an event
  std::vector<void *> tables;
  brgemm_tc_take_capture_allocs(tables);
  tables.insert(tables.end(), t_ctx.capture_tables.begin(), t_ctx.capture_tables.end());
  t_ctx.capture_tables.clear();
  if (!tables.empty()) {
    ThreadCtx::Retired r;
    TPP_CUDA_CHECK(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
    TPP_CUDA_CHECK(cudaEventRecord(r.ev, t_ctx.stream));
    r.ptrs.swap(tables);
    t_ctx.retired.push_back(std::move(r));
  }
  if (!t_ctx.retired.empty()) retire_lazy_allocs(false);
}

// TPP_XSMM_HOST_PROFILE=1: where the host time of a BRGEMM invoke goes (printed at thread exit)
const bool g_host_prof = getenv("TPP_XSMM_HOST_PROFILE") != nullptr;
struct HostProf {
  uint64_t n = 0, resolve_ns = 0, launch_ns = 0, total_ns = 0;
  ~HostProf() {
    if (n)
      fprintf(stderr, "tpp-xsmm-cuda host profile: %llu brgemm invokes, per invoke: resolve %.2f us, launch %.2f us, "
                      "total %.2f us\n", (unsigned long long)n, resolve_ns / 1e3 / n, launch_ns / 1e3 / n,
              total_ns / 1e3 / n);
  }
};
thread_local HostProf t_prof;
inline uint64_t now_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

void gemm_family_invoke(const KernelDesc *d, int64_t dtype, void *pA, int64_t offA, void *pB, int64_t offB, void *pC,
                        int64_t offC, void *pD, int64_t offD, int64_t batch) {
  if (dtype != d->dtype) fail("invoke data type does not match the dispatched kernel");
  if (batch < 0) batch = 0;
  const bool vnni_b = (d->gemm_flags & XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI) != 0;
  const bool beta0 = (d->gemm_flags & XSMM_GEMM_FLAG_BETA_0) != 0;
  Operand ops[4];
  // A: batch x (m x k, pitch lda); staged as one row of the whole span
  const int64_t nb = batch > 0 ? batch : 1;
  ops[0].aligned = pA; ops[0].elem = elem_ptr(dtype, pA, offA);
  ops[0].rows = 1; ops[0].width = (nb - 1) * d->stride_a + (d->m - 1) * d->lda + d->k; ops[0].ld = ops[0].width;
  ops[0].is_input = batch > 0;
  ops[1].aligned = pB; ops[1].elem = elem_ptr(dtype, pB, offB);
  ops[1].rows = 1;
  const int64_t vf = d->vnni_factor > 0 ? d->vnni_factor : 2;
  ops[1].width = vnni_b ? (nb - 1) * d->stride_b + ((d->k / vf - 1) * d->ldb + d->n) * vf
                        : (nb - 1) * d->stride_b + (d->k - 1) * d->ldb + d->n;
  ops[1].ld = ops[1].width;
  ops[1].is_input = batch > 0;
  ops[2].aligned = pC; ops[2].elem = elem_ptr(dtype, pC, offC);
  ops[2].rows = d->m; ops[2].width = d->n; ops[2].ld = d->ldc;
  ops[2].is_input = !beta0; ops[2].is_output = true;
  const bool has_d = d->op == OpClass::FusedBrgemm && d->binary_kind != 0 && pD != nullptr;
  if (has_d) {
    ops[3].aligned = pD; ops[3].elem = elem_ptr(dtype, pD, offD);
    const int64_t bf = d->binary_flags;
    if (bf & 4) { ops[3].rows = 1; ops[3].width = d->n; ops[3].ld = d->n; }
    else if (bf & 1) { ops[3].rows = 1; ops[3].width = d->m; ops[3].ld = d->m; }
    else if (bf & 16) { ops[3].rows = 1; ops[3].width = 1; ops[3].ld = 1; }
    else { ops[3].rows = d->m; ops[3].width = d->n; ops[3].ld = d->ldc; }
    ops[3].is_input = true;
  }
  cudaStream_t stream = t_ctx.stream;
  StagedCall sc{ops, 4, esize(dtype)};
  const uint64_t t0 = g_host_prof ? now_ns() : 0;
  stage_in(sc, stream);
  const uint64_t t1 = g_host_prof ? now_ns() : 0;

  GemmArgs g;
  g.A = ops[0].dev; g.B = ops[1].dev; g.C = ops[2].dev; g.D = has_d ? ops[3].dev : nullptr;
  g.batch = batch;
  {
    const size_t es = esize(dtype);
    g.a_independent = !t_ctx.recently_written(ops[0].dev, (size_t)ops[0].width * es);
    g.b_independent = !t_ctx.recently_written(ops[1].dev, (size_t)ops[1].width * es) &&
                      (!has_d || !t_ctx.recently_written(ops[3].dev, (size_t)((ops[3].rows - 1) * ops[3].ld + ops[3].width) * es));
    t_ctx.note_output(ops[2].dev, (size_t)((d->m - 1) * d->ldc + d->n) * es);
  }
  if (d->impl == KernelImpl::BrgemmTC && !t_ctx.capturing) {
    if (++t_ctx.pdl_run >= kPdlWindow) {   // close the window: this launch waits for everything before it
      g.pdl = false;
      t_ctx.pdl_run = 0;
    }
  } else {
    t_ctx.pdl_run = 0;                     // generic kernels / captured work are launched in plain stream order
  }
  if (t_ctx.recording() && t_ctx.held_zero.d) {
    const ThreadCtx::HeldZero z = t_ctx.held_zero;
    if (!sc.any_host && !beta0 && d->dtype == z.d->dtype && z.out == ops[2].dev && z.d->m == d->m && z.d->n == d->n &&
        z.d->ldo == d->ldc && t_ctx.pending_tiles.empty()) {
      t_ctx.held_zero = {};                      // C = 0; C += A B   ==   C = A B
      d = fused_variant(d, false, false, true);
    } else {
      flush_pending();                           // earlier BRGEMMs, then the zero, then this invoke
    }
  }
  // (bf16 invokes the tensor-core kernels can take are fused by flush_pending(); everything else - f32, shapes only the
  // generic kernel takes - is recorded too, so that a regular grid of such tile invokes becomes one batched launch)
  if (t_ctx.recording() && !sc.any_host && batch > 0) {
    flush_tiles();
    t_ctx.pending.push_back({d, g});   // launched (possibly fused with its neighbours) by flush_pending()
    if (!t_ctx.capturing && t_ctx.pending.size() >= kLazyQueueMax) flush_pending();   // lazy mode: bounded queue
    return;
  }
  flush_pending();
  issue_gemm(d, g, stream);
  if (g_host_prof) {
    const uint64_t t2 = now_ns();
    t_prof.n++; t_prof.resolve_ns += t1 - t0; t_prof.launch_ns += t2 - t1; t_prof.total_ns += t2 - t0;
  }
  stage_out(sc, stream);
}

int bcast_mode_unary(int64_t flags) {
  return flags == XSMM_UNARY_FLAG_BCAST_ROW      ? kBcastRow
         : flags == XSMM_UNARY_FLAG_BCAST_COL    ? kBcastCol
         : flags == XSMM_UNARY_FLAG_BCAST_SCALAR ? kBcastScalar
                                                 : kBcastNone;
}

void set_footprint(Operand &o, int mode, int64_t m, int64_t n, int64_t ld) {
  switch (mode) {
  case kBcastRow: o.rows = m; o.width = 1; o.ld = ld > 0 ? ld : 1; break;
  case kBcastCol: o.rows = 1; o.width = n; o.ld = n; break;
  case kBcastScalar: o.rows = 1; o.width = 1; o.ld = 1; break;
  default: o.rows = m; o.width = n; o.ld = ld; break;
  }
}

} // namespace

// ================================ dispatch ========================================

extern "C" int64_t xsmm_gemm_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                                      int64_t ldc, int64_t flags) {
  return gemm_family_dispatch(OpClass::Gemm, dtype, m, n, k, lda, ldb, ldc, 0, 0, flags, 0, 0, 0, 0);
}

extern "C" int64_t xsmm_brgemm_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                                        int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t flags) {
  return gemm_family_dispatch(OpClass::Brgemm, dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags, 0, 0, 0, 0);
}

extern "C" int64_t xsmm_fused_brgemm_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
                                              int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
                                              int64_t gemm_flags, int64_t unary_flags, int64_t unary_kind,
                                              int64_t binary_flags, int64_t binary_kind) {
  return gemm_family_dispatch(OpClass::FusedBrgemm, dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags,
                              unary_flags, unary_kind, binary_flags, binary_kind);
}

extern "C" int64_t xsmm_unary_dispatch(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                                       int64_t flags) {
  ensure_cuda();
  bool ok = (dtype == kF32 || dtype == kBF16) && m > 0 && n > 0 && ldi >= 0 && ldo > 0;
  KernelImpl impl = KernelImpl::Eltwise;
  const char *name = "";
  switch (kind) {
  case XSMM_UNARY_IDENTITY: name = "unary_identity"; break;
  case XSMM_UNARY_ZERO: name = "unary_zero"; break;
  case XSMM_UNARY_RELU: name = "unary_relu"; break;
  case XSMM_UNARY_TRANSPOSE: impl = KernelImpl::Transpose; name = "unary_transpose_64x64"; ok = ok && flags == 0 && ldi >= n && ldo >= m; break;
  case XSMM_UNARY_VNNI2: impl = KernelImpl::Vnni2Pack; name = "unary_vnni2_pack"; ok = ok && flags == 0 && dtype == kBF16 && (m % 2) == 0 && ldi >= n && ldo >= n; break;
  case XSMM_UNARY_UNVNNI2_EXT: impl = KernelImpl::Vnni2Unpack; name = "unary_vnni2_unpack"; ok = ok && flags == 0 && dtype == kBF16 && (m % 2) == 0 && ldi >= n && ldo >= n; break;
  case XSMM_UNARY_VNNI4: impl = KernelImpl::Vnni4Pack; name = "unary_vnni4_pack"; ok = ok && flags == 0 && dtype == kBF16 && (m % 4) == 0 && ldi >= n && ldo >= n; break;
  case XSMM_UNARY_UNVNNI4_EXT: impl = KernelImpl::Vnni4Unpack; name = "unary_vnni4_unpack"; ok = ok && flags == 0 && dtype == kBF16 && (m % 4) == 0 && ldi >= n && ldo >= n; break;
  default: ok = false;
  }
  if (impl == KernelImpl::Eltwise) {
    ok = ok && (flags == 0 || flags == 2 || flags == 4 || flags == 8) && ldo >= n;
    if (flags == 0 && kind != XSMM_UNARY_ZERO) ok = ok && ldi >= n;
  }
  if (!ok) {
    fprintf(stderr, "failed to generate unary func\nop_type: %lld\nflags: %lld\n", (long long)kind, (long long)flags);
    fprintf(stderr, "M: %lld\nN: %lld\ndtype: %lld\nldi: %lld\nldo: %lld\n", (long long)m, (long long)n,
            (long long)dtype, (long long)ldi, (long long)ldo);
    exit(-1);
  }
  Key key{{(int64_t)OpClass::Unary, kind, dtype, m, n, ldi, ldo, flags, 0, 0, 0, 0, 0, 0, 0, 0}};
  return cached(key, [&](KernelDesc &d) {
    d.op = OpClass::Unary;
    d.impl = impl;
    d.dtype = dtype; d.kind = kind; d.m = m; d.n = n; d.ldi = ldi; d.ldo = ldo; d.flags = flags;
    snprintf(d.name, sizeof(d.name), "%s_%s", name, dtype == kF32 ? "f32" : "bf16");
  });
}

extern "C" int64_t xsmm_binary_dispatch(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldiLhs,
                                        int64_t ldiRhs, int64_t ldo, int64_t flags) {
  ensure_cuda();
  bool ok = (dtype == kF32 || dtype == kBF16) && m > 0 && n > 0 && kind >= 1 && kind <= 4 && ldo >= n;
  const int64_t f0 = flags & (1 | 4 | 16), f1 = flags & (2 | 8 | 32);
  ok = ok && (f0 == 0 || f0 == 1 || f0 == 4 || f0 == 16) && (f1 == 0 || f1 == 2 || f1 == 8 || f1 == 32) &&
       (flags & ~63ll) == 0;
  if (f0 == 0) ok = ok && ldiLhs >= n;
  if (f1 == 0) ok = ok && ldiRhs >= n;
  if (!ok) {
    fprintf(stderr, "failed to generate binary func\nop_type: %lld\nflags: %lld\n", (long long)kind, (long long)flags);
    fprintf(stderr, "M: %lld\nN: %lld\ndtype: %lld\nldi: %lld\nldi2: %lld\nldo: %lld\n", (long long)m, (long long)n,
            (long long)dtype, (long long)ldiLhs, (long long)ldiRhs, (long long)ldo);
    exit(-1);
  }
  Key key{{(int64_t)OpClass::Binary, kind, dtype, m, n, ldiLhs, ldiRhs, ldo, flags, 0, 0, 0, 0, 0, 0, 0}};
  return cached(key, [&](KernelDesc &d) {
    static const char *names[] = {"", "add", "mul", "sub", "div"};
    d.op = OpClass::Binary;
    d.impl = KernelImpl::Eltwise;
    d.dtype = dtype; d.kind = kind; d.m = m; d.n = n; d.ldi = ldiLhs; d.ldi2 = ldiRhs; d.ldo = ldo; d.flags = flags;
    snprintf(d.name, sizeof(d.name), "binary_%s_%s", names[kind], dtype == kF32 ? "f32" : "bf16");
  });
}

extern "C" int64_t xsmm_intel_amx_tile_config_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
                                                       int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
                                                       int64_t flags) {
  // AMX tile registers do not exist on a GPU (lib/TPP/Transforms/IntelAMXTileConfig.cpp:32-139
  // inserts these unconditionally for bf16): one shared no-op descriptor.
  (void)m; (void)n; (void)k; (void)lda; (void)ldb; (void)ldc; (void)stride_a; (void)stride_b; (void)flags;
  Key key{{(int64_t)OpClass::TileConfig, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
  return cached(key, [&](KernelDesc &d) {
    d.op = OpClass::TileConfig;
    d.impl = KernelImpl::Noop;
    d.dtype = dtype;
    snprintf(d.name, sizeof(d.name), "amx_tile_config_noop");
  });
}

// ================================= invoke ==========================================

extern "C" void xsmm_gemm_invoke(int64_t dtype, int64_t addr, void *alignedPtrA, int64_t offsetA, void *alignedPtrB,
                                 int64_t offsetB, void *alignedPtrC, int64_t offsetC) {
  NvtxRange nvtx("xsmm_gemm_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::Gemm);
  gemm_family_invoke(d, dtype, alignedPtrA, offsetA, alignedPtrB, offsetB, alignedPtrC, offsetC, nullptr, 0, 1);
}

extern "C" void xsmm_brgemm_invoke(int64_t dtype, int64_t addr, void *alignedPtrA, int64_t offsetA, void *alignedPtrB,
                                   int64_t offsetB, void *alignedPtrC, int64_t offsetC, int64_t numBatches) {
  NvtxRange nvtx("xsmm_brgemm_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::Brgemm, OpClass::Gemm);
  gemm_family_invoke(d, dtype, alignedPtrA, offsetA, alignedPtrB, offsetB, alignedPtrC, offsetC, nullptr, 0,
                     numBatches);
}

extern "C" void xsmm_fused_brgemm_invoke(int64_t dtype, int64_t addr, void *alignedPtrA, int64_t offsetA,
                                         void *alignedPtrB, int64_t offsetB, void *alignedPtrC, int64_t offsetC,
                                         void *alignedPtrD, int64_t offsetD, int64_t numBatches) {
  NvtxRange nvtx("xsmm_fused_brgemm_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::FusedBrgemm);
  gemm_family_invoke(d, dtype, alignedPtrA, offsetA, alignedPtrB, offsetB, alignedPtrC, offsetC, alignedPtrD, offsetD,
                     numBatches);
}

namespace {
// launch one unary TPP on resolved device operands
void launch_unary(const KernelDesc *d, const char *in, char *out, bool use_imm, float imm, cudaStream_t stream) {
  switch (d->impl) {
  case KernelImpl::Transpose:
    launch_transpose(in, out, d->m, d->n, d->ldi, d->ldo, (int)esize(d->dtype), stream);
    break;
  case KernelImpl::Vnni2Pack:
    launch_vnni2_pack(in, out, d->m, d->n, d->ldi, d->ldo, stream);
    break;
  case KernelImpl::Vnni2Unpack:
    launch_vnni2_unpack(in, out, d->m, d->n, d->ldi, d->ldo, stream);
    break;
  case KernelImpl::Vnni4Pack:
    launch_vnni4_pack(in, out, d->m, d->n, d->ldi, d->ldo, stream);
    break;
  case KernelImpl::Vnni4Unpack:
    launch_vnni4_unpack(in, out, d->m, d->n, d->ldi, d->ldo, stream);
    break;
  default: {
    EltwiseArgs a;
    a.in0 = in; a.out = out;
    a.m = d->m; a.n = d->n; a.ld0 = d->ldi; a.ldo = d->ldo;
    a.mode0 = use_imm ? kBcastImm : bcast_mode_unary(d->flags);
    a.imm = imm;
    a.op = d->kind == XSMM_UNARY_ZERO ? kOpZero : d->kind == XSMM_UNARY_RELU ? kOpRelu : kOpIdentity;
    a.dtype = d->dtype;
    launch_eltwise(a, stream);
  }
  }
}

// Do two pitched rectangles of bytes share a byte? Exact when both have the same pitch (tiles of one matrix interleave
// in address space without touching), conservative (bounding ranges) otherwise.
bool rects_overlap(const char *a, int64_t a_rows, int64_t a_w, int64_t a_ld, const char *b, int64_t b_rows, int64_t b_w,
                   int64_t b_ld) {
  const char *a_hi = a + (a_rows - 1) * a_ld + a_w, *b_hi = b + (b_rows - 1) * b_ld + b_w;
  if (!(a < b_hi && b < a_hi)) return false;
  if (a_ld != b_ld || a_ld <= 0) return true;
  if (b < a) { std::swap(a, b); std::swap(a_rows, b_rows); std::swap(a_w, b_w); }
  const int64_t delta = b - a, dr = delta / a_ld, dc = delta % a_ld;   // b's first byte sits at row dr, column dc of a's grid
  if (dr < a_rows && dc < a_w) return true;
  if (dc + b_w > a_ld && dr + 1 < a_rows) return true;                 // b's rows wrap into the next grid row, column 0
  return false;
}

struct TileRects { const char *in; char *out; int64_t in_rows, in_w, in_ld, out_rows, out_w, out_ld; };
TileRects tile_rects(const KernelDesc *d, const char *in, char *out) {
  const int64_t es = (int64_t)esize(d->dtype);
  if (d->impl == KernelImpl::Transpose) return {in, out, d->m, d->n * es, d->ldi * es, d->n, d->m * es, d->ldo * es};
  return {in, out, d->m, d->n * es, d->ldi * es, d->m, d->n * es, d->ldo * es};
}

// the zero that no BRGEMM absorbed is launched where it was issued
void flush_held_zero() {
  if (!t_ctx.held_zero.d) return;
  const ThreadCtx::HeldZero z = t_ctx.held_zero;
  t_ctx.held_zero = {};
  launch_unary(z.d, nullptr, z.out, false, 0.f, t_ctx.stream);
  t_ctx.note_output(z.out, (size_t)((z.d->m - 1) * z.d->ldo + z.d->n) * esize(z.d->dtype));
  t_ctx.last_kernel = z.d->name;
  count_launch();
}

// Does a run of tile moves walk a regular grid - tile t = i * J + j at in0 + i * in_outer + j * in_inner ->
// out0 + i * out_outer + j * out_inner (byte steps)? J = the length of the first stretch of constant steps.
struct TileGrid { int64_t J, I, in_inner, in_outer, out_inner, out_outer; };
bool detect_tile_grid(const std::vector<PendingTile> &list, TileGrid *tg) {
  const int64_t n_t = (int64_t)list.size();
  if (n_t < 2) return false;
  const int64_t in_inner = list[1].in - list[0].in, out_inner = list[1].out - list[0].out;
  int64_t J = n_t;
  for (int64_t i = 1; i < n_t; ++i)
    if (list[i].in - list[i - 1].in != in_inner || list[i].out - list[i - 1].out != out_inner) { J = i; break; }
  if ((n_t % J) != 0) return false;
  const int64_t I = n_t / J;
  const int64_t in_outer = I > 1 ? list[J].in - list[0].in : 0, out_outer = I > 1 ? list[J].out - list[0].out : 0;
  for (int64_t t = 0; t < n_t; ++t)
    if (list[t].in != list[0].in + (t / J) * in_outer + (t % J) * in_inner ||
        list[t].out != list[0].out + (t / J) * out_outer + (t % J) * out_inner)
      return false;
  *tg = {J, I, in_inner, in_outer, out_inner, out_outer};
  return true;
}

// Launch the tile moves recorded during graph capture: four or more become ONE batched kernel reading a device table
// of (in, out) pointers, fewer are launched as they would have been.
void flush_tiles() {
  if (t_ctx.pending_tiles.empty()) return;
  std::vector<PendingTile> list;
  list.swap(t_ctx.pending_tiles);
  t_ctx.tiles_by_in.clear();
  t_ctx.tiles_by_out.clear();
  cudaStream_t stream = t_ctx.stream;
  const KernelDesc *d = list[0].d;
  if (list.size() < 4) {
    for (const PendingTile &t : list) {
      launch_unary(d, t.in, t.out, false, 0.f, stream);
      count_launch();
    }
    t_ctx.last_kernel = d->name;
    return;
  }
  const int64_t es = (int64_t)esize(d->dtype);
  std::vector<TilePtrs> host(list.size());
  bool vec_ok = d->impl != KernelImpl::Transpose && ((d->n * es) % 16) == 0 && ((d->ldi * es) % 16) == 0 &&
                ((d->ldo * es) % 16) == 0;
  for (size_t i = 0; i < list.size(); ++i) {
    host[i] = {list[i].in, list[i].out};
    vec_ok = vec_ok && aligned16(list[i].in) && aligned16(list[i].out);
  }
  // A run that walks a regular grid (tile t = i * J + j at in0 + i * in_outer + j * in_inner -> out0 + i * out_outer +
  // j * out_inner: what a lowered tensor.pack / unpack emits) needs no table at all: source and destination are one
  // rank-4 tensor each and the copy is TMA to TMA (tile_grid.cu)
  static const bool grid_off = [] { const char *e = getenv("TPP_XSMM_TILE_GRID"); return e && e[0] == '0'; }();
  if (vec_ok && !grid_off) {
    TileGrid tg;
    if (detect_tile_grid(list, &tg) && launch_tile_grid(list[0].in, list[0].out, tg.J, tg.I, tg.in_inner, tg.in_outer, tg.out_inner,
                                                        tg.out_outer, d->m, d->n, d->ldi, d->ldo, (int)es, stream)) {
      const int64_t I = tg.I, J = tg.J;
      static thread_local char gname[96];
      snprintf(gname, sizeof(gname), "%s_batch%zu_tma%lldx%lld", d->name, list.size(), (long long)I, (long long)J);
      t_ctx.last_kernel = gname;
      count_launch();
      return;
    }
  }
  // the table is written now, outside the capture; the graph only holds the kernel that reads it
  static thread_local cudaStream_t side = nullptr;
  if (!side) TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  TilePtrs *table = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&table, host.size() * sizeof(TilePtrs)));
  TPP_CUDA_CHECK(cudaMemcpyAsync(table, host.data(), host.size() * sizeof(TilePtrs), cudaMemcpyHostToDevice, side));
  TPP_CUDA_CHECK(cudaStreamSynchronize(side));
  t_ctx.capture_tables.push_back(table);
  launch_tile_batch(table, (int64_t)host.size(), d->impl == KernelImpl::Transpose, d->m, d->n, d->ldi, d->ldo, (int)es,
                    vec_ok, stream);
  static thread_local char name[96];
  snprintf(name, sizeof(name), "%s_batch%zu", d->name, host.size());
  t_ctx.last_kernel = name;
  count_launch();
}
}  // namespace

namespace {
// ---- runtime CombineXsmmOp (capture only) --------------------------------------------------------------------
// When the compiler's fusion pass does not fire (lib/TPP/Transforms/CombineXsmmPass.cpp:31-145 needs the
// brgemm -> binary add(bcast_col_in0) -> unary relu chain on ONE buffer inside one block), a layer reaches the ABI as
// [unary zero(C)] -> brgemm(A, B, C) -> binary add(bias, C, C) -> unary relu(C, C): three or four launches and as many
// trips of C through memory ("unfused MLP layer"). During graph capture the BRGEMM is still
// pending when its consumers arrive, so the same rewrite is done here: the add / relu are folded into the pending
// invoke's descriptor (post-ops on the f32 accumulator, ONE rounding - exactly what the compiler's fused op computes,
// docs/TPPDialect.md:288-300), a zero that is overwritten by a beta=1 BRGEMM of the same tile becomes beta_0.
// Returns the descriptor of `d` with the given epilogue / beta, from the ordinary dispatch cache.
const KernelDesc *fused_variant(const KernelDesc *d, bool add_bias, bool relu, bool beta0) {
  const int64_t gflags = beta0 ? (d->gemm_flags | XSMM_GEMM_FLAG_BETA_0) : d->gemm_flags;
  const bool bias = add_bias || d->binary_kind == XSMM_BINARY_ADD;
  const bool act = relu || d->unary_kind == XSMM_UNARY_RELU;
  const OpClass op = (bias || act) ? OpClass::FusedBrgemm : d->op;
  const int64_t h = gemm_family_dispatch(op, d->dtype, d->m, d->n, d->k, d->lda, d->ldb, d->ldc, d->stride_a, d->stride_b, gflags,
                                         0, act ? XSMM_UNARY_RELU : XSMM_UNARY_NONE, bias ? XSMM_BINARY_FLAG_BCAST_COL_IN_0 : 0,
                                         bias ? XSMM_BINARY_ADD : XSMM_BINARY_NONE);
  return reinterpret_cast<const KernelDesc *>(h);
}

// the last pending BRGEMM if its output tile is exactly [C, m x n, pitch ld] of this dtype
PendingGemm *pending_producer_of(int64_t dtype, const char *C, int64_t m, int64_t n, int64_t ld) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_COMBINE"); return e && e[0] == '0'; }();
  if (off || !t_ctx.recording() || t_ctx.pending.empty() || !t_ctx.pending_tiles.empty()) return nullptr;
  PendingGemm &p = t_ctx.pending.back();
  if (p.d->dtype != dtype || static_cast<const char *>(p.g.C) != C || p.d->m != m || p.d->n != n || p.d->ldc != ld) return nullptr;
  return &p;
}

}  // namespace

static void unary_invoke_impl(const KernelDesc *d, int64_t dtype, void *pIn, int64_t offIn, bool use_imm, float imm,
                              void *pOut, int64_t offOut) {
  // tile moves (plain identity copy / transpose) issued during graph capture are collected, see flush_tiles()
  const bool batchable = t_ctx.recording() && !use_imm &&
                         (d->impl == KernelImpl::Transpose ||
                          (d->impl == KernelImpl::Eltwise && d->kind == XSMM_UNARY_IDENTITY && d->flags == 0));
  if (dtype != d->dtype) fail("invoke data type does not match the dispatched kernel");
  if (t_ctx.recording() && !use_imm && d->impl == KernelImpl::Eltwise && d->flags == 0 && d->dtype == kBF16) {
    if (d->kind == XSMM_UNARY_RELU && pIn == pOut && offIn == offOut && d->ldi == d->ldo) {
      // relu(C, C) right after the BRGEMM that produces C: becomes that invoke's epilogue
      Resolved r = resolve(pOut, elem_ptr(dtype, pOut, offOut));
      PendingGemm *p = r.where != Where::HostPlain ? pending_producer_of(dtype, r.dev, d->m, d->n, d->ldo) : nullptr;
      if (p && p->d->unary_kind == XSMM_UNARY_NONE && (p->d->op != OpClass::FusedBrgemm || p->d->binary_kind == XSMM_BINARY_NONE ||
                                                      (p->d->binary_kind == XSMM_BINARY_ADD && (p->d->binary_flags & 4)))) {
        p->d = fused_variant(p->d, false, true, (p->d->gemm_flags & XSMM_GEMM_FLAG_BETA_0) != 0);
        return;
      }
    }
    if (d->kind == XSMM_UNARY_ZERO) {
      // zero(C): held back; dropped if the very next invoke is a beta=1 BRGEMM that overwrites exactly this tile
      Resolved r = resolve(pOut, elem_ptr(dtype, pOut, offOut));
      if (r.where != Where::HostPlain) {
        if (t_ctx.held_zero.d) flush_pending();   // one zero is held at a time (pending BRGEMMs first, then that zero)
        flush_tiles();
        t_ctx.held_zero = {d, r.dev};
        return;
      }
    }
  }
  if (batchable) {
    std::vector<PendingTile> keep;
    keep.swap(t_ctx.pending_tiles);   // flush_pending() must not launch the run this invoke may still join
    flush_pending();
    keep.swap(t_ctx.pending_tiles);
  } else {
    flush_pending();
  }
  Operand ops[2];
  const int mode = bcast_mode_unary(d->flags);
  const bool reads_input = d->kind != XSMM_UNARY_ZERO && !use_imm;
  if (reads_input) {
    ops[0].aligned = pIn; ops[0].elem = elem_ptr(dtype, pIn, offIn);
    ops[0].is_input = true;
  }
  ops[1].aligned = pOut; ops[1].elem = elem_ptr(dtype, pOut, offOut);
  ops[1].is_output = true;
  switch (d->impl) {
  case KernelImpl::Transpose:
    ops[0].rows = d->m; ops[0].width = d->n; ops[0].ld = d->ldi;
    ops[1].rows = d->n; ops[1].width = d->m; ops[1].ld = d->ldo;
    break;
  case KernelImpl::Vnni2Pack:
    ops[0].rows = d->m; ops[0].width = d->n; ops[0].ld = d->ldi;
    ops[1].rows = d->m / 2; ops[1].width = 2 * d->n; ops[1].ld = 2 * d->ldo;
    break;
  case KernelImpl::Vnni2Unpack:
    ops[0].rows = d->m / 2; ops[0].width = 2 * d->n; ops[0].ld = 2 * d->ldi;
    ops[1].rows = d->m; ops[1].width = d->n; ops[1].ld = d->ldo;
    break;
  case KernelImpl::Vnni4Pack:
    ops[0].rows = d->m; ops[0].width = d->n; ops[0].ld = d->ldi;
    ops[1].rows = d->m / 4; ops[1].width = 4 * d->n; ops[1].ld = 4 * d->ldo;
    break;
  case KernelImpl::Vnni4Unpack:
    ops[0].rows = d->m / 4; ops[0].width = 4 * d->n; ops[0].ld = 4 * d->ldi;
    ops[1].rows = d->m; ops[1].width = d->n; ops[1].ld = d->ldo;
    break;
  default:
    set_footprint(ops[0], mode, d->m, d->n, d->ldi);
    ops[1].rows = d->m; ops[1].width = d->n; ops[1].ld = d->ldo;
  }
  // outputs with gaps (ld > width) are staged with 2-D copies so host bytes in
  // the gaps are never touched
  cudaStream_t stream = t_ctx.stream;
  StagedCall sc{ops, 2, esize(dtype)};
  stage_in(sc, stream);
  if (batchable && !sc.any_host) {
    // joins the pending run if it has the same descriptor and neither reads nor writes anything the run writes (nor
    // writes anything the run reads); otherwise the run is launched first and this invoke starts a new one
    const TileRects nr = tile_rects(d, ops[0].dev, ops[1].dev);
    constexpr size_t kMaxRun = 1u << 17;
    bool join = t_ctx.pending_tiles.empty() || (t_ctx.pending_tiles[0].d == d && t_ctx.pending_tiles.size() < kMaxRun);
    if (join && !t_ctx.pending_tiles.empty()) {
      // byte extents of a source / destination rectangle of this descriptor; a recorded rectangle starting at `lo` can
      // only touch [a, a_hi) if lo lies in (a - extent, a_hi)
      const int64_t in_ext = (nr.in_rows - 1) * nr.in_ld + nr.in_w, out_ext = (nr.out_rows - 1) * nr.out_ld + nr.out_w;
      auto hits = [&](const std::multimap<const char *, uint32_t> &index, int64_t ext, bool index_is_out, const char *a,
                      int64_t a_rows, int64_t a_w, int64_t a_ld) {
        const char *a_hi = a + (a_rows - 1) * a_ld + a_w;
        for (auto it = index.upper_bound(a - ext); it != index.end() && it->first < a_hi; ++it) {
          const PendingTile &pt = t_ctx.pending_tiles[it->second];
          const TileRects pr = tile_rects(d, pt.in, pt.out);
          if (index_is_out ? rects_overlap(a, a_rows, a_w, a_ld, pr.out, pr.out_rows, pr.out_w, pr.out_ld)
                           : rects_overlap(a, a_rows, a_w, a_ld, pr.in, pr.in_rows, pr.in_w, pr.in_ld))
            return true;
        }
        return false;
      };
      join = !hits(t_ctx.tiles_by_out, out_ext, true, nr.out, nr.out_rows, nr.out_w, nr.out_ld) &&    // write after write
             !hits(t_ctx.tiles_by_in, in_ext, false, nr.out, nr.out_rows, nr.out_w, nr.out_ld) &&     // write after read
             !hits(t_ctx.tiles_by_out, out_ext, true, nr.in, nr.in_rows, nr.in_w, nr.in_ld);          // read after write
    }
    if (!join) flush_tiles();
    t_ctx.tiles_by_in.emplace(ops[0].dev, (uint32_t)t_ctx.pending_tiles.size());
    t_ctx.tiles_by_out.emplace(ops[1].dev, (uint32_t)t_ctx.pending_tiles.size());
    t_ctx.pending_tiles.push_back({d, ops[0].dev, ops[1].dev});
    t_ctx.note_output(ops[1].dev, (size_t)((ops[1].rows - 1) * ops[1].ld + ops[1].width) * esize(dtype));
    return;
  }
  flush_tiles();
  (void)mode;
  t_ctx.pdl_run = 0;   // a plain launch: ordered after everything before it
  launch_unary(d, ops[0].dev, ops[1].dev, use_imm, imm, stream);
  t_ctx.note_output(ops[1].dev, (size_t)((ops[1].rows - 1) * ops[1].ld + ops[1].width) * esize(dtype));
  t_ctx.last_kernel = d->name;
  count_launch();
  stage_out(sc, stream);
}

extern "C" void xsmm_unary_invoke(int64_t dtype, int64_t addr, void *alignedPtrIn, int64_t offsetIn,
                                  void *alignedPtrOut, int64_t offsetOut) {
  NvtxRange nvtx("xsmm_unary_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::Unary);
  unary_invoke_impl(d, dtype, alignedPtrIn, offsetIn, false, 0.f, alignedPtrOut, offsetOut);
}

extern "C" void xsmm_unary_scalar_invoke(int64_t dtype, int64_t addr, float scalar, void *alignedPtrOut,
                                         int64_t offsetOut) {
  NvtxRange nvtx("xsmm_unary_scalar_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::Unary);
  if (d->impl != KernelImpl::Eltwise) fail("xsmm_unary_scalar_invoke: only identity/zero/relu take a scalar input");
  unary_invoke_impl(d, dtype, nullptr, 0, true, scalar, alignedPtrOut, offsetOut);
}

extern "C" void xsmm_binary_invoke(int64_t dtype, int64_t addr, void *alignedPtrLhs, int64_t offsetLhs,
                                   void *alignedPtrRhs, int64_t offsetRhs, void *alignedPtrOut, int64_t offsetOut) {
  NvtxRange nvtx("xsmm_binary_invoke");
  const KernelDesc *d = desc_of(addr, OpClass::Binary);
  if (dtype != d->dtype) fail("invoke data type does not match the dispatched kernel");
  if (t_ctx.recording() && d->dtype == kBF16 && d->kind == XSMM_BINARY_ADD && d->flags == XSMM_BINARY_FLAG_BCAST_COL_IN_0 &&
      alignedPtrRhs == alignedPtrOut && offsetRhs == offsetOut && d->ldi2 == d->ldo) {
    // add(bias[bcast_col_in0], C, C) right after the BRGEMM that produces C: becomes that invoke's epilogue
    Resolved rc = resolve(alignedPtrOut, elem_ptr(dtype, alignedPtrOut, offsetOut));
    Resolved rb = resolve(alignedPtrLhs, elem_ptr(dtype, alignedPtrLhs, offsetLhs));
    PendingGemm *p = (rc.where != Where::HostPlain && rb.where != Where::HostPlain)
                         ? pending_producer_of(dtype, rc.dev, d->m, d->n, d->ldo) : nullptr;
    if (p && p->d->unary_kind == XSMM_UNARY_NONE && (p->d->op != OpClass::FusedBrgemm || p->d->binary_kind == XSMM_BINARY_NONE)) {
      p->d = fused_variant(p->d, true, false, (p->d->gemm_flags & XSMM_GEMM_FLAG_BETA_0) != 0);
      p->g.D = rb.dev;
      return;
    }
  }
  flush_pending();
  const int64_t f = d->flags;
  const int mode0 = (f & 1) ? kBcastRow : (f & 4) ? kBcastCol : (f & 16) ? kBcastScalar : kBcastNone;
  const int mode1 = (f & 2) ? kBcastRow : (f & 8) ? kBcastCol : (f & 32) ? kBcastScalar : kBcastNone;
  Operand ops[3];
  ops[0].aligned = alignedPtrLhs; ops[0].elem = elem_ptr(dtype, alignedPtrLhs, offsetLhs); ops[0].is_input = true;
  ops[1].aligned = alignedPtrRhs; ops[1].elem = elem_ptr(dtype, alignedPtrRhs, offsetRhs); ops[1].is_input = true;
  ops[2].aligned = alignedPtrOut; ops[2].elem = elem_ptr(dtype, alignedPtrOut, offsetOut); ops[2].is_output = true;
  set_footprint(ops[0], mode0, d->m, d->n, d->ldi);
  set_footprint(ops[1], mode1, d->m, d->n, d->ldi2);
  ops[2].rows = d->m; ops[2].width = d->n; ops[2].ld = d->ldo;
  cudaStream_t stream = t_ctx.stream;
  StagedCall sc{ops, 3, esize(dtype)};
  stage_in(sc, stream);
  EltwiseArgs a;
  a.in0 = ops[0].dev; a.in1 = ops[1].dev; a.out = ops[2].dev;
  a.m = d->m; a.n = d->n; a.ld0 = d->ldi; a.ld1 = d->ldi2; a.ldo = d->ldo;
  a.mode0 = mode0; a.mode1 = mode1;
  a.op = kOpAdd + (int)(d->kind - 1);
  a.dtype = dtype;
  t_ctx.pdl_run = 0;
  launch_eltwise(a, stream);
  t_ctx.note_output(ops[2].dev, (size_t)((ops[2].rows - 1) * ops[2].ld + ops[2].width) * esize(dtype));
  t_ctx.last_kernel = d->name;
  count_launch();
  stage_out(sc, stream);
}

extern "C" void xsmm_intel_amx_tile_config_invoke(int64_t dtype, int64_t addr, void *alignedPtrA, int64_t offset) {
  (void)dtype; (void)addr; (void)alignedPtrA; (void)offset; // nothing to configure on a GPU
}

// ================================ perf timers ========================================

extern "C" int64_t perf_start_timer(void) {
  if (g_cuda_ready.load() && !t_ctx.capturing) {   // earlier async (or lazily queued) work is not ours to time
    flush_pending();
    TPP_CUDA_CHECK(cudaDeviceSynchronize());
    retire_lazy_allocs(true);
  }
  auto timestamp = std::chrono::high_resolution_clock::now();
  return timestamp.time_since_epoch().count();
}

extern "C" double perf_stop_timer(int64_t startTimestamp) {
  if (g_cuda_ready.load() && !t_ctx.capturing) {   // invokes are asynchronous launches (lazy mode: possibly still queued)
    flush_pending();
    TPP_CUDA_CHECK(cudaDeviceSynchronize());
    retire_lazy_allocs(true);
  }
  auto stop = std::chrono::high_resolution_clock::now();
  std::chrono::high_resolution_clock::time_point start{std::chrono::high_resolution_clock::duration{startTimestamp}};
  return std::chrono::duration_cast<std::chrono::duration<double>>(stop - start).count();
}

extern "C" int libxsmm_cpuid_dot_pack_factor(int datatype) {
  // libxsmm_datatype: F32 = 1, BF16 = 2 (include/TPP/Dialect/Xsmm/XsmmEnum.td:13-18)
  if (datatype != (int)kBF16) return 1;
  const char *env = getenv("TPP_XSMM_VNNI");
  if (env && env[0] == '0') return 0; // odd factors disable VNNI packing (VNNIUtils.cpp:41-43)
  if (env && env[0] == '4') return 4; // mlir-gen --vnni=4 layouts ([K/4][N][4]; benchmarks/config/omp/mlir-bf16.json:65-125)
  return 2;
}

// ================================ CUDA extensions =====================================

extern "C" void xsmm_cuda_set_stream(void *stream) {
  if (t_ctx.capturing) { // keep recording on the capture stream; the new stream takes effect after graph_end
    t_ctx.saved_stream = static_cast<cudaStream_t>(stream);
    return;
  }
  if (t_ctx.stream != static_cast<cudaStream_t>(stream)) {
    flush_pending();   // lazy mode: what was queued belongs to the old stream
    for (auto &r : t_ctx.recent_out) r = {};
    t_ctx.pdl_run = kPdlWindow;   // first BRGEMM on the new stream: plain stream order
  }   // dependency tracking is per stream; cross-stream order is the caller's
  t_ctx.stream = static_cast<cudaStream_t>(stream);
}
extern "C" void *xsmm_cuda_get_stream(void) { return t_ctx.stream; }

extern "C" void *xsmm_cuda_stream_create(void) {
  ensure_cuda();
  cudaStream_t s = nullptr;
  TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return s;
}
extern "C" void xsmm_cuda_stream_destroy(void *stream) {
  if (!stream) return;
  if (t_ctx.stream == static_cast<cudaStream_t>(stream)) t_ctx.stream = nullptr;
  TPP_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  TPP_CUDA_CHECK(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
}

extern "C" void xsmm_cuda_sync(void) {
  if (!g_cuda_ready.load()) return;
  if (!t_ctx.capturing) flush_pending();   // lazy mode: queued invokes are launched first
  TPP_CUDA_CHECK(cudaDeviceSynchronize());
  if (!t_ctx.capturing) retire_lazy_allocs(true);
}

// ---- function-local temporaries (xsmm_cuda_mark_temporary) ------------------------------------------------------------
namespace {
std::shared_mutex g_temp_mutex;
std::vector<std::pair<const char *, const char *>> g_temporaries;   // device byte ranges [lo, hi)
}  // namespace

namespace tpp {
bool range_is_temporary(const void *p, size_t bytes) {
  const char *lo = static_cast<const char *>(p), *hi = lo + bytes;
  std::shared_lock<std::shared_mutex> lock(g_temp_mutex);
  for (const auto &r : g_temporaries)
    if (r.first <= lo && hi <= r.second) return true;
  return false;
}
}  // namespace tpp

extern "C" void xsmm_cuda_mark_temporary(void *ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return;
  const Resolved r = resolve(ptr, ptr);
  if (r.where == Where::HostPlain) return;   // plain host memory never reaches a fused kernel: nothing to mark
  std::unique_lock<std::shared_mutex> lock(g_temp_mutex);
  for (auto &e : g_temporaries)
    if (e.first == r.dev) { e.second = r.dev + bytes; return; }
  g_temporaries.emplace_back(r.dev, r.dev + bytes);
}

extern "C" void xsmm_cuda_unmark_temporary(void *ptr) {
  if (!ptr) return;
  const Resolved r = resolve(ptr, ptr);
  if (r.where == Where::HostPlain) return;
  std::unique_lock<std::shared_mutex> lock(g_temp_mutex);
  for (size_t i = 0; i < g_temporaries.size(); ++i)
    if (g_temporaries[i].first == r.dev) {
      g_temporaries[i] = g_temporaries.back();
      g_temporaries.pop_back();
      return;
    }
}

extern "C" void xsmm_cuda_set_lazy(int64_t on) {
  if (!t_ctx.capturing && !on) flush_pending();
  t_ctx.lazy_init = true;
  t_ctx.lazy = on != 0;
}

extern "C" void xsmm_cuda_stream_sync(void) {
  if (!g_cuda_ready.load()) return;
  if (!t_ctx.capturing) flush_pending();
  TPP_CUDA_CHECK(cudaStreamSynchronize(t_ctx.stream));
  if (t_ctx.up_stream && !t_ctx.capturing) {   // everything this thread issued, asynchronous copies included
    TPP_CUDA_CHECK(cudaStreamSynchronize(t_ctx.up_stream));
    TPP_CUDA_CHECK(cudaStreamSynchronize(t_ctx.down_stream));
  }
}

extern "C" int64_t xsmm_cuda_register_host(void *host, int64_t bytes, int64_t upload) {
  ensure_cuda();
  if (!host || bytes <= 0) return -1;
  Mirror m;
  m.host = static_cast<char *>(host);
  m.bytes = (size_t)bytes;
  m.pinned_here = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess;
  if (!m.pinned_here) cudaGetLastError(); // already pinned (e.g. by torch) or not pinnable: copies still work
  void *dev = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&dev, (size_t)bytes));
  m.dev = static_cast<char *>(dev);
  if (upload) TPP_CUDA_CHECK(cudaMemcpyAsync(m.dev, m.host, m.bytes, cudaMemcpyHostToDevice, t_ctx.stream));
  {
    std::unique_lock<std::shared_mutex> lock(g_mirror_mutex);
    g_mirrors[reinterpret_cast<uintptr_t>(host)] = m;
    g_mirror_count.store((int)g_mirrors.size());
  }
  return 0;
}

extern "C" int64_t xsmm_cuda_unregister_host(void *host) {
  Mirror m;
  {
    std::unique_lock<std::shared_mutex> lock(g_mirror_mutex);
    auto it = g_mirrors.find(reinterpret_cast<uintptr_t>(host));
    if (it == g_mirrors.end()) return -1;
    m = it->second;
    g_mirrors.erase(it);
    g_mirror_count.store((int)g_mirrors.size());
  }
  TPP_CUDA_CHECK(cudaDeviceSynchronize());
  TPP_CUDA_CHECK(cudaFree(m.dev));
  if (m.pinned_here) cudaHostUnregister(m.host);
  return 0;
}

extern "C" int64_t xsmm_cuda_update_device(void *host, int64_t bytes) {
  Mirror m;
  if (!find_mirror(host, &m) || static_cast<char *>(host) + bytes > m.host + m.bytes) return -1;
  flush_pending();
  TPP_CUDA_CHECK(cudaMemcpyAsync(m.dev + (static_cast<char *>(host) - m.host), host, (size_t)bytes,
                                 cudaMemcpyHostToDevice, t_ctx.stream));
  return 0;
}

extern "C" int64_t xsmm_cuda_update_host(void *host, int64_t bytes) {
  Mirror m;
  if (!find_mirror(host, &m) || static_cast<char *>(host) + bytes > m.host + m.bytes) return -1;
  flush_pending();
  TPP_CUDA_CHECK(cudaMemcpyAsync(host, m.dev + (static_cast<char *>(host) - m.host), (size_t)bytes,
                                 cudaMemcpyDeviceToHost, t_ctx.stream));
  return 0;
}

namespace {
void ensure_copy_streams() {
  if (t_ctx.up_stream) return;
  TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&t_ctx.up_stream, cudaStreamNonBlocking));
  TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&t_ctx.down_stream, cudaStreamNonBlocking));
  for (cudaEvent_t *e : {&t_ctx.ev_up, &t_ctx.ev_compute, &t_ctx.ev_fork, &t_ctx.ev_join})
    TPP_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  for (auto &d : t_ctx.downloads) TPP_CUDA_CHECK(cudaEventCreateWithFlags(&d.ev, cudaEventDisableTiming));
}
} // namespace

// Upload on the thread's upload stream: ordered after earlier uploads only, so it overlaps kernels that are already
// queued; every invoke issued AFTER this call waits for it. During graph capture the copy becomes a parallel branch
// of the graph (root node, or behind the previous upload).
extern "C" int64_t xsmm_cuda_upload_async(void *host, int64_t bytes) {
  Mirror m;
  if (!find_mirror(host, &m) || static_cast<char *>(host) + bytes > m.host + m.bytes) return -1;
  flush_pending();
  ensure_copy_streams();
  if (t_ctx.capturing && !t_ctx.up_in_capture) {   // fork: pull the upload stream into the capture
    TPP_CUDA_CHECK(cudaEventRecord(t_ctx.ev_fork, t_ctx.stream));
    TPP_CUDA_CHECK(cudaStreamWaitEvent(t_ctx.up_stream, t_ctx.ev_fork, 0));
    t_ctx.up_in_capture = true;
  }
  TPP_CUDA_CHECK(cudaMemcpyAsync(m.dev + (static_cast<char *>(host) - m.host), host, (size_t)bytes,
                                 cudaMemcpyHostToDevice, t_ctx.up_stream));
  TPP_CUDA_CHECK(cudaEventRecord(t_ctx.ev_up, t_ctx.up_stream));
  t_ctx.up_pending = true;
  return 0;
}

// Download on the thread's download stream: starts when the invokes issued BEFORE this call have finished and runs
// under whatever is issued after it. Outside a capture, xsmm_cuda_wait_host(host) blocks until the bytes are in host
// memory; a captured download is complete when the graph launch is (xsmm_cuda_stream_sync).
extern "C" int64_t xsmm_cuda_download_async(void *host, int64_t bytes) {
  Mirror m;
  if (!find_mirror(host, &m) || static_cast<char *>(host) + bytes > m.host + m.bytes) return -1;
  flush_pending();
  ensure_copy_streams();
  TPP_CUDA_CHECK(cudaEventRecord(t_ctx.ev_compute, t_ctx.stream));
  TPP_CUDA_CHECK(cudaStreamWaitEvent(t_ctx.down_stream, t_ctx.ev_compute, 0));
  if (t_ctx.capturing) t_ctx.down_in_capture = true;
  TPP_CUDA_CHECK(cudaMemcpyAsync(host, m.dev + (static_cast<char *>(host) - m.host), (size_t)bytes,
                                 cudaMemcpyDeviceToHost, t_ctx.down_stream));
  if (!t_ctx.capturing) {
    ThreadCtx::Download &d = t_ctx.downloads[t_ctx.download_pos++ % kDownloadRing];
    d.host = host;
    TPP_CUDA_CHECK(cudaEventRecord(d.ev, t_ctx.down_stream));
  }
  return 0;
}

// Block until the most recent xsmm_cuda_download_async(host, ...) of this thread has delivered its bytes.
// Returns -1 if the thread has no such download on record (only the last kDownloadRing are kept): the caller must then
// drain the stream (xsmm_cuda_stream_sync) before it reuses the buffer.
extern "C" int64_t xsmm_cuda_wait_host(void *host) {
  for (int i = 1; i <= kDownloadRing && i <= t_ctx.download_pos; ++i) {
    ThreadCtx::Download &d = t_ctx.downloads[(t_ctx.download_pos - i) % kDownloadRing];
    if (d.host == host && d.ev) {
      TPP_CUDA_CHECK(cudaEventSynchronize(d.ev));
      return 0;
    }
  }
  return -1;
}

extern "C" void *xsmm_cuda_device_ptr(void *host) {
  Mirror m;
  if (!find_mirror(host, &m)) return nullptr;
  return m.dev + (static_cast<char *>(host) - m.host);
}

extern "C" int64_t xsmm_cuda_graph_begin(void) {
  ensure_cuda();
  if (t_ctx.capturing) return -1;
  flush_pending();   // lazy mode: earlier invokes are not part of the graph
  t_ctx.saved_stream = t_ctx.stream;
  if (t_ctx.stream == nullptr) { // the legacy default stream cannot be captured
    if (!t_ctx.capture_stream) TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&t_ctx.capture_stream, cudaStreamNonBlocking));
    t_ctx.stream = t_ctx.capture_stream;
  }
  // relaxed: pointer queries / attribute calls of the first invoke are legal during capture
  TPP_CUDA_CHECK(cudaStreamBeginCapture(t_ctx.stream, cudaStreamCaptureModeRelaxed));
  t_ctx.capturing = true;
  t_ctx.captured_launches = 0;
  t_ctx.up_in_capture = t_ctx.down_in_capture = false;
  t_ctx.up_pending = false;
  return 0;
}

extern "C" int64_t xsmm_cuda_graph_end(void) {
  if (!t_ctx.capturing) return 0;
  flush_pending();
  // join the copy branches back into the capture stream
  if (t_ctx.up_in_capture) {
    TPP_CUDA_CHECK(cudaEventRecord(t_ctx.ev_join, t_ctx.up_stream));
    TPP_CUDA_CHECK(cudaStreamWaitEvent(t_ctx.stream, t_ctx.ev_join, 0));
  }
  if (t_ctx.down_in_capture) {
    TPP_CUDA_CHECK(cudaEventRecord(t_ctx.ev_join, t_ctx.down_stream));
    TPP_CUDA_CHECK(cudaStreamWaitEvent(t_ctx.stream, t_ctx.ev_join, 0));
  }
  t_ctx.up_in_capture = t_ctx.down_in_capture = false;
  t_ctx.up_pending = false;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(t_ctx.stream, &graph);
  t_ctx.capturing = false;
  t_ctx.stream = t_ctx.saved_stream;
  std::vector<void *> tables;
  brgemm_tc_take_capture_allocs(tables);
  tables.insert(tables.end(), t_ctx.capture_tables.begin(), t_ctx.capture_tables.end());
  t_ctx.capture_tables.clear();
  if (e != cudaSuccess || !graph) {
    fprintf(stderr, "tpp-xsmm-cuda: graph capture failed: %s\n", cudaGetErrorString(e));
    cudaGetLastError();
    for (void *t : tables) cudaFree(t);
    return 0;
  }
  GraphHandle *gh = new GraphHandle();
  gh->device_allocs.swap(tables);
  gh->launches = t_ctx.captured_launches;
  snprintf(gh->last_kernel, sizeof(gh->last_kernel), "%s", t_ctx.last_kernel ? t_ctx.last_kernel : "");
  e = cudaGraphInstantiate(&gh->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    fprintf(stderr, "tpp-xsmm-cuda: graph instantiate failed: %s\n", cudaGetErrorString(e));
    for (void *t : gh->device_allocs) cudaFree(t);
    delete gh;
    return 0;
  }
  return reinterpret_cast<int64_t>(gh);
}

extern "C" void xsmm_cuda_graph_launch(int64_t graph) {
  NvtxRange nvtx("xsmm_cuda_graph_launch");
  GraphHandle *gh = reinterpret_cast<GraphHandle *>(graph);
  if (!gh || gh->magic != 0x47525048u) fail("xsmm_cuda_graph_launch: not a graph handle");
  flush_pending();   // orders the launch after this thread's pending upload_async
  TPP_CUDA_CHECK(cudaGraphLaunch(gh->exec, t_ctx.stream));
  {
    thread_local char replayed[96];
    memcpy(replayed, gh->last_kernel, sizeof(replayed));
    t_ctx.last_kernel = replayed;
  }
  g_launches.fetch_add(gh->launches, std::memory_order_relaxed);
}

extern "C" void xsmm_cuda_graph_destroy(int64_t graph) {
  GraphHandle *gh = reinterpret_cast<GraphHandle *>(graph);
  if (!gh || gh->magic != 0x47525048u) return;
  cudaGraphExecDestroy(gh->exec);
  for (void *t : gh->device_allocs) cudaFree(t);   // callers destroy a graph only after its replays have completed
  gh->magic = 0;
  delete gh;
}

extern "C" int64_t xsmm_cuda_launch_count(void) { return g_launches.load(); }
extern "C" const char *xsmm_cuda_last_kernel(void) { return t_ctx.last_kernel; }
extern "C" const char *xsmm_cuda_handle_kernel(int64_t addr) {
  const KernelDesc *d = reinterpret_cast<const KernelDesc *>(addr);
  return (d && d->magic == kDescMagic) ? d->name : "";
}
extern "C" int64_t xsmm_cuda_debug_rects_overlap(const void *a, int64_t a_rows, int64_t a_width, int64_t a_ld,
                                                 const void *b, int64_t b_rows, int64_t b_width, int64_t b_ld) {
  return rects_overlap(static_cast<const char *>(a), a_rows, a_width, a_ld, static_cast<const char *>(b), b_rows, b_width,
                       b_ld) ? 1 : 0;
}
// Debug / test hook (no device needed): what the capture path folds a run of tile invokes into. The invokes share one
// bf16 descriptor (m, n, k, lda, ldb, ldc, stride_a, stride_b, flags) and batch count; invoke t uses operand offsets
// a_off[t], b_off[t], c_off[t], d_off[t] (elements; d_off may be NULL). out[0..7] = grid_n, grid_k, a_step, b_step,
// c_step_n, c_step_k, d_step, number of invokes folded into the first layer.
extern "C" int64_t xsmm_cuda_debug_fold_grid(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                                             int64_t stride_a, int64_t stride_b, int64_t flags, int64_t batch, int64_t num,
                                             const int64_t *a_off, const int64_t *b_off, const int64_t *c_off,
                                             const int64_t *d_off, int64_t *out) {
  KernelDesc d;
  d.op = OpClass::FusedBrgemm;
  d.impl = KernelImpl::BrgemmTC;
  const bool f32 = (flags & (1ll << 40)) != 0;   // hook-only bit: fold f32 invokes (4-byte elements)
  flags &= ~(1ll << 40);
  d.dtype = f32 ? kF32 : kBF16;
  d.m = m; d.n = n; d.k = k; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.stride_a = stride_a; d.stride_b = stride_b;
  d.gemm_flags = flags;
  d.vnni_factor = (flags & XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI) ? 2 : 0;
  const int64_t hes = f32 ? 4 : 2;
  char *const A = reinterpret_cast<char *>(0x100000000ull), *const B = reinterpret_cast<char *>(0x200000000ull),
             *const C = reinterpret_cast<char *>(0x300000000ull), *const D = reinterpret_cast<char *>(0x400000000ull);
  std::vector<PendingGemm> list((size_t)num);
  for (int64_t t = 0; t < num; ++t) {
    list[(size_t)t].d = &d;
    GemmArgs &g = list[(size_t)t].g;
    g.A = A + hes * a_off[t]; g.B = B + hes * b_off[t]; g.C = C + hes * c_off[t];
    g.D = d_off ? D + hes * d_off[t] : nullptr;
    g.batch = batch;
  }
  Layer L;
  const size_t folded = num > 0 ? fold_grid(list, 0, &L) : 0;
  out[0] = L.g.grid_n; out[1] = L.g.grid_k; out[2] = L.g.a_step; out[3] = L.g.b_step; out[4] = L.g.c_step_n;
  out[5] = L.g.c_step_k; out[6] = L.g.d_step; out[7] = (int64_t)folded;
  return 0;
}
extern "C" int64_t xsmm_cuda_debug_tile_grid(int64_t num, const int64_t *in_off, const int64_t *out_off, int64_t *out) {
  char *const in = reinterpret_cast<char *>(0x100000000ull), *const o = reinterpret_cast<char *>(0x200000000ull);
  std::vector<PendingTile> list((size_t)num);
  for (int64_t t = 0; t < num; ++t) list[(size_t)t] = {nullptr, in + in_off[t], o + out_off[t]};
  TileGrid tg{};
  const bool ok = detect_tile_grid(list, &tg);
  out[0] = tg.J; out[1] = tg.I; out[2] = tg.in_inner; out[3] = tg.in_outer; out[4] = tg.out_inner; out[5] = tg.out_outer;
  return ok ? 1 : 0;
}
extern "C" int64_t xsmm_cuda_abi_version(void) { return 1; }
extern "C" void xsmm_cuda_debug_dump_trace(void) { brgemm_tc_dump_trace(); }
