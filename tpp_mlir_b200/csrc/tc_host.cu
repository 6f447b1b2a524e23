// tc_host.cu - host-side helpers shared by the tcgen05 kernel families (declared in tc_common.cuh): tensor-map
// encoding through the driver entry point, device memory owned by the graph being captured, operand-hazard tests of
// layer chains, the name of the last launch, and the debug trace dump.
#include <atomic>
#include <cstdarg>
#include <utility>

#include "tc_common.cuh"

namespace tpp {
namespace tc {

unsigned long long *g_trace_buf = nullptr;
int g_trace_next = 0;
int g_trace_ctas[kTraceRing] = {0};
int g_chain_trace_ctas = 0, g_chain_trace_layers = 0;
bool g_chain_trace_ft = false;
unsigned long long *g_pc_trace = nullptr;
int g_pc_trace_ctas = 0;

namespace {
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
      fprintf(stderr, "tpp-xsmm-cuda: cuTensorMapEncodeTiled is not available from the driver\n");
      exit(-1);
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 3-D bf16 tensor map: dims (inner, rows, batch), strides in elements for rows and batch.
} // namespace

bool encode_map(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t ld,
                uint64_t stride, uint32_t box_inner, uint32_t box_rows, uint32_t box_batch) {
  cuuint64_t dims[3] = {inner, rows, batch};
  // a size-1 batch dimension may carry any legal stride
  uint64_t bstride = stride * 2;
  if (batch <= 1 || bstride == 0) bstride = ld * 2;
  cuuint64_t strides[2] = {ld * 2, bstride};
  cuuint32_t box[3] = {box_inner, box_rows, box_batch};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 4-D bf16 tensor map over the activation matrix: dims (k within a 64-wide k-block, row, k-block, batch element) so
// that ONE box covers several k-blocks of the same rows: shared memory receives [batch][k-block][row][64], i.e.
// consecutive 128-byte-swizzled k-block tiles. The k-block dimension has a 128-byte stride (smaller than the row
// stride): TMA only requires strides to be multiples of 16 bytes.
bool encode_map_x4(CUtensorMap *map, const void *base, uint64_t k, uint64_t rows, uint64_t batch, uint64_t ld,
                   uint64_t stride, uint32_t box_rows, uint32_t box_kb, uint32_t box_b) {
  cuuint64_t dims[4] = {BLOCK_K, rows, k / BLOCK_K, batch};
  uint64_t bstride = stride * 2;
  if (batch <= 1 || bstride == 0) bstride = ld * 2;
  cuuint64_t strides[3] = {ld * 2, BLOCK_K * 2, bstride};
  cuuint32_t box[4] = {BLOCK_K, box_rows, box_kb, box_b};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


bool encode_map_nd(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides,
                   const uint32_t *box, int swizzle_bytes) {
  if (rank < 1 || rank > 5) return false;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i] * 2;
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                       : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// rank-4 map over 4-byte units (dtype-agnostic data movement): dims[0] in 4-byte units, strides in BYTES, no swizzle
bool encode_map_u32_4d(CUtensorMap *map, const void *base, const uint64_t *dims, const uint64_t *strides_bytes,
                       const uint32_t *box) {
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], estr[4];
  for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i < 3; ++i) gs[i] = strides_bytes[i];
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(base), gd, gs, bx, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int bin_mode_from_flags(int64_t f) {
  if (f & 4) return kBcastCol;
  if (f & 1) return kBcastRow;
  if (f & 16) return kBcastScalar;
  return kBcastNone;
}


// ---- the name of this thread's last tcgen05 launch ----------------------------------------------------------------
namespace {
thread_local char t_last_name[96] = "brgemm_tc_bf16";
thread_local int t_extra_launches = 0;
}
void note_extra_launch() { ++t_extra_launches; }
void set_last_name(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_last_name, sizeof(t_last_name), fmt, ap);
  va_end(ap);
}


// ---- device memory owned by the graph being captured ----------------------------------------------------------
// Everything a captured kernel node reads or spins on (descriptor tables, arrival counters, split-K workspaces) is
// allocated here, written / zeroed on a private non-capturing stream BEFORE the node can ever run, and handed to the
// graph handle at xsmm_cuda_graph_end (brgemm_tc_take_capture_allocs), which frees it with the graph. Nothing a graph
// references is shared with direct launches, so no later launch can free or re-zero it under a replay.
namespace {
thread_local std::vector<void *> t_capture_allocs;
}
cudaStream_t table_stream() {
  thread_local cudaStream_t st = nullptr;
  if (!st) TPP_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  return st;
}
bool stream_is_capturing(cudaStream_t stream) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cs);
  return cs != cudaStreamCaptureStatusNone;
}
// zero-filled device words, complete (not merely enqueued) when this returns: never a node of somebody's graph
void *alloc_zeroed(size_t bytes) {
  void *p = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&p, bytes));
  TPP_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, table_stream()));
  TPP_CUDA_CHECK(cudaStreamSynchronize(table_stream()));
  return p;
}
void *capture_owned_zeroed(size_t bytes) {
  void *p = alloc_zeroed(bytes);
  t_capture_allocs.push_back(p);
  return p;
}
// split-K exchange workspace of the capture in progress: kernels of one captured stream are serialised, so they share
// it; when a later launch needs more, a new one is allocated and the old one stays alive with the graph
namespace {
struct CaptureWs { float *ptr = nullptr; size_t bytes = 0; };
thread_local CaptureWs t_capture_ws;
}
float *capture_owned_ws(size_t need) {
  if (need > t_capture_ws.bytes) {
    void *p = nullptr;
    const size_t want = need < (4u << 20) ? (4u << 20) : need;
    TPP_CUDA_CHECK(cudaMalloc(&p, want));
    t_capture_allocs.push_back(p);
    t_capture_ws.ptr = static_cast<float *>(p);
    t_capture_ws.bytes = want;
  }
  return t_capture_ws.ptr;
}


void capture_adopt(void *p) { t_capture_allocs.push_back(p); }
void *capture_owned_table(const void *host, size_t bytes) {
  void *p = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&p, bytes));
  TPP_CUDA_CHECK(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, table_stream()));
  TPP_CUDA_CHECK(cudaStreamSynchronize(table_stream()));
  t_capture_allocs.push_back(p);
  return p;
}

namespace {
// has this process launched through the library on more than one (thread, stream)?
bool concurrent_launchers_seen(cudaStream_t stream) {
  static std::mutex mu;
  static std::vector<std::pair<size_t, cudaStream_t>> seen;
  static std::atomic<bool> many{false};
  if (many.load(std::memory_order_relaxed)) return true;
  thread_local size_t me = 0;
  thread_local cudaStream_t last = reinterpret_cast<cudaStream_t>(~(uintptr_t)0);
  if (me != 0 && last == stream) return false;   // fast path: same launcher as last time
  std::lock_guard<std::mutex> lock(mu);
  if (me == 0) { static size_t next_id = 0; me = ++next_id; }
  last = stream;
  bool known = false;
  for (auto &e : seen) known = known || (e.first == me && e.second == stream);
  if (!known) seen.emplace_back(me, stream);
  if (seen.size() > 1) many.store(true);
  return seen.size() > 1;
}
}  // namespace

bool prepare_resident_launch(const void *kernel, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attrs, bool only_if_concurrent) {
  static const int coop_env = [] { const char *e = getenv("TPP_XSMM_COOP"); return e ? atoi(e) : -1; }();   // 0 never, 1 always
  const bool concurrent = concurrent_launchers_seen(cfg->stream);
  const bool coop = coop_env == 0 ? false : coop_env == 1 ? true : (!only_if_concurrent || concurrent);
  const size_t ctas = (size_t)cfg->gridDim.x * cfg->gridDim.y * cfg->gridDim.z;
  size_t cluster = 1;
  for (unsigned i = 0; i < cfg->numAttrs; ++i)
    if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension)
      cluster = (size_t)cfg->attrs[i].val.clusterDim.x * cfg->attrs[i].val.clusterDim.y * cfg->attrs[i].val.clusterDim.z;
  size_t capacity = 0;
  if (cluster > 1) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    capacity = (size_t)n * cluster;
  } else {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)cfg->blockDim.x, cfg->dynamicSmemBytes) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    capacity = (size_t)per_sm * sms;
  }
  if (capacity < ctas) return false;
  // (the driver accepts programmatic serialisation + cooperative together, but the launches do not overlap: measured)
  if (coop && attrs[0].id == cudaLaunchAttributeProgrammaticStreamSerialization) {
    attrs[0].id = cudaLaunchAttributeCooperative;
    attrs[0].val.cooperative = 1;
  }
  return true;
}

unsigned long long *trace_ring() {
  if (!g_trace_buf) {
    const size_t bytes = sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS;
    g_trace_buf = static_cast<unsigned long long *>(alloc_zeroed(bytes));
  }
  return g_trace_buf;
}

LayerRanges layer_ranges(const KernelDesc &d, const GemmArgs &g) {
  const int64_t nb = g.batch > 0 ? g.batch : 1;
  const bool vnni = (d.gemm_flags & 2048) != 0;
  const int64_t a_tile = (nb - 1) * d.stride_a + (d.m - 1) * d.lda + d.k;
  const int64_t vf = d.vnni_factor > 0 ? d.vnni_factor : 2;
  const int64_t b_tile = vnni ? (nb - 1) * d.stride_b + ((d.k / vf - 1) * d.ldb + d.n) * vf
                              : (nb - 1) * d.stride_b + (d.k - 1) * d.ldb + d.n;
  const int64_t c_tile = (d.m - 1) * d.ldc + d.n;
  LayerRanges r;
  r.a = bf16_range(g.A, (g.grid_n - 1) * g.a_step + a_tile);
  r.b = bf16_range(g.B, (g.grid_k - 1) * g.b_step + b_tile);
  r.c = bf16_range(g.C, (g.grid_n - 1) * g.c_step_n + (g.grid_k - 1) * g.c_step_k + c_tile);
  r.d = g.D ? bf16_range(g.D, (g.grid_k - 1) * g.d_step + d.n) : ByteRange{nullptr, nullptr};
  return r;
}

// operand footprints of one chain: inputs (first layer's A, every layer's B and D) and outputs (every layer's C)
void chain_ranges(const KernelDesc *const *descs, const GemmArgs *args, int L, std::vector<ByteRange> &in,
                  std::vector<ByteRange> &out) {
  for (int l = 0; l < L; ++l) {
    const LayerRanges r = layer_ranges(*descs[l], args[l]);
    if (l == 0) in.push_back(r.a);
    in.push_back(r.b);
    if (r.d.lo) in.push_back(r.d);
    out.push_back(r.c);
  }
}

// No layer's weights / bias overlap ANY layer's output, no two outputs overlap, and the chain's input is not one of its
// outputs (byte ranges, not pointer equality: an operand that starts inside another layer's C is a hazard too).
bool chain_operands_hazard_free(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  LayerRanges r[8];
  if (L > 8) return false;
  for (int l = 0; l < L; ++l) r[l] = layer_ranges(*descs[l], args[l]);
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < L; ++j) {
      if (overlaps(r[l].b, r[j].c)) return false;
      if (r[l].d.lo && overlaps(r[l].d, r[j].c)) return false;
      if (j != l && overlaps(r[l].c, r[j].c)) return false;
    }
  for (int j = 0; j < L; ++j)
    if (overlaps(r[0].a, r[j].c)) return false;
  return true;
}
} // namespace tc

using namespace tc;

const char *brgemm_tc_last_name() { return t_last_name; }
int brgemm_tc_take_extra_launches() {
  const int n = t_extra_launches;
  t_extra_launches = 0;
  return n;
}

void brgemm_tc_take_capture_allocs(std::vector<void *> &out) {
  out.insert(out.end(), t_capture_allocs.begin(), t_capture_allocs.end());
  t_capture_allocs.clear();
  t_capture_ws = {};
}


// Debug (TPP_XSMM_TC_TRACE=2): wall-clock timeline of the traced launches, oldest first.
void brgemm_tc_dump_trace() {
  if (g_pc_trace && g_pc_trace_ctas) {
    TPP_CUDA_CHECK(cudaDeviceSynchronize());
    const int n = g_pc_trace_ctas;
    std::vector<unsigned long long> h((size_t)n * PC_TRACE_SLOTS);
    TPP_CUDA_CHECK(cudaMemcpy(h.data(), g_pc_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    auto med = [&](int sl, int parity) {   // median over the CTAs of one parity (0 = leaders, 1 = peers, 2 = all)
      std::vector<double> v;
      for (int c = 0; c < n; ++c) {
        if (parity < 2 && (c & 1) != parity) continue;
        const unsigned long long *r = &h[(size_t)c * PC_TRACE_SLOTS];
        if (r[sl] && r[60]) v.push_back((double)((long long)r[sl] - (long long)r[60]));
      }
      if (v.empty()) return 0.0;
      std::sort(v.begin(), v.end());
      return v[v.size() / 2];
    };
    fprintf(stderr, "pair-chain-trace %d CTAs; median SM clocks since CTA start; end=%.0f\n", n, med(61, 2));
    for (int t = 0; t < 12; ++t)
      fprintf(stderr, "  tile %2d: mma_start=%.0f mma_issued=%.0f acc_ready=%.0f stored=%.0f (peer: acc_ready=%.0f stored=%.0f)\n",
              t, med(4 * t, 0), med(4 * t + 1, 0), med(4 * t + 2, 0), med(4 * t + 3, 0), med(4 * t + 2, 1), med(4 * t + 3, 1));
    for (int l = 1; l < 4; ++l)
      if (med(48 + 2 * l, 2) > 0)
        fprintf(stderr, "  layer %d input: producer waits from %.0f to %.0f\n", l, med(48 + 2 * l, 2), med(49 + 2 * l, 2));
    return;
  }
  if (!g_trace_buf) return;
  TPP_CUDA_CHECK(cudaDeviceSynchronize());
  if (g_chain_trace_ctas && g_chain_trace_ft) {
    const int n_ctas = g_chain_trace_ctas;
    std::vector<unsigned long long> h((size_t)n_ctas * FT_TRACE_SLOTS);
    TPP_CUDA_CHECK(cudaMemcpy(h.data(), g_trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    auto med = [&](int sl, int ref) {   // median over CTAs (the ~20 CTAs that start early on idle SMs skew a mean)
      std::vector<double> v;
      for (int c = 0; c < n_ctas; ++c) {
        const unsigned long long *r = &h[(size_t)c * FT_TRACE_SLOTS];
        if (r[sl] && r[ref]) v.push_back((double)((long long)r[sl] - (long long)r[ref]));
      }
      if (v.empty()) return 0.0;
      std::sort(v.begin(), v.end());
      return v[v.size() / 2];
    };
    fprintf(stderr, "ft-chain-trace %d passes, %d CTAs; median SM clocks since the PDL wait passed: cta_start=%.0f end=%.0f\n",
            g_chain_trace_layers, n_ctas, med(0, 1), med(2, 1));
    for (int p = 0; p < 9 && p < g_chain_trace_layers; ++p)
      fprintf(stderr, "  pass %d: e0(inputs_ready|xchg_done)=%.0f e1(x_issued|sender_start)=%.0f e2(mma_start|pushed)=%.0f acc_ready=%.0f stored=%.0f arrived=%.0f\n", p,
              med(8 + 6 * p, 1), med(9 + 6 * p, 1), med(10 + 6 * p, 1), med(11 + 6 * p, 1), med(12 + 6 * p, 1),
              med(13 + 6 * p, 1));
    return;
  }
  if (g_chain_trace_ctas) {
    const int n_ctas = g_chain_trace_ctas;
    std::vector<unsigned long long> h((size_t)n_ctas * TRACE_SLOTS);
    TPP_CUDA_CHECK(cudaMemcpy(h.data(), g_trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    static const char *names[13] = {"", "L1_released", "L1_A_issued", "L1_data1", "L1_data_all", "L0_gridbar_passed", "",
                                    "L1_acc_ready", "L1_pushed", "L1_cluster", "L1_stored", "L1_gridbar_passed", "end"};
    fprintf(stderr, "chain-trace %d layers, %d CTAs: avg clocks since CTA start:", g_chain_trace_layers, n_ctas);
    for (int sl : {5, 1, 2, 3, 4, 7, 8, 9, 10, 11, 12}) {
      double sum = 0;
      int cnt = 0;
      for (int c = 0; c < n_ctas; ++c) {
        const unsigned long long *r = &h[(size_t)c * TRACE_SLOTS];
        if (r[sl] && r[0] && r[sl] > r[0]) { sum += (double)(r[sl] - r[0]); ++cnt; }
      }
      fprintf(stderr, " %s=%.0f", names[sl], cnt ? sum / cnt : 0.0);
    }
    fprintf(stderr, "\n");
    return;
  }
  std::vector<unsigned long long> h((size_t)kTraceRing * kTraceRingCtas * TRACE_SLOTS);
  TPP_CUDA_CHECK(cudaMemcpy(h.data(), g_trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  struct Row { unsigned long long start, wait_min, wait_max, end; int slot; };
  std::vector<Row> rows;
  for (int s = 0; s < kTraceRing; ++s) {
    if (!g_trace_ctas[s]) continue;
    Row r{~0ull, ~0ull, 0, 0, s};
    for (int c = 0; c < g_trace_ctas[s]; ++c) {
      const unsigned long long *t = &h[((size_t)s * kTraceRingCtas + c) * TRACE_SLOTS];
      if (!t[15]) continue;
      if (t[15] < r.start) r.start = t[15];
      if (t[13] && t[13] < r.wait_min) r.wait_min = t[13];
      if (t[13] > r.wait_max) r.wait_max = t[13];
      if (t[14] > r.end) r.end = t[14];
    }
    if (r.end) rows.push_back(r);
  }
  std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.start < b.start; });
  const size_t first = rows.size() > 12 ? rows.size() - 12 : 0;
  for (size_t i = first; i < rows.size(); ++i) {
    const Row &r = rows[i];
    const unsigned long long t0 = rows[first].start;
    fprintf(stderr, "tc-timeline slot %3d: first CTA start %+7lld ns, PDL wait passed %lld..%lld, last CTA end %lld ns "
                    "(kernel span %lld ns)\n", r.slot, (long long)(r.start - t0), (long long)(r.wait_min - t0),
            (long long)(r.wait_max - t0), (long long)(r.end - t0), (long long)(r.end - r.start));
  }
}


} // namespace tpp
