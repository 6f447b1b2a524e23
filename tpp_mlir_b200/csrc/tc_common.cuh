// tc_common.cuh - what the tcgen05 kernels of this backend share: tile constants, the parameter block of the
// stand-alone BRGEMM kernels, the fused epilogue, and the host-side helpers (tensor-map encoding, memory owned by
// the graph being captured, operand-hazard tests, debug trace buffers). Each kernel family lives in its own
// translation unit: brgemm_tc.cu (per-layer kernels), mlp_chain.cu / mlp_chain_ft.cu / mlp_chain_pair.cu (fused
// layer chains); tc_host.cu holds the definitions of the host helpers declared here.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace tpp {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one swizzle row of A
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KiB
constexpr int B_CHUNK_BYTES = BLOCK_K * 64 * 2;           // 64 k-rows x 128 bytes = 8 KiB
constexpr int NUM_THREADS = 192;
constexpr int RECV_BYTES = BLOCK_M * 64 * 4;              // split-K exchange: S slots x 128 rows x (64/S) f32 = 32 KiB

struct TcParams {
  void *C;
  const void *D;
  int64_t m, n, ldc;
  int32_t k_iters;      // ceil(k / BLOCK_K)
  int32_t total_iters;  // batch * k_iters
  int32_t split_k;      // cluster size along the reduction (1, 2 or 4)
  int32_t beta0, bin_kind, bin_mode, relu;
  int32_t c_vec_ok;     // C base 16B aligned and ldc % 8 == 0
  int32_t b_early;      // B (and D) do not depend on in-flight kernels: fetch B before the PDL wait
  unsigned int *flags;  // split-K arrival counters per tile (flag-synchronised exchange, SPLITK == 3)
  float *ws;            // split-K exchange through L2: [tile][owner][src][128][64/S] f32 (SPLITK == 2)
  unsigned long long *trace;   // TPP_XSMM_TC_TRACE: per-CTA clock stamps (nullptr in normal runs)
};

constexpr int TRACE_SLOTS = 16;
// stamp slot `slot` of this CTA's trace row with the SM clock (slot 0 additionally gets %globaltimer in slot 15)
__device__ __forceinline__ void trace_stamp(const TcParams &p, int slot) {
  if (p.trace) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    p.trace[(size_t)cta * TRACE_SLOTS + slot] = clock64();
    if (slot == 0 || slot == 2 || slot == 11) {   // wall-clock (ns) of CTA start / PDL wait passed / CTA end
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.trace[(size_t)cta * TRACE_SLOTS + (slot == 0 ? 15 : slot == 2 ? 13 : 14)] = gt;
    }
  }
}

template <int BLOCK_N> struct SmemLayout {
  static constexpr int kBChunks = BLOCK_N / 64;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBChunks * B_CHUNK_BYTES;
};

// Fused epilogue on NC consecutive f32 accumulator columns of one row: (+C) -> binary(D) -> relu -> bf16.
template <int NC>
__device__ __forceinline__ void epilogue_store(float (&v)[NC], const TcParams &p, int64_t row, int64_t col0,
                                               const float *bias_pref = nullptr) {
  const uint16_t *Dp = static_cast<const uint16_t *>(p.D);
  uint16_t *crow = static_cast<uint16_t *>(p.C) + row * p.ldc + col0;
  const bool full = (col0 + NC <= p.n);
  if (!p.beta0) {
    if (full && p.c_vec_ok) {
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        const uint4 cv = *reinterpret_cast<const uint4 *>(crow + g * 8);
        const uint32_t w[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          v[g * 8 + 2 * h] += __uint_as_float(w[h] << 16);
          v[g * 8 + 2 * h + 1] += __uint_as_float(w[h] & 0xffff0000u);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < NC; ++e)
        if (col0 + e < p.n) v[e] += bf16_bits_to_f32(crow[e]);
    }
  }
  if (p.bin_kind) {
    if (bias_pref) {                                            // bias was prefetched during the main loop
#pragma unroll
      for (int e = 0; e < NC; ++e) v[e] += bias_pref[e];
    } else if (p.bin_mode == kBcastCol && p.bin_kind == 1 && full) {   // the MLP case: bias vector add
#pragma unroll
      for (int e = 0; e < NC; ++e) v[e] += bf16_bits_to_f32(__ldg(Dp + col0 + e));
    } else {
#pragma unroll
      for (int e = 0; e < NC; ++e) {
        if (col0 + e < p.n) {
          const int64_t di = p.bin_mode == kBcastCol   ? col0 + e
                             : p.bin_mode == kBcastRow ? row
                             : p.bin_mode == kBcastNone ? row * p.ldc + col0 + e
                                                        : 0;
          const float d = bf16_bits_to_f32(__ldg(Dp + di));
          v[e] = p.bin_kind == 1 ? v[e] + d : p.bin_kind == 2 ? v[e] * d : p.bin_kind == 3 ? v[e] - d : v[e] / d;
        }
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int e = 0; e < NC; ++e) v[e] = relu_f32(v[e]);
  }
  if (full && p.c_vec_ok) {
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) {
      uint4 o;
      o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
      o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
      o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
      o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
      *reinterpret_cast<uint4 *>(crow + g * 8) = o;
    }
  } else {
#pragma unroll
    for (int e = 0; e < NC; ++e)
      if (col0 + e < p.n) crow[e] = f32_to_bf16_bits(v[e]);
  }
}

// VNNI-2 converter warps (mlp_chain_pair.cu, mlp_chain_ft.cu): lanes 2p, 2p + 1 of a half-warp rewrite the 8 columns
// [8 g, 8 g + 8) of one raw k-pair row (pieces 2 g, 2 g + 1 of its sixteen 16-byte pieces) into chunk g of two swizzled
// tile rows. g as a function of the lane: pair j of quarter Q takes j + 4 ((j & 1) ^ Q) - {0,5,2,7} and {4,1,6,3}.
__device__ __forceinline__ int vnni_group_of_lane(int lane) {
  const int j = (lane >> 1) & 3, q = (lane >> 3) & 1;
  return j + 4 * ((j & 1) ^ q);
}

// ---- host helpers (tc_host.cu) ---------------------------------------------------------------------------------
// 3-D bf16 tensor map: dims (inner, rows, batch), strides in elements for rows and batch; 128-byte swizzle
bool encode_map(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t ld,
                uint64_t stride, uint32_t box_inner, uint32_t box_rows, uint32_t box_batch = 1);
// 4-D bf16 tensor map over an activation matrix: (k within a 64-wide k-block, row, k-block, batch element)
bool encode_map_x4(CUtensorMap *map, const void *base, uint64_t k, uint64_t rows, uint64_t batch, uint64_t ld,
                   uint64_t stride, uint32_t box_rows, uint32_t box_kb, uint32_t box_b);
// generic bf16 tensor map of rank <= 5: dims / strides (elements; strides[0] is the stride of dims[1]) / box
bool encode_map_nd(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides,
                   const uint32_t *box, int swizzle_bytes);
// rank-4 map over 4-byte units (dtype-agnostic data movement): dims[0] in 4-byte units, strides in BYTES, no swizzle
bool encode_map_u32_4d(CUtensorMap *map, const void *base, const uint64_t *dims, const uint64_t *strides_bytes,
                       const uint32_t *box);
int bin_mode_from_flags(int64_t f);

// Device memory owned by the graph being captured. Everything a captured kernel node reads or spins on (descriptor
// tables, arrival counters, split-K workspaces) is allocated here, written / zeroed on a private non-capturing stream
// BEFORE the node can ever run, and handed to the graph handle at xsmm_cuda_graph_end
// (brgemm_tc_take_capture_allocs), which frees it with the graph. Nothing a graph references is shared with direct
// launches, so no later launch can free or re-zero it under a replay.
cudaStream_t table_stream();
bool stream_is_capturing(cudaStream_t stream);
void *alloc_zeroed(size_t bytes);              // zero-filled, complete (not merely enqueued) on return
void *capture_owned_zeroed(size_t bytes);
void capture_adopt(void *p);                   // a device allocation the caller filled itself
void *capture_owned_table(const void *host, size_t bytes);   // device copy of a host table, complete on return
float *capture_owned_ws(size_t need);          // split-K exchange workspace of the capture in progress

// Kernels whose CTAs wait for each other through global-memory counters (flag-synchronised split-K, the chain kernels'
// layer barriers) need every CTA of the grid resident at once. They are launched COOPERATIVELY - the driver then
// guarantees co-residency even when other streams hold SMs, instead of a spin that could only trap - after an occupancy
// query has shown that the grid fits; when it does not, the caller takes a path without cross-CTA waits.
// prepare_resident_launch: puts the cooperative attribute into attrs[0] (replacing programmatic serialisation, the two do
// not combine) unless TPP_XSMM_COOP=0, and returns false if the grid cannot be co-resident on this device.
// `only_if_concurrent`: keep the ordinary (programmatic) launch while this process has only ever launched on ONE
// (thread, stream) - nothing else of ours can hold SMs then, and back-to-back launches keep their PDL overlap (a
// cooperative launch serialises them: cfg2 1130 -> 1037 TF/s) - and go cooperative from the first sign of a second one.
bool prepare_resident_launch(const void *kernel, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attrs,
                             bool only_if_concurrent = false);

void set_last_name(const char *fmt, ...);      // this thread's last tcgen05 launch (xsmm_cuda_last_kernel)
void note_extra_launch();                      // a helper kernel launched besides the one the caller counts

struct ByteRange { const char *lo, *hi; };
inline bool overlaps(const ByteRange &a, const ByteRange &b) { return a.lo < b.hi && b.lo < a.hi; }
inline ByteRange bf16_range(const void *p, int64_t elems) {
  const char *c = static_cast<const char *>(p);
  return ByteRange{c, c + elems * 2};
}
// byte ranges one layer (a plain invoke or a grid of tile invokes) reads and writes; B in VNNI-2 layout included
struct LayerRanges { ByteRange a, b, c, d; };
LayerRanges layer_ranges(const KernelDesc &d, const GemmArgs &g);
// operand footprints of one chain: inputs (first layer's A, every layer's B and D) and outputs (every layer's C)
void chain_ranges(const KernelDesc *const *descs, const GemmArgs *args, int L, std::vector<ByteRange> &in,
                  std::vector<ByteRange> &out);
// No layer's weights / bias overlap ANY layer's output, no two outputs overlap, the chain's input is not an output
bool chain_operands_hazard_free(const KernelDesc *const *descs, const GemmArgs *args, int L);

// debug traces (TPP_XSMM_TC_TRACE): buffers the kernels stamp, dumped by brgemm_tc_dump_trace()
constexpr int kTraceRing = 128, kTraceRingCtas = 256;
constexpr int FT_TRACE_SLOTS = 64;    // [0] CTA start, [1] PDL wait passed, [2] end, [8 + 6 p + e] events of pass p < 9
constexpr int PC_TRACE_SLOTS = 64;    // [4t+0] MMA tile start, [4t+1] MMAs issued, [4t+2] accumulator ready, [4t+3] tile stored
                                      // (t < 12); [48+2l], [49+2l] producer waits for / has layer l's input; [60] start, [61] end
extern unsigned long long *g_trace_buf;
extern int g_trace_next;
extern int g_trace_ctas[kTraceRing];
extern int g_chain_trace_ctas, g_chain_trace_layers;
extern bool g_chain_trace_ft;
extern unsigned long long *g_pc_trace;   // pair-per-chain kernel stamps (TPP_XSMM_TC_TRACE=4)
extern int g_pc_trace_ctas;
unsigned long long *trace_ring();   // allocates g_trace_buf on first use

// chain-kernel families (each in its own translation unit); shared shape limits
constexpr int CHAIN_MAX_LAYERS = 4;

} // namespace tc
} // namespace tpp
