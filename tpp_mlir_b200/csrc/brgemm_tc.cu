// brgemm_tc.cu - batch-reduce GEMM / fused BRGEMM on the sm_100a tensor cores.
//
// Replaces the JIT-ed libxsmm BRGEMM the reference dispatches in
// runtime/Xsmm/XsmmRunnerUtils.cpp:308-361 (brgemm) and :385-457 (fused brgemm,
// C = relu(C_in*beta + sum_b A_b*B_b + bias)) for bf16 operands.
//
// Design (B200-first, not a translation of the CPU microkernel):
//   * one CTA per 128 x BLOCK_N output tile; the reduction - every k-block of
//     every batch element - accumulates into ONE f32 accumulator in TMEM
//     ("batch-reduce" == one TMEM tile, many TMA stages);
//   * A (row-major [b][m][k], K-major for UMMA) and B (row-major [b][k][n],
//     MN-major for UMMA) are fetched by TMA through 3-D tensor maps
//     (k|n, m|k, batch) with 128-byte swizzle, so lda/ldb/stride_a/stride_b - the
//     BRGEMM "reduction stride" - live in the tensor map and out-of-bounds rows /
//     columns / k are zero-filled by hardware (any m, n, k works);
//   * warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one
//     elected thread) + TMEM allocator, warps 2..5 = epilogue (tcgen05.ld ->
//     beta / binary(D) / relu in f32 -> one RNE rounding -> 16-byte stores);
//   * full/empty mbarrier ring of STAGES smem slots between TMA and MMA,
//     tcgen05.commit releases a slot and finally signals the epilogue;
//   * when the output has too few tiles to fill 148 SMs (the 256 x 1024 MLP
//     layer has 32), the (batch x k) reduction is split across a thread-block
//     CLUSTER of 2 or 4 CTAs; partial accumulators are exchanged through
//     distributed shared memory (each CTA owns BLOCK_N/S columns of the tile,
//     receives the other CTAs' partials with st.shared::cluster pushes) and the
//     owner applies the fused epilogue - no HBM round trip, no second kernel;
//   * programmatic dependent launch: everything up to the first global access
//     (barrier init, TMEM alloc, tensor-map prefetch) overlaps the previous
//     kernel's tail (griddepcontrol).
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace tpp {

namespace {

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

// Split-K exchange (BLOCK_N == 64): cluster rank r owns columns [r*NC, (r+1)*NC) of the tile, NC = 64 / S.
// Every CTA pushes the slices it does not own into the owner's receive buffer, slot = sender rank:
//   recv[slot][row][NC f32], 16-byte chunks XOR-swizzled by the row so that both the remote stores and
//   the owner's loads are bank-conflict free.
template <int NC>
__device__ __forceinline__ uint32_t recv_offset(int slot, int row, int chunk) {
  constexpr int NCH = NC / 4;                       // 16-byte chunks per row
  const int sw = NCH == 4 ? ((row >> 1) & 3) : (row & (NCH - 1));
  return static_cast<uint32_t>(slot * (BLOCK_M * NC * 4) + row * (NC * 4) + ((chunk ^ sw) << 4));
}

template <int NC>
__device__ __forceinline__ void splitk_epilogue(const TcParams &p, uint32_t tmem_acc, uint32_t recv_base, int q,
                                                int lane, int64_t m0, int64_t n0, uint32_t rank, bool has_acc) {
  constexpr int S = 64 / NC;
  constexpr int NCH = NC / 4;
  const int row_in_tile = q * 32 + lane;
  float own[NC];
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t r[32];
    if (has_acc) {
      ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) r[e] = 0u;
    }
#pragma unroll
    for (int part = 0; part < 32 / NC; ++part) {     // the 32-column chunk holds 32/NC owner slices of NC columns
      const uint32_t owner = static_cast<uint32_t>(c / NC + part);
      if (owner == rank) {
#pragma unroll
        for (int e = 0; e < NC; ++e) own[e] = __uint_as_float(r[part * NC + e]);
      } else {
        const uint32_t remote = ptx::mapa(recv_base, owner);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
          ptx::st_cluster_v4(remote + recv_offset<NC>((int)rank, row_in_tile, j),
                             __uint_as_float(r[part * NC + 4 * j]), __uint_as_float(r[part * NC + 4 * j + 1]),
                             __uint_as_float(r[part * NC + 4 * j + 2]), __uint_as_float(r[part * NC + 4 * j + 3]));
      }
    }
  }
  const int64_t row = m0 + row_in_tile;
  const int64_t col0 = n0 + (int64_t)rank * NC;
  // bias for the owned columns: requested before the barrier so its latency hides behind it
  float bias[NC];
  const bool pref = p.bin_kind == 1 && p.bin_mode == kBcastCol && col0 + NC <= p.n;
  if (pref) {
    const uint16_t *Dp = static_cast<const uint16_t *>(p.D) + col0;
#pragma unroll
    for (int e = 0; e < NC; ++e) bias[e] = bf16_bits_to_f32(__ldg(Dp + e));
  }
  // all partials of this cluster have landed in their owners' shared memory
  if (threadIdx.x == 64) trace_stamp(p, 8);
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (threadIdx.x == 64) trace_stamp(p, 9);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    if (static_cast<uint32_t>(s) == rank) continue;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float4 t;
      const uint32_t a = recv_base + recv_offset<NC>(s, row_in_tile, j);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(a));
      own[4 * j] += t.x; own[4 * j + 1] += t.y; own[4 * j + 2] += t.z; own[4 * j + 3] += t.w;
    }
  }
  if (row < p.m && col0 < p.n) epilogue_store<NC>(own, p, row, col0, pref ? bias : nullptr);
  if (threadIdx.x == 64) trace_stamp(p, 10);
}

// Split-K exchange through L2 (SPLITK == 2). DSMEM moves ~17 B/clk/SM (measured: 24 KiB in + 24 KiB out took
// ~4400 clk including the barrier); the SM<->L2 path is several times wider. Every CTA stores the slices it does
// not own to a small f32 workspace that stays L2-resident, laid out [tile][owner][src][16-byte chunk][row] so that
// one warp store / load instruction covers 512 contiguous bytes, meets its cluster at the cluster barrier (its
// release/acquire at cluster scope orders the global stores for the other CTAs of the cluster; no gpu-scope fence
// is needed), and the owner reads its S-1 incoming slices back - all loads in flight before the first add.
template <int NC>
__device__ __forceinline__ void splitk_epilogue_l2(const TcParams &p, uint32_t tmem_acc, int q, int lane, int64_t m0,
                                                   int64_t n0, uint32_t rank, bool has_acc) {
  constexpr int S = 64 / NC;
  constexpr int NCH = NC / 4;
  const int row_in_tile = q * 32 + lane;
  const int64_t row = m0 + row_in_tile;
  const int64_t col0 = n0 + (int64_t)rank * NC;
  const size_t tile = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
  float4 *ws_tile = reinterpret_cast<float4 *>(p.ws) + tile * (size_t)(S * S * NCH * BLOCK_M);
  float own[NC];
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t r[32];
    if (has_acc) {
      ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) r[e] = 0u;
    }
#pragma unroll
    for (int part = 0; part < 32 / NC; ++part) {
      const uint32_t owner = static_cast<uint32_t>(c / NC + part);
      if (owner == rank) {
#pragma unroll
        for (int e = 0; e < NC; ++e) own[e] = __uint_as_float(r[part * NC + e]);
      } else {
        float4 *dst = ws_tile + (size_t)(owner * S + rank) * NCH * BLOCK_M + row_in_tile;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
          dst[j * BLOCK_M] = make_float4(__uint_as_float(r[part * NC + 4 * j]), __uint_as_float(r[part * NC + 4 * j + 1]),
                                         __uint_as_float(r[part * NC + 4 * j + 2]),
                                         __uint_as_float(r[part * NC + 4 * j + 3]));
      }
    }
  }
  // bias for the owned columns: requested before the barrier so its latency hides behind it
  uint32_t bias_raw[NC / 2];
  const bool pref = p.bin_kind == 1 && p.bin_mode == kBcastCol && col0 + NC <= p.n &&
                    ((reinterpret_cast<uintptr_t>(p.D) + col0 * 2) & 15) == 0;
  if (pref) {
    const uint4 *Dp = reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(p.D) + col0);
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) {
      const uint4 w = __ldg(Dp + g);
      bias_raw[4 * g] = w.x; bias_raw[4 * g + 1] = w.y; bias_raw[4 * g + 2] = w.z; bias_raw[4 * g + 3] = w.w;
    }
  }
  if (threadIdx.x == 64) trace_stamp(p, 8);
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (threadIdx.x == 64) trace_stamp(p, 9);
  float4 in[(S - 1) * NCH];
#pragma unroll
  for (int k = 0; k < S - 1; ++k) {   // the S-1 other ranks, starting after our own (static register indices)
    const uint32_t s = (rank + 1 + k) & (S - 1);
    const float4 *src = ws_tile + (size_t)(rank * S + s) * NCH * BLOCK_M + row_in_tile;
#pragma unroll
    for (int j = 0; j < NCH; ++j) in[k * NCH + j] = __ldcg(src + j * BLOCK_M);
  }
#pragma unroll
  for (int k = 0; k < S - 1; ++k)
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      own[4 * j] += in[k * NCH + j].x; own[4 * j + 1] += in[k * NCH + j].y;
      own[4 * j + 2] += in[k * NCH + j].z; own[4 * j + 3] += in[k * NCH + j].w;
    }
  float bias[NC];
  if (pref) {
#pragma unroll
    for (int e = 0; e < NC / 2; ++e) {
      bias[2 * e] = __uint_as_float(bias_raw[e] << 16);
      bias[2 * e + 1] = __uint_as_float(bias_raw[e] & 0xffff0000u);
    }
  }
  if (row < p.m && col0 < p.n) epilogue_store<NC>(own, p, row, col0, pref ? bias : nullptr);
  if (threadIdx.x == 64) trace_stamp(p, 10);
}

// Split-K exchange through L2 for the wide tiles (BLOCK_N = 128 / 256, S = 2 or 4, NC = BLOCK_N / S >= 32 columns per
// owner). Same workspace layout as above. The owner does not keep its own slice in registers across the barrier:
// after the barrier it re-reads it from TMEM 32 columns at a time, adds the S-1 incoming slices and stores.
template <int BLOCK_N>
__device__ __forceinline__ void splitk_epilogue_l2_wide(const TcParams &p, uint32_t tmem_acc, int q, int lane,
                                                        int64_t m0, int64_t n0, uint32_t rank, bool has_acc,
                                                        unsigned tile_x = blockIdx.x, unsigned tiles_x = gridDim.x,
                                                        unsigned tile_y = blockIdx.y, bool flag_sync = false) {
  const int S = p.split_k;
  const int NC = BLOCK_N / S;          // columns per owner (>= 32)
  const int NCH = NC / 4;              // 16-byte chunks per owner row
  const int row_in_tile = q * 32 + lane;
  const int64_t row = m0 + row_in_tile;
  const size_t tile = tile_x + (size_t)tiles_x * tile_y;
  float4 *ws_tile = reinterpret_cast<float4 *>(p.ws) + tile * (size_t)(S * BLOCK_N / 4 * BLOCK_M);
  const uint32_t lane_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  // phase 1: every 32-column chunk this CTA does not own goes to its owner's slot [owner][src = rank]
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += 32) {
    const uint32_t owner = static_cast<uint32_t>(c / NC);
    if (owner == rank) continue;       // warp-uniform
    uint32_t r[32];
    if (has_acc) {
      ptx::tmem_ld_32x32(lane_addr + c, r);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) r[e] = 0u;
    }
    float4 *dst = ws_tile + ((size_t)(owner * S + rank) * NCH + (c % NC) / 4) * BLOCK_M + row_in_tile;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dst[j * BLOCK_M] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  }
  if (threadIdx.x == 64) trace_stamp(p, 8);
  if (!flag_sync) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  } else {
    // The S CTAs of this tile are NOT in one cluster (8-CTA clusters of pairs only fit 15 at a time on a B200,
    // measured): they meet at a monotonically increasing arrival counter in global memory instead. All of them
    // are co-resident (the launcher keeps such grids within one wave of 1-CTA-per-SM kernels), so spinning is safe.
    __threadfence();                                      // my partial sums are visible device-wide ...
    asm volatile("bar.sync 1, 128;" ::: "memory");        // ... for all 128 epilogue threads of this CTA
    if (threadIdx.x == 64) {
      unsigned int *cnt = p.flags + tile;
      const unsigned int old = atomicAdd(cnt, 1u);
      const unsigned int target = (old / (unsigned)S + 1u) * (unsigned)S;
      unsigned int seen, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
        if (++spins > (1u << 22)) __trap();   // co-residency assumption broken: fail loudly, never hang
      } while (seen < target);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  if (threadIdx.x == 64) trace_stamp(p, 9);
  // phase 2: owned columns, 32 at a time
#pragma unroll 1
  for (int c = 0; c < NC; c += 32) {
    const int64_t col0 = n0 + (int64_t)rank * NC + c;
    if (col0 >= p.n) break;            // warp-uniform
    float4 in[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < S - 1) {
        const uint32_t src_rank = (rank + 1 + k) & (S - 1);
        const float4 *src = ws_tile + ((size_t)(rank * S + src_rank) * NCH + c / 4) * BLOCK_M + row_in_tile;
#pragma unroll
        for (int j = 0; j < 8; ++j) in[k][j] = __ldcg(src + j * BLOCK_M);
      }
    }
    uint32_t r[32];
    if (has_acc) {
      ptx::tmem_ld_32x32(lane_addr + rank * NC + c, r);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) r[e] = 0u;
    }
    float v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < S - 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j] += in[k][j].x; v[4 * j + 1] += in[k][j].y; v[4 * j + 2] += in[k][j].z; v[4 * j + 3] += in[k][j].w;
        }
      }
    }
    if (row < p.m) epilogue_store<32>(v, p, row, col0);
  }
  if (threadIdx.x == 64) trace_stamp(p, 10);
}

// SPLITK: 0 = one CTA per tile, 1 = cluster split-K with DSMEM exchange, 2 = cluster split-K with L2 exchange
// MC = 1: 2 x 2 (x S) thread-block clusters with TMA multicast. The two CTAs of a cluster row (same m-tile) each
// fetch one 64-row half of the A stage and multicast it to both; the two CTAs of a cluster column (same n-tile)
// each fetch half of the B chunks and multicast them. Every SM then pulls only half of its operand bytes over its
// SM<->L2 link, which is what bounds the single-CTA kernel (24 % of tensor peak at 2048 x 1024 x 1024).
template <int BLOCK_N, int STAGES, int SPLITK, int MC = 0>
__global__ void __launch_bounds__(NUM_THREADS, (SPLITK && BLOCK_N == 64) ? 2 : 1)
brgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  using L = SmemLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t recv_base = smem_b + STAGES * L::kBChunks * B_CHUNK_BYTES;     // SPLITK only
  const uint32_t bar_base = recv_base + (SPLITK == 1 ? RECV_BYTES : 0);
  const uint32_t full_bar = bar_base;                 // STAGES x 8 bytes
  const uint32_t empty_bar = bar_base + STAGES * 8;   // STAGES x 8 bytes
  const uint32_t accum_bar = bar_base + 2 * STAGES * 8;
  const uint32_t tmem_slot = accum_bar + 8;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;

  // this CTA's share of the (batch x k-block) reduction
  uint32_t rank = 0;
  int32_t it_begin = 0, it_end = p.total_iters;
  // multicast geometry: cluster = (2 n-tiles, 2 m-tiles, S); rank in cluster = cx + 2*cy + 4*cz
  uint32_t cx = 0, cy = 0;
  uint16_t a_mask = 0, b_mask = 0, free_mask = 0;
  if constexpr (MC) {
    const uint32_t cr = ptx::cluster_ctarank();
    cx = cr & 1; cy = (cr >> 1) & 1;
    const uint32_t zbase = cr & ~3u;
    a_mask = static_cast<uint16_t>((1u << (zbase + 2 * cy)) | (1u << (zbase + 2 * cy + 1)));   // same m-tile: both cx
    b_mask = static_cast<uint16_t>((1u << (zbase + cx)) | (1u << (zbase + cx + 2)));           // same n-tile: both cy
    free_mask = a_mask | b_mask;   // the CTAs whose producers write into this CTA's stages
  }
  if constexpr (SPLITK) {
    rank = blockIdx.z;             // gridDim.z == cluster z extent == split_k
    it_begin = (int32_t)(((int64_t)p.total_iters * rank) / p.split_k);
    it_end = (int32_t)(((int64_t)p.total_iters * (rank + 1)) / p.split_k);
  }
  const int32_t num_iters = it_end - it_begin;
  if (threadIdx.x == 0) trace_stamp(p, 0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      // with multicast a stage is refilled by this CTA and by its row / column partner: all three must have
      // seen their MMAs retire before anyone overwrites it
      ptx::mbar_init(empty_bar + 8 * s, MC ? 3 : 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, BLOCK_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  if constexpr (MC) {
    // partners' barriers must exist before the first multicast / remote arrive
    __syncwarp();
    ptx::cluster_arrive();
    ptx::cluster_wait();
  } else {
    __syncthreads();
  }
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  if (threadIdx.x == 0) trace_stamp(p, 1);

  // PDL: let the next kernel in the stream start its own prologue ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ... and, when the host knows that B was not produced by one of the kernels that may still be
  // running (MLP weights), fetch the B tiles of the first STAGES k-blocks already now: they overlap
  // the previous layer's epilogue instead of sitting on this layer's critical path.
  int32_t b_prefetched = 0;
  if (p.b_early && !MC) b_prefetched = num_iters < STAGES ? num_iters : STAGES;
  if (warp == 0 && lane == 0) {
    for (int32_t i = 0; i < b_prefetched; ++i) {
      const int32_t it = it_begin + i;
      const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
      ptx::mbar_arrive_expect_tx(full_bar + 8 * i, L::kStageBytes);
#pragma unroll
      for (int c = 0; c < L::kBChunks; ++c)
        ptx::tma_load_3d(smem_b + (i * L::kBChunks + c) * B_CHUNK_BYTES, &tmB, full_bar + 8 * i, n0 + c * 64,
                         kb * BLOCK_K, b);
    }
  }
  // wait until everything the previous kernels wrote (our A operand is the previous layer's C) is
  // visible before the first dependent global access.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) trace_stamp(p, 2);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int32_t i = 0; i < num_iters; ++i) {
        const int32_t it = it_begin + i;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
        if (i >= b_prefetched) {
          ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
          ptx::mbar_arrive_expect_tx(full_bar + 8 * s, L::kStageBytes);
        }
        if constexpr (MC) {
          // my half of A (64 rows) to both CTAs of my cluster row; my half of the B chunks to my cluster column
          ptx::tma_load_3d_mc(smem_a + s * A_STAGE_BYTES + cx * (A_STAGE_BYTES / 2), &tmA, full_bar + 8 * s,
                              kb * BLOCK_K, m0 + (int32_t)cx * (BLOCK_M / 2), b, a_mask);
#pragma unroll
          for (int c = 0; c < L::kBChunks / 2; ++c) {
            const int cc = (int)cy * (L::kBChunks / 2) + c;
            ptx::tma_load_3d_mc(smem_b + (s * L::kBChunks + cc) * B_CHUNK_BYTES, &tmB, full_bar + 8 * s, n0 + cc * 64,
                                kb * BLOCK_K, b, b_mask);
          }
        } else {
          ptx::tma_load_3d(smem_a + s * A_STAGE_BYTES, &tmA, full_bar + 8 * s, kb * BLOCK_K, m0, b);
          if (i >= b_prefetched) {
#pragma unroll
            for (int c = 0; c < L::kBChunks; ++c)
              ptx::tma_load_3d(smem_b + (s * L::kBChunks + c) * B_CHUNK_BYTES, &tmB, full_bar + 8 * s, n0 + c * 64,
                               kb * BLOCK_K, b);
          }
        }
        if (i == 0) trace_stamp(p, 3);
      }
      trace_stamp(p, 4);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, BLOCK_N, /*A K-major*/ 0, /*B MN-major*/ 1);
      for (int32_t i = 0; i < num_iters; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        ptx::mbar_wait(full_bar + 8 * s, ph);
        ptx::tc_fence_after_sync();
        if (i == 0) trace_stamp(p, 5);
        const uint32_t a_addr = smem_a + s * A_STAGE_BYTES;
        const uint32_t b_addr = smem_b + s * L::kBChunks * B_CHUNK_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          // A: K-major, 8-row groups 1024 B apart; one UMMA_K slice = 32 B inside the swizzled row
          const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
          // B: MN-major, 64-column atoms B_CHUNK_BYTES apart (LBO), 8-k-row groups 1024 B apart (SBO);
          // one UMMA_K slice = 16 k-rows = 2048 B
          const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
          ptx::umma_bf16(tmem_acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        // frees the slot when these MMAs retire (in this CTA and, with multicast, at both partners)
        if constexpr (MC) ptx::umma_commit_mc(empty_bar + 8 * s, free_mask);
        else ptx::umma_commit(empty_bar + 8 * s);
      }
      if (num_iters > 0) ptx::umma_commit(accum_bar);   // accumulator complete
      trace_stamp(p, 6);
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    if (num_iters > 0) {
      ptx::mbar_wait(accum_bar, 0);
      ptx::tc_fence_after_sync();
    }
    if (threadIdx.x == 64) trace_stamp(p, 7);
    if constexpr (!SPLITK) {
      const int64_t row = (int64_t)m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        const int64_t col0 = (int64_t)n0 + c;
        if (col0 >= p.n) break;   // warp-uniform
        uint32_t r[32];
        if (num_iters > 0) {
          ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0u;
        }
        if (row < p.m) {
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
          epilogue_store<32>(v, p, row, col0);
        }
      }
    }
  }

  if constexpr (SPLITK) {
    // every thread of every CTA in the cluster reaches the cluster barrier inside / next to the exchange
    if (warp >= 2) {
      const int q = warp & 3;
      if constexpr (SPLITK == 1) {
        if (p.split_k == 4)
          splitk_epilogue<16>(p, tmem_acc, recv_base, q, lane, m0, n0, rank, num_iters > 0);
        else
          splitk_epilogue<32>(p, tmem_acc, recv_base, q, lane, m0, n0, rank, num_iters > 0);
      } else if constexpr (BLOCK_N == 64) {
        if (p.split_k == 4)
          splitk_epilogue_l2<16>(p, tmem_acc, q, lane, m0, n0, rank, num_iters > 0);
        else
          splitk_epilogue_l2<32>(p, tmem_acc, q, lane, m0, n0, rank, num_iters > 0);
      } else {
        splitk_epilogue_l2_wide<BLOCK_N>(p, tmem_acc, q, lane, m0, n0, rank, num_iters > 0);
      }
    } else {
      __syncwarp();   // lane 0 ran the producer / MMA loop; the cluster barrier is warp-aligned
      ptx::cluster_arrive();
      ptx::cluster_wait();
    }
  }

  ptx::tc_fence_before_sync();
  if constexpr (MC) {
    // partners may still multicast into / arrive on this CTA's shared memory until their main loops end
    __syncwarp();
    ptx::cluster_arrive();
    ptx::cluster_wait();
  }
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(p, 11);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, BLOCK_N);
  }
}

// ---- CTA-pair kernel (cta_group::2) ---------------------------------------------------------------------
// One 256 x BLOCK_N output tile per pair of CTAs (two SMs of a TPC): CTA r of the pair stages A rows
// [128r, 128r+128) and B columns [r*BLOCK_N/2, (r+1)*BLOCK_N/2) of every k-block in ITS shared memory; the leader
// (r = 0) issues tcgen05.mma.cta_group::2 with M = 256, which reads both CTAs' operands and accumulates rows
// 128r.. into CTA r's TMEM. Each SM therefore receives 16 KiB + BLOCK_N*64 B per k-block instead of
// 16 KiB + BLOCK_N*128 B: for BLOCK_N = 256 that is 32 KiB per 512 MMA clocks = 62 B/clk, inside what the SM<->L2
// link delivers, where the single-CTA 128 x 256 tile needs 94 B/clk (measured 61 % of tensor peak, link-bound).
// Cluster = (2, 1, S): the pair along x, optional split-K along z with the L2 workspace exchange.
template <int BLOCK_N, int STAGES, int SPLITK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
brgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  constexpr int HALF_N = BLOCK_N / 2;                 // B columns staged by each CTA
  constexpr int kBChunks = HALF_N / 64;
  constexpr int kStageBytes = A_STAGE_BYTES + kBChunks * B_CHUNK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_b + STAGES * kBChunks * B_CHUNK_BYTES;
  const uint32_t full_bar = bar_base;                 // used in the leader only (both CTAs' bytes land on it)
  const uint32_t empty_bar = bar_base + STAGES * 8;   // one per CTA, released by the pair's MMA commits
  const uint32_t accum_bar = bar_base + 2 * STAGES * 8;
  const uint32_t tmem_slot = accum_bar + 8;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();      // = peer + 2 * z
  const uint32_t peer = crank & 1;                    // 0 = leader
  const uint32_t leader_rank = crank & ~1u;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << leader_rank);
  const int32_t m0 = blockIdx.x * BLOCK_M;            // this CTA's 128 rows (blockIdx.x = 2 * pair + peer)
  const int32_t n0 = blockIdx.y * BLOCK_N;

  uint32_t rank = 0;
  int32_t it_begin = 0, it_end = p.total_iters;
  if constexpr (SPLITK) {
    rank = blockIdx.z;
    it_begin = (int32_t)(((int64_t)p.total_iters * rank) / p.split_k);
    it_end = (int32_t)(((int64_t)p.total_iters * (rank + 1)) / p.split_k);
  }
  const int32_t num_iters = it_end - it_begin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, BLOCK_N);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();   // both CTAs' barriers and TMEM exist before any remote signal / pair MMA
  ptx::cluster_wait();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own halves into own smem, bytes counted on the LEADER's full barrier =====
    if (lane == 0) {
      const uint32_t leader_full = ptx::mapa(full_bar, leader_rank);
      for (int32_t i = 0; i < num_iters; ++i) {
        const int32_t it = it_begin + i;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
        ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
        if (peer == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2 * kStageBytes);   // both CTAs' bytes
        ptx::tma_load_3d_pair(smem_a + s * A_STAGE_BYTES, &tmA, leader_full + 8 * s, kb * BLOCK_K, m0, b);
#pragma unroll
        for (int c = 0; c < kBChunks; ++c)
          ptx::tma_load_3d_pair(smem_b + (s * kBChunks + c) * B_CHUNK_BYTES, &tmB, leader_full + 8 * s,
                                n0 + (int32_t)peer * HALF_N + c * 64, kb * BLOCK_K, b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader only =====
    if (lane == 0 && peer == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BLOCK_M, BLOCK_N, /*A K-major*/ 0, /*B MN-major*/ 1);
      for (int32_t i = 0; i < num_iters; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        ptx::mbar_wait(full_bar + 8 * s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = smem_a + s * A_STAGE_BYTES;
        const uint32_t b_addr = smem_b + s * kBChunks * B_CHUNK_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
          const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
          ptx::umma_bf16_pair(tmem_acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit_pair(empty_bar + 8 * s, pair_mask);   // frees the slot in both CTAs
      }
      if (num_iters > 0) ptx::umma_commit_pair(accum_bar, pair_mask);   // both epilogues may start
    }
  } else {
    // ===== epilogue (both CTAs, each on its own 128 rows of TMEM) =====
    const int q = warp & 3;
    if (num_iters > 0) {
      ptx::mbar_wait(accum_bar, 0);
      ptx::tc_fence_after_sync();
    }
    if constexpr (!SPLITK) {
      const int64_t row = (int64_t)m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        const int64_t col0 = (int64_t)n0 + c;
        if (col0 >= p.n) break;
        uint32_t r[32];
        if (num_iters > 0) {
          ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0u;
        }
        if (row < p.m) {
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
          epilogue_store<32>(v, p, row, col0);
        }
      }
    }
  }

  if constexpr (SPLITK) {
    if (warp >= 2) {
      const int q = warp & 3;
      splitk_epilogue_l2_wide<BLOCK_N>(p, tmem_acc, q, lane, m0, n0, rank, num_iters > 0, blockIdx.x, gridDim.x,
                                       blockIdx.y, /*flag_sync=*/SPLITK == 3);
    } else if constexpr (SPLITK != 3) {
      __syncwarp();
      ptx::cluster_arrive();
      ptx::cluster_wait();
    }
  }

  // the peer must not exit (nor free TMEM) while the leader's MMAs still read its shared memory / write its TMEM
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_acc, BLOCK_N);
  }
}

// ---- fused chain kernel: L consecutive BRGEMM layers in ONE persistent launch ----------------------------
// SURVEY.md section 8(f) item 2. The batch-256 MLP layer is latency-bound: per layer ~0.9 us of
// kernel hand-off plus ~4.4 us of kernel, most of it waiting. When a captured invoke sequence (xsmm_cuda_graph_*)
// contains layers whose C is the next layer's A, the runtime launches this kernel instead: same tiling as the
// stand-alone kernel (128 x 64 tiles, 4-CTA split-K clusters, L2 exchange), but
//   * the weight tiles of ALL layers (iters_per_cta x 8 KiB per layer) are fetched at kernel start and stay in
//     shared memory: no weight traffic on the critical path of layers 1..L-1;
//   * layers are separated by a grid-wide arrival counter instead of a kernel boundary (all CTAs are co-resident:
//     <= 148 CTAs, 1 per SM), so there is no launch hand-off and no per-layer setup;
//   * TMEM, barriers and the A stages are allocated once.
constexpr int CHAIN_MAX_LAYERS = 4;
constexpr int CHAIN_IPC = 4;   // (batch x k-block) iterations per CTA and layer == A stages

struct ChainParams {
  CUtensorMap tmA[CHAIN_MAX_LAYERS], tmB[CHAIN_MAX_LAYERS];
  TcParams layer[CHAIN_MAX_LAYERS];
  unsigned int *grid_counter;   // monotonic arrival counter shared by all CTAs of this grid size
  int num_layers;
  int weights_early;            // no layer's weights / bias are written by in-flight kernels
  unsigned long long *trace;    // TPP_XSMM_TC_TRACE=1: clock stamps of layer 1 (nullptr in normal runs)
};

__device__ __forceinline__ void chain_stamp(const ChainParams &cp, int slot) {
  if (cp.trace) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    cp.trace[(size_t)cta * TRACE_SLOTS + slot] = clock64();
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                                     // CHAIN_IPC stages x 16 KiB
  const uint32_t smem_w = smem_base + CHAIN_IPC * A_STAGE_BYTES;         // L x CHAIN_IPC tiles x 8 KiB
  const uint32_t bar_base = smem_w + CHAIN_MAX_LAYERS * CHAIN_IPC * B_CHUNK_BYTES;
  const uint32_t a_full = bar_base;                                      // CHAIN_IPC
  const uint32_t w_full = bar_base + 8 * CHAIN_IPC;                      // CHAIN_MAX_LAYERS
  const uint32_t acc_bar = w_full + 8 * CHAIN_MAX_LAYERS;
  const uint32_t layer_bar = acc_bar + 8;
  const uint32_t tmem_slot = layer_bar + 8;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * 64;
  const uint32_t rank = blockIdx.z;                   // k-slice of this CTA (cluster = (1,1,4))
  const int L = cp.num_layers;
  const unsigned int G = gridDim.x * gridDim.y * gridDim.z;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < L; ++l) {
      ptx::prefetch_tensormap(&cp.tmA[l]);
      ptx::prefetch_tensormap(&cp.tmB[l]);
      ptx::mbar_init(w_full + 8 * l, 1);
    }
    for (int s = 0; s < CHAIN_IPC; ++s) ptx::mbar_init(a_full + 8 * s, 1);
    ptx::mbar_init(acc_bar, 1);
    ptx::mbar_init(layer_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  if (threadIdx.x == 0) chain_stamp(cp, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // weights of every layer: this CTA's k-slice (CHAIN_IPC k-blocks) x its 64 columns
  auto issue_weights = [&]() {
    for (int l = 0; l < L; ++l) {
      const TcParams &p = cp.layer[l];
      ptx::mbar_arrive_expect_tx(w_full + 8 * l, CHAIN_IPC * B_CHUNK_BYTES);
      for (int i = 0; i < CHAIN_IPC; ++i) {
        const int32_t it = (int32_t)rank * CHAIN_IPC + i;
        const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
        ptx::tma_load_3d(smem_w + (l * CHAIN_IPC + i) * B_CHUNK_BYTES, &cp.tmB[l], w_full + 8 * l, n0, kb * BLOCK_K, b);
      }
    }
  };
  if (warp == 0 && lane == 0 && cp.weights_early) issue_weights();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (warp == 0 && lane == 0 && !cp.weights_early) issue_weights();

  for (int l = 0; l < L; ++l) {
    const TcParams &p = cp.layer[l];
    const uint32_t par = l & 1;
    if (warp == 0) {
      // ===== producer: this layer's A k-slice (previous layer's output once the whole grid has stored it) =====
      if (lane == 0) {
        if (l > 0) {
          ptx::mbar_wait(layer_bar, (l - 1) & 1);
          // Y(l-1) was written by other SMs with generic-proxy stores, fenced at gpu scope before the arrival
          // counter moved and acquired by this CTA's thread 64: it is in L2, which is where TMA reads from.
          if (l == 1) chain_stamp(cp, 1);
        }
        for (int i = 0; i < CHAIN_IPC; ++i) {
          const int32_t it = (int32_t)rank * CHAIN_IPC + i;
          const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
          ptx::mbar_arrive_expect_tx(a_full + 8 * i, A_STAGE_BYTES);
          ptx::tma_load_3d(smem_a + i * A_STAGE_BYTES, &cp.tmA[l], a_full + 8 * i, kb * BLOCK_K, m0, b);
        }
        if (l == 1) chain_stamp(cp, 2);
      }
      __syncwarp();
      ptx::cluster_arrive();   // the split-K exchange barrier of this layer (all threads of the cluster)
      ptx::cluster_wait();
    } else if (warp == 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, 64, 0, 1);
        ptx::mbar_wait(w_full + 8 * l, 0);
        for (int i = 0; i < CHAIN_IPC; ++i) {
          ptx::mbar_wait(a_full + 8 * i, par);
          ptx::tc_fence_after_sync();
          if (l == 1 && i == 0) chain_stamp(cp, 3);
          if (l == 1 && i == CHAIN_IPC - 1) chain_stamp(cp, 4);
          const uint32_t a_addr = smem_a + i * A_STAGE_BYTES;
          const uint32_t b_addr = smem_w + (l * CHAIN_IPC + i) * B_CHUNK_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
            const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
            ptx::umma_bf16(tmem_acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(acc_bar);
      }
      __syncwarp();
      ptx::cluster_arrive();
      ptx::cluster_wait();
    } else {
      // ===== epilogue: split-K exchange + fused bias/ReLU/store, then the grid-wide layer barrier =====
      const int q = warp & 3;
      ptx::mbar_wait(acc_bar, par);
      ptx::tc_fence_after_sync();
      if (l == 1 && threadIdx.x == 64) chain_stamp(cp, 7);
      splitk_epilogue_l2<16>(p, tmem_acc, q, lane, m0, n0, rank, true);
      if (l + 1 < L) {
        ptx::tc_fence_before_sync();                       // TMEM reads done before the next layer's MMAs overwrite it
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all 128 epilogue threads have issued their Y(l) stores
        if (threadIdx.x == 64) {
          // one gpu-scope fence by the arriving thread: the CTA barrier above ordered the other threads' stores
          // before it (cumulativity), so they are visible device-wide before the counter moves
          __threadfence();
          const unsigned int old = atomicAdd(cp.grid_counter, 1u);
          const unsigned int target = (old / G + 1u) * G;
          unsigned int seen, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cp.grid_counter) : "memory");
            if (++spins > (1u << 22)) __trap();            // co-residency assumption broken: fail loudly, never hang
          } while (seen < target);
          if (l == 1) chain_stamp(cp, 11);
          if (l == 0) chain_stamp(cp, 5);
          ptx::mbar_arrive(layer_bar);                     // release the producer for layer l+1
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) chain_stamp(cp, 12);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 64);
  }
}

// ---- feature-major chain kernel: no split-K, one 16-CTA software barrier per layer ---------------------------
// Second design of the fused chain, replacing the split-K clusters above for the MLP shape:
// the split-K chain spends more than half of a layer in its two synchronisations (cluster barrier around the f32
// partial exchange, grid barrier around the layer). Here one CTA owns a (32 batch rows) x (64 features) output
// tile over the FULL reduction and computes it transposed, D^T[64 features x 32 rows] = W^T x X^T with
// tcgen05.mma M = 64, N = 32 (A = the weight tile exactly as TMA delivers it, MN-major; B = the activation tile,
// K-major), so
//   * there is no partial-sum exchange at all: one accumulator (32 TMEM columns), one rounding, one store;
//   * a batch tile's n/64 CTAs only depend on each other (rows of the MLP are independent): the layer boundary is
//     an arrival counter per batch tile (8 independent groups of 16 CTAs), not a grid-wide barrier;
//   * TMEM lane = feature, so the bias is one scalar per thread and all four epilogue warps have work;
//   * operands move in few, large TMA boxes: the activation slice through a 4-D map (k-in-block, row, k-block,
//     batch) whose box covers 4 k-blocks x 32 rows = 16 KiB, the weights in 32 KiB boxes (the measured cost of one
//     TMA issue + barrier hand-off is ~150-600 clk, far more than the 100 clk of MMA work per k-block);
//   * the kernel runs a list of PASSES (one pass = one layer of one chain). The operands of pass p+1 stream into
//     the 16 weight / activation slots as pass p's MMAs retire them (tcgen05.commit per group of 4 slots), i.e.
//     under the MMAs, the epilogue and the barrier latency of pass p;
//   * several INDEPENDENT chains captured in one graph (the benchmark's rotating operand sets, or any batch of
//     forward passes on different buffers) become ONE launch whose pass list interleaves two chains (A.L0, B.L0,
//     A.L1, B.L1, ...): while chain A's layer output travels store -> fence -> counter -> poll (~2500 clk, most of
//     a layer when one chain runs alone), the tensor pipe works on chain B; there is no kernel boundary (1.5-1.8 us
//     of programmatic-launch hand-off) between forward passes, and the next chain's first-layer operands load
//     under the previous chain's last layer.
// Per pass and CTA the tensor pipe reads 192 KiB of operands from shared memory and TMA writes 192 KiB into it:
// at 128 B/clk that is ~3000 clk, the bound of this tiling (measured: tcgen05.mma time = operand bytes / 128 B/clk,
// scripts/probes/umma_rate.cu); L2 -> SM delivery of the same 192 KiB runs at ~53 B/clk per SM with all SMs pulling.
constexpr int FT_M = 64;                          // features per CTA  (UMMA M)
constexpr int FT_N = 32;                          // batch rows per CTA (UMMA N)
constexpr int FT_KB = 16;                         // k-block slots: (batch x k) reduction of at most 16 x 64
constexpr int FT_GROUP = 4;                       // k-blocks per TMA box / barrier
constexpr int FT_NG = FT_KB / FT_GROUP;
constexpr int FT_X_BYTES = FT_N * BLOCK_K * 2;    // 4 KiB
constexpr int FT_W_BYTES = BLOCK_K * FT_M * 2;    // 8 KiB
constexpr int FT_CTR_STRIDE = 32;                 // one 128-byte line per batch-tile counter
constexpr int FT_CTR_SLOT = 160 * FT_CTR_STRIDE;  // counters of chain slot s start at s * FT_CTR_SLOT
constexpr int FT_MAX_WAYS = 4;                    // chains interleaved in one launch (= counter slots)
constexpr int FT_MAX_PASSES = 64;

struct alignas(64) FtPass {
  CUtensorMap tmX, tmW;
  void *C;
  const void *D;
  int64_t ldc;
  int32_t k_iters;          // k-blocks per batch element
  int32_t groups;           // (batch x k-blocks) / FT_GROUP
  uint8_t has_bias, relu;
  uint8_t arrive;           // a later pass reads this pass's output: arrive on the slot's counter after storing
  uint8_t x_dep;            // X is the output of an earlier pass of the same chain slot: wait for wait_arrivals
  uint8_t slot;             // chain slot (< FT_MAX_WAYS) = which counter set this pass's chain uses
  uint8_t pad[3];
  uint32_t wait_arrivals;   // arrivals per CTA on the slot's counter (this launch) that must be visible before X loads
};

struct FtParams {
  FtPass pass[FT_MAX_PASSES];
  unsigned int *counters;   // [slot][batch tile][FT_CTR_STRIDE]: monotonic, multiples of gridDim.x between launches
  int num_passes;
  int weights_early;        // no weight / bias is produced by in-flight kernels: fetch pass 0's before the PDL wait
  int x0_early;             // same for pass 0's activations
  int proxy_fence;
  int w_multicast;          // launched as (1,2,1) clusters: the two batch tiles of a cluster share each weight box
  // split-K-2 variant only: counter values are derived from a per-slot launch epoch instead of being read back
  unsigned int *epoch;      // [FT_MAX_WAYS] barriers completed per counter by earlier launches; [FT_MAX_WAYS] exit ticket
  unsigned int arrivals_total[FT_MAX_WAYS];   // barriers per counter this launch adds to each slot
  unsigned long long *trace;
};

constexpr int FT_TRACE_SLOTS = 64;    // [0] CTA start, [1] PDL wait passed, [2] end, [8 + 6 p + e] events of pass p < 9
__device__ __forceinline__ void ft_stamp(unsigned long long *trace, int slot) {
  if (trace) trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * FT_TRACE_SLOTS + slot] = clock64();
}
__device__ __forceinline__ void ft_stamp_pass(unsigned long long *trace, int p, int e) {
  if (trace && p < 9) trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * FT_TRACE_SLOTS + 8 + 6 * p + e] = clock64();
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_chain_ft_kernel(const __grid_constant__ FtParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // FT_KB x 4 KiB
  const uint32_t smem_w = smem_base + FT_KB * FT_X_BYTES;              // FT_KB x 8 KiB
  const uint32_t bar_base = smem_w + FT_KB * FT_W_BYTES;
  const uint32_t x_full = bar_base;                                    // [FT_NG]
  const uint32_t w_full = bar_base + 8 * FT_NG;                        // [FT_NG]
  const uint32_t w_empty = bar_base + 16 * FT_NG;                      // [FT_NG] group's X and W slots consumed
  const uint32_t acc_full = bar_base + 24 * FT_NG;                     // [2] accumulator (pass parity) complete
  const uint32_t acc_free = acc_full + 16;                             // [2] accumulator read out by the epilogue
  const uint32_t tmem_slot = acc_free + 16;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t n0 = blockIdx.x * FT_M;              // first feature of this CTA
  const int32_t m0 = blockIdx.y * FT_N;              // first batch row of this CTA
  const unsigned int G = gridDim.x;                  // CTAs per batch tile == arrivals per barrier
  unsigned int *counter0 = cp.counters + (size_t)blockIdx.y * FT_CTR_STRIDE;
  const int P = cp.num_passes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&cp.pass[0].tmX);
    ptx::prefetch_tensormap(&cp.pass[0].tmW);
    for (int g = 0; g < FT_NG; ++g) {
      ptx::mbar_init(x_full + 8 * g, 1);
      ptx::mbar_init(w_full + 8 * g, 1);
      ptx::mbar_init(w_empty + 8 * g, cp.w_multicast ? 2 : 1);   // multicast: both CTAs of the cluster retire a group
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 4);           // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * FT_N);            // two 32-column accumulators, alternating by pass
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  // weight multicast: the peer's barriers must exist before this CTA's first multicast box can complete on them
  const uint32_t crank = cp.w_multicast ? ptx::cluster_ctarank() : 0u;
  if (cp.w_multicast) ptx::cluster_sync();
  if (threadIdx.x == 0) ft_stamp(cp.trace, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== producer: the whole warp walks the (uniform) control flow, one elected lane issues =====
    // box coordinates (batch element, k-block) of group g's first slot, without integer division
    auto group_coords = [&](const FtPass &ps, int g, int32_t &b, int32_t &kb) {
      b = 0;
      kb = g * FT_GROUP;
      while (kb >= ps.k_iters) { kb -= ps.k_iters; ++b; }
    };
    auto issue_w = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        // every CTA expects the whole box on its own barrier; with multicast only cluster rank (g & 1) fetches it,
        // and the box lands in both CTAs' slots (same offsets) and completes on both CTAs' barriers
        ptx::mbar_arrive_expect_tx(w_full + 8 * g, FT_GROUP * FT_W_BYTES);
        if (!cp.w_multicast)
          ptx::tma_load_3d(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b);
        else if ((uint32_t)(g & 1) == crank)
          ptx::tma_load_3d_mc(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b,
                              (uint16_t)0x3);
      }
      __syncwarp();
    };
    auto issue_x = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(x_full + 8 * g, FT_GROUP * FT_X_BYTES);
        ptx::tma_load_4d(smem_x + g * (FT_GROUP * FT_X_BYTES), &ps.tmX, x_full + 8 * g, 0, m0, kb, b);
      }
      __syncwarp();
    };
    const bool x0_early = cp.weights_early && cp.x0_early && !cp.pass[0].x_dep;
    auto early_loads = [&]() {
      for (int g = 0; g < cp.pass[0].groups; ++g) {  // group by group: the MMAs start on the first 48 KiB
        issue_w(0, g);
        if (x0_early) issue_x(0, g);
      }
      // the next passes' weights: this CTA's share of the feature tile's slice goes to L2 now
      for (int p = 1; p < P && p < 3; ++p) {
        const FtPass &ps = cp.pass[p];
        for (int g = (int)blockIdx.y; g < ps.groups; g += (int)gridDim.y) {
          int32_t b, kb;
          group_coords(ps, g, b, kb);
          if (ptx::elect_one()) ptx::tma_prefetch_3d(&ps.tmW, n0, kb * BLOCK_K, b);
          __syncwarp();
        }
      }
    };
    if (cp.weights_early) early_loads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) ft_stamp(cp.trace, 1);
    if (!cp.weights_early) early_loads();
    // arrivals of this launch so far are < G on either counter (nobody passes a barrier without this CTA)
    unsigned int base[FT_MAX_WAYS];
#pragma unroll
    for (int sl = 0; sl < FT_MAX_WAYS; ++sl) base[sl] = (ld_acquire_gpu(counter0 + sl * FT_CTR_SLOT) / G) * G;
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      if (p + 1 < P && lane == 0) {
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmX);
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmW);
      }
      const unsigned int *ctr = counter0 + (int)ps.slot * FT_CTR_SLOT;
      unsigned int target = ps.wait_arrivals * G;
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl)
        if (sl == (int)ps.slot) target += base[sl];
      bool ready = !ps.x_dep;
      int x_next = (p == 0 && x0_early) ? ps.groups : 0;     // X groups issued so far
      for (int g = 0; g < ps.groups; ++g) {
        if (p > 0) ptx::mbar_wait(w_empty + 8 * g, (p - 1) & 1);   // slots of group g retired by pass p-1's MMAs
        if (p > 0) issue_w(p, g);                    // pass 0's weights were issued by early_loads()
        // relaxed (L2-coherent) polls: an acquire on every iteration costs a fence per poll. The data this flag
        // guards was fenced to L2 by its writers before they arrived, and it is only read by TMA (L2, never L1),
        // issued after the check - a control dependency the hardware does not speculate across.
        if (!ready) ready = (int)(ld_relaxed_gpu(ctr) - target) >= 0;   // one non-blocking look per group
        if (ready) {
          if (x_next == 0 && ps.x_dep) {
            if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
            if (lane == 0) ft_stamp_pass(cp.trace, p, 0);
          }
          for (; x_next <= g; ++x_next) issue_x(p, x_next);
        }
      }
      if (!ready) {
        unsigned int spins = 0;
        while ((int)(ld_relaxed_gpu(ctr) - target) < 0) {
          if (++spins > (1u << 22)) __trap();        // co-residency assumption broken: fail loudly, never hang
        }
        if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");   // generic stores (other SMs) -> TMA reads
        if (lane == 0) ft_stamp_pass(cp.trace, p, 0);
      }
      for (; x_next < ps.groups; ++x_next) issue_x(p, x_next);
      if (lane == 0) ft_stamp_pass(cp.trace, p, 1);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: uniform control flow, one elected lane issues =====
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(FT_M, FT_N, 1, 0);     // A (weights) MN-major, B (X) K-major
    // descriptors of slot 0 / k-step 0; every other (slot, k-step) is a constant added to the 14-bit address field
    const uint64_t da0 = ptx::umma_smem_desc_sw128(smem_w, FT_W_BYTES, 1024);
    const uint64_t db0 = ptx::umma_smem_desc_sw128(smem_x, 16, 1024);
    for (int p = 0; p < P; ++p) {
      const int NG = cp.pass[p].groups;
      const uint32_t par = p & 1;
      const uint32_t acc = tmem_acc + par * FT_N;
      if (p >= 2) {                                  // the epilogue of pass p-2 has read this accumulator out
        ptx::mbar_wait(acc_free + 8 * par, ((p >> 1) - 1) & 1);
        ptx::tc_fence_after_sync();
      }
#pragma unroll
      for (int g = 0; g < FT_NG; ++g) {
        if (g < NG) {
          ptx::mbar_wait(w_full + 8 * g, par);
          ptx::mbar_wait(x_full + 8 * g, par);
          ptx::tc_fence_after_sync();
          if (g == 0 && lane == 0) ft_stamp_pass(cp.trace, p, 2);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < FT_GROUP; ++j) {
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                const uint64_t da = da0 + (uint64_t)(((g * FT_GROUP + j) * FT_W_BYTES + kk * (UMMA_K * 128)) >> 4);
                const uint64_t db = db0 + (uint64_t)(((g * FT_GROUP + j) * FT_X_BYTES + kk * (UMMA_K * 2)) >> 4);
                ptx::umma_bf16(acc, da, db, idesc, (g > 0 || j > 0 || kk > 0) ? 1u : 0u);
              }
            }
            // the group's slots may be refilled with the next pass's tiles (multicast: tell both CTAs of the cluster)
            if (cp.w_multicast) ptx::umma_commit_mc(w_empty + 8 * g, (uint16_t)0x3);
            else ptx::umma_commit(w_empty + 8 * g);
            if (g == NG - 1) ptx::umma_commit(acc_full + 8 * par);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lanes 32q + (0..15) hold features 16q + (0..15); columns = the 32 batch rows =====
    const int q = warp & 3;
    const int f = 16 * q + (lane & 15);
    const bool active = lane < 16;
    const int ep_tid0 = 64;                          // first epilogue thread: the one that arrives for the CTA
    auto load_bias = [&](int p) -> uint16_t {
      return (p < P && cp.pass[p].has_bias) ? __ldg(static_cast<const uint16_t *>(cp.pass[p].D) + n0 + f) : (uint16_t)0;
    };
    // this thread's bias of the first pass, requested before the wait when the parameters are not produced in flight
    uint16_t bias_next = 0;
    if (cp.weights_early) bias_next = load_bias(0);
    // everything before this point only read memory; no store may precede the previous kernel's completion
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!cp.weights_early) bias_next = load_bias(0);
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      const float bias = bf16_bits_to_f32(bias_next);
      bias_next = load_bias(p + 1);                  // in flight while this pass's accumulator completes
      const uint32_t par = p & 1;
      ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
      ptx::tc_fence_after_sync();
      if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 3);
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * FT_N, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
      if (active) {
        uint16_t *out = static_cast<uint16_t *>(ps.C) + (int64_t)m0 * ps.ldc + n0 + f;
        if (ps.relu) {
#pragma unroll
          for (int j = 0; j < FT_N; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(relu_f32(__uint_as_float(r[j]) + bias));
        } else {
#pragma unroll
          for (int j = 0; j < FT_N; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(__uint_as_float(r[j]) + bias);
        }
      }
      if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 4);
      if (ps.arrive) {
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all epilogue threads have issued their stores
        if (threadIdx.x == ep_tid0) {
          // one gpu-scope release by the arriving thread; the CTA barrier ordered the other threads' stores before it
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter0 + (int)ps.slot * FT_CTR_SLOT) : "memory");
          ft_stamp_pass(cp.trace, p, 5);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (cp.w_multicast) ptx::cluster_sync();           // the peer may still signal this CTA's barriers until it is done too
  if (threadIdx.x == 0) ft_stamp(cp.trace, 2);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 2 * FT_N);
  }
}

// ---- split-K variants (S = 2, 4) of the pass kernel -----------------------------------------------------------------
// A pass of the kernel above is bound by the bytes one SM receives (192 KiB at ~50 B/clk); the tensor pipe and even
// shared memory have slack. Here a cluster of S CTAs shares a (64 features) x (32 S batch rows) tile and splits the
// reduction S ways: per pass a CTA receives 128/S KiB of weights + 64 KiB of activations (S = 2: 128 KiB, S = 4:
// 96 KiB) and the S partial accumulators meet through distributed shared memory: each CTA owns 32 of the rows and
// pushes the other rows of its partial (8 KiB per peer) into the peers' receive buffers with st.async, whose bytes
// complete_tx on the RECEIVER's mbarrier - no global-memory round trip, no cluster-wide barrier, no release/acquire
// round trip (a release.cluster arrive after plain st.shared::cluster stores cost ~3000 clk per pass). Everything else
// (pass list, interleaved chains, slot retirement by tcgen05.commit, arrival counters) is unchanged; a consumer CTA
// (feature tile, batch tile, k-slice z) waits for the 16 CTAs that produce its slice of the features.
template <int S> struct FS {
  static constexpr int N = 32 * S;                  // batch rows per tile (UMMA N); a CTA stores 32 of them
  static constexpr int KB = FT_KB / S;              // k-block slots per CTA
  static constexpr int NG = 4;                      // groups (TMA boxes / barriers) per pass
  static constexpr int GROUP = KB / NG;             // k-blocks per group
  static constexpr int X_BYTES = N * BLOCK_K * 2;   // one k-block of activations
  static constexpr int RECV_BYTES = FT_M * 32 * 4;  // one peer's partial for my 32 rows: 64 features x 32 f32
  static constexpr int SMEM = KB * (X_BYTES + FT_W_BYTES) + 2 * (S - 1) * RECV_BYTES + (3 * NG + 8) * 8 + 16 + 1024;
};
constexpr int F2_THREADS = 352;                   // producer, MMA issuer, 4 finisher warps, 4 sender warps, arriver

template <int S>
__global__ void __launch_bounds__(F2_THREADS, 1) mlp_chain_fts_kernel(const __grid_constant__ FtParams cp) {
  constexpr int F2_N = FS<S>::N, F2_KB = FS<S>::KB, F2_NG = FS<S>::NG, F2_GROUP = FS<S>::GROUP;
  constexpr int F2_X_BYTES = FS<S>::X_BYTES, F2_RECV_BYTES = FS<S>::RECV_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // F2_KB x 8 KiB
  const uint32_t smem_w = smem_base + F2_KB * F2_X_BYTES;              // F2_KB x 8 KiB
  const uint32_t smem_recv = smem_w + F2_KB * FT_W_BYTES;              // [pass parity][sender rank slot] x 8 KiB
  const uint32_t bar_base = smem_recv + 2 * (S - 1) * F2_RECV_BYTES;
  const uint32_t x_full = bar_base;                                    // [F2_NG]
  const uint32_t w_full = bar_base + 8 * F2_NG;
  const uint32_t w_empty = bar_base + 16 * F2_NG;
  const uint32_t acc_full = bar_base + 24 * F2_NG;                     // [2]
  const uint32_t acc_free = acc_full + 16;                             // [2]
  const uint32_t xchg_full = acc_free + 16;                            // [2] the peer's partial has landed in recv[parity]
  const uint32_t recv_free = xchg_full + 16;                           // [2] (local) my finishers are done with recv[parity]
  const uint32_t tmem_slot = recv_free + 16;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t n0 = blockIdx.x * FT_M;              // first feature of this CTA
  const int32_t m0 = blockIdx.y * F2_N;              // first batch row of the pair's tile
  const uint32_t z = blockIdx.z;                     // k-slice of this CTA == its rank in the (1,1,S) cluster
  const unsigned int G = gridDim.x;                  // arrivals per barrier: gridDim.x/S feature tiles x S k-slices
  // the counter this CTA waits on: its batch tile, ITS k-slice of the next layer's reduction
  unsigned int *wait_ctr0 = cp.counters + (size_t)(blockIdx.y * S + z) * FT_CTR_STRIDE;
  // the counter this CTA arrives on: its batch tile, the k-slice its features belong to
  unsigned int *arrive_ctr0 = cp.counters + (size_t)(blockIdx.y * S + blockIdx.x / (gridDim.x / S)) * FT_CTR_STRIDE;
  const int P = cp.num_passes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&cp.pass[0].tmX);
    ptx::prefetch_tensormap(&cp.pass[0].tmW);
    for (int g = 0; g < F2_NG; ++g) {
      ptx::mbar_init(x_full + 8 * g, 1);
      ptx::mbar_init(w_full + 8 * g, 1);
      ptx::mbar_init(w_empty + 8 * g, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 8);           // one arrival per finisher and per sender warp
      ptx::mbar_init(xchg_full + 8 * b, 1);          // one expect_tx arrival (mine); the peers' st.async bytes complete it
      ptx::mbar_init(recv_free + 8 * b, 4);          // one arrival per finisher warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * F2_N);            // two 64-column accumulators, alternating by pass
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  ptx::cluster_sync();                               // the peer's barriers exist before anything is pushed to it
  if (threadIdx.x == 0) ft_stamp(cp.trace, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== producer =====
    // box coordinates (batch element, k-block) of group g's first k-block: global k-block index z * 8 + 2 g
    auto group_coords = [&](const FtPass &ps, int g, int32_t &b, int32_t &kb) {
      b = 0;
      kb = (int32_t)z * F2_KB + g * F2_GROUP;
      while (kb >= ps.k_iters) { kb -= ps.k_iters; ++b; }
    };
    auto issue_w = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(w_full + 8 * g, F2_GROUP * FT_W_BYTES);
        ptx::tma_load_3d(smem_w + g * (F2_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b);
      }
      __syncwarp();
    };
    auto issue_x = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(x_full + 8 * g, F2_GROUP * F2_X_BYTES);
        ptx::tma_load_4d(smem_x + g * (F2_GROUP * F2_X_BYTES), &ps.tmX, x_full + 8 * g, 0, m0, kb, b);
      }
      __syncwarp();
    };
    const bool x0_early = cp.weights_early && cp.x0_early && !cp.pass[0].x_dep;
    auto early_loads = [&]() {
      for (int g = 0; g < F2_NG; ++g) {
        issue_w(0, g);
        if (x0_early) issue_x(0, g);
      }
    };
    if (cp.weights_early) early_loads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) ft_stamp(cp.trace, 1);
    if (!cp.weights_early) early_loads();
    unsigned int base[FT_MAX_WAYS];
#pragma unroll
    for (int sl = 0; sl < FT_MAX_WAYS; ++sl) base[sl] = ld_acquire_gpu(cp.epoch + sl) * G;
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      if (p + 1 < P && lane == 0) {
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmX);
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmW);
      }
      const unsigned int *ctr = wait_ctr0 + (int)ps.slot * FT_CTR_SLOT;
      unsigned int target = ps.wait_arrivals * G;
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl)
        if (sl == (int)ps.slot) target += base[sl];
      bool ready = !ps.x_dep;
      int x_next = (p == 0 && x0_early) ? F2_NG : 0;
      for (int g = 0; g < F2_NG; ++g) {
        if (p > 0) ptx::mbar_wait(w_empty + 8 * g, (p - 1) & 1);
        if (p > 0) issue_w(p, g);
        if (!ready) ready = (int)(ld_relaxed_gpu(ctr) - target) >= 0;
        if (ready) {
          if (x_next == 0 && ps.x_dep) {
            if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
          }
          for (; x_next <= g; ++x_next) issue_x(p, x_next);
        }
      }
      if (!ready) {
        unsigned int spins = 0;
        while ((int)(ld_relaxed_gpu(ctr) - target) < 0) {
          if (++spins > (1u << 22)) __trap();        // co-residency assumption broken: fail loudly, never hang
        }
        if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
      }
      for (; x_next < F2_NG; ++x_next) issue_x(p, x_next);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(FT_M, F2_N, 1, 0);     // A (weights) MN-major, B (X) K-major
    const uint64_t da0 = ptx::umma_smem_desc_sw128(smem_w, FT_W_BYTES, 1024);
    const uint64_t db0 = ptx::umma_smem_desc_sw128(smem_x, 16, 1024);
    for (int p = 0; p < P; ++p) {
      const uint32_t par = p & 1;
      const uint32_t acc = tmem_acc + par * F2_N;
      if (p >= 2) {
        ptx::mbar_wait(acc_free + 8 * par, ((p >> 1) - 1) & 1);
        ptx::tc_fence_after_sync();
      }
#pragma unroll
      for (int g = 0; g < F2_NG; ++g) {
        ptx::mbar_wait(w_full + 8 * g, par);
        ptx::mbar_wait(x_full + 8 * g, par);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < F2_GROUP; ++j) {
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              const uint64_t da = da0 + (uint64_t)(((g * F2_GROUP + j) * FT_W_BYTES + kk * (UMMA_K * 128)) >> 4);
              const uint64_t db = db0 + (uint64_t)(((g * F2_GROUP + j) * F2_X_BYTES + kk * (UMMA_K * 2)) >> 4);
              ptx::umma_bf16(acc, da, db, idesc, (g > 0 || j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(w_empty + 8 * g);
          if (g == F2_NG - 1) ptx::umma_commit(acc_full + 8 * par);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: lanes 0..15 of quarter q hold features 16q + lane; columns = the tile's 64 batch rows.
    // This CTA finishes rows [32z, 32z+32). SENDER warps (6..9) push the other 32 columns of the partial into the
    // peer's receive buffer and signal it; FINISHER warps (2..5) add the peer's partial to their own 32 columns, apply
    // bias / ReLU, round once and store. Two warp sets, so that waiting for the peer never delays what the peer waits for.
    const int q = warp & 3;
    const int f = 16 * q + (lane & 15);
    const bool active = lane < 16;
    if (warp == 10) {
      // ===== arriver: publishes a pass's output for the finishers (they only bar.arrive), so the ~1100-clk gpu-scope
      // fence is off their critical path =====
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int p = 0; p < P; ++p) {
        const FtPass &ps = cp.pass[p];
        if (!ps.arrive) continue;
        asm volatile("bar.sync 1, 160;" ::: "memory");     // the 128 finisher threads have issued this pass's stores
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(arrive_ctr0 + (int)ps.slot * FT_CTR_SLOT) : "memory");
          ft_stamp_pass(cp.trace, p, 5);
        }
        __syncwarp();
      }
    } else if (warp >= 6) {
      for (int p = 0; p < P; ++p) {
        const uint32_t par = p & 1;
        ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
        ptx::tc_fence_after_sync();
        uint32_t oth[S - 1][32];
#pragma unroll
        for (int r = 1; r < S; ++r)                   // the 32 columns (rows of the tile) owned by cluster rank z ^ r
          ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * F2_N + 32 * (z ^ (uint32_t)r),
                             oth[r - 1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
        if (threadIdx.x == 192) ft_stamp_pass(cp.trace, p, 1);
        // Overwriting the peer's recv[par] (last used by pass p-2) needs no signal from the peer: this push waits
        // until MY finishers have consumed pass p-1, i.e. received the peer's push of pass p-1, which the peer only
        // sent after ITS finishers had consumed pass p-2 (same rule on its side). All local, no cross-SM release.
        if (p >= 1) ptx::mbar_wait(recv_free + 8 * ((p - 1) & 1), ((p - 1) >> 1) & 1);
        if (active) {
          // st.async: every 16-byte store carries its own completion (complete_tx on the peer's barrier); a
          // release.cluster arrive after plain st.shared::cluster stores cost ~3000 clk per pass here
#pragma unroll
          for (int r = 1; r < S; ++r) {
            const uint32_t peer = z ^ (uint32_t)r;     // the peer files my partial under slot r - 1 (it sees me as peer ^ r)
            const uint32_t remote = ptx::mapa(smem_recv + (par * (S - 1) + (r - 1)) * F2_RECV_BYTES + (uint32_t)f * 128u, peer);
            const uint32_t remote_bar = ptx::mapa(xchg_full + 8 * par, peer);
#pragma unroll
            for (int j = 0; j < 8; ++j)                // 16-byte chunks XOR-swizzled by the feature: no bank conflicts
              ptx::st_async_v4(remote + (uint32_t)((j ^ (f & 7)) << 4), remote_bar, oth[r - 1][4 * j],
                               oth[r - 1][4 * j + 1], oth[r - 1][4 * j + 2], oth[r - 1][4 * j + 3]);
          }
        }
        if (threadIdx.x == 192) ft_stamp_pass(cp.trace, p, 2);
      }
    } else {
      const int ep_tid0 = 64;
      auto load_bias = [&](int p) -> uint16_t {
        return (p < P && cp.pass[p].has_bias) ? __ldg(static_cast<const uint16_t *>(cp.pass[p].D) + n0 + f) : (uint16_t)0;
      };
      uint16_t bias_next = 0;
      if (cp.weights_early) bias_next = load_bias(0);
      asm volatile("griddepcontrol.wait;" ::: "memory");   // no store before the previous kernel has completed
      if (!cp.weights_early) bias_next = load_bias(0);
      for (int p = 0; p < P; ++p) {
        const FtPass &ps = cp.pass[p];
        const float bias = bf16_bits_to_f32(bias_next);
        bias_next = load_bias(p + 1);
        const uint32_t par = p & 1;
        ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
        ptx::tc_fence_after_sync();
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 3);
        uint32_t own[32];
        ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * F2_N + 32 * z, own);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
        const uint32_t recv = smem_recv + par * (S - 1) * F2_RECV_BYTES + (uint32_t)f * 128u;   // this feature's 32 f32
        if (threadIdx.x == ep_tid0) ptx::mbar_arrive_expect_tx(xchg_full + 8 * par, (S - 1) * F2_RECV_BYTES);
        ptx::mbar_wait(xchg_full + 8 * par, (p >> 1) & 1);
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 0);
        float v[32];
        if (active) {
          // partial of k-slice z ^ r sits in slot r - 1; summation order own + (z^1) + (z^2) + (z^3): fixed per CTA,
          // hence deterministic (S = 2: IEEE addition is commutative, both CTAs of a pair even round identically)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(own[j]);
#pragma unroll
          for (int r = 1; r < S; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                           : "r"(recv + (uint32_t)((r - 1) * F2_RECV_BYTES) + (uint32_t)((j ^ (f & 7)) << 4)));
              v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(recv_free + 8 * par);   // this warp has consumed recv[par] of pass p
        if (active) {
          uint16_t *out = static_cast<uint16_t *>(ps.C) + (int64_t)(m0 + 32 * (int32_t)z) * ps.ldc + n0 + f;
          if (ps.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(relu_f32(v[j] + bias));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(v[j] + bias);
          }
        }
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 4);
        if (ps.arrive) asm volatile("bar.arrive 1, 160;" ::: "memory");   // stores issued; the arriver warp publishes them
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();                               // the peer may still push to / signal this CTA until it is done too
  if (threadIdx.x == 0) {
    ft_stamp(cp.trace, 2);
    // launch epoch: the last CTA to leave publishes how many barriers every counter has completed
    const unsigned int n_ctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int ticket = atomicAdd(cp.epoch + FT_MAX_WAYS, 1u);
    if (ticket == n_ctas - 1) {
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl) cp.epoch[sl] += cp.arrivals_total[sl];
      cp.epoch[FT_MAX_WAYS] = 0;
      __threadfence();
    }
  }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 2 * F2_N);
  }
}

// ---- pair-per-chain kernel: one CTA pair runs a whole layer chain, many chains side by side ---------------------------
// Third design of the fused chain, for launches that carry MANY independent chains (a captured graph of the
// benchmark's rotating operand sets, a batch of requests, or the 256-row blocks of a large-batch MLP: rows are
// independent through all layers). The pass kernels above spread ONE layer over 128 SMs in 64 x 64 tiles: every SM then
// receives 128 KiB of operands per layer pass, 16 MiB per layer over all SMs for 2.5 MiB of unique data, and the
// chip-wide L2 -> SM throughput (~6300 B/clk) bounds a pass at ~2700 clk no matter how the latencies are hidden.
// Here a work item (one chain x one block of 256 batch rows) belongs to ONE pair of CTAs (two SMs of a TPC,
// tcgen05.mma.cta_group::2, M = 256) which walks the layers and, per layer, the 256-column output tiles:
//   * per tile and k-block each CTA stages its 128 activation rows (16 KiB) and HALF of the 256 weight columns (16 KiB);
//     a layer costs 4 MiB of L2 -> SM traffic per item instead of 16 MiB, every weight byte is fetched exactly once;
//   * CTA r only ever reads the activation rows it wrote itself (rows 128 r .. 128 r + 127 of the item), so a layer
//     boundary needs no cross-SM synchronisation at all: the epilogue thread that issues the CTA's TMA stores waits
//     for their completion and arrives on a LOCAL mbarrier per output tile; the CTA's producer waits for tile i / 4
//     before reduction step i of the next layer's first tile. No counters, no co-residency assumption, nothing to
//     spin on across SMs; the last epilogue of a layer hides behind 12 of the next tile's 16 reduction steps;
//   * the next layer's weights do not depend on anything: their box of a ring slot is always issued BEFORE that wait
//     (same mbarrier, expect_tx covers both operands);
//   * TMEM holds two 256-column accumulators: the epilogue of tile t (tcgen05.ld -> bias from shared memory -> ReLU ->
//     bf16 -> swizzled staging buffer -> TMA store) runs under the MMAs of tile t + 1;
//   * L2 eviction-priority hints keep the activations (re-read once per output tile) resident under the weight stream;
//   * pairs are independent: the grid is min(items, 74) pairs, pair p takes items p, p + pairs, ...
// Layer descriptors (three tensor maps + epilogue parameters per layer) live in a device table written once at capture.
// History of the measurements that shaped it: profiles/kernel_trace_r1.txt.
constexpr int PC_STAGES = 6;
constexpr int PC_BLOCK_N = 256;                       // output columns per tile (UMMA N)
constexpr int PC_HALF_N = PC_BLOCK_N / 2;             // weight columns staged by each CTA
constexpr int PC_W_CHUNKS = PC_HALF_N / 64;           // 64-column TMA boxes per CTA and k-block
constexpr int PC_STAGE_BYTES = A_STAGE_BYTES + PC_W_CHUNKS * B_CHUNK_BYTES;   // 32 KiB
constexpr int PC_ROWS = 2 * BLOCK_M;                  // batch rows per work item
constexpr int PC_OUT_COLS = 64;                       // columns per TMA store box (128 bytes: one swizzle row)
constexpr int PC_OUT_BYTES = BLOCK_M * PC_OUT_COLS * 2;   // 16 KiB staging buffer, two of them
constexpr int PC_BIAS_BYTES = PC_BLOCK_N * 2;         // one tile's bias slice, two of them
constexpr int PC_MAX_TILES = 16;                      // output tiles per layer (n <= 4096): one "stored" barrier each
constexpr int PC_SMEM = PC_STAGES * PC_STAGE_BYTES + 2 * PC_OUT_BYTES + 2 * PC_BIAS_BYTES +
                        (2 * PC_STAGES + 4 + PC_MAX_TILES) * 8 + 16 + 1024;

struct alignas(128) PcLayer {
  CUtensorMap tmX;          // activations: (k, row, batch element), box 64 x 128
  CUtensorMap tmW;          // weights: (n, k, batch element), box 64 x 64
  CUtensorMap tmC;          // output: (n, row, 1), box 64 x 128 (TMA store from the swizzled staging buffer)
  void *C;
  const void *D;            // bias vector or nullptr
  int64_t ldc;
  int32_t k_iters;          // k-blocks per batch element
  int32_t total_iters;      // batch x k_iters
  int32_t n_tiles;          // n / 256
  int32_t n;
  int32_t relu;
  int32_t pad[3];
};
struct PcItem {
  int32_t layer0, num_layers, row0, pad;
};
struct PcParams {
  const PcLayer *layers;
  const PcItem *items;
  int32_t num_items;
  int32_t prefetch_w;          // L2 prefetch of the next tile's weight boxes (TPP_XSMM_CHAIN_PAIR_PREFETCH=1)
  int32_t l2_hints;            // L2 eviction-priority hints on the TMA loads / stores (TPP_XSMM_CHAIN_PAIR_HINTS=0: off)
  unsigned long long *trace;   // TPP_XSMM_TC_TRACE=4: clock stamps of each CTA's first item (nullptr in normal runs)
};
constexpr int PC_TRACE_SLOTS = 64;   // [4t+0] MMA tile start, [4t+1] MMAs issued, [4t+2] accumulator ready, [4t+3] tile stored
                                     // (t < 12); [48+2l], [49+2l] producer waits for / has layer l's input; [60] start, [61] end
__device__ __forceinline__ void pc_stamp(const PcParams &cp, int slot) {
  if (cp.trace) cp.trace[(size_t)blockIdx.x * PC_TRACE_SLOTS + slot] = clock64();
}

__device__ __forceinline__ void tensormap_acquire(const void *map) {
  // the table was written by a host copy: make it visible to the tensor-map proxy of this SM before the first use
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_chain_pair_kernel(const PcParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                                       // PC_STAGES x 16 KiB
  const uint32_t smem_w = smem_base + PC_STAGES * A_STAGE_BYTES;           // PC_STAGES x 2 x 8 KiB
  const uint32_t smem_out = smem_w + PC_STAGES * PC_W_CHUNKS * B_CHUNK_BYTES;   // 2 x 16 KiB output staging
  const uint32_t smem_bias = smem_out + 2 * PC_OUT_BYTES;                  // 2 x 512 B: bias slice of tile t / t + 1
  const uint32_t bar_base = smem_bias + 2 * PC_BIAS_BYTES;
  const uint32_t full_bar = bar_base;                                      // leader's: both CTAs' bytes land on it
  const uint32_t empty_bar = bar_base + PC_STAGES * 8;                     // per CTA, released by the pair's MMA commits
  const uint32_t acc_full = bar_base + 2 * PC_STAGES * 8;                  // [2] per CTA: accumulator complete
  const uint32_t acc_free = acc_full + 16;                                 // [2] leader's: both epilogues have read it out
  const uint32_t tile_done = acc_free + 16;                                // [PC_MAX_TILES] per CTA: my rows of output tile j are stored
  const uint32_t tmem_slot = tile_done + 8 * PC_MAX_TILES;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t peer = ptx::cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PC_STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 8);            // one arrival per epilogue warp of both CTAs
    }
    for (int j = 0; j < PC_MAX_TILES; ++j) ptx::mbar_init(tile_done + 8 * j, 1);   // the thread that issues the TMA stores
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 2 * PC_BLOCK_N);  // all 512 columns: two accumulators
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();   // both CTAs' barriers and TMEM exist before any remote signal / pair MMA
  ptx::cluster_wait();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) pc_stamp(cp, 60);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows / own weight columns into own smem, bytes counted on the LEADER =====
    if (lane == 0) {
      const uint32_t leader_full = ptx::mapa(full_bar, 0);
      // L2 residency (measured with ncu before the hints: 1.34 GB of DRAM reads per launch for 1.01 GB of operands -
      // the weight stream evicted activations between their four re-reads): weights are used once -> evict_first;
      // activations are re-read once per output tile -> evict_last until the layer's last tile, whose read demotes them
      const uint64_t pol_first = ptx::l2_policy_evict_first(), pol_last = ptx::l2_policy_evict_last();
      const bool hints = cp.l2_hints != 0;
      const bool prefetch_w = cp.prefetch_w != 0;
      int s = 0;
      uint32_t ph = 0, done_ph = 0;                   // done_ph bit j: parity of tile_done[j]'s next phase
      for (int item = pair; item < cp.num_items; item += num_pairs) {
        const PcItem it = cp.items[item];
        const int32_t row0 = it.row0 + (int32_t)peer * BLOCK_M;
        for (int l = 0; l < it.num_layers; ++l) {
          const PcLayer *L = cp.layers + it.layer0 + l;
          tensormap_acquire(&L->tmX);
          tensormap_acquire(&L->tmW);
          const int32_t k_iters = L->k_iters, total = L->total_iters, n_tiles = L->n_tiles;
          int32_t ready = 0;                          // output tiles of layer l - 1 (my rows) known to be stored
          for (int32_t j = 0; j < n_tiles; ++j) {
            const int32_t wcol = j * PC_BLOCK_N + (int32_t)peer * PC_HALF_N;
            const uint64_t pol_x = (j + 1 < n_tiles) ? pol_last : pol_first;
            int32_t b = 0, kb = 0;
            for (int32_t i = 0; i < total; ++i) {
              ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
              if (peer == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2 * PC_STAGE_BYTES);   // both CTAs' bytes
              // the weights depend on nothing: their boxes go out before any wait for the previous layer
#pragma unroll
              for (int c = 0; c < PC_W_CHUNKS; ++c) {
                if (hints)
                  ptx::tma_load_3d_pair_hint(smem_w + (s * PC_W_CHUNKS + c) * B_CHUNK_BYTES, &L->tmW, leader_full + 8 * s,
                                             wcol + c * 64, kb * BLOCK_K, b, pol_first);
                else
                  ptx::tma_load_3d_pair(smem_w + (s * PC_W_CHUNKS + c) * B_CHUNK_BYTES, &L->tmW, leader_full + 8 * s,
                                        wcol + c * 64, kb * BLOCK_K, b);
              }
              if (prefetch_w && (j & 1) == 0 && j + 1 < n_tiles) {
                // the same k-rows of the NEXT tile's weight columns go to L2 now: DRAM sees runs of 1 KiB per row
                // (this pair's two tiles) instead of 512 B, and the odd tiles' weight loads hit L2
#pragma unroll
                for (int c = 0; c < PC_W_CHUNKS; ++c)
                  ptx::tma_prefetch_3d(&L->tmW, wcol + PC_BLOCK_N + c * 64, kb * BLOCK_K, b);
              }
              if (l > 0 && j == 0) {
                // reduction step i reads columns [64 i, 64 i + 64) of the previous layer's output = its tile i / 4:
                // only the last four steps of the first tile have to wait for the previous layer's last epilogue
                const int32_t need = (i * BLOCK_K) / PC_BLOCK_N;
                while (ready <= need) {
                  const bool last = ready + 1 == (total * BLOCK_K) / PC_BLOCK_N;
                  if (last && item == pair) pc_stamp(cp, 48 + 2 * l);
                  ptx::mbar_wait(tile_done + 8 * ready, (done_ph >> ready) & 1u);
                  if (last && item == pair) pc_stamp(cp, 49 + 2 * l);
                  done_ph ^= 1u << ready;
                  ++ready;
                  asm volatile("fence.proxy.async;" ::: "memory");
                }
              }
              if (hints)
                ptx::tma_load_3d_pair_hint(smem_a + s * A_STAGE_BYTES, &L->tmX, leader_full + 8 * s, kb * BLOCK_K, row0, b,
                                           pol_x);
              else
                ptx::tma_load_3d_pair(smem_a + s * A_STAGE_BYTES, &L->tmX, leader_full + 8 * s, kb * BLOCK_K, row0, b);
              if (++kb == k_iters) { kb = 0; ++b; }
              if (++s == PC_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader only =====
    if (lane == 0 && peer == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(PC_ROWS, PC_BLOCK_N, /*A K-major*/ 0, /*B MN-major*/ 1);
      const uint16_t pair_mask = 3;
      int s = 0;
      uint32_t ph = 0, t = 0;
      for (int item = pair; item < cp.num_items; item += num_pairs) {
        const PcItem it = cp.items[item];
        for (int l = 0; l < it.num_layers; ++l) {
          const PcLayer *L = cp.layers + it.layer0 + l;
          const int32_t total = L->total_iters, n_tiles = L->n_tiles;
          for (int32_t j = 0; j < n_tiles; ++j, ++t) {
            const uint32_t buf = t & 1;
            if (t >= 2) {                              // both epilogues have read tile t - 2 out of this accumulator
              ptx::mbar_wait_cluster(acc_free + 8 * buf, ((t >> 1) - 1) & 1);
              ptx::tc_fence_after_sync();
            }
            const uint32_t acc = tmem_acc + buf * PC_BLOCK_N;
            if (t < 12) pc_stamp(cp, 4 * t);
            for (int32_t i = 0; i < total; ++i) {
              ptx::mbar_wait(full_bar + 8 * s, ph);
              ptx::tc_fence_after_sync();
              const uint32_t a_addr = smem_a + s * A_STAGE_BYTES;
              const uint32_t b_addr = smem_w + s * PC_W_CHUNKS * B_CHUNK_BYTES;
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
                const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
                ptx::umma_bf16_pair(acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
              }
              ptx::umma_commit_pair(empty_bar + 8 * s, pair_mask);   // frees the slot in both CTAs
              if (++s == PC_STAGES) { s = 0; ph ^= 1; }
            }
            ptx::umma_commit_pair(acc_full + 8 * buf, pair_mask);    // both epilogues may start
            if (t < 12) pc_stamp(cp, 4 * t + 1);
          }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs, each on its own 128 rows / TMEM lanes) =====
    // TMEM lane = row: a thread owns one output row. Its bf16 results go to a 128-byte-swizzled staging buffer
    // (64 columns x 128 rows), which one thread hands to TMA as a store box: full 128-byte lines leave the SM instead of
    // 16-byte pieces of 32 different lines per warp instruction (measured: direct stores cost 15.7k clk per tile, twice
    // the tile's MMA time).
    const int q = warp & 3;
    const int r_in = q * 32 + lane;                   // row within this CTA's 128
    const uint32_t leader_acc_free = ptx::mapa(acc_free, 0);
    const uint32_t lane_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const bool issuer = threadIdx.x == 64;
    const uint64_t pol_first = ptx::l2_policy_evict_first(), pol_last = ptx::l2_policy_evict_last();
    const bool hints = cp.l2_hints != 0;
    const uint32_t row_off = (uint32_t)r_in * 128u;
    const uint32_t sw = (uint32_t)(r_in & 7);
    uint32_t t = 0, g = 0;                            // tiles / store boxes handled so far
    // The bias slice of a tile (256 bf16) is staged in shared memory one tile ahead: thread i fetches columns 2i, 2i+1
    // of the NEXT tile into a register before it starts on the current one and parks it in the other half of the
    // buffer afterwards, so no global-load latency (an L2 / HBM miss every time: bias slices are never reused by an SM)
    // sits inside the column loop. Measured before: 8 exposed misses per tile, epilogue 10-20k clk for 8k clk of MMAs.
    auto fetch_bias = [&](const void *D, int32_t j) -> uint32_t {
      return D ? __ldg(reinterpret_cast<const uint32_t *>(static_cast<const uint16_t *>(D) + (size_t)j * PC_BLOCK_N) + r_in) : 0u;
    };
    if (pair < cp.num_items) {
      const PcItem it0 = cp.items[pair];
      const uint32_t b0 = fetch_bias(cp.layers[it0.layer0].D, 0);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_bias + (uint32_t)r_in * 4u), "r"(b0) : "memory");
    }
    for (int item = pair; item < cp.num_items; item += num_pairs) {
      const PcItem it = cp.items[item];
      const int32_t row0 = it.row0 + (int32_t)peer * BLOCK_M;
      for (int l = 0; l < it.num_layers; ++l) {
        const PcLayer *L = cp.layers + it.layer0 + l;
        if (issuer) tensormap_acquire(&L->tmC);
        const void *Dp = L->D;
        const bool relu = L->relu != 0;
        const int32_t n_tiles = L->n_tiles;
        for (int32_t j = 0; j < n_tiles; ++j, ++t) {
          const uint32_t buf = t & 1;
          // next tile's bias: same layer / next layer / first layer of this pair's next item
          uint32_t bias_next = 0;
          if (j + 1 < n_tiles) bias_next = fetch_bias(Dp, j + 1);
          else if (l + 1 < it.num_layers) bias_next = fetch_bias(L[1].D, 0);
          else if (item + num_pairs < cp.num_items) bias_next = fetch_bias(cp.layers[cp.items[item + num_pairs].layer0].D, 0);
          ptx::mbar_wait(acc_full + 8 * buf, (t >> 1) & 1);
          ptx::tc_fence_after_sync();
          if (t < 12 && issuer) pc_stamp(cp, 4 * t + 2);
          const uint32_t bias_s = smem_bias + buf * PC_BIAS_BYTES;
#pragma unroll 1
          for (int c = 0; c < PC_BLOCK_N; c += PC_OUT_COLS, ++g) {
            const uint32_t sbuf = smem_out + (g & 1) * PC_OUT_BYTES;
            // the store box issued two boxes ago has been read out of this staging buffer
            if (issuer) ptx::bulk_wait_group_read<1>();
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int h = 0; h < PC_OUT_COLS; h += 32) {
              uint32_t r[32];
              ptx::tmem_ld_32x32(lane_addr + buf * PC_BLOCK_N + c + h, r);
              uint32_t bw[16];
#pragma unroll
              for (int u = 0; u < 4; ++u)             // warp-uniform address: a broadcast read
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bw[4 * u]), "=r"(bw[4 * u + 1]), "=r"(bw[4 * u + 2]), "=r"(bw[4 * u + 3])
                             : "r"(bias_s + (uint32_t)((c + h) * 2 + u * 16)));
              ptx::tmem_ld_wait();
              if (c + h == PC_BLOCK_N - 32) {         // the accumulator is in registers: hand it back to the MMA issuer
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(leader_acc_free + 8 * buf);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint32_t o[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  float lo = __uint_as_float(r[8 * u + 2 * w]), hi = __uint_as_float(r[8 * u + 2 * w + 1]);
                  if (Dp) {
                    lo += __uint_as_float(bw[4 * u + w] << 16);
                    hi += __uint_as_float(bw[4 * u + w] & 0xffff0000u);
                  }
                  if (relu) { lo = relu_f32(lo); hi = relu_f32(hi); }
                  o[w] = pack_bf16x2(lo, hi);
                }
                const uint32_t chunk = (uint32_t)(h / 8 + u);   // 16-byte chunk of the 128-byte row
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(sbuf + row_off + ((chunk ^ sw) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                             : "memory");
              }
            }
            ptx::fence_proxy_async();                 // my shared-memory writes -> the async proxy (TMA store)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (issuer) {
              // a layer output that the next layer re-reads four times stays in L2; the chain's result does not
              if (!hints) ptx::tma_store_3d(&L->tmC, sbuf, j * PC_BLOCK_N + c, row0, 0);
              else ptx::tma_store_3d_hint(&L->tmC, sbuf, j * PC_BLOCK_N + c, row0, 0,
                                          l + 1 < it.num_layers ? pol_last : pol_first);
              ptx::bulk_commit_group();
              if (c == 0 && j > 0 && l + 1 < it.num_layers) {
                // every store group but the one just committed is complete: tile j - 1 (my rows) is in L2
                ptx::bulk_wait_group<1>();
                ptx::mbar_arrive(tile_done + 8 * (j - 1));
              }
            }
          }
          // the other half of the bias buffer was last read during tile t - 1: every thread is past that
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_bias + (buf ^ 1u) * PC_BIAS_BYTES + (uint32_t)r_in * 4u), "r"(bias_next)
                       : "memory");
          if (t < 12 && issuer) pc_stamp(cp, 4 * t + 3);
        }
        if (l + 1 < it.num_layers && issuer) {
          // the layer's last tile: its stores are the only ones outstanding
          ptx::bulk_wait_group<0>();
          ptx::mbar_arrive(tile_done + 8 * (n_tiles - 1));
        }
      }
    }
    if (issuer) ptx::bulk_wait_group<0>();
  }

  // the peer must not exit (nor free TMEM) while the leader's MMAs still read its shared memory / write its TMEM
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (threadIdx.x == 0) pc_stamp(cp, 61);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_acc, 2 * PC_BLOCK_N);
  }
}

// ---- host side ----------------------------------------------------------------

constexpr int kTraceRing = 128, kTraceRingCtas = 256;
unsigned long long *g_trace_buf = nullptr;
int g_trace_next = 0;
int g_trace_ctas[kTraceRing] = {0};
int g_chain_trace_ctas = 0, g_chain_trace_layers = 0;
bool g_chain_trace_ft = false;
unsigned long long *g_pc_trace = nullptr;   // pair-per-chain kernel stamps (TPP_XSMM_TC_TRACE=4)
int g_pc_trace_ctas = 0;

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
bool encode_map(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t ld,
                uint64_t stride, uint32_t box_inner, uint32_t box_rows, uint32_t box_batch = 1) {
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

template <int BLOCK_N, int STAGES, int SPLITK> constexpr int smem_bytes() {
  return STAGES * SmemLayout<BLOCK_N>::kStageBytes + (SPLITK == 1 ? RECV_BYTES : 0) + (2 * STAGES + 1) * 8 + 16 + 1024;
}

// programmatic dependent launch of the next direct BRGEMM launch (false: plain stream order, see GemmArgs::pdl)
thread_local int t_pdl_allowed = 1;

template <int BLOCK_N, int STAGES, int SPLITK, int MC = 0>
void launch_cfg(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = smem_bytes<BLOCK_N, STAGES, SPLITK>();
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(brgemm_tc_kernel<BLOCK_N, STAGES, SPLITK, MC>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[na].val.programmaticStreamSerializationAllowed = t_pdl_allowed;
  ++na;
  if (SPLITK || MC) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = MC ? 2 : 1;
    attrs[na].val.clusterDim.y = MC ? 2 : 1;
    attrs[na].val.clusterDim.z = SPLITK ? (unsigned)p.split_k : 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, brgemm_tc_kernel<BLOCK_N, STAGES, SPLITK, MC>, tmA, tmB, p));
}

template <int BLOCK_N, int STAGES, int SPLITK>
void launch_cfg_pair(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + (BLOCK_N / 128) * B_CHUNK_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(brgemm_tc2_kernel<BLOCK_N, STAGES, SPLITK>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = t_pdl_allowed;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = SPLITK == 2 ? (unsigned)p.split_k : 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, brgemm_tc2_kernel<BLOCK_N, STAGES, SPLITK>, tmA, tmB, p));
}

int bin_mode_from_flags(int64_t f) {
  if (f & 4) return kBcastCol;
  if (f & 1) return kBcastRow;
  if (f & 16) return kBcastScalar;
  return kBcastNone;
}

} // namespace

// Shape-level eligibility, decided once at dispatch.
bool brgemm_tc_supported(const KernelDesc &d) {
  if (d.dtype != kBF16) return false;
  if (d.gemm_flags & (2048 | 8192)) return false;           // VNNI-B / VNNI-C: generic kernel
  if ((d.lda % 8) != 0 || (d.ldb % 8) != 0) return false;   // TMA strides are multiples of 16 bytes
  if (d.op != OpClass::Gemm && ((d.stride_a % 8) != 0 || (d.stride_b % 8) != 0)) return false;
  if (d.m > (1ll << 31) || d.n > (1ll << 31) || d.k > (1ll << 31)) return false;
  return true;
}

// ---- tile / split selection ----------------------------------------------------------------------------
// Clocks one CTA spends per 64-wide k-block: the MMA itself (M=128: BLOCK_N/2 clk per UMMA_K=16 step) or, more
// often, the SM's ingest of the A+B stage over the SM<->L2 link (~50 B/clk measured), whichever is larger.
static double kblock_clocks(int bn) {
  const double mma = 4.0 * bn / 2.0;
  const double ingest = (A_STAGE_BYTES + bn * 128.0) / 50.0;
  return mma > ingest ? mma : ingest;
}
static int split_for(int64_t tiles, int64_t total_iters) {
  int split = 1;
  while (split < 4 && tiles * (split * 2) <= 148 && total_iters >= 2 * (split * 2)) split *= 2;
  return split;
}
// Small cost model, evaluated per launch (the batch count is a runtime argument):
//   time ~ waves x [ (k-blocks per CTA) x clocks per k-block + split-K exchange ] ,
// the reduction may be split over a cluster of up to 4 CTAs while the grid stays within one wave (148 SMs).
// Wide tiles raise the arithmetic intensity per SM (the SM<->L2 link is the limiter: cfg2 went from 27 % to 61 %
// of tensor peak with 128x256 tiles), narrow tiles + split-K fill the machine when the output has few tiles
// (the 256 x 1024 MLP layer).
static void choose_tile(const KernelDesc &d, int64_t total_iters, int *bn_out, int *split_out, int *mc_out) {
  const int64_t tiles_m = (d.m + BLOCK_M - 1) / BLOCK_M;
  // Multicast clusters are implemented and parity-tested but OFF by default: measured on B200 they do not help
  // (cfg2 990 -> 404 TF/s because the cluster limit caps split-K at 2; cfg5 384 -> 375 TF/s). Multicast saves L2
  // reads, not the bytes each SM must receive over its own SM<->L2 link, and that link (~50-64 B/clk) is what
  // bounds this kernel; halving the per-SM bytes needs cta_group::2 MMAs (next step).
  static const bool mc_off = [] { const char *e = getenv("TPP_XSMM_MULTICAST"); return !(e && e[0] == '1'); }();
  int best = 64, best_split = 1, best_mc = 0;
  double best_cost = 1e300;
  for (int bn : {256, 128, 64}) {
    if (bn > 64 && d.n <= bn / 2) continue;
    const int64_t tiles_n = (d.n + bn - 1) / bn;
    const int64_t tiles = tiles_m * tiles_n;
    for (int mc = 0; mc <= 1; ++mc) {
      // 2 x 2 multicast clusters: wide tiles only, even tile counts, and at most 2-way split-K (cluster <= 8)
      if (mc && (mc_off || bn == 64 || (tiles_m & 1) || (tiles_n & 1))) continue;
      int split = split_for(tiles, total_iters);
      if (mc && split > 2) split = 2;
      const double waves = (double)((tiles * split + 147) / 148);
      const double per_cta_iters = (double)((total_iters + split - 1) / split);
      const double mma = 4.0 * bn / 2.0;
      const double ingest = (A_STAGE_BYTES + bn * 128.0) / (mc ? 2.0 : 1.0) / 50.0;
      double cost = per_cta_iters * (mma > ingest ? mma : ingest) + (mc ? 1000.0 : 0.0);
      if (split > 1) cost += 1500.0 + 2.0 * (BLOCK_M * bn * 4.0) * (split - 1) / split / 50.0;   // barrier + ws out/in
      cost *= waves;
      if (cost < best_cost) { best_cost = cost; best = bn; best_split = split; best_mc = mc; }
    }
  }
  // CTA pairs (cta_group::2): one 256 x bn tile per pair, each SM receives A (128 x 64) + half of B per k-block
  static const bool pair_on = [] { const char *e = getenv("TPP_XSMM_PAIR"); return !(e && e[0] == '0'); }();
  if (pair_on) {
    const int64_t tiles_m256 = (d.m + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    for (int bn : {256, 128}) {
      if (d.n <= bn / 2) continue;
      const int64_t ctas1 = 2 * tiles_m256 * ((d.n + bn - 1) / bn);
      int split = 1;
      while (split < 4 && ctas1 * (split * 2) <= 148 && total_iters >= 2 * (split * 2)) split *= 2;
      const double waves = (double)((ctas1 * split + 147) / 148);
      const double per_cta_iters = (double)((total_iters + split - 1) / split);
      const double mma = 4.0 * bn / 2.0;
      const double ingest = (A_STAGE_BYTES + bn * 64.0) / 50.0;
      double cost = per_cta_iters * (mma > ingest ? mma : ingest) + 1000.0;
      if (split > 1) cost += 1500.0 + 2.0 * (BLOCK_M * bn * 4.0) * (split - 1) / split / 50.0;
      cost *= waves;
      if (cost < best_cost) { best_cost = cost; best = bn; best_split = split; best_mc = 2; }
    }
  }
  *bn_out = best;
  *split_out = best_split;
  *mc_out = best_mc;
}

thread_local char t_last_name[64] = "brgemm_tc_bf16";

// ---- device memory owned by the graph being captured ----------------------------------------------------------
// Everything a captured kernel node reads or spins on (descriptor tables, arrival counters, split-K workspaces) is
// allocated here, written / zeroed on a private non-capturing stream BEFORE the node can ever run, and handed to the
// graph handle at xsmm_cuda_graph_end (brgemm_tc_take_capture_allocs), which frees it with the graph. Nothing a graph
// references is shared with direct launches, so no later launch can free or re-zero it under a replay.
thread_local std::vector<void *> t_capture_allocs;
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
struct CaptureWs { float *ptr = nullptr; size_t bytes = 0; };
thread_local CaptureWs t_capture_ws;
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

// Device scratch of the split-K exchange and the chain kernel's grid counters. One instance per (host thread, stream):
// launches of one thread on one stream are serialised and may share it; a thread that pipelines work over several
// streams (xsmm_cuda_stream_create + xsmm_cuda_set_stream) gets a private copy per stream, so kernels that overlap in
// time never share a workspace. Graphs bake these pointers in: replay a graph on the stream it was captured for.
struct StreamScratch {
  cudaStream_t stream = nullptr;
  float *ws = nullptr;
  size_t ws_bytes = 0;
  unsigned int *flags = nullptr;
};
StreamScratch &scratch_for(cudaStream_t stream) {
  thread_local std::vector<StreamScratch *> all;
  for (StreamScratch *s : all)
    if (s->stream == stream) return *s;
  all.push_back(new StreamScratch());
  all.back()->stream = stream;
  return *all.back();
}
const char *brgemm_tc_last_name() { return t_last_name; }

void brgemm_tc_configure(KernelDesc &d) {
  // the tile shape is chosen per launch (choose_tile); the descriptor only records the family
  d.block_n = 0;
  d.stages = 0;
  snprintf(d.name, sizeof(d.name), "brgemm_tc_bf16_128xNx64");
}

bool launch_brgemm_tc(const KernelDesc &d, const GemmArgs &g, cudaStream_t stream) {
  if (!aligned16(g.A) || !aligned16(g.B)) return false;
  t_pdl_allowed = g.pdl ? 1 : 0;
  const int64_t batch = g.batch;
  if (batch > (1ll << 31)) return false;
  const int32_t k_iters = (int32_t)((d.k + BLOCK_K - 1) / BLOCK_K);
  int block_n = 64, split = 1, mc = 0;
  choose_tile(d, batch * k_iters, &block_n, &split, &mc);
  {
    static const char *env_bn = getenv("TPP_XSMM_BLOCK_N");   // tuning overrides, read once
    static const char *env = getenv("TPP_XSMM_SPLITK");
    if (env_bn && (atoi(env_bn) == 64 || atoi(env_bn) == 128 || atoi(env_bn) == 256)) {
      block_n = atoi(env_bn);
      mc = 0;
      split = split_for(((d.n + block_n - 1) / block_n) * ((d.m + BLOCK_M - 1) / BLOCK_M), batch * k_iters);
    }
    if (env) split = atoi(env);
    if (split != 2 && split != 4) split = 1;
    if (mc == 1 && split > 2) split = 2;
  }
  // A tensor map is a pure function of (descriptor, operand address, batch, box): cache the encoded
  // pair per thread so steady-state invokes (the same memrefs over and over) skip the driver call.
  struct MapCacheEntry {
    const KernelDesc *desc = nullptr;
    const void *A = nullptr, *B = nullptr;
    int64_t batch = -1;
    int mc = -1;
    bool ok = false;
    CUtensorMap tmA, tmB;
  };
  constexpr int kMapCache = 256;
  thread_local MapCacheEntry t_maps[kMapCache];
  const uintptr_t ha = reinterpret_cast<uintptr_t>(g.A), hb = reinterpret_cast<uintptr_t>(g.B);
  MapCacheEntry &e = t_maps[((ha >> 7) ^ (ha >> 19) ^ (hb >> 9) ^ (hb >> 23) ^ (uintptr_t)batch) & (kMapCache - 1)];
  if (e.desc != &d || e.A != g.A || e.B != g.B || e.batch != batch || e.mc != mc) {
    const uint64_t nb = batch > 0 ? (uint64_t)batch : 1;
    e.desc = &d; e.A = g.A; e.B = g.B; e.batch = batch; e.mc = mc;
    // with multicast every CTA fetches one 64-row half of the A stage
    e.ok = encode_map(&e.tmA, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K,
                      mc == 1 ? BLOCK_M / 2 : BLOCK_M) &&
           encode_map(&e.tmB, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64,
                      BLOCK_K);
  }
  if (!e.ok) return false;
  const CUtensorMap &tmA = e.tmA, &tmB = e.tmB;

  TcParams p;
  p.C = g.C;
  p.D = g.D;
  p.m = d.m; p.n = d.n; p.ldc = d.ldc;
  p.k_iters = k_iters;
  p.total_iters = (int32_t)(batch * p.k_iters);
  p.beta0 = (d.gemm_flags & 4) != 0;
  p.bin_kind = (d.op == OpClass::FusedBrgemm && g.D) ? (int)d.binary_kind : 0;
  p.bin_mode = bin_mode_from_flags(d.binary_flags);
  p.relu = d.op == OpClass::FusedBrgemm && d.unary_kind == 5;
  p.c_vec_ok = aligned16(g.C) && (d.ldc % 8) == 0;
  static const bool b_early_off = [] { const char *e = getenv("TPP_XSMM_B_EARLY"); return e && e[0] == '0'; }();
  p.b_early = (g.b_independent && !b_early_off) ? 1 : 0;
  // TPP_XSMM_TC_TRACE=1 (debug): synchronous launch with per-CTA clock stamps, summary on stderr
  // TPP_XSMM_TC_TRACE=2: asynchronous, every launch stamps its own slot of a ring; xsmm_cuda_debug_dump_trace()
  // prints the wall-clock timeline (kernel overlap under PDL / graph replay)
  static const int trace_mode = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e ? atoi(e) : 0; }();
  static const bool trace_on = trace_mode != 0;
  unsigned long long *&trace_buf = g_trace_buf;
  constexpr int kTraceCtas = kTraceRingCtas;
  if (trace_on && !trace_buf) {
    TPP_CUDA_CHECK(cudaMalloc(&trace_buf, sizeof(unsigned long long) * kTraceRing * kTraceCtas * TRACE_SLOTS));
    TPP_CUDA_CHECK(cudaMemset(trace_buf, 0, sizeof(unsigned long long) * kTraceRing * kTraceCtas * TRACE_SLOTS));
  }
  p.trace = nullptr;

  p.split_k = split;

  dim3 grid((unsigned)((d.n + block_n - 1) / block_n), (unsigned)((d.m + BLOCK_M - 1) / BLOCK_M), (unsigned)split);
  if (mc == 2)   // CTA pairs: x = 2 * (256-row tiles) so that the two CTAs of a pair are cluster ranks 2i, 2i+1
    grid = dim3((unsigned)(2 * ((d.m + 2 * BLOCK_M - 1) / (2 * BLOCK_M))), (unsigned)((d.n + block_n - 1) / block_n),
                (unsigned)split);
  const int n_ctas = (int)(grid.x * grid.y * grid.z);
  // exchange path of the split-K partials: the L2 workspace (default; 23.1 us per MLP step) or DSMEM
  // (TPP_XSMM_XCHG=d; 26.7 us: st.shared::cluster moves only ~17 B/clk/SM)
  static const bool xchg_dsmem = [] { const char *e = getenv("TPP_XSMM_XCHG"); return e && e[0] == 'd'; }();
  p.ws = nullptr;
  p.flags = nullptr;
  if (split > 1 && (!xchg_dsmem || block_n != 64)) {
    // per-(thread, stream) workspace: launches on one stream are serialised and may share it
    const size_t need = (size_t)n_ctas * BLOCK_M * block_n * sizeof(float);
    const bool capturing = stream_is_capturing(stream);
    constexpr int kFlagTiles = 4096;
    if (mc == 2 && n_ctas / split > kFlagTiles) return false;
    if (capturing) {
      // a captured node never shares scratch with direct launches (which may grow = free theirs): graph-owned memory
      p.ws = capture_owned_ws(need);
      if (mc == 2) p.flags = static_cast<unsigned int *>(capture_owned_zeroed(sizeof(unsigned int) * (size_t)(n_ctas / split)));
    } else {
      StreamScratch &sc = scratch_for(stream);
      float *&ws = sc.ws;
      size_t &ws_bytes = sc.ws_bytes;
      if (need > ws_bytes) {
        if (ws) { TPP_CUDA_CHECK(cudaDeviceSynchronize()); TPP_CUDA_CHECK(cudaFree(ws)); }
        const size_t want = need < (8u << 20) ? (8u << 20) : need;
        TPP_CUDA_CHECK(cudaMalloc(&ws, want));
        ws_bytes = want;
      }
      p.ws = ws;
      // arrival counters of the flag-synchronised exchange: zeroed once (complete before the first launch), only ever
      // incremented; one region per S so that every counter is a multiple of S between launches
      unsigned int *&flags = sc.flags;
      if (mc == 2) {
        if (!flags) flags = static_cast<unsigned int *>(alloc_zeroed(sizeof(unsigned int) * 2 * kFlagTiles));
        p.flags = flags + (split == 4 ? kFlagTiles : 0);
      }
    }
  }
  if (trace_mode == 1 && n_ctas <= kTraceCtas) {
    TPP_CUDA_CHECK(cudaMemsetAsync(trace_buf, 0, sizeof(unsigned long long) * n_ctas * TRACE_SLOTS, stream));
    p.trace = trace_buf;
  } else if (trace_mode == 2 && n_ctas <= kTraceCtas) {
    const int slot = g_trace_next++ % kTraceRing;
    g_trace_ctas[slot] = n_ctas;
    p.trace = trace_buf + (size_t)slot * kTraceCtas * TRACE_SLOTS;
  }
  snprintf(t_last_name, sizeof(t_last_name), "brgemm_tc_bf16_%dx%dx64%s%s", mc == 2 ? 256 : 128, block_n,
           split == 1 ? "" : split == 2 ? "_splitk2" : "_splitk4", mc == 1 ? "_mc2x2" : mc == 2 ? "_2cta" : "");
  if (mc == 2) {
    if (block_n == 256) {
      if (split > 1) launch_cfg_pair<256, 6, 3>(tmA, tmB, p, grid, stream);
      else launch_cfg_pair<256, 6, 0>(tmA, tmB, p, grid, stream);
    } else {
      if (split > 1) launch_cfg_pair<128, 8, 3>(tmA, tmB, p, grid, stream);
      else launch_cfg_pair<128, 8, 0>(tmA, tmB, p, grid, stream);
    }
  } else
  switch (block_n) {
  case 256:
    if (mc == 1 && split > 1) launch_cfg<256, 4, 2, 1>(tmA, tmB, p, grid, stream);
    else if (mc == 1) launch_cfg<256, 4, 0, 1>(tmA, tmB, p, grid, stream);
    else if (split > 1) launch_cfg<256, 4, 2>(tmA, tmB, p, grid, stream);
    else launch_cfg<256, 4, 0>(tmA, tmB, p, grid, stream);
    break;
  case 128:
    if (mc == 1 && split > 1) launch_cfg<128, 6, 2, 1>(tmA, tmB, p, grid, stream);
    else if (mc == 1) launch_cfg<128, 6, 0, 1>(tmA, tmB, p, grid, stream);
    else if (split > 1) launch_cfg<128, 6, 2>(tmA, tmB, p, grid, stream);
    else launch_cfg<128, 6, 0>(tmA, tmB, p, grid, stream);
    break;
  default:
    if (split > 1 && xchg_dsmem) launch_cfg<64, 3, 1>(tmA, tmB, p, grid, stream);   // 105 KiB smem: two CTAs per SM
    else if (split > 1) launch_cfg<64, 4, 2>(tmA, tmB, p, grid, stream);            //  97 KiB smem: two CTAs per SM
    else launch_cfg<64, 8, 0>(tmA, tmB, p, grid, stream);
  }
  if (p.trace && trace_mode == 1) {
    static int dumps = 0;
    std::vector<unsigned long long> h((size_t)n_ctas * TRACE_SLOTS);
    TPP_CUDA_CHECK(cudaStreamSynchronize(stream));
    TPP_CUDA_CHECK(cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (dumps++ % 64 == 40) {   // a steady-state launch
      double avg[12] = {0};
      unsigned long long gmin = ~0ull, gmax = 0;
      for (int c = 0; c < n_ctas; ++c) {
        const unsigned long long *r = &h[(size_t)c * TRACE_SLOTS];
        for (int s = 1; s < 12; ++s) avg[s] += r[s] ? (double)(r[s] - r[0]) : 0.0;
        if (r[15] < gmin) gmin = r[15];
        if (r[15] > gmax) gmax = r[15];
      }
      fprintf(stderr, "tc-trace %s grid=(%u,%u,%u): CTA start spread %llu ns; avg clocks since CTA start:", d.name, grid.x,
              grid.y, grid.z, gmax - gmin);
      static const char *names[12] = {"", "setup", "pdl_wait", "tma1", "tma_all", "data1", "mma_issued", "acc_ready",
                                      "pushed", "cluster", "stored", "end"};
      for (int s = 1; s < 12; ++s) fprintf(stderr, " %s=%.0f", names[s], avg[s] / n_ctas);
      fprintf(stderr, "\n");
    }
  }
  return true;
}

// ---- fused chain launch -------------------------------------------------------------------------------------
namespace {
struct ByteRange { const char *lo, *hi; };
inline bool overlaps(const ByteRange &a, const ByteRange &b) { return a.lo < b.hi && b.lo < a.hi; }
inline ByteRange bf16_range(const void *p, int64_t elems) {
  const char *c = static_cast<const char *>(p);
  return ByteRange{c, c + elems * 2};
}
// No layer's weights / bias overlap ANY layer's output, no two outputs overlap, and the chain's input is not one of its
// outputs (byte ranges, not pointer equality: an operand that starts inside another layer's C is a hazard too).
bool chain_operands_hazard_free(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  ByteRange outs[8], bs[8], ds[8];
  if (L > 8) return false;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const int64_t nb = args[l].batch > 0 ? args[l].batch : 1;
    outs[l] = bf16_range(args[l].C, (d.m - 1) * d.ldc + d.n);
    bs[l] = bf16_range(args[l].B, (nb - 1) * d.stride_b + (d.k - 1) * d.ldb + d.n);
    ds[l] = args[l].D ? bf16_range(args[l].D, d.n) : ByteRange{nullptr, nullptr};
  }
  const KernelDesc &d0 = *descs[0];
  const int64_t nb0 = args[0].batch > 0 ? args[0].batch : 1;
  const ByteRange in0 = bf16_range(args[0].A, (nb0 - 1) * d0.stride_a + (d0.m - 1) * d0.lda + d0.k);
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < L; ++j) {
      if (overlaps(bs[l], outs[j])) return false;
      if (ds[l].lo && overlaps(ds[l], outs[j])) return false;
      if (j != l && overlaps(outs[l], outs[j])) return false;
    }
  for (int j = 0; j < L; ++j)
    if (overlaps(in0, outs[j])) return false;
  return true;
}
}  // namespace

// True if layers[0..L) can run in mlp_chain_kernel: every layer is a bf16 tensor-core BRGEMM with beta_0, the same
// m and n, exactly 4 x CHAIN_IPC (batch x k-block) iterations, and layer l+1 reads layer l's C as its A.
// Kernel-independent part: layers[0..L) form a chain - every layer a bf16 tensor-core BRGEMM with beta_0 on the same m
// rows, layer l+1 reads exactly layer l's C as its A, no weight / bias / output buffer is written inside the chain.
// Each chain kernel adds its own shape constraints on top (brgemm_chain_supported, chain_ft_supported,
// chain_pair_supported); a linked chain that no kernel takes is launched layer by layer.
bool brgemm_chain_linked(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == '0'; }();
  if (off || L < 2 || L > CHAIN_MAX_LAYERS) return false;
  const KernelDesc &d0 = *descs[0];
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    if (d.impl != KernelImpl::BrgemmTC || !(d.gemm_flags & 4) || d.m != d0.m) return false;
    if ((d.k % BLOCK_K) != 0 || args[l].batch < 1) return false;
    if (!aligned16(args[l].A) || !aligned16(args[l].B) || !aligned16(args[l].C) || (d.ldc % 8) != 0) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && !(d.binary_kind == 1 && (d.binary_flags & 4))) return false;
    if (l > 0) {
      if (args[l].A != args[l - 1].C || d.lda != descs[l - 1]->ldc) return false;
      if (args[l].batch * d.k != descs[l - 1]->n) return false;
      if (args[l].batch > 1 && d.stride_a != d.k) return false;   // batch element b = columns [b k, b k + k) of C(l-1)
    }
  }
  return chain_operands_hazard_free(descs, args, L);
}

bool brgemm_chain_supported(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == '0'; }();
  if (off || L < 2 || L > CHAIN_MAX_LAYERS) return false;
  const KernelDesc &d0 = *descs[0];
  const int64_t tiles = ((d0.n + 63) / 64) * ((d0.m + BLOCK_M - 1) / BLOCK_M);
  if (tiles * 4 > 128) return false;   // 4-CTA clusters: 33 fit at a time (measured); all CTAs must be co-resident
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    if (d.impl != KernelImpl::BrgemmTC || !(d.gemm_flags & 4) || d.m != d0.m || d.n != d0.n) return false;
    const int64_t k_iters = (d.k + BLOCK_K - 1) / BLOCK_K;
    if ((d.k % BLOCK_K) != 0 || args[l].batch * k_iters != 4 * CHAIN_IPC) return false;
    if (!aligned16(args[l].A) || !aligned16(args[l].B) || !aligned16(args[l].C) || (d.ldc % 8) != 0) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && !(d.binary_kind == 1 && (d.binary_flags & 4))) return false;
    if (l > 0) {
      // the chain link: A(l) is exactly C(l-1), viewed with the same leading dimension
      if (args[l].A != args[l - 1].C || d.lda != descs[l - 1]->ldc) return false;
      if (args[l].batch * d.k != descs[l - 1]->n) return false;
      if (args[l].batch > 1 && d.stride_a != d.k) return false;
    }
  }
  return chain_operands_hazard_free(descs, args, L);   // weights / bias must not be produced inside the chain
}

// Feature-major chain (mlp_chain_ft_kernel): additionally needs m % 32 == 0, n % 64 == 0, a reduction of exactly
// FT_KB k-blocks per layer, bias-add (bcast_col) or no binary, and (m/32) x (n/64) <= 148 co-resident CTAs.
static bool chain_ft_supported(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == 's'; }();
  if (off) return false;
  const KernelDesc &d0 = *descs[0];
  if ((d0.m % FT_N) != 0 || (d0.n % FT_M) != 0) return false;
  if ((d0.m / FT_N) * (d0.n / FT_M) > 148) return false;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const int64_t k_iters = d.k / BLOCK_K, iters = args[l].batch * k_iters;
    if (iters != FT_KB) return false;
    // a group of 4 k-block slots must be one TMA box: 4 k-blocks of one batch element, or whole batch elements
    if (!((k_iters % FT_GROUP) == 0 || k_iters == 1 || k_iters == 2)) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && args[l].D == nullptr) return false;
  }
  return true;
}

// split-K variants (mlp_chain_fts_kernel<S>): 32 S-row batch tiles, a multiple of S feature tiles, and a reduction
// whose S slices are made of whole TMA boxes. Returns the largest usable S in {4, 2}, or 1.
static int chain_ft_split(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  // S = 4 is implemented and parity-clean but slower than S = 2 (7.97 vs 5.71 us per forward): its 24 KiB of st.async
  // pushes per pass move at ~8 B/clk and become the bound. TPP_XSMM_CHAIN_SPLIT=4 enables it, =1 disables split-K.
  static const int max_split = [] { const char *e = getenv("TPP_XSMM_CHAIN_SPLIT"); return e ? atoi(e) : 2; }();
  const KernelDesc &d0 = *descs[0];
  for (int S = 4; S >= 2; S /= 2) {
    if (S > max_split) continue;
    const int rows = 32 * S, kb = FT_KB / S, group = kb / 4;
    if ((d0.m % rows) != 0 || ((d0.n / FT_M) % S) != 0) continue;
    if ((d0.m / rows) * (d0.n / FT_M) * S > 148) continue;
    bool ok = true;
    for (int l = 0; l < L && ok; ++l) {
      const int64_t k_iters = descs[l]->k / BLOCK_K;
      if (!(k_iters == 1 || (k_iters % group) == 0)) ok = false;            // a box = `group` k-blocks of one batch element
      if (k_iters > kb && (k_iters % kb) != 0) ok = false;                   // a k-slice divides a batch element ...
      if (k_iters < kb && (kb % k_iters) != 0) ok = false;                   // ... or is whole batch elements
    }
    (void)args;
    if (ok) return S;
  }
  return 1;
}

namespace {
// operand footprints of one chain: inputs (first layer's A, every layer's B and D) and outputs (every layer's C)
void chain_ranges(const KernelDesc *const *descs, const GemmArgs *args, int L, std::vector<ByteRange> &in,
                  std::vector<ByteRange> &out) {
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    const int64_t nb = g.batch > 0 ? g.batch : 1;
    auto rng = [](const void *p, int64_t elems) {
      const char *c = static_cast<const char *>(p);
      return ByteRange{c, c + elems * 2};
    };
    if (l == 0) in.push_back(rng(g.A, (nb - 1) * d.stride_a + (d.m - 1) * d.lda + d.k));
    in.push_back(rng(g.B, (nb - 1) * d.stride_b + (d.k - 1) * d.ldb + d.n));
    if (g.D) in.push_back(rng(g.D, d.n));
    out.push_back(rng(g.C, (d.m - 1) * d.ldc + d.n));
  }
}
}  // namespace

// Launch chains [0, num_chains) - chain c is layers [first[c], first[c] + len[c]) of descs / args, each already accepted
// by brgemm_chain_supported - as ONE feature-major launch, interleaving pairs of chains. Only a prefix of mutually
// independent, identically tiled chains is taken. Returns the number of chains launched (0: not applicable).
int launch_brgemm_chains_ft(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                            int num_chains, cudaStream_t stream) {
  if (num_chains < 1 || !brgemm_chain_supported(descs + first[0], args + first[0], len[0]) ||
      !chain_ft_supported(descs + first[0], args + first[0], len[0]))
    return 0;
  static const bool multi_off = [] { const char *e = getenv("TPP_XSMM_CHAIN_MULTI"); return e && e[0] == '0'; }();
  const KernelDesc &d0 = *descs[first[0]];
  int split = chain_ft_split(descs + first[0], args + first[0], len[0]);
  // ---- which chains go into this launch ----
  int take = 1, passes = len[0];
  {
    std::vector<ByteRange> in_all, out_all;
    chain_ranges(descs + first[0], args + first[0], len[0], in_all, out_all);
    while (!multi_off && take < num_chains) {
      const int c = take;
      const KernelDesc &d = *descs[first[c]];
      if (d.m != d0.m || d.n != d0.n || passes + len[c] > FT_MAX_PASSES) break;
      if (!brgemm_chain_supported(descs + first[c], args + first[c], len[c]) ||
          !chain_ft_supported(descs + first[c], args + first[c], len[c]))
        break;
      if (chain_ft_split(descs + first[c], args + first[c], len[c]) != split) break;
      std::vector<ByteRange> in, out;
      chain_ranges(descs + first[c], args + first[c], len[c], in, out);
      bool indep = true;
      for (const ByteRange &o : out) {
        for (const ByteRange &x : in_all) indep = indep && !overlaps(o, x);
        for (const ByteRange &x : out_all) indep = indep && !overlaps(o, x);
      }
      for (const ByteRange &i : in)
        for (const ByteRange &x : out_all) indep = indep && !overlaps(i, x);
      if (!indep) break;
      in_all.insert(in_all.end(), in.begin(), in.end());
      out_all.insert(out_all.end(), out.begin(), out.end());
      passes += len[c];
      ++take;
    }
  }
  // one or two chains per launch have nothing to hide the exchange latency behind: the full-K kernel is faster there
  // (11.0 vs 14.7 us for a single forward); every split-K shape is also a full-K shape
  if (take < 3) split = 1;
  const bool split2 = split > 1;
  // ---- the pass list: `ways` chains at a time interleaved layer by layer; a chain's counter slot is its position in
  // the tuple. One chain's layer-to-layer latency (store, fence, counter, poll, TMA: ~5000 clk) is longer than one
  // pass (~3000-4000 clk), so three chains are needed to keep the tensor pipe busy. ----
  static const int ways = [] {
    const char *e = getenv("TPP_XSMM_CHAIN_WAYS");
    const int w = e ? atoi(e) : 3;
    return w < 1 ? 1 : w > FT_MAX_WAYS ? FT_MAX_WAYS : w;
  }();
  static FtParams cp;   // ~20 KiB: too large for the stack of a small thread; launches are serialised per thread anyway
  static std::mutex cp_mutex;
  std::lock_guard<std::mutex> lock(cp_mutex);
  memset(&cp, 0, sizeof(cp));
  uint32_t arrivals[FT_MAX_WAYS] = {0, 0, 0, 0};
  int np = 0;
  bool weights_early = true;
  auto add_pass = [&](int c, int l, int slot) -> bool {
    const KernelDesc &d = *descs[first[c] + l];
    const GemmArgs &g = args[first[c] + l];
    FtPass &ps = cp.pass[np];
    const uint64_t nb = (uint64_t)g.batch;
    const uint32_t k_iters = (uint32_t)(d.k / BLOCK_K);
    const uint32_t grp = split == 4 ? FS<4>::GROUP : split == 2 ? FS<2>::GROUP : FT_GROUP;
    const uint32_t gk = k_iters >= grp ? grp : k_iters, gb = grp / gk;   // box = gk k-blocks x gb batch elements
    if (!encode_map_x4(&ps.tmX, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a,
                       32 * split, gk, gb) ||
        !encode_map(&ps.tmW, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, FT_M,
                    BLOCK_K * gk, gb))
      return false;
    ps.C = g.C;
    ps.D = g.D;
    ps.ldc = d.ldc;
    ps.k_iters = (int32_t)k_iters;
    ps.groups = FT_NG;
    ps.has_bias = (d.op == OpClass::FusedBrgemm && g.D && d.binary_kind == 1) ? 1 : 0;
    ps.relu = (d.op == OpClass::FusedBrgemm && d.unary_kind == 5) ? 1 : 0;
    ps.slot = (uint8_t)slot;
    ps.x_dep = l > 0 ? 1 : 0;
    ps.wait_arrivals = arrivals[ps.slot];
    ps.arrive = l + 1 < len[c] ? 1 : 0;
    if (ps.arrive) ++arrivals[ps.slot];
    if (!g.b_independent) weights_early = false;
    ++np;
    return true;
  };
  bool ok = true;
  for (int c = 0; c < take && ok; c += ways) {
    const int nc = std::min(ways, take - c);
    int maxL = 0;
    for (int j = 0; j < nc; ++j) maxL = std::max(maxL, len[c + j]);
    for (int l = 0; l < maxL && ok; ++l)
      for (int j = 0; j < nc && ok; ++j)
        if (l < len[c + j]) ok = add_pass(c + j, l, j);
  }
  if (!ok) {
    static bool warned = false;
    if (!warned) fprintf(stderr, "tpp-xsmm-cuda: feature-major chain: tensor map encode failed, using the split-K chain\n");
    warned = true;
    return 0;
  }
  dim3 grid((unsigned)(d0.n / FT_M), (unsigned)(d0.m / (32 * split)), (unsigned)split);
  const int n_ctas = (int)(grid.x * grid.y * grid.z);
  // arrival counters of THIS kernel node (the chain kernels only ever run inside a capture): zero-filled before the
  // node exists, owned by the graph, monotonic across its replays - every counter stays a multiple of the group size
  // between launches; the split-K-2 variant's launch epoch (+ exit ticket) lives behind its counters
  unsigned int *counters = static_cast<unsigned int *>(
      capture_owned_zeroed(sizeof(unsigned int) * (FT_MAX_WAYS * FT_CTR_SLOT + 2 * FT_MAX_WAYS)));
  cp.counters = counters;
  cp.epoch = counters + FT_MAX_WAYS * FT_CTR_SLOT;
  for (int sl = 0; sl < FT_MAX_WAYS; ++sl) cp.arrivals_total[sl] = arrivals[sl];
  cp.num_passes = np;
  cp.weights_early = weights_early ? 1 : 0;
  static const bool x0_off = [] { const char *e = getenv("TPP_XSMM_CHAIN_X0"); return e && e[0] == '0'; }();
  cp.x0_early = (args[first[0]].a_independent && !x0_off) ? 1 : 0;
  // fence.proxy.async between the flag observation and the TMA reads costs ~0.3 us per layer and is not needed for
  // data that other SMs fenced to L2 (TMA reads L2); TPP_XSMM_CHAIN_PROXY_FENCE=1 turns it on
  static const bool pf = [] { const char *e = getenv("TPP_XSMM_CHAIN_PROXY_FENCE"); return e && e[0] == '1'; }();
  cp.proxy_fence = pf ? 1 : 0;
  constexpr int smem1 = FT_KB * (FT_X_BYTES + FT_W_BYTES) + (3 * FT_NG + 4) * 8 + 16 + 1024;
  const int smem = split == 4 ? FS<4>::SMEM : split == 2 ? FS<2>::SMEM : smem1;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_fts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS<2>::SMEM));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_fts_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS<4>::SMEM));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(split2 ? F2_THREADS : NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  // TPP_XSMM_CHAIN_MC=1: weight multicast across pairs of batch tiles ((1,2,1) clusters). It halves the L2 reads of
  // the weights but not the bytes each SM receives, and a pass is bound by the latter (~47-53 B/clk per SM):
  // measured 7.15 us (multicast) vs 7.12 us (unicast) per forward, so it stays off by default.
  static const bool mc_on = [] { const char *e = getenv("TPP_XSMM_CHAIN_MC"); return e && e[0] == '1'; }();
  cp.w_multicast = (!split2 && mc_on && (grid.y % 2) == 0) ? 1 : 0;
  if (cp.w_multicast || split2) {
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = 1;
    attrs[1].val.clusterDim.y = split2 ? 1 : 2;
    attrs[1].val.clusterDim.z = split2 ? (unsigned)split : 1;
    cfg.numAttrs = 2;
  }
  static const bool trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 3; }();
  if (trace_on) {
    if (!g_trace_buf) {
      TPP_CUDA_CHECK(cudaMalloc(&g_trace_buf, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemsetAsync(g_trace_buf, 0, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS, stream));
    }
    cp.trace = g_trace_buf;
    g_chain_trace_ctas = n_ctas;
    g_chain_trace_layers = np;
    g_chain_trace_ft = true;
  }
  if (split == 4) TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_fts_kernel<4>, cp));
  else if (split == 2) TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_fts_kernel<2>, cp));
  else TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_ft_kernel, cp));
  const char *tile = split == 4 ? "ft64x128_splitk4" : split == 2 ? "ft64x64_splitk2" : "ft64x32_fullk";
  if (take == 1) snprintf(t_last_name, sizeof(t_last_name), "mlp_chain_bf16_%dlayers_%s", len[0], tile);
  else snprintf(t_last_name, sizeof(t_last_name), "mlp_chain_bf16_%dx%dlayers_%s", take, len[0], tile);
  return take;
}

// ---- pair-per-chain launch ---------------------------------------------------------------------------------------------
namespace {
bool chain_pair_supported(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  const KernelDesc &d0 = *descs[0];
  if ((d0.m % PC_ROWS) != 0 || d0.m > (1 << 30)) return false;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    if ((d.n % PC_BLOCK_N) != 0 || d.n > PC_MAX_TILES * PC_BLOCK_N) return false;
    if ((d.k % BLOCK_K) != 0 || args[l].batch < 1 || (d.ldc % 8) != 0) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && args[l].D == nullptr) return false;
    if (d.op == OpClass::FusedBrgemm && d.unary_kind != 0 && d.unary_kind != 5) return false;
    if (args[l].D && !aligned16(args[l].D)) return false;   // the epilogue reads the bias in 16-byte words
  }
  return true;
}
}  // namespace

void brgemm_tc_take_capture_allocs(std::vector<void *> &out) {
  out.insert(out.end(), t_capture_allocs.begin(), t_capture_allocs.end());
  t_capture_allocs.clear();
  t_capture_ws = CaptureWs();
}

// Launch a prefix of chains [0, num_chains) as ONE launch of mlp_chain_pair_kernel: every chain is cut into blocks of
// 256 batch rows, every block is a work item of one CTA pair. Taken only when the launch carries enough items to
// occupy a useful share of the 74 pairs (a single pair needs ~50 us for a 3 x 1024^2 chain; the pass kernels above
// finish a lone chain in ~11 us). Returns the number of chains launched (0: not applicable).
int launch_brgemm_chains_pair(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                              int num_chains, cudaStream_t stream) {
  static const int min_items = [] {
    const char *e = getenv("TPP_XSMM_CHAIN_PAIR_MIN");   // 0 disables the kernel
    return e ? atoi(e) : 12;
  }();
  if (min_items <= 0 || num_chains < 1) return 0;
  int take = 0;
  int64_t items = 0, layers = 0;
  {
    std::vector<ByteRange> in_all, out_all;
    while (take < num_chains) {
      const int c = take;
      if (!chain_pair_supported(descs + first[c], args + first[c], len[c])) break;
      std::vector<ByteRange> in, out;
      chain_ranges(descs + first[c], args + first[c], len[c], in, out);
      bool indep = true;
      for (const ByteRange &o : out) {
        for (const ByteRange &x : in_all) indep = indep && !overlaps(o, x);
        for (const ByteRange &x : out_all) indep = indep && !overlaps(o, x);
      }
      for (const ByteRange &i : in)
        for (const ByteRange &x : out_all) indep = indep && !overlaps(i, x);
      if (!indep) break;
      in_all.insert(in_all.end(), in.begin(), in.end());
      out_all.insert(out_all.end(), out.begin(), out.end());
      items += descs[first[c]]->m / PC_ROWS;
      layers += len[c];
      ++take;
    }
  }
  if (take == 0 || items < min_items) return 0;
  std::vector<PcLayer> hl((size_t)layers);
  std::vector<PcItem> hi((size_t)items);
  size_t nl = 0, ni = 0;
  for (int c = 0; c < take; ++c) {
    const int32_t layer0 = (int32_t)nl;
    for (int l = 0; l < len[c]; ++l) {
      const KernelDesc &d = *descs[first[c] + l];
      const GemmArgs &g = args[first[c] + l];
      PcLayer &pl = hl[nl++];
      memset(&pl, 0, sizeof(pl));
      const uint64_t nb = (uint64_t)g.batch;
      if (!encode_map(&pl.tmX, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K,
                      BLOCK_M) ||
          !encode_map(&pl.tmW, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64, BLOCK_K) ||
          !encode_map(&pl.tmC, g.C, (uint64_t)d.n, (uint64_t)d.m, 1, (uint64_t)d.ldc, 0, PC_OUT_COLS, BLOCK_M))
        return 0;
      pl.C = g.C;
      pl.D = (d.op == OpClass::FusedBrgemm && g.D && d.binary_kind == 1) ? g.D : nullptr;
      pl.ldc = d.ldc;
      pl.k_iters = (int32_t)(d.k / BLOCK_K);
      pl.total_iters = (int32_t)(g.batch * (d.k / BLOCK_K));
      pl.n_tiles = (int32_t)(d.n / PC_BLOCK_N);
      pl.n = (int32_t)d.n;
      pl.relu = (d.op == OpClass::FusedBrgemm && d.unary_kind == 5) ? 1 : 0;
    }
    for (int64_t r = 0; r < descs[first[c]]->m; r += PC_ROWS) {
      PcItem &pi = hi[ni++];
      pi.layer0 = layer0;
      pi.num_layers = len[c];
      pi.row0 = (int32_t)r;
      pi.pad = 0;
    }
  }
  // the table is written now (not captured): a graph replay only launches the kernel that reads it
  const size_t lbytes = hl.size() * sizeof(PcLayer), ibytes = (hi.size() * sizeof(PcItem) + 127) & ~(size_t)127;
  char *table = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&table, lbytes + ibytes));
  TPP_CUDA_CHECK(cudaMemcpyAsync(table, hl.data(), lbytes, cudaMemcpyHostToDevice, table_stream()));
  TPP_CUDA_CHECK(cudaMemcpyAsync(table + lbytes, hi.data(), hi.size() * sizeof(PcItem), cudaMemcpyHostToDevice, table_stream()));
  TPP_CUDA_CHECK(cudaStreamSynchronize(table_stream()));
  t_capture_allocs.push_back(table);
  PcParams cp;
  cp.layers = reinterpret_cast<const PcLayer *>(table);
  cp.items = reinterpret_cast<const PcItem *>(table + lbytes);
  cp.num_items = (int32_t)items;
  static const bool hints_on = [] { const char *e = getenv("TPP_XSMM_CHAIN_PAIR_HINTS"); return !(e && e[0] == '0'); }();
  cp.l2_hints = hints_on ? 1 : 0;
  static const bool prefetch_on = [] { const char *e = getenv("TPP_XSMM_CHAIN_PAIR_PREFETCH"); return e && e[0] == '1'; }();
  cp.prefetch_w = prefetch_on ? 1 : 0;
  cp.trace = nullptr;
  static const bool pc_trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 4; }();
  if (pc_trace_on) {
    if (!g_pc_trace) {
      TPP_CUDA_CHECK(cudaMalloc(&g_pc_trace, sizeof(unsigned long long) * 2 * 148 * PC_TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemset(g_pc_trace, 0, sizeof(unsigned long long) * 2 * 148 * PC_TRACE_SLOTS));
    }
    cp.trace = g_pc_trace;
  }
  static const int max_pairs = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char *e = getenv("TPP_XSMM_CHAIN_PAIRS");
    const int p = e ? atoi(e) : sms / 2;
    return p < 1 ? 1 : p;
  }();
  // balanced: the fewest pairs that still need the minimal number of rounds
  const int rounds = (int)((items + max_pairs - 1) / max_pairs);
  const int pairs = (int)((items + rounds - 1) / rounds);
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  g_pc_trace_ctas = 2 * pairs;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = PC_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_pair_kernel, cp));
  snprintf(t_last_name, sizeof(t_last_name), "mlp_chain_bf16_%dx%dlayers_pair256x256", (int)items, len[0]);
  return take;
}

bool launch_brgemm_chain(const KernelDesc *const *descs, const GemmArgs *args, int L, cudaStream_t stream) {
  if (!brgemm_chain_supported(descs, args, L)) return false;
  {
    const int first = 0;
    if (launch_brgemm_chains_ft(descs, args, &first, &L, 1, stream) == 1) return true;
  }
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  const KernelDesc &d0 = *descs[0];
  dim3 grid((unsigned)((d0.n + 63) / 64), (unsigned)((d0.m + BLOCK_M - 1) / BLOCK_M), 4);
  const int n_ctas = (int)(grid.x * grid.y * grid.z);
  // per-(thread, stream) exchange workspace + grid counters (same life cycle as the stand-alone kernel's workspace)
  float *ws = capture_owned_ws((size_t)148 * BLOCK_M * 64 * sizeof(float));
  unsigned int *counters = static_cast<unsigned int *>(capture_owned_zeroed(sizeof(unsigned int) * 256));
  cp.grid_counter = counters + n_ctas;   // one counter per grid size: always a multiple of G between launches
  cp.num_layers = L;
  cp.weights_early = 1;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    const uint64_t nb = (uint64_t)g.batch;
    if (!encode_map(&cp.tmA[l], g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K,
                    BLOCK_M) ||
        !encode_map(&cp.tmB[l], g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64, BLOCK_K))
      return false;
    TcParams &p = cp.layer[l];
    p.C = g.C; p.D = g.D;
    p.m = d.m; p.n = d.n; p.ldc = d.ldc;
    p.k_iters = (int32_t)(d.k / BLOCK_K);
    p.total_iters = 4 * CHAIN_IPC;
    p.split_k = 4;
    p.beta0 = 1;
    p.bin_kind = (d.op == OpClass::FusedBrgemm && g.D) ? (int)d.binary_kind : 0;
    p.bin_mode = bin_mode_from_flags(d.binary_flags);
    p.relu = d.op == OpClass::FusedBrgemm && d.unary_kind == 5;
    p.c_vec_ok = 1;
    p.b_early = 0;
    p.flags = nullptr;
    p.ws = ws;
    p.trace = nullptr;
    if (!g.b_independent) cp.weights_early = 0;
  }
  constexpr int smem = CHAIN_IPC * A_STAGE_BYTES + CHAIN_MAX_LAYERS * CHAIN_IPC * B_CHUNK_BYTES +
                       (CHAIN_IPC + CHAIN_MAX_LAYERS + 2) * 8 + 16 + 1024;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 1;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 4;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  // TPP_XSMM_TC_TRACE=3: the chain kernels stamp SM clocks of layer 1 into a buffer that is baked into the captured
  // graph; xsmm_cuda_debug_dump_trace() prints the averages of the last replay (no synchronisation here: this
  // function runs inside a stream capture)
  static const bool trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 3; }();
  if (trace_on) {
    if (!g_trace_buf) {
      TPP_CUDA_CHECK(cudaMalloc(&g_trace_buf, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemsetAsync(g_trace_buf, 0, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS, stream));
    }
    cp.trace = g_trace_buf;
    if (L > 1) cp.layer[1].trace = g_trace_buf;   // splitk_epilogue_l2 stamps slots 8 (pushed) 9 (cluster) 10 (stored)
    g_chain_trace_ctas = n_ctas;
    g_chain_trace_layers = L;
  }
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_kernel, cp));
  snprintf(t_last_name, sizeof(t_last_name), "mlp_chain_bf16_%dlayers_128x64x64_splitk4", L);
  return true;
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
