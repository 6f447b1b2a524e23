// tc_splitk.cuh - split-K exchange epilogues of the tcgen05 BRGEMM kernels (DSMEM and L2-workspace variants); shared by
// brgemm_tc.cu and mlp_chain.cu.
#pragma once
#include "tc_common.cuh"

namespace tpp {
namespace tc {

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


} // namespace tc
} // namespace tpp
