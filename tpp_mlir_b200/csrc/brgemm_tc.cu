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
#include "tc_common.cuh"
#include "tc_splitk.cuh"

namespace tpp {
using namespace tc;

namespace {

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
// VNNI: B is VNNI-2 packed ([k/2][n][2], what the compiler emits for bf16 by default). TMA drops the raw rows of the
// CTA's columns into the stage (no swizzle, counted on a per-CTA "raw" barrier) and eight converter warps rewrite them in
// place into the swizzled MN-major tile - the scheme of the pair-per-chain kernel (mlp_chain_pair.cu) - and arrive on the
// leader's "full" barrier: no un-interleave pass through HBM in front of the GEMM.
constexpr int TC2_CONV_WARPS = 8;
constexpr int TC2_THREADS_VNNI = NUM_THREADS + 32 * TC2_CONV_WARPS;
template <int BLOCK_N, int STAGES, int SPLITK, bool VNNI = false>
__global__ void __launch_bounds__(VNNI ? TC2_THREADS_VNNI : NUM_THREADS, 1)
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
  const uint32_t raw_full = accum_bar + 8;            // [STAGES] per CTA (VNNI): my raw B rows have landed
  const uint32_t tmem_slot = raw_full + STAGES * 8;
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
      // VNNI: besides the producer's expect_tx arrival, the converter warps of BOTH CTAs arrive once their part is in place
      ptx::mbar_init(full_bar + 8 * s, VNNI ? 1 + 2 * TC2_CONV_WARPS : 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
      ptx::mbar_init(raw_full + 8 * s, 1);
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
        // both CTAs' bytes: A, plus B unless the converter warps deliver it
        if (peer == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, VNNI ? 2 * A_STAGE_BYTES : 2 * kStageBytes);
        ptx::tma_load_3d_pair(smem_a + s * A_STAGE_BYTES, &tmA, leader_full + 8 * s, kb * BLOCK_K, m0, b);
        if constexpr (VNNI) {
          // raw [k/2][n][2] rows of my columns: (element of the row | k pair | batch element), counted on MY raw barrier
          ptx::mbar_arrive_expect_tx(raw_full + 8 * s, kBChunks * B_CHUNK_BYTES);
#pragma unroll
          for (int c = 0; c < kBChunks; ++c)
            ptx::tma_load_3d(smem_b + (s * kBChunks + c) * B_CHUNK_BYTES, &tmB, raw_full + 8 * s,
                             2 * (n0 + (int32_t)peer * HALF_N + c * 64), kb * (BLOCK_K / 2), b);
        } else {
#pragma unroll
          for (int c = 0; c < kBChunks; ++c)
            ptx::tma_load_3d_pair(smem_b + (s * kBChunks + c) * B_CHUNK_BYTES, &tmB, leader_full + 8 * s,
                                  n0 + (int32_t)peer * HALF_N + c * 64, kb * BLOCK_K, b);
        }
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
  } else if (VNNI && warp >= 6) {
    // ===== VNNI-2 converters (both CTAs): raw rows in the stage -> swizzled MN-major tile, in place (see mlp_chain_pair.cu) =====
    const int cw = warp - 6;
    const int row_sub = lane >> 4, half = lane & 1, g8 = vnni_group_of_lane(lane);
    const uint32_t leader_full = ptx::mapa(full_bar, leader_rank);
    constexpr int UNITS = 2 * kBChunks;               // 16-byte pieces per thread and k-block: 2 row groups x chunks
    // software pipeline: the raw rows of k-block i + 1 are fetched into registers before k-block i is rewritten, fenced
    // and signalled, so that the next stage's load latency hides behind this stage's store -> fence -> arrive
    uint4 v[UNITS], vn[UNITS];
    auto fetch = [&](int32_t i, uint4 (&dst)[UNITS]) {
      const uint32_t s = (uint32_t)(i % STAGES), ph = (uint32_t)(i / STAGES) & 1u;
      if (lane == 0) ptx::mbar_wait(raw_full + 8 * s, ph);   // one poller per warp
      __syncwarp();
#pragma unroll
      for (int u = 0; u < UNITS; ++u) {
        const uint32_t R = (uint32_t)((u & 1) * 16 + cw * 2 + row_sub);     // raw row = k pair of the k-block
        const uint32_t src = smem_b + (s * kBChunks + (u >> 1)) * B_CHUNK_BYTES + R * 256u + (uint32_t)(2 * g8 + half) * 16u;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dst[u].x), "=r"(dst[u].y), "=r"(dst[u].z), "=r"(dst[u].w) : "r"(src));
      }
    };
    if (num_iters > 0) fetch(0, vn);
    for (int32_t i = 0; i < num_iters; ++i) {
      const uint32_t s = (uint32_t)(i % STAGES);
#pragma unroll
      for (int u = 0; u < UNITS; ++u) v[u] = vn[u];
      __syncwarp();                                   // every lane has read its rows before any lane overwrites them
      if (i + 1 < num_iters) fetch(i + 1, vn);        // a different stage: nothing this warp still has to write
#pragma unroll
      for (int u = 0; u < UNITS; ++u) {
        const uint32_t lo0 = __byte_perm(v[u].x, v[u].y, 0x5410), lo1 = __byte_perm(v[u].z, v[u].w, 0x5410);
        const uint32_t hi0 = __byte_perm(v[u].x, v[u].y, 0x7632), hi1 = __byte_perm(v[u].z, v[u].w, 0x7632);
        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, half ? lo0 : hi0, 1);
        const uint32_t r1 = __shfl_xor_sync(0xffffffffu, half ? lo1 : hi1, 1);
        const uint32_t o0 = half ? r0 : lo0, o1 = half ? r1 : lo1, o2 = half ? hi0 : r0, o3 = half ? hi1 : r1;
        const uint32_t krow = 2u * (uint32_t)((u & 1) * 16 + cw * 2 + row_sub) + (uint32_t)half;   // k row of the 64 x 64 chunk
        const uint32_t base = smem_b + (s * kBChunks + (u >> 1)) * B_CHUNK_BYTES;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(base + krow * 128u + (((uint32_t)g8 ^ (krow & 7u)) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                     : "memory");
      }
      ptx::fence_proxy_async();                       // my shared-memory writes -> the async proxy (the pair's MMAs)
      __syncwarp();
      if (lane == 0) {
        if (peer == 0) ptx::mbar_arrive(full_bar + 8 * s);
        else ptx::mbar_arrive_remote_relaxed(leader_full + 8 * s);
      }
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
    if (warp >= 2 && warp < 6) {
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


// ---- host side ----------------------------------------------------------------
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
  cudaLaunchAttribute attrs[3];
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

template <int BLOCK_N, int STAGES, int SPLITK, bool VNNI = false>
bool launch_cfg_pair(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + (BLOCK_N / 128) * B_CHUNK_BYTES) + (3 * STAGES + 1) * 8 + 16 + 1024;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(brgemm_tc2_kernel<BLOCK_N, STAGES, SPLITK, VNNI>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(VNNI ? TC2_THREADS_VNNI : NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[3];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = t_pdl_allowed;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = SPLITK == 2 ? (unsigned)p.split_k : 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  if (SPLITK == 3) {
    // the k-slices of a tile meet at an arrival counter in global memory: all CTAs must be resident. The launcher keeps
    // such grids within one wave; the cooperative launch makes that a guarantee (other streams may hold SMs)
    if (!prepare_resident_launch(reinterpret_cast<const void *>(brgemm_tc2_kernel<BLOCK_N, STAGES, SPLITK, VNNI>), &cfg, attrs,
                                 /*only_if_concurrent=*/true))
      return false;
  }
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, brgemm_tc2_kernel<BLOCK_N, STAGES, SPLITK, VNNI>, tmA, tmB, p));
  return true;
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
  // VNNI-2 packed B (g.B points at [batch][k/2][ldb][2], `d` is the flat twin of the caller's descriptor): only the CTA-pair
  // kernel converts it in shared memory; whole k-blocks of k pairs, raw rows on 16-byte boundaries. Otherwise the
  // caller un-interleaves B into a scratch buffer first.
  // The rewrite adds 32 KiB of shared-memory traffic to the 80 KiB a k-block already moves and shared memory is this
  // kernel's bound (measured on 1024^3 x 16: 980 instead of 575 clocks per k-block, 43.7 us against 30.4 us flat), so a long
  // reduction is cheaper with ONE un-interleave pass through HBM in front (43.5 us); the in-kernel path is for the short
  // ones, where that extra launch is what costs (TPP_XSMM_VNNI_NATIVE=1: always, =0: never).
  const bool vnni = g.b_vnni2;
  if (vnni) {
    static const int native_mode = [] { const char *e = getenv("TPP_XSMM_VNNI_NATIVE"); return e ? atoi(e) : -1; }();
    if (mc != 2 || (d.k % BLOCK_K) != 0 || (d.ldb % 4) != 0 || (d.stride_b % 8) != 0 || (d.n % 8) != 0) return false;
    if (native_mode == 0 || (native_mode < 0 && batch * k_iters / split > 24)) return false;
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
  const int map_kind = vnni ? 3 : mc;
  if (e.desc != &d || e.A != g.A || e.B != g.B || e.batch != batch || e.mc != map_kind) {
    const uint64_t nb = batch > 0 ? (uint64_t)batch : 1;
    e.desc = &d; e.A = g.A; e.B = g.B; e.batch = batch; e.mc = map_kind;
    // with multicast every CTA fetches one 64-row half of the A stage
    e.ok = encode_map(&e.tmA, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K,
                      mc == 1 ? BLOCK_M / 2 : BLOCK_M);
    if (e.ok && vnni) {
      // raw VNNI-2 rows: (element of the [n][2] row | k pair | batch element); a box is 64 columns x 2 = 256 bytes per k
      // pair, 32 k pairs; no swizzle (the converter warps produce the swizzled tile)
      const uint64_t dims[3] = {2 * (uint64_t)d.n, (uint64_t)d.k / 2, nb};
      const uint64_t str[2] = {2 * (uint64_t)d.ldb, nb > 1 ? (uint64_t)d.stride_b : 2 * (uint64_t)d.ldb};
      const uint32_t box[3] = {128, BLOCK_K / 2, 1};
      e.ok = encode_map_nd(&e.tmB, g.B, 3, dims, str, box, 0);
    } else if (e.ok) {
      e.ok = encode_map(&e.tmB, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64, BLOCK_K);
    }
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
  set_last_name("brgemm_tc_bf16_%dx%dx64%s%s%s", mc == 2 ? 256 : 128, block_n,
           split == 1 ? "" : split == 2 ? "_splitk2" : "_splitk4", mc == 1 ? "_mc2x2" : mc == 2 ? "_2cta" : "",
           vnni ? "_vnni2" : "");
  if (mc == 2) {
    bool ok = true;
    if (split > 1)
      ok = vnni ? (block_n == 256 ? launch_cfg_pair<256, 6, 3, true>(tmA, tmB, p, grid, stream)
                                  : launch_cfg_pair<128, 8, 3, true>(tmA, tmB, p, grid, stream))
                : (block_n == 256 ? launch_cfg_pair<256, 6, 3>(tmA, tmB, p, grid, stream)
                                  : launch_cfg_pair<128, 8, 3>(tmA, tmB, p, grid, stream));
    if (split == 1 || !ok) {   // no split, or the split grid cannot be co-resident here: one CTA pair per tile, whole reduction
      p.split_k = 1;
      grid.z = 1;
      p.ws = nullptr;
      p.flags = nullptr;
      if (vnni) {
        if (block_n == 256) launch_cfg_pair<256, 6, 0, true>(tmA, tmB, p, grid, stream);
        else launch_cfg_pair<128, 8, 0, true>(tmA, tmB, p, grid, stream);
      } else if (block_n == 256) launch_cfg_pair<256, 6, 0>(tmA, tmB, p, grid, stream);
      else launch_cfg_pair<128, 8, 0>(tmA, tmB, p, grid, stream);
      if (!ok) set_last_name("brgemm_tc_bf16_256x%dx64_2cta%s", block_n, vnni ? "_vnni2" : "");
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


} // namespace tpp
