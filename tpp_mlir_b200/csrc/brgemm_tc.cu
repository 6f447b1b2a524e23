// brgemm_tc.cu - batch-reduce GEMM / fused BRGEMM on the sm_100a tensor cores.
//
// Replaces the JIT-ed libxsmm BRGEMM the reference dispatches in
// runtime/Xsmm/XsmmRunnerUtils.cpp:308-361 (brgemm) and :385-457 (fused brgemm,
// C = relu(C_in*beta + sum_b A_b*B_b + bias)) for bf16 operands.
//
// Design (B200-first, not a translation of the CPU microkernel):
//   * one CTA per 128 x BLOCK_N output tile; the WHOLE reduction - every k-block
//     of every batch element - accumulates into ONE f32 accumulator in TMEM
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
//     tcgen05.commit releases a slot and finally signals the epilogue.
#include <cuda.h>

#include <mutex>

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

struct TcParams {
  void *C;
  const void *D;
  int64_t m, n, ldc;
  int32_t k_iters;      // ceil(k / BLOCK_K)
  int32_t total_iters;  // batch * k_iters
  int32_t beta0, bin_kind, bin_mode, relu;
  int32_t c_vec_ok;     // C base 16B aligned and ldc % 8 == 0
};

template <int BLOCK_N> struct SmemLayout {
  static constexpr int kBChunks = BLOCK_N / 64;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBChunks * B_CHUNK_BYTES;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
brgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  using L = SmemLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_b + STAGES * L::kBChunks * B_CHUNK_BYTES;
  const uint32_t full_bar = bar_base;                 // STAGES x 8 bytes
  const uint32_t empty_bar = bar_base + STAGES * 8;   // STAGES x 8 bytes
  const uint32_t accum_bar = bar_base + 2 * STAGES * 8;
  const uint32_t tmem_slot = accum_bar + 8;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;

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
    ptx::tmem_alloc(tmem_slot, BLOCK_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int32_t it = 0; it < p.total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
        ptx::mbar_arrive_expect_tx(full_bar + 8 * s, L::kStageBytes);
        const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
        ptx::tma_load_3d(smem_a + s * A_STAGE_BYTES, &tmA, full_bar + 8 * s, kb * BLOCK_K, m0, b);
#pragma unroll
        for (int c = 0; c < L::kBChunks; ++c)
          ptx::tma_load_3d(smem_b + (s * L::kBChunks + c) * B_CHUNK_BYTES, &tmB, full_bar + 8 * s, n0 + c * 64,
                           kb * BLOCK_K, b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, BLOCK_N, /*A K-major*/ 0, /*B MN-major*/ 1);
      for (int32_t it = 0; it < p.total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(full_bar + 8 * s, ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = smem_a + s * A_STAGE_BYTES;
        const uint32_t b_addr = smem_b + s * L::kBChunks * B_CHUNK_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          // A: K-major, 8-row groups 1024 B apart; one UMMA_K slice = 32 B inside the swizzled row
          const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
          // B: MN-major, 64-column atoms B_CHUNK_BYTES apart (LBO), 8-k-row groups 1024 B apart (SBO);
          // one UMMA_K slice = 16 k-rows = 2048 B
          const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
          ptx::umma_bf16(tmem_acc, da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar + 8 * s);   // frees the slot when these MMAs retire
      }
      ptx::umma_commit(accum_bar);             // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int64_t row = (int64_t)m0 + q * 32 + lane;
    if (p.total_iters > 0) {
      ptx::mbar_wait(accum_bar, 0);
      ptx::tc_fence_after_sync();
    }
    const uint16_t *Dp = static_cast<const uint16_t *>(p.D);
    uint16_t *Cp = static_cast<uint16_t *>(p.C);
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      const int64_t col0 = (int64_t)n0 + c;
      if (col0 >= p.n) break;   // warp-uniform
      uint32_t r[32];
      if (p.total_iters > 0) {
        ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
      if (row < p.m) {
        uint16_t *crow = Cp + row * p.ldc + col0;
        const bool full = (col0 + 32 <= p.n);
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        if (!p.beta0) {
          if (full && p.c_vec_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
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
            for (int e = 0; e < 32; ++e)
              if (col0 + e < p.n) v[e] += bf16_bits_to_f32(crow[e]);
          }
        }
        if (p.bin_kind) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
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
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = relu_f32(v[e]);
        }
        if (full && p.c_vec_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
            o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
            o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
            o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4 *>(crow + g * 8) = o;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col0 + e < p.n) crow[e] = f32_to_bf16_bits(v[e]);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, BLOCK_N);
  }
}

// ---- host side ----------------------------------------------------------------

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
                uint64_t stride, uint32_t box_inner, uint32_t box_rows) {
  cuuint64_t dims[3] = {inner, rows, batch};
  // a size-1 batch dimension may carry any legal stride
  uint64_t bstride = stride * 2;
  if (batch <= 1 || bstride == 0) bstride = ld * 2;
  cuuint64_t strides[2] = {ld * 2, bstride};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BLOCK_N, int STAGES> constexpr int smem_bytes() {
  return STAGES * SmemLayout<BLOCK_N>::kStageBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
}

template <int BLOCK_N, int STAGES>
void launch_cfg(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = smem_bytes<BLOCK_N, STAGES>();
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(brgemm_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem));
  });
  brgemm_tc_kernel<BLOCK_N, STAGES><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, p);
  TPP_CUDA_CHECK(cudaGetLastError());
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

void brgemm_tc_configure(KernelDesc &d) {
  // Largest BLOCK_N that still yields >= ~3/4 of the SMs worth of CTAs; otherwise the
  // smallest tile (most CTAs).
  const int64_t tiles_m = (d.m + BLOCK_M - 1) / BLOCK_M;
  int bn = 64;
  for (int cand : {256, 128}) {
    if (tiles_m * ((d.n + cand - 1) / cand) >= 110) { bn = cand; break; }
  }
  d.block_n = bn;
  d.stages = bn == 64 ? 8 : bn == 128 ? 6 : 4;
  snprintf(d.name, sizeof(d.name), "brgemm_tc_bf16_128x%dx64_s%d", bn, d.stages);
}

bool launch_brgemm_tc(const KernelDesc &d, const GemmArgs &g, cudaStream_t stream) {
  if (!aligned16(g.A) || !aligned16(g.B)) return false;
  const int64_t batch = g.batch;
  if (batch > (1ll << 31)) return false;
  CUtensorMap tmA, tmB;
  const uint64_t nb = batch > 0 ? (uint64_t)batch : 1;
  if (!encode_map(&tmA, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K, BLOCK_M))
    return false;
  if (!encode_map(&tmB, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64, BLOCK_K))
    return false;

  TcParams p;
  p.C = g.C;
  p.D = g.D;
  p.m = d.m; p.n = d.n; p.ldc = d.ldc;
  p.k_iters = (int32_t)((d.k + BLOCK_K - 1) / BLOCK_K);
  p.total_iters = (int32_t)(batch * p.k_iters);
  p.beta0 = (d.gemm_flags & 4) != 0;
  p.bin_kind = (d.op == OpClass::FusedBrgemm && g.D) ? (int)d.binary_kind : 0;
  p.bin_mode = bin_mode_from_flags(d.binary_flags);
  p.relu = d.op == OpClass::FusedBrgemm && d.unary_kind == 5;
  p.c_vec_ok = aligned16(g.C) && (d.ldc % 8) == 0;

  dim3 grid((unsigned)((d.n + d.block_n - 1) / d.block_n), (unsigned)((d.m + BLOCK_M - 1) / BLOCK_M), 1);
  switch (d.block_n) {
  case 256: launch_cfg<256, 4>(tmA, tmB, p, grid, stream); break;
  case 128: launch_cfg<128, 6>(tmA, tmB, p, grid, stream); break;
  default: launch_cfg<64, 8>(tmA, tmB, p, grid, stream); break;
  }
  return true;
}

} // namespace tpp
