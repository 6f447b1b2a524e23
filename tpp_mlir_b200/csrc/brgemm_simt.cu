// brgemm_simt.cu - generic batch-reduce GEMM on the FFMA pipe.
//
// The correctness path for everything the tcgen05 kernel does not take:
//   * f32 operands (kind::tf32 would break the 1e-5 f32 tolerance of the
//     reference's tests, so f32 is computed with true fp32 FMAs),
//   * VNNI-packed B ([K/2][N][2], lib/TPP/Transforms/Utils/VNNIUtils.cpp:75-78),
//   * leading dimensions / strides / base pointers that TMA cannot express
//     (not multiples of 16 bytes), e.g. the 6x6x6 and 4x4x4 shapes of
//     test/BF16/Integration/xsmm-brgemm-bf16.mlir / xsmm-ternary-bf16.mlir.
// Semantics: SURVEY.md Appendix A / runtime/Xsmm/XsmmRunnerUtils.cpp:288-457.
// f32 accumulation over all batches and k, post-ops on the accumulator, a single
// rounding at the store.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace tpp {

namespace {

constexpr int BK = 16;   // k per shared-memory step; the CTA tile BM x BN is a template parameter (64 x 64, or 32 x 32 for
                         // the 32-wide tiles of the reference's default tiling: a 64 x 64 CTA would idle on 3/4 of it)

struct SimtParams {
  const void *A, *B, *D;
  void *C;
  int64_t m, n, k, lda, ldb, ldc, stride_a, stride_b, batch;
  int beta0, vnni_b, bin_kind, bin_mode, relu;
  // a grid of tile BRGEMMs in one launch (GemmArgs::grid_*): blockIdx.z = tile index, tile (i, j) of the grid works on
  // A + i a_step, B + j b_step, C + i c_step_n + j c_step_k, D + j d_step (elements); 1 x 1: a plain invoke
  int grid_k;
  int64_t a_step, b_step, c_step_n, c_step_k, d_step, tile0;
};

template <typename T> __device__ __forceinline__ float ldf(const T *p, int64_t i) {
  if constexpr (sizeof(T) == 4) return p[i]; else return bf16_bits_to_f32(p[i]);
}

template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256) brgemm_simt_kernel(SimtParams p) {
  constexpr int TM = BM / 16, TN = BN / 16;          // outputs per thread (16 x 16 threads)
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int64_t tile = p.tile0 + blockIdx.z, ti = tile / p.grid_k, tj = tile - ti * p.grid_k;
  const T *A = static_cast<const T *>(p.A) + ti * p.a_step;
  const T *B = static_cast<const T *>(p.B) + tj * p.b_step;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;

  for (int64_t b = 0; b < p.batch; ++b) {
    const T *Ab = A + b * p.stride_a;
    const T *Bb = B + b * p.stride_b;
    for (int64_t k0 = 0; k0 < p.k; k0 += BK) {
      // A tile: BM rows x 16 k; thread loads BM / 16 elements, k fastest (coalesced on k)
#pragma unroll
      for (int e = 0; e < BM / 16; ++e) {
        const int lin = tid + e * 256;
        const int r = lin / BK, kk = lin % BK;
        const int64_t i = i0 + r, kq = k0 + kk;
        As[kk][r] = (i < p.m && kq < p.k) ? ldf(Ab, i * p.lda + kq) : 0.f;
      }
      // B tile: 16 k x BN cols; j fastest
#pragma unroll
      for (int e = 0; e < BN / 16; ++e) {
        const int lin = tid + e * 256;
        const int kk = lin / BN, c = lin % BN;
        const int64_t j = j0 + c, kq = k0 + kk;
        float v = 0.f;
        if (j < p.n && kq < p.k) {
          v = p.vnni_b ? ldf(Bb, ((kq / p.vnni_b) * p.ldb + j) * p.vnni_b + (kq % p.vnni_b)) : ldf(Bb, kq * p.ldb + j);
        }
        Bs[kk][c] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], bb[TN];
#pragma unroll
        for (int r = 0; r < TM; ++r) a[r] = As[kk][ty * TM + r];
#pragma unroll
        for (int c = 0; c < TN; ++c) bb[c] = Bs[kk][tx * TN + c];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
          for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
      }
      __syncthreads();
    }
  }

  T *C = static_cast<T *>(p.C) + ti * p.c_step_n + tj * p.c_step_k;
  const T *D = p.D ? static_cast<const T *>(p.D) + tj * p.d_step : nullptr;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int64_t i = i0 + ty * TM + r;
    if (i >= p.m) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int64_t j = j0 + tx * TN + c;
      if (j >= p.n) continue;
      float v = acc[r][c];
      if (!p.beta0) v += ldf(C, i * p.ldc + j);
      if (p.bin_kind) {
        const int64_t di = p.bin_mode == kBcastCol   ? j
                           : p.bin_mode == kBcastRow ? i
                           : p.bin_mode == kBcastNone ? i * p.ldc + j
                                                      : 0;
        const float d = ldf(D, di);
        v = p.bin_kind == 1 ? v + d : p.bin_kind == 2 ? v * d : p.bin_kind == 3 ? v - d : v / d;
      }
      if (p.relu) v = relu_f32(v);
      if constexpr (sizeof(T) == 4) C[i * p.ldc + j] = v; else C[i * p.ldc + j] = f32_to_bf16_bits(v);
    }
  }
}

} // namespace

// D's broadcast mode from the fused binary flags (operand 0 is D)
static int bin_mode_from_flags(int64_t f) {
  if (f & 4) return kBcastCol;
  if (f & 1) return kBcastRow;
  if (f & 16) return kBcastScalar;
  return kBcastNone;
}

void launch_brgemm_simt(const KernelDesc &d, const GemmArgs &g, cudaStream_t stream) {
  SimtParams p;
  p.A = g.A; p.B = g.B; p.C = g.C; p.D = g.D;
  p.m = d.m; p.n = d.n; p.k = d.k; p.lda = d.lda; p.ldb = d.ldb; p.ldc = d.ldc;
  p.stride_a = d.stride_a; p.stride_b = d.stride_b; p.batch = g.batch;
  p.beta0 = (d.gemm_flags & 4) != 0;
  p.vnni_b = ((d.gemm_flags & 2048) != 0 && d.dtype == kBF16) ? (d.vnni_factor == 4 ? 4 : 2) : 0;   // VNNI factor, 0 = flat B
  p.bin_kind = (d.op == OpClass::FusedBrgemm && g.D) ? (int)d.binary_kind : 0;
  p.bin_mode = bin_mode_from_flags(d.binary_flags);
  p.relu = d.op == OpClass::FusedBrgemm && d.unary_kind == 5;
  // a layer folded from a regular grid of tile invokes runs as ONE launch, one z-slice per tile
  const int64_t tiles = (int64_t)g.grid_n * g.grid_k;
  p.grid_k = g.grid_k;
  p.a_step = g.grid_n > 1 ? g.a_step : 0;
  p.b_step = g.grid_k > 1 ? g.b_step : 0;
  p.c_step_n = g.grid_n > 1 ? g.c_step_n : 0;
  p.c_step_k = g.grid_k > 1 ? g.c_step_k : 0;
  p.d_step = g.grid_k > 1 ? g.d_step : 0;
  for (int64_t t0 = 0; t0 < tiles; t0 += 65535) {
    p.tile0 = t0;
    const bool small = d.m <= 32 && d.n <= 32;
    const int bm = small ? 32 : 64, bn = small ? 32 : 64;
    dim3 grid((unsigned)((d.n + bn - 1) / bn), (unsigned)((d.m + bm - 1) / bm), (unsigned)std::min<int64_t>(tiles - t0, 65535));
    if (d.dtype == kF32) {
      if (small) brgemm_simt_kernel<float, 32, 32><<<grid, 256, 0, stream>>>(p);
      else brgemm_simt_kernel<float, 64, 64><<<grid, 256, 0, stream>>>(p);
    } else {
      if (small) brgemm_simt_kernel<uint16_t, 32, 32><<<grid, 256, 0, stream>>>(p);
      else brgemm_simt_kernel<uint16_t, 64, 64><<<grid, 256, 0, stream>>>(p);
    }
    TPP_CUDA_CHECK(cudaGetLastError());
  }
}

} // namespace tpp
